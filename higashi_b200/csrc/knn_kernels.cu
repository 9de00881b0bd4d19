// knn_kernels.cu -- kNN cell graph on the device: Higashi.get_cell_neighbor (Higashi_wrapper.py:1303-1324).
//
// reference:  distance = sklearn pairwise_distances(v, 'euclidean')  (float32 in: squared norms and the cross term are
//             accumulated in float64, clipped at 0, diagonal zeroed, cast to float32, sqrt in float32)
//             distance /= np.quantile(np.sort(distance)[:, 1:k1].ravel(), 0.25)
//             neighbor = argsort(distance)[:k1] ; w = exp(-distance[neighbor]) / sum
// Here: one block per cell computes its distance row (fp64 accumulation, fp32 sqrt), then extracts the k1 smallest
// (ties -> lower index).  The 25 % quantile of the n*(k1-1) selected distances and the weights are finished on the host
// from the [n, k1] result (a few kilobytes).
#include <algorithm>
#include <cmath>
#include <vector>

#include "common.cuh"

__global__ void knn_rows_kernel(const float* __restrict__ emb, int n, int dim, int k1, float* __restrict__ scratch,
                                int32_t* __restrict__ nbr, float* __restrict__ dist) {
    extern __shared__ double sx[];               // this cell's embedding in fp64
    __shared__ float red_v[32];
    __shared__ int red_i[32];
    const int i = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
    for (int t = tid; t < dim; t += blockDim.x) sx[t] = (double)emb[(size_t)i * dim + t];
    __syncthreads();
    double xx = 0.0;
    for (int t = 0; t < dim; ++t) xx += sx[t] * sx[t];
    float* row = scratch + (size_t)i * n;
    for (int j = tid; j < n; j += blockDim.x) {
        const float* y = emb + (size_t)j * dim;
        double yy = 0.0, xy = 0.0;
        for (int t = 0; t < dim; ++t) { double v = (double)__ldg(y + t); yy += v * v; xy += sx[t] * v; }
        double d2 = xx + yy - 2.0 * xy;
        if (d2 < 0.0) d2 = 0.0;
        float f = (j == i) ? 0.0f : (float)d2;
        row[j] = sqrtf(f);
    }
    __syncthreads();
    for (int r = 0; r < k1; ++r) {
        float bv = INFINITY; int bi = 0x7fffffff;
        for (int j = tid; j < n; j += blockDim.x) {
            float v = row[j];
            if (v < bv || (v == bv && j < bi)) { bv = v; bi = j; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            float ov = __shfl_xor_sync(0xffffffffu, bv, o); int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov < bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        if (lane == 0) { red_v[warp] = bv; red_i[warp] = bi; }
        __syncthreads();
        if (warp == 0) {
            bv = (lane < nw) ? red_v[lane] : INFINITY; bi = (lane < nw) ? red_i[lane] : 0x7fffffff;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                float ov = __shfl_xor_sync(0xffffffffu, bv, o); int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ov < bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
            }
            if (lane == 0) {
                nbr[(size_t)i * k1 + r] = bi; dist[(size_t)i * k1 + r] = bv;
                row[bi] = INFINITY;              // remove from the next rounds
            }
        }
        __syncthreads();
    }
}

// numpy.quantile(a, 0.25) with the default 'linear' method on a float32 array
static float np_quantile25(std::vector<float>& a) {
    std::sort(a.begin(), a.end());
    const size_t N = a.size();
    double pos = 0.25 * (double)(N - 1);
    size_t lo = (size_t)std::floor(pos);
    size_t hi = std::min(lo + 1, N - 1);
    double t = pos - (double)lo;
    double A = a[lo], B = a[hi];
    double r = (t >= 0.5) ? B - (B - A) * (1.0 - t) : A + (B - A) * t;     // numpy _lerp
    return (float)r;
}

int knn_build(hsg_ctx* c, const float* embed_host, int dim, int k1, int32_t* nbr_out, float* w_out) {
    const int n = c->n_cells;
    HSG_CHECK(k1 >= 1 && k1 <= n, "k+1 must be in [1, n_cells]");
    float *d_emb = nullptr, *d_scratch = nullptr, *d_dist = nullptr;
    int32_t* d_nbr = nullptr;
    // every error path frees what was allocated before it (the n x n scratch is the allocation most likely to fail)
    cudaError_t e = cudaMalloc(&d_emb, (size_t)n * dim * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&d_scratch, (size_t)n * n * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&d_dist, (size_t)n * k1 * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&d_nbr, (size_t)n * k1 * sizeof(int32_t));
    if (e == cudaSuccess) e = cudaMemcpy(d_emb, embed_host, (size_t)n * dim * sizeof(float), cudaMemcpyHostToDevice);
    std::vector<float> dist((size_t)n * k1);
    if (e == cudaSuccess) {
        knn_rows_kernel<<<n, 256, dim * sizeof(double), c->stream>>>(d_emb, n, dim, k1, d_scratch, d_nbr, d_dist);
        c->launches++;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(dist.data(), d_dist, dist.size() * sizeof(float), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(nbr_out, d_nbr, (size_t)n * k1 * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(d_emb); cudaFree(d_scratch); cudaFree(d_dist); cudaFree(d_nbr);
    if (e != cudaSuccess) { hsg_set_error(std::string("knn: ") + cudaGetErrorString(e)); return 1; }
    float scale = 1.0f;
    if (k1 > 1) {
        std::vector<float> sel;
        sel.reserve((size_t)n * (k1 - 1));
        for (int i = 0; i < n; ++i)
            for (int r = 1; r < k1; ++r) sel.push_back(dist[(size_t)i * k1 + r]);
        scale = np_quantile25(sel);
    }
    for (int i = 0; i < n; ++i) {
        float sum = 0.f;
        for (int r = 0; r < k1; ++r) {
            float v = expf(-(dist[(size_t)i * k1 + r] / scale));
            w_out[(size_t)i * k1 + r] = v;
            sum += v;
        }
        for (int r = 0; r < k1; ++r) w_out[(size_t)i * k1 + r] /= sum;
    }
    return 0;
}
