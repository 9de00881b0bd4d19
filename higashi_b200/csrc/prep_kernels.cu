// prep_kernels.cu -- once-per-model kernels: off-hook embedding tables, cell/bin static tokens,
// cell-invariant pair bias.  None of these is on the per-triplet hot loop.
#include "common.cuh"

// ---------------------------------------------------------------------------------------------
// off-hook table: out[r,:] = act(x[r,:] . W^T + b)      (Modules.py:235-253, 605-631)
// ---------------------------------------------------------------------------------------------
template <int R>
__global__ void embed_table_kernel(const float* __restrict__ x, int64_t rows, int in_dim,
                                   const float* __restrict__ w, const float* __restrict__ b, int d,
                                   int apply_swish, float* __restrict__ out) {
    extern __shared__ float sx[];               // [R][in_dim]
    const int64_t r0 = (int64_t)blockIdx.x * R;
    for (int i = threadIdx.x; i < R * in_dim; i += blockDim.x) {
        int rr = i / in_dim, t = i - rr * in_dim;
        sx[i] = (r0 + rr < rows) ? x[(r0 + rr) * in_dim + t] : 0.0f;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    for (int j = warp; j < d; j += nwarps) {
        float acc[R];
#pragma unroll
        for (int rr = 0; rr < R; ++rr) acc[rr] = 0.0f;
        const float* wj = w + (int64_t)j * in_dim;
        for (int t = lane; t < in_dim; t += 32) {
            float wv = __ldg(wj + t);
#pragma unroll
            for (int rr = 0; rr < R; ++rr) acc[rr] = fmaf(wv, sx[rr * in_dim + t], acc[rr]);
        }
#pragma unroll
        for (int rr = 0; rr < R; ++rr) acc[rr] = warp_sum(acc[rr]);
        if (lane == 0) {
            float bj = b[j];
#pragma unroll
            for (int rr = 0; rr < R; ++rr) {
                if (r0 + rr < rows) {
                    float v = acc[rr] + bj;
                    out[(r0 + rr) * d + j] = apply_swish ? swish_acc(v) : v;
                }
            }
        }
    }
}

int launch_embed_table(hsg_ctx* c, const float* x, int64_t rows, int in_dim, const float* w, const float* b,
                       int apply_swish, float* out, cudaStream_t s) {
    constexpr int R = 4;
    size_t smem = (size_t)R * in_dim * sizeof(float);
    HSG_CHECK(smem <= 200 * 1024, "embed_table: in_dim too large for the staging buffer");
    HSG_CUDA(cudaFuncSetAttribute(embed_table_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    unsigned grid = (unsigned)((rows + R - 1) / R);
    embed_table_kernel<R><<<grid, 256, smem, s>>>(x, rows, in_dim, w, b, c->d, apply_swish, out);
    c->launches++;
    HSG_CUDA(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------
// static tokens: for a table row e (cell or bin embedding)
//   V = w_vs . LN3(e)                      (Modules.py:1037,1045)
//   S = layer_norm2( fc2 . V )             (Modules.py:1089, 753)
//   optionally Q = w_qs . LN1(e), K = w_ks . LN2(e)  (cell token: dynamic == static == E_cell)
// One block (128 threads) per row.
// ---------------------------------------------------------------------------------------------
template <int D>
__global__ void static_tokens_kernel(const float* __restrict__ E, int64_t rows,
                                     const float* __restrict__ g1, const float* __restrict__ b1,
                                     const float* __restrict__ g2, const float* __restrict__ b2,
                                     const float* __restrict__ g3, const float* __restrict__ b3,
                                     const float* __restrict__ wqT, const float* __restrict__ wkT,
                                     const float* __restrict__ wvT, const float* __restrict__ fc2T,
                                     const float* __restrict__ lg2, const float* __restrict__ lb2,
                                     float* __restrict__ Q, float* __restrict__ K, float* __restrict__ V,
                                     float* __restrict__ S) {
    __shared__ float xhat[D];
    __shared__ float sv[HSG_HD];
    __shared__ float red[8];
    __shared__ float sraw[D];
    const int64_t row = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float e = (tid < D) ? E[row * D + tid] : 0.0f;
    // mean
    float s = warp_sum(e);
    if (lane == 0) red[warp] = s;
    __syncthreads();
    float mean = (red[0] + red[1] + red[2] + red[3]) / (float)D;
    __syncthreads();
    float dv = (tid < D) ? (e - mean) : 0.0f;
    float s2 = warp_sum(dv * dv);
    if (lane == 0) red[warp] = s2;
    __syncthreads();
    float var = (red[0] + red[1] + red[2] + red[3]) / (float)D;
    float rstd = 1.0f / sqrtf(var + HSG_LN_EPS);
    if (tid < D) xhat[tid] = dv * rstd;
    __syncthreads();
    // Q, K (only for cell tokens), V: thread j -> output j of 128
    float q = 0.f, k = 0.f, v = 0.f;
    for (int t = 0; t < D; ++t) {
        float xh = xhat[t];
        v = fmaf(wvT[t * HSG_HD + tid], fmaf(xh, g3[t], b3[t]), v);
        if (Q != nullptr) {
            q = fmaf(wqT[t * HSG_HD + tid], fmaf(xh, g1[t], b1[t]), q);
            k = fmaf(wkT[t * HSG_HD + tid], fmaf(xh, g2[t], b2[t]), k);
        }
    }
    V[row * HSG_HD + tid] = v;
    sv[tid] = v;
    if (Q != nullptr) { Q[row * HSG_HD + tid] = q; K[row * HSG_HD + tid] = k; }
    __syncthreads();
    // static' = fc2 . V  then top-level layer_norm2
    float sr = 0.f;
    if (tid < D) {
        for (int t = 0; t < HSG_HD; ++t) sr = fmaf(fc2T[t * D + tid], sv[t], sr);
    }
    __syncthreads();
    float m2 = warp_sum((tid < D) ? sr : 0.f);
    if (lane == 0) red[warp] = m2;
    __syncthreads();
    float mean2 = (red[0] + red[1] + red[2] + red[3]) / (float)D;
    __syncthreads();
    float d2 = (tid < D) ? (sr - mean2) : 0.f;
    float v2 = warp_sum(d2 * d2);
    if (lane == 0) red[warp] = v2;
    __syncthreads();
    float var2 = (red[0] + red[1] + red[2] + red[3]) / (float)D;
    float rstd2 = 1.0f / sqrtf(var2 + HSG_LN_EPS);
    if (tid < D) S[row * D + tid] = fmaf(d2 * rstd2, lg2[tid], lb2[tid]);
    (void)sraw;
}

int launch_static_tokens(hsg_ctx* c, const float* E, int64_t rows, float* Q, float* K, float* V, float* S,
                         cudaStream_t s) {
    if (rows == 0) return 0;
    float** w = c->w;
#define ARGS E, rows, w[HSG_W_ALN1_G], w[HSG_W_ALN1_B], w[HSG_W_ALN2_G], w[HSG_W_ALN2_B], w[HSG_W_ALN3_G],   \
             w[HSG_W_ALN3_B], c->dw.wqT, c->dw.wkT, c->dw.wvT, c->dw.fc2T, w[HSG_W_LN2_G], w[HSG_W_LN2_B], Q, K, V, S
    if (c->d == 64) static_tokens_kernel<64><<<(unsigned)rows, 128, 0, s>>>(ARGS);
    else if (c->d == 128) static_tokens_kernel<128><<<(unsigned)rows, 128, 0, s>>>(ARGS);
    else { hsg_set_error("d_model must be 64 or 128"); return 2; }
#undef ARGS
    c->launches++;
    HSG_CUDA(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------
// pair bias: extra_proba2([attr[b_i] | attr[b_j] | 0_cf])   (Modules.py:735-740, only_model=True)
// ---------------------------------------------------------------------------------------------
__global__ void pair_bias_kernel(const int2* __restrict__ pairs, int64_t P, int n_cells, int C, int cf,
                                 const float* __restrict__ attr, const float* __restrict__ w0,
                                 const float* __restrict__ b0, const float* __restrict__ w1,
                                 const float* __restrict__ b1, float* __restrict__ out) {
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    int2 pr = pairs[p];
    const int A = C + 1, in_w = 2 * A + cf;
    const float* ai = attr + (int64_t)(n_cells + 1 + pr.x) * A;
    const float* aj = attr + (int64_t)(n_cells + 1 + pr.y) * A;
    float o = b1[0];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        float acc = 0.0f;
        const float* wr = w0 + u * in_w;
        for (int t = 0; t < A; ++t) acc = fmaf(wr[t], ai[t], acc);
        for (int t = 0; t < A; ++t) acc = fmaf(wr[A + t], aj[t], acc);
        acc += b0[u];
        o = fmaf(w1[u], swish_acc(acc), o);
    }
    out[p] = o;
}

int launch_pair_bias(hsg_ctx* c, cudaStream_t s) {
    if (c->P == 0) return 0;
    unsigned grid = (unsigned)((c->P + 255) / 256);
    pair_bias_kernel<<<grid, 256, 0, s>>>(c->pairs, c->P, c->n_cells, c->n_chrom, c->cf, c->w[HSG_W_ATTR],
                                           c->w[HSG_W_XP2_W0], c->w[HSG_W_XP2_B0], c->w[HSG_W_XP2_W1],
                                           c->w[HSG_W_XP2_B1], c->pair_bias);
    c->launches++;
    HSG_CUDA(cudaGetLastError());
    return 0;
}
