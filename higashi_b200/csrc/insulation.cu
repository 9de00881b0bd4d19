// insulation.cu -- first consumer of the imputed scores on the device: single-cell insulation score and TAD
// boundary calls straight from the score block of a batch of cells, without ever materialising the n x n contact maps.
//
//   scTAD.py:91-104   m[x, y] += proba ; m = m + m^T ; m *= mask ; m = sqrt_norm(m) ; score = insulation_score(m) ;
//                     score[discard_rows] = 1 ; border = call_tads(score)
//   sqrt_norm         Higashi_analysis.py:581-588    m[a, b] / (sqrt(rowsum a) sqrt(rowsum b)), non-finite -> 0
//   insulation_score  Higashi_TAD.py:7-19            sum(m[i-w:i, i+1:min(n-1, i+w+1)]) / sum(m[i-w:i+w+1, i-w:i+w+1]), nan -> 1
//   call_tads         Higashi_TAD.py:22-27           strict local minima over +-order bins (edges clipped), 0 < score < 0.999
//
// The matrix is never built: a pair (a < b) with normalised value v contributes v to the numerator of the bins
// i in [max(a+1, b-w), min(a+w, b-1)] (b <= n-2, the reference's column bound) and 2v to the denominator of the bins
// i in [b-w, a+w] -- two range updates on per-cell difference arrays, then one prefix sum per cell.
//
// The difference arrays hold 64-bit FIXED-POINT integers (v * 2^40, v <= 1 after sqrt_norm): integer atomics make the result
// independent of the order the atomics land in, and -- what the boundary calls need -- sums that are mathematically zero or
// mathematically equal come out exactly zero / exactly equal (the reference's last two bins have an empty numerator block and
// score exactly 0, which `score > 0` then rejects; a floating-point prefix sum would leave a 1e-17 residue there).
// Wrap-around arithmetic is exact as long as every window sum fits: (2w+1)^2 * 2 * 2^40 < 2^63  <=>  w <= 1000.
#include "common.cuh"

#define INS_FIX_SCALE 1099511627776.0      // 2^40
__device__ __forceinline__ void fix_add(unsigned long long* p, long long q) { atomicAdd(p, (unsigned long long)q); }

// rowsum[cell][a] += v, rowsum[cell][b] += v   (symmetric matrix, masked rows / columns dropped)
__global__ void ins_rowsum_kernel(const float* __restrict__ scores, int64_t ld, int64_t off, const int2* __restrict__ pairs, int64_t P,
                                  int goff, int n, int n_cells, const unsigned char* __restrict__ keep, double* __restrict__ rowsum) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int cell = blockIdx.y;
    if (i >= P || cell >= n_cells) return;
    const int2 pr = pairs[off + i];
    const int a = pr.x - goff, b = pr.y - goff;
    if (keep && (!keep[a] || !keep[b])) return;
    const double v = (double)scores[(int64_t)cell * ld + off + i];
    if (v == 0.0) return;
    atomicAdd(rowsum + (int64_t)cell * n + a, v);
    if (b != a) atomicAdd(rowsum + (int64_t)cell * n + b, v); else atomicAdd(rowsum + (int64_t)cell * n + a, v);
}

// range updates of the difference arrays: diff[cell][0][.] numerator, diff[cell][1][.] denominator (n + 1 entries each)
__global__ void ins_diff_kernel(const float* __restrict__ scores, int64_t ld, int64_t off, const int2* __restrict__ pairs, int64_t P,
                                int goff, int n, int n_cells, int w, const unsigned char* __restrict__ keep,
                                const double* __restrict__ rowsum, unsigned long long* __restrict__ diff) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int cell = blockIdx.y;
    if (i >= P || cell >= n_cells) return;
    const int2 pr = pairs[off + i];
    int a = pr.x - goff, b = pr.y - goff;
    if (keep && (!keep[a] || !keep[b])) return;
    if (a > b) { int t = a; a = b; b = t; }
    const double raw = (double)scores[(int64_t)cell * ld + off + i];
    const double ra = rowsum[(int64_t)cell * n + a], rb = rowsum[(int64_t)cell * n + b];
    if (raw == 0.0 || !(ra > 0.0) || !(rb > 0.0)) return;          // sqrt_norm: division by a zero coverage -> 0
    const double v = raw / sqrt(ra) / sqrt(rb);
    if (!isfinite(v)) return;
    const long long q = llrint(v * INS_FIX_SCALE);
    if (q == 0) return;
    unsigned long long* num = diff + (int64_t)cell * 2 * (n + 1);
    unsigned long long* den = num + (n + 1);
    if (a == b) {                                                    // a diagonal entry counts once, in the denominator only
        const int lo = max(0, a - w), hi = min(n - 1, a + w);
        fix_add(den + lo, 2 * q); fix_add(den + hi + 1, -2 * q);     // m + m^T doubles the diagonal
        return;
    }
    {   // denominator: both (a, b) and (b, a) lie inside the window of bin i  <=>  b - w <= i <= a + w
        const int lo = max(0, b - w), hi = min(n - 1, a + w);
        if (lo <= hi) { fix_add(den + lo, 2 * q); fix_add(den + hi + 1, -2 * q); }
    }
    if (b <= n - 2) {   // numerator: a in [i - w, i), b in [i + 1, min(n - 1, i + w + 1))
        const int lo = max(a + 1, b - w), hi = min(a + w, b - 1);
        if (lo <= hi) { fix_add(num + lo, q); fix_add(num + hi + 1, -q); }
    }
}

// one block per cell: prefix sums, ratio, discarded rows, boundary calls
__global__ void ins_finish_kernel(const unsigned long long* __restrict__ diff, int n, int order, const unsigned char* __restrict__ discard,
                                  float* __restrict__ score_out, unsigned char* __restrict__ border_out, double* __restrict__ score64) {
    const int cell = blockIdx.x;
    const unsigned long long* num = diff + (int64_t)cell * 2 * (n + 1);
    const unsigned long long* den = num + (n + 1);
    double* sc = score64 + (int64_t)cell * n;
    if (threadIdx.x == 0) {                       // n is a few thousand: one serial (exact, integer) scan per cell
        unsigned long long sn = 0, sd = 0;
        for (int i = 0; i < n; ++i) {
            sn += num[i]; sd += den[i];
            double v = (double)(long long)sn / (double)(long long)sd;
            if (isnan(v) || sd == 0) v = 1.0;     // 0 / 0 -> nan -> 1 (the numerator is a subset of the denominator)
            if (discard && discard[i]) v = 1.0;
            sc[i] = v;
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double v = sc[i];
        bool ok = v < 0.999 && v > 0.0;
        for (int k = 1; k <= order && ok; ++k) {
            const int r = min(i + k, n - 1), l = max(i - k, 0);
            ok = v < sc[r] && v < sc[l];
        }
        score_out[(int64_t)cell * n + i] = (float)v;
        border_out[(int64_t)cell * n + i] = ok ? 1 : 0;
    }
}

extern "C" int hsg_insulation(hsg_ctx* c, int sel_slot, const float* scores_dev, int n_batch, int window_bins, int tad_order,
                              const unsigned char* keep_host, const unsigned char* discard_host, float* score_out_host,
                              unsigned char* border_out_host, void* stream) {
    HSG_CHECK(c && c->d > 0, "context not initialised");
    HSG_CUDA(cudaSetDevice(c->device));
    HSG_CHECK(sel_slot >= 0 && sel_slot < (int)c->sel_chroms.size(), "bad chromosome slot (hsg_set_pairs first)");
    HSG_CHECK(scores_dev && n_batch > 0 && window_bins > 0 && tad_order >= 0 && score_out_host && border_out_host, "bad arguments");
    HSG_CHECK(window_bins <= 1000, "insulation window above 1000 bins");
    cudaStream_t s = (cudaStream_t)stream;
    const int chrom = c->sel_chroms[sel_slot], n = c->bins[chrom], goff = c->bin_off[chrom];
    const int64_t off = c->sel_off[sel_slot], P = c->sel_P[sel_slot];
    double *rowsum = nullptr, *sc64 = nullptr;
    unsigned long long* diff = nullptr;
    float* d_score = nullptr;
    unsigned char *d_border = nullptr, *d_keep = nullptr, *d_discard = nullptr;
    auto cleanup = [&]() {
        cudaFree(rowsum); cudaFree(diff); cudaFree(sc64); cudaFree(d_score); cudaFree(d_border); cudaFree(d_keep); cudaFree(d_discard);
    };
    const size_t nb = (size_t)n_batch;
#define INS_CUDA(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { cleanup(); hsg_set_error(std::string("CUDA error: ") + cudaGetErrorString(e_)); return 1; } } while (0)
    INS_CUDA(cudaMalloc((void**)&rowsum, nb * n * sizeof(double)));
    INS_CUDA(cudaMalloc((void**)&diff, nb * 2 * (n + 1) * sizeof(unsigned long long)));
    INS_CUDA(cudaMalloc((void**)&sc64, nb * n * sizeof(double)));
    INS_CUDA(cudaMalloc((void**)&d_score, nb * n * sizeof(float)));
    INS_CUDA(cudaMalloc((void**)&d_border, nb * n));
    INS_CUDA(cudaMemsetAsync(rowsum, 0, nb * n * sizeof(double), s));
    INS_CUDA(cudaMemsetAsync(diff, 0, nb * 2 * (n + 1) * sizeof(unsigned long long), s));
    if (keep_host) {
        INS_CUDA(cudaMalloc((void**)&d_keep, n));
        INS_CUDA(cudaMemcpyAsync(d_keep, keep_host, n, cudaMemcpyHostToDevice, s));
    }
    if (discard_host) {
        INS_CUDA(cudaMalloc((void**)&d_discard, n));
        INS_CUDA(cudaMemcpyAsync(d_discard, discard_host, n, cudaMemcpyHostToDevice, s));
    }
    if (P > 0) {
        dim3 grid((unsigned)((P + 255) / 256), (unsigned)n_batch);
        ins_rowsum_kernel<<<grid, 256, 0, s>>>(scores_dev, c->P, off, c->pairs, P, goff, n, n_batch, d_keep, rowsum);
        c->launches++;
        INS_CUDA(cudaGetLastError());
        ins_diff_kernel<<<grid, 256, 0, s>>>(scores_dev, c->P, off, c->pairs, P, goff, n, n_batch, window_bins, d_keep, rowsum, diff);
        c->launches++;
        INS_CUDA(cudaGetLastError());
    }
    ins_finish_kernel<<<n_batch, 256, 0, s>>>(diff, n, tad_order, d_discard, d_score, d_border, sc64);
    c->launches++;
    INS_CUDA(cudaGetLastError());
    INS_CUDA(cudaMemcpyAsync(score_out_host, d_score, nb * n * sizeof(float), cudaMemcpyDeviceToHost, s));
    INS_CUDA(cudaMemcpyAsync(border_out_host, d_border, nb * n, cudaMemcpyDeviceToHost, s));
    INS_CUDA(cudaStreamSynchronize(s));
#undef INS_CUDA
    cleanup();
    return 0;
}
