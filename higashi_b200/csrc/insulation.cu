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

// rowsum[cell][a] += v, rowsum[cell][b] += v   (symmetric matrix, masked rows / columns dropped), in 64-bit fixed point
// (v * 2^30; |score| < 1e5 keeps a row of 8192 bins inside 63 bits) so that shared-memory and global integer atomics can be mixed
// and the sums do not depend on their order.
// A block owns INS_CHUNK consecutive pairs.  Pairs are stored row-major (a fixed, b ascending), so a chunk touches the bins of a
// narrow window behind its first row: both contributions go to a shared-memory window first (out-of-window bins straight to
// global memory) and the window is flushed with one global atomic per touched bin -- ~8x fewer global atomics than one per pair.
// The lanes of a warp mostly share `a`: their contributions to rowsum[a] are summed with a segmented shuffle scan and the last
// lane of each run issues ONE atomic; the b's of a run are consecutive bins (no contention inside the warp).
#define INS_CHUNK 4096
#define INS_WIN 1024
#define INS_ROW_SCALE 1073741824.0          // 2^30
__global__ void __launch_bounds__(256) ins_rowsum_kernel(const float* __restrict__ scores, int64_t ld, int64_t off,
                                                         const int2* __restrict__ pairs, int64_t P, int goff, int n, int n_cells,
                                                         const unsigned char* __restrict__ keep, unsigned long long* __restrict__ rowsum) {
    __shared__ unsigned long long win[INS_WIN];
    const int cell = blockIdx.y, lane = threadIdx.x & 31;
    const int64_t chunk0 = (int64_t)blockIdx.x * INS_CHUNK;
    for (int j = threadIdx.x; j < INS_WIN; j += 256) win[j] = 0ull;
    const int base = pairs[off + chunk0].x - goff;          // chunk0 < P by construction of the grid
    __syncthreads();
    unsigned long long* rs = rowsum + (int64_t)cell * n;
    for (int k = 0; k < INS_CHUNK / 256; ++k) {
        const int64_t i = chunk0 + (int64_t)k * 256 + threadIdx.x;
        int a = -1, b = -1;
        long long q = 0;
        if (i < P) {
            const int2 pr = pairs[off + i];
            a = pr.x - goff; b = pr.y - goff;
            if (!keep || (keep[a] && keep[b])) q = llrint((double)scores[(int64_t)cell * ld + off + i] * INS_ROW_SCALE);
        }
        if (q != 0) {                                       // a diagonal pair (a == b) counts twice: once here, once below
            const unsigned ib = (unsigned)(b - base);
            if (ib < INS_WIN) atomicAdd(&win[ib], (unsigned long long)q); else atomicAdd(rs + b, (unsigned long long)q);
        }
        // segmented inclusive scan over runs of equal a (lanes are in pair order): the LAST lane of a run holds its sum
        long long sum = q;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const long long up = __shfl_up_sync(0xffffffffu, sum, d);
            const int au = __shfl_up_sync(0xffffffffu, a, d);
            if (lane >= d && au == a) sum += up;
        }
        const int an = __shfl_down_sync(0xffffffffu, a, 1);
        if (a >= 0 && (lane == 31 || an != a) && sum != 0) {
            const unsigned ia = (unsigned)(a - base);
            if (ia < INS_WIN) atomicAdd(&win[ia], (unsigned long long)sum); else atomicAdd(rs + a, (unsigned long long)sum);
        }
    }
    __syncthreads();
    for (int j = threadIdx.x; j < INS_WIN; j += 256) {
        const unsigned long long v = win[j];
        if (v != 0ull && base + j < n) atomicAdd(rs + base + j, v);
    }
}

// inv[cell][a] = 1 / sqrt(rowsum[cell][a])  (0 where the coverage is not positive: sqrt_norm maps those rows to 0)
__global__ void ins_invsqrt_kernel(const unsigned long long* __restrict__ rowsum, int64_t total, double* __restrict__ inv) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const double r = (double)(long long)rowsum[i] * (1.0 / INS_ROW_SCALE);
    inv[i] = (r > 0.0) ? 1.0 / sqrt(r) : 0.0;
}

// range updates of the difference arrays: diff[cell][0][.] numerator, diff[cell][1][.] denominator (n + 1 entries each)
__global__ void ins_diff_kernel(const float* __restrict__ scores, int64_t ld, int64_t off, const int2* __restrict__ pairs, int64_t P,
                                int goff, int n, int n_cells, int w, const unsigned char* __restrict__ keep,
                                const double* __restrict__ inv, unsigned long long* __restrict__ diff) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int cell = blockIdx.y;
    if (i >= P || cell >= n_cells) return;
    const int2 pr = pairs[off + i];
    int a = pr.x - goff, b = pr.y - goff;
    if (a > b) { int t = a; a = b; b = t; }
    if (b - a > 2 * w) return;                                       // outside every window: no numerator, no denominator range
    if (keep && (!keep[a] || !keep[b])) return;
    const double raw = (double)scores[(int64_t)cell * ld + off + i];
    const double ia = inv[(int64_t)cell * n + a], ib = inv[(int64_t)cell * n + b];
    if (raw == 0.0 || ia == 0.0 || ib == 0.0) return;               // sqrt_norm: division by a zero coverage -> 0
    const double v = raw * ia * ib;
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
    // exact (integer, wrap-around) prefix sums of both difference arrays: every thread scans a contiguous segment, the 256 segment
    // totals are scanned through shared memory, then every thread finishes its segment
    __shared__ unsigned long long tot_n[256], tot_d[256];
    const int seg = (n + 255) / 256, i0 = min(n, (int)threadIdx.x * seg), i1 = min(n, i0 + seg);
    unsigned long long sn = 0, sd = 0;
    for (int i = i0; i < i1; ++i) { sn += num[i]; sd += den[i]; }
    tot_n[threadIdx.x] = sn; tot_d[threadIdx.x] = sd;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long an = 0, ad = 0;
        for (int t = 0; t < 256; ++t) { const unsigned long long xn = tot_n[t], xd = tot_d[t]; tot_n[t] = an; tot_d[t] = ad; an += xn; ad += xd; }
    }
    __syncthreads();
    sn = tot_n[threadIdx.x]; sd = tot_d[threadIdx.x];
    for (int i = i0; i < i1; ++i) {
        sn += num[i]; sd += den[i];
        double v = (double)(long long)sn / (double)(long long)sd;
        if (isnan(v) || sd == 0) v = 1.0;         // 0 / 0 -> nan -> 1 (the numerator is a subset of the denominator)
        if (discard && discard[i]) v = 1.0;
        sc[i] = v;
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
    HSG_CHECK(n_batch <= 65535, "at most 65535 cells per hsg_insulation call");
    cudaStream_t s = (cudaStream_t)stream;
    const int chrom = c->sel_chroms[sel_slot], n = c->bins[chrom], goff = c->bin_off[chrom];
    const int64_t off = c->sel_off[sel_slot], P = c->sel_P[sel_slot];
    // one grow-only workspace per context: [rowsum | diff] (zeroed per call) | inv | score64 | score | border | keep | discard
    const size_t nb = (size_t)n_batch, cells_n = nb * n;
    auto up = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const size_t o_rowsum = 0, o_diff = o_rowsum + up(cells_n * 8), o_inv = o_diff + up(nb * 2 * (n + 1) * 8),
                 o_sc64 = o_inv + up(cells_n * 8), o_score = o_sc64 + up(cells_n * 8), o_border = o_score + up(cells_n * 4),
                 o_keep = o_border + up(cells_n), o_disc = o_keep + up(n), total = o_disc + up(n);
    if (c->ins_ws_bytes < total) {
        if (c->ins_ws) { HSG_CUDA(cudaStreamSynchronize(s)); cudaFree(c->ins_ws); c->ins_ws = nullptr; c->ins_ws_bytes = 0; }
        HSG_CUDA(cudaMalloc(&c->ins_ws, total));
        c->ins_ws_bytes = total;
    }
    unsigned char* ws = (unsigned char*)c->ins_ws;
    unsigned long long* rowsum = (unsigned long long*)(ws + o_rowsum);
    unsigned long long* diff = (unsigned long long*)(ws + o_diff);
    double* inv = (double*)(ws + o_inv);
    double* sc64 = (double*)(ws + o_sc64);
    float* d_score = (float*)(ws + o_score);
    unsigned char* d_border = ws + o_border;
    unsigned char* d_keep = keep_host ? ws + o_keep : nullptr;
    unsigned char* d_discard = discard_host ? ws + o_disc : nullptr;
    HSG_CUDA(cudaMemsetAsync(ws, 0, o_inv, s));
    if (keep_host) HSG_CUDA(cudaMemcpyAsync(d_keep, keep_host, n, cudaMemcpyHostToDevice, s));
    if (discard_host) HSG_CUDA(cudaMemcpyAsync(d_discard, discard_host, n, cudaMemcpyHostToDevice, s));
    if (P > 0) {
        dim3 grid((unsigned)((P + 255) / 256), (unsigned)n_batch), grid_rs((unsigned)((P + INS_CHUNK - 1) / INS_CHUNK), (unsigned)n_batch);
        ins_rowsum_kernel<<<grid_rs, 256, 0, s>>>(scores_dev, c->P, off, c->pairs, P, goff, n, n_batch, d_keep, rowsum);
        c->launches++;
        HSG_CUDA(cudaGetLastError());
        ins_invsqrt_kernel<<<(unsigned)((cells_n + 255) / 256), 256, 0, s>>>(rowsum, (int64_t)cells_n, inv);
        c->launches++;
        HSG_CUDA(cudaGetLastError());
        ins_diff_kernel<<<grid, 256, 0, s>>>(scores_dev, c->P, off, c->pairs, P, goff, n, n_batch, window_bins, d_keep, inv, diff);
        c->launches++;
        HSG_CUDA(cudaGetLastError());
    }
    ins_finish_kernel<<<n_batch, 256, 0, s>>>(diff, n, tad_order, d_discard, d_score, d_border, sc64);
    c->launches++;
    HSG_CUDA(cudaGetLastError());
    HSG_CUDA(cudaMemcpyAsync(score_out_host, d_score, cells_n * sizeof(float), cudaMemcpyDeviceToHost, s));
    HSG_CUDA(cudaMemcpyAsync(border_out_host, d_border, cells_n, cudaMemcpyDeviceToHost, s));
    HSG_CUDA(cudaStreamSynchronize(s));
    return 0;
}
