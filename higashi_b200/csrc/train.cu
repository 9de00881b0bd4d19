// train.cu -- training-side scorer: three-head forward, loss, and the full backward pass (fp32).
//
// Reference: Hyper_SAGNN.forward (Modules.py:726-783) with the GraphSAGE encoder on-hook
// (GraphSageEncoder_with_weights.forward_on_hook, Modules.py:1485-1553; MeanAggregator_with_weights.forward :1337-1362),
// the loss heads of Higashi.forward_batch_hyperedge (Higashi_wrapper.py:861-913) and autograd's backward of the same
// graph (loss.backward(), Higashi_wrapper.py:1006).  Hyperedge batches are small (1.5k-7.7k triplets), so this path is a
// sequence of plain fp32 kernels over [3B, d] activations that stay in HBM/L2; parameters and gradients live in two flat
// device buffers that the host (torch Adam, NCCL allreduce) owns.
//
// Token rows: row 3r+t is token t of hyperedge r (t = 0 cell, 1 bin_i, 2 bin_j); bin-token rows 2r, 2r+1 follow the
// reference's to_neighs row order.
#include <cstring>

#include <algorithm>
#include "common.cuh"

#define TR_CHECK_LAUNCH(c) do { (c)->launches++; HSG_CUDA(cudaGetLastError()); } while (0)

// ------------------------------------------------------------------------------------------------ sgemm
// C[M,N] = sum_k a(m,k) b(k,n) (+ beta*C);  a(m,k) = TA ? A[k*lda+m] : A[m*lda+k];  b(k,n) = TB ? B[n*ldb+k] : B[k*ldb+n]
//
// Every GEMM of the training step is tall and skinny (M = 3B rows of activations, N and K in {32, 64, 128}, or the transposed
// weight-gradient shape with the long dimension as the reduction), so one launch is a few MFLOP and a handful of dependent
// global-load latencies.  64 x 64 output tile per block of 256 threads, 4 x 4 outputs per thread, K in steps of 32:
//   * both operand tiles live in shared memory k-major ([k][m], [k][n]) so that a thread reads its 4 rows / 4 columns with one
//     128-bit load each per k (the A read is a broadcast: 3 wavefronts per 16 FMAs);
//   * operands that are m-/n-major in global memory are transposed on the way in; the 4-float groups of a row are XOR-swizzled
//     with k/4, which makes the transposing scalar stores conflict-free without padding (padding would break the 128-bit reads);
//   * global loads are 128-bit and fully coalesced whenever the leading dimensions / base pointers allow it, guarded scalar
//     loads otherwise (ragged feature widths);
//   * the next k tile is fetched into registers before the current one is multiplied (one __syncthreads per tile);
//   * weight-gradient shapes slice the reduction over gridDim.z and combine with atomics (only launched with beta == 1).
#define SG_BM 64
#define SG_BN 64
#define SG_BK 32

// tile element (kk, x) with x the m / n index inside the tile -> shared-memory offset (floats)
__device__ __forceinline__ int sg_off(int kk, int x) { return kk * 64 + ((((x >> 2) ^ (kk >> 2)) & 15) << 2) + (x & 3); }

// Fetch one [SG_BK x 64] operand tile into registers (8 floats per thread).  XMAJOR: element (k, x) at P[k*ld + x] (x contiguous),
// else at P[x*ld + k] (k contiguous).  Thread mapping: XMAJOR  idx -> kk = idx/16, x4 = idx%16 ; else idx -> x = idx/8, k4 = idx%8.
template <bool XMAJOR>
__device__ __forceinline__ void sg_fetch(float (&r)[8], const float* __restrict__ P, int ld, int x0, int X, int k0, int K, bool vec) {
#pragma unroll
    for (int t = 0; t < 2; ++t) {
        const int idx = threadIdx.x + t * 256;
        int kk, xx;
        if (XMAJOR) { kk = idx >> 4; xx = (idx & 15) << 2; } else { xx = idx >> 3; kk = (idx & 7) << 2; }
        const int k = k0 + kk, x = x0 + xx;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (XMAJOR) {
            if (k < K) {
                const float* q = P + (size_t)k * ld + x;
                if (vec && x + 3 < X) v = __ldg(reinterpret_cast<const float4*>(q));
                else { if (x < X) v.x = __ldg(q); if (x + 1 < X) v.y = __ldg(q + 1); if (x + 2 < X) v.z = __ldg(q + 2); if (x + 3 < X) v.w = __ldg(q + 3); }
            }
        } else {
            if (x < X) {
                const float* q = P + (size_t)x * ld + k;
                if (vec && k + 3 < K) v = __ldg(reinterpret_cast<const float4*>(q));
                else { if (k < K) v.x = __ldg(q); if (k + 1 < K) v.y = __ldg(q + 1); if (k + 2 < K) v.z = __ldg(q + 2); if (k + 3 < K) v.w = __ldg(q + 3); }
            }
        }
        r[t * 4] = v.x; r[t * 4 + 1] = v.y; r[t * 4 + 2] = v.z; r[t * 4 + 3] = v.w;
    }
}
template <bool XMAJOR>
__device__ __forceinline__ void sg_stash(const float (&r)[8], float* __restrict__ S) {
#pragma unroll
    for (int t = 0; t < 2; ++t) {
        const int idx = threadIdx.x + t * 256;
        if (XMAJOR) {
            const int kk = idx >> 4, xx = (idx & 15) << 2;
            *reinterpret_cast<float4*>(S + sg_off(kk, xx)) = make_float4(r[t * 4], r[t * 4 + 1], r[t * 4 + 2], r[t * 4 + 3]);
        } else {
            const int xx = idx >> 3, kk = (idx & 7) << 2;          // the 4 values are 4 consecutive k of one row / column
#pragma unroll
            for (int j = 0; j < 4; ++j) S[sg_off(kk + j, xx)] = r[t * 4 + j];
        }
    }
}

template <bool TA, bool TB>
__global__ void __launch_bounds__(256) sgemm_kernel(int M, int N, int K, const float* __restrict__ A, int lda,
                                                    const float* __restrict__ B, int ldb, float* __restrict__ C, int ldc,
                                                    float beta, int vecA, int vecB) {
    __shared__ __align__(16) float sA[2][SG_BK * SG_BM], sB[2][SG_BK * SG_BN];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int m0 = blockIdx.y * SG_BM, n0 = blockIdx.x * SG_BN;
    // split-K: gridDim.z slices of the reduction, combined with atomics
    const int kchunk = ((K + (int)gridDim.z - 1) / (int)gridDim.z + SG_BK - 1) & ~(SG_BK - 1);
    const int kbeg = blockIdx.z * kchunk;
    K = min(K, kbeg + kchunk);
    float acc[4][4] = {};
    float ra[8], rb[8];
    if (kbeg < K) {
        sg_fetch<TA>(ra, A, lda, m0, M, kbeg, K, vecA != 0);
        sg_fetch<!TB>(rb, B, ldb, n0, N, kbeg, K, vecB != 0);
    }
    int buf = 0;
    for (int k0 = kbeg; k0 < K; k0 += SG_BK) {
        sg_stash<TA>(ra, sA[buf]);
        sg_stash<!TB>(rb, sB[buf]);
        __syncthreads();
        if (k0 + SG_BK < K) {                     // next tile in flight while this one is multiplied
            sg_fetch<TA>(ra, A, lda, m0, M, k0 + SG_BK, K, vecA != 0);
            sg_fetch<!TB>(rb, B, ldb, n0, N, k0 + SG_BK, K, vecB != 0);
        }
        const float* a_s = sA[buf];
        const float* b_s = sB[buf];
#pragma unroll
        for (int kk = 0; kk < SG_BK; ++kk) {
            const float4 av = *reinterpret_cast<const float4*>(a_s + kk * 64 + ((ty ^ (kk >> 2)) << 2));
            const float4 bv = *reinterpret_cast<const float4*>(b_s + kk * 64 + ((tx ^ (kk >> 2)) << 2));
            const float a[4] = {av.x, av.y, av.z, av.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                acc[i][0] = fmaf(a[i], bv.x, acc[i][0]); acc[i][1] = fmaf(a[i], bv.y, acc[i][1]);
                acc[i][2] = fmaf(a[i], bv.z, acc[i][2]); acc[i][3] = fmaf(a[i], bv.w, acc[i][3]);
            }
        }
        buf ^= 1;                                 // the other buffer was last read two tiles ago: one barrier per tile suffices
    }
    if (kbeg >= K && gridDim.z > 1) return;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n < N) {
                float* p = C + (size_t)m * ldc + n;
                if (gridDim.z > 1) atomicAdd(p, acc[i][j]);
                else *p = (beta != 0.f) ? fmaf(beta, *p, acc[i][j]) : acc[i][j];
            }
        }
    }
}

static inline bool sg_vec_ok(const float* P, int ld) { return (((uintptr_t)P) & 15) == 0 && (ld & 3) == 0; }

static int gemm(hsg_ctx* c, cudaStream_t s, bool ta, bool tb, int M, int N, int K, const float* A, int lda, const float* B,
                int ldb, float* C, int ldc, float beta) {
    if (M <= 0 || N <= 0) return 0;
    dim3 grid((N + SG_BN - 1) / SG_BN, (M + SG_BM - 1) / SG_BM);
    // weight-gradient shapes (tiny M x N, long K): slice the reduction so the grid covers the SMs
    if (beta == 1.0f && K >= 512 && grid.x * grid.y < 148) {
        int want = (2 * 148 + grid.x * grid.y - 1) / (grid.x * grid.y);
        grid.z = (unsigned)std::max(1, std::min(want, K / 128));
    }
    const int va = sg_vec_ok(A, lda) ? 1 : 0, vb = sg_vec_ok(B, ldb) ? 1 : 0;
    if (!ta && tb) sgemm_kernel<false, true><<<grid, 256, 0, s>>>(M, N, K, A, lda, B, ldb, C, ldc, beta, va, vb);
    else if (!ta && !tb) sgemm_kernel<false, false><<<grid, 256, 0, s>>>(M, N, K, A, lda, B, ldb, C, ldc, beta, va, vb);
    else if (ta && !tb) sgemm_kernel<true, false><<<grid, 256, 0, s>>>(M, N, K, A, lda, B, ldb, C, ldc, beta, va, vb);
    else sgemm_kernel<true, true><<<grid, 256, 0, s>>>(M, N, K, A, lda, B, ldb, C, ldc, beta, va, vb);
    TR_CHECK_LAUNCH(c);
    return 0;
}
// y = x W^T            : linear forward      (x [M,K], W [N,K])
#define LIN_FWD(M, N, K, x, ldx, W, ldw, y, ldy, beta) gemm(c, s, false, true, M, N, K, x, ldx, W, ldw, y, ldy, beta)
// dx (+)= dy W         : (dy [M,N], W [N,K])
#define LIN_DX(M, N, K, dy, ldy, W, ldw, dx, ldx, beta) gemm(c, s, false, false, M, K, N, dy, ldy, W, ldw, dx, ldx, beta)
// dW += dy^T x         : (dy [M,N], x [M,K]) -> [N,K]
#define LIN_DW(M, N, K, dy, ldy, x, ldx, dW, ldw) gemm(c, s, true, false, N, K, M, dy, ldy, x, ldx, dW, ldw, 1.0f)

// ------------------------------------------------------------------------------------------------ elementwise helpers
__device__ __forceinline__ float dswish(float x) {            // d/dx [x sigmoid(x)]
    float sg = 1.0f / (1.0f + expf(-x));
    return sg * (1.0f + x * (1.0f - sg));
}
// y = act(x + bias[col]);  optionally keep the pre-activation
__global__ void bias_act_kernel(float* __restrict__ x, const float* __restrict__ bias, int64_t n, int cols, int act,
                                float* __restrict__ pre) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float v = x[i] + (bias ? bias[i % cols] : 0.f);
    if (pre) pre[i] = v;
    x[i] = act ? swish_acc(v) : v;
}
// db[col] += sum_rows dy[row, col]
// block = 64 columns x 4 row lanes; gridDim.y row slices (COLSUM_SLICES) so that the grid covers the SMs
#define COLSUM_SLICES 128
__global__ void colsum_kernel(const float* __restrict__ dy, int rows, int cols, float* __restrict__ db) {
    __shared__ float part[4][64];
    const int cl = threadIdx.x & 63, rl = threadIdx.x >> 6;
    const int col = blockIdx.x * 64 + cl;
    float s = 0.f;
    if (col < cols)
        for (int r = blockIdx.y * 4 + rl; r < rows; r += gridDim.y * 4) s += dy[(size_t)r * cols + col];
    part[rl][cl] = s;
    __syncthreads();
    if (rl == 0 && col < cols) atomicAdd(db + col, (part[0][cl] + part[1][cl]) + (part[2][cl] + part[3][cl]));
}
// inverted dropout with a counter-based mask (same (seed, tag, index) -> same mask in forward and backward)
__device__ __forceinline__ float keep_scale(unsigned long long seed, unsigned tag, unsigned long long i, float p) {
    unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (i + 1) + ((unsigned long long)tag << 56);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; z ^= z >> 31;      // splitmix64
    float u = (float)(z >> 40) * (1.0f / 16777216.0f);
    return u >= p ? 1.0f / (1.0f - p) : 0.0f;
}
__global__ void dropout_kernel(float* __restrict__ x, int64_t n, float p, unsigned long long seed, unsigned tag) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] *= keep_scale(seed, tag, (unsigned long long)i, p);
}
// dpre = [dropout mask *] dy * swish'(pre) in place, and db[col] += sum_rows dpre[row, col] in the same pass (the bias gradient of
// the layer whose activation this is).  Same tiling as colsum_kernel; p_drop > 0 applies the forward pass's dropout mask first.
__global__ void dswish_colsum_kernel(float* __restrict__ dy, const float* __restrict__ pre, int rows, int cols, float* __restrict__ db,
                                     float p_drop, unsigned long long seed, unsigned tag) {
    __shared__ float part[4][64];
    const int cl = threadIdx.x & 63, rl = threadIdx.x >> 6;
    const int col = blockIdx.x * 64 + cl;
    float s = 0.f;
    if (col < cols)
        for (int r = blockIdx.y * 4 + rl; r < rows; r += gridDim.y * 4) {
            const size_t i = (size_t)r * cols + col;
            float g = dy[i];
            if (p_drop > 0.f) g *= keep_scale(seed, tag, (unsigned long long)i, p_drop);
            g *= dswish(pre[i]);
            dy[i] = g;
            s += g;
        }
    part[rl][cl] = s;
    __syncthreads();
    if (rl == 0 && col < cols) atomicAdd(db + col, (part[0][cl] + part[1][cl]) + (part[2][cl] + part[3][cl]));
}
// y = dropout(act(x + bias[col])) in one pass (the forward of pff_n1's hidden layer); keeps the pre-activation
__global__ void bias_act_dropout_kernel(float* __restrict__ x, const float* __restrict__ bias, int64_t n, int cols, int act,
                                        float* __restrict__ pre, float p_drop, unsigned long long seed, unsigned tag) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float v = x[i] + (bias ? bias[i % cols] : 0.f);
    if (pre) pre[i] = v;
    v = act ? swish_acc(v) : v;
    x[i] = v * keep_scale(seed, tag, (unsigned long long)i, p_drop);
}
__global__ void axpy_kernel(float* __restrict__ y, const float* __restrict__ x, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] += x[i];
}
static inline unsigned nblk(int64_t n, int t = 256) { return (unsigned)((n + t - 1) / t); }

// ------------------------------------------------------------------------------------------------ LayerNorm
// warp per row; out = (x - mu) * rs * g + b ; stats saved for the backward
template <int D>
__global__ void ln_fwd_kernel(const float* __restrict__ x, int rows, const float* __restrict__ g, const float* __restrict__ b,
                              float* __restrict__ out, float* __restrict__ mu, float* __restrict__ rs) {
    int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rows) return;
    constexpr int V = D / 32;
    float v[V], s = 0.f;
#pragma unroll
    for (int q = 0; q < V; ++q) { v[q] = x[(size_t)row * D + q * 32 + lane]; s += v[q]; }
    float m = warp_sum(s) / (float)D, s2 = 0.f;
#pragma unroll
    for (int q = 0; q < V; ++q) { v[q] -= m; s2 = fmaf(v[q], v[q], s2); }
    float r = 1.0f / sqrtf(warp_sum(s2) / (float)D + HSG_LN_EPS);
#pragma unroll
    for (int q = 0; q < V; ++q) out[(size_t)row * D + q * 32 + lane] = fmaf(v[q] * r, g[q * 32 + lane], b[q * 32 + lane]);
    if (lane == 0) { mu[row] = m; rs[row] = r; }
}
// dx (+)= LN backward ; dg, db accumulated with one atomic per column per block
template <int D>
__global__ void ln_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, int rows, const float* __restrict__ g,
                              const float* __restrict__ mu, const float* __restrict__ rs, float* __restrict__ dx, int accumulate,
                              float* __restrict__ dg, float* __restrict__ db) {
    __shared__ float sg[D], sb[D];
    for (int i = threadIdx.x; i < D; i += blockDim.x) { sg[i] = 0.f; sb[i] = 0.f; }
    __syncthreads();
    int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    constexpr int V = D / 32;
    if (row < rows) {
        float xh[V], dyg[V], a = 0.f, bsum = 0.f;
        const float m = mu[row], r = rs[row];
#pragma unroll
        for (int q = 0; q < V; ++q) {
            int col = q * 32 + lane;
            float d = dy[(size_t)row * D + col];
            xh[q] = (x[(size_t)row * D + col] - m) * r;
            dyg[q] = d * g[col];
            a += dyg[q]; bsum = fmaf(dyg[q], xh[q], bsum);
            atomicAdd(&sg[col], d * xh[q]);
            atomicAdd(&sb[col], d);
        }
        a = warp_sum(a) / (float)D; bsum = warp_sum(bsum) / (float)D;
#pragma unroll
        for (int q = 0; q < V; ++q) {
            float v = r * (dyg[q] - a - xh[q] * bsum);
            float* p = dx + (size_t)row * D + q * 32 + lane;
            *p = accumulate ? (*p + v) : v;
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < D; i += blockDim.x) { atomicAdd(dg + i, sg[i]); atomicAdd(db + i, sb[i]); }
}
static int ln_fwd(hsg_ctx* c, cudaStream_t s, const float* x, int rows, const float* g, const float* b, float* out, float* mu,
                  float* rs) {
    unsigned grid = (rows + 7) / 8;
    if (c->d == 64) ln_fwd_kernel<64><<<grid, 256, 0, s>>>(x, rows, g, b, out, mu, rs);
    else ln_fwd_kernel<128><<<grid, 256, 0, s>>>(x, rows, g, b, out, mu, rs);
    TR_CHECK_LAUNCH(c);
    return 0;
}
static int ln_bwd(hsg_ctx* c, cudaStream_t s, const float* dy, const float* x, int rows, const float* g, const float* mu,
                  const float* rs, float* dx, int accumulate, float* dg, float* db) {
    unsigned grid = (rows + 7) / 8;
    if (c->d == 64) ln_bwd_kernel<64><<<grid, 256, 0, s>>>(dy, x, rows, g, mu, rs, dx, accumulate, dg, db);
    else ln_bwd_kernel<128><<<grid, 256, 0, s>>>(dy, x, rows, g, mu, rs, dx, accumulate, dg, db);
    TR_CHECK_LAUNCH(c);
    return 0;
}

// ------------------------------------------------------------------------------------------------ node embedding
// CSR SpMM over the bin-token rows: neigh[row] = sum_e val[e] * E[col[e]]        (Modules.py:1356-1360)
template <int D>
__global__ void spmm_fwd_kernel(const int* __restrict__ indptr, const int* __restrict__ rowcnt, const int* __restrict__ cols,
                                const float* __restrict__ vals, const float* __restrict__ E, int rows, float* __restrict__ out) {
    int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rows) return;
    constexpr int V = D / 32;
    float acc[V] = {};
    const int e_end = rowcnt ? indptr[row] + rowcnt[row] : indptr[row + 1];
    for (int e = indptr[row]; e < e_end; ++e) {
        float w = vals[e];
        const float* er = E + (size_t)cols[e] * D;
#pragma unroll
        for (int q = 0; q < V; ++q) acc[q] = fmaf(w, er[q * 32 + lane], acc[q]);
    }
#pragma unroll
    for (int q = 0; q < V; ++q) out[(size_t)row * D + q * 32 + lane] = acc[q];
}
template <int D>
__global__ void spmm_bwd_kernel(const int* __restrict__ indptr, const int* __restrict__ rowcnt, const int* __restrict__ cols,
                                const float* __restrict__ vals, const float* __restrict__ dout, int rows, float* __restrict__ dE) {
    int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rows) return;
    constexpr int V = D / 32;
    float d[V];
#pragma unroll
    for (int q = 0; q < V; ++q) d[q] = dout[(size_t)row * D + q * 32 + lane];
    const int e_end = rowcnt ? indptr[row] + rowcnt[row] : indptr[row + 1];
    for (int e = indptr[row]; e < e_end; ++e) {
        float w = vals[e];
        float* er = dE + (size_t)cols[e] * D;
#pragma unroll
        for (int q = 0; q < V; ++q) atomicAdd(er + q * 32 + lane, w * d[q]);
    }
}
// dynamic / static assembly: row 3r = cell table row, rows 3r+1, 3r+2 = (comb | self) of the two bins
__global__ void assemble_kernel(const int64_t* __restrict__ x, int B, int d, int n_cells, int bin_base, const float* __restrict__ E_cell,
                                const float* __restrict__ Eall, const float* __restrict__ comb, float* __restrict__ dyn,
                                float* __restrict__ sta, float* __restrict__ selfb, const float* __restrict__ cellrows = nullptr) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)B * 3 * d) return;
    int f = (int)(i % d); int64_t row = i / d; int r = (int)(row / 3), t = (int)(row % 3);
    if (t == 0) {
        float v = cellrows ? cellrows[(size_t)r * d + f] : E_cell[(size_t)(x[r * 3] - 1) * d + f];
        dyn[i] = v; sta[i] = v;
    } else {
        int lb = (int)(x[r * 3 + t] - bin_base);             // local bin of the chromosome
        float v = Eall[(size_t)lb * d + f];
        sta[i] = v;
        if (selfb) selfb[((size_t)r * 2 + (t - 1)) * d + f] = v;
        dyn[i] = comb ? comb[((size_t)r * 2 + (t - 1)) * d + f] : v;
    }
}
// Xg[r, :] = X[cell(r), :]  (cell feature rows of the batch, stage 1)
__global__ void gather_cell_rows_kernel(const int64_t* __restrict__ x, int B, int F, const float* __restrict__ X, float* __restrict__ Xg) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)B * F) return;
    int r = (int)(i / F), f = (int)(i % F);
    Xg[i] = X[(size_t)(x[r * 3] - 1) * F + f];
}
// dcell[r, :] = ddyn[3r, :] + dsta[3r, :]   (stage 1: dynamic == static)
__global__ void cell_grad_rows_kernel(const float* __restrict__ ddyn, const float* __restrict__ dsta, int B, int d, float* __restrict__ dcell) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)B * d) return;
    int r = (int)(i / d), f = (int)(i % d);
    dcell[i] = ddyn[((size_t)r * 3) * d + f] + dsta[((size_t)r * 3) * d + f];
}
// auto-encoder reconstruction: a2 = swish(swish(enc)) ; dE = drecon-side gradient through both swishes
__global__ void swish2_fwd_kernel(const float* __restrict__ enc, int64_t n, float* __restrict__ a1, float* __restrict__ a2) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { float u = swish_acc(enc[i]); a1[i] = u; a2[i] = swish_acc(u); }
}
__global__ void swish2_bwd_kernel(float* __restrict__ da, const float* __restrict__ enc, const float* __restrict__ a1, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) da[i] *= dswish(a1[i]) * dswish(enc[i]);
}
// diff = recon + bias - target ; loss += mean(diff^2) ; drecon = 2 diff / N
__global__ void recon_mse_kernel(float* __restrict__ recon, const float* __restrict__ bias, const float* __restrict__ target, int64_t n,
                                 int F, float* __restrict__ loss) {
    __shared__ float red[32];
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    float l = 0.f;
    if (i < n) {
        float df = recon[i] + bias[i % F] - target[i];
        l = df * df / (float)n;
        recon[i] = 2.0f * df / (float)n;
    }
    l = warp_sum(l);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = l;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.f;
        v = warp_sum(v);
        if (threadIdx.x == 0) atomicAdd(loss, v);
    }
}
__global__ void sumsq_kernel(const float* __restrict__ v, int64_t n, float* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    float l = (i < n) ? v[i] * v[i] : 0.f;
    l = warp_sum(l);
    if ((threadIdx.x & 31) == 0) atomicAdd(out, l);
}
__global__ void axpy_scaled_kernel(float* __restrict__ y, const float* __restrict__ x, float a, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = fmaf(a, x[i], y[i]);
}
// scatter the gradients of the bin rows back: dself (per bin token) -> dEall ; ddyn/dsta rows -> dcomb / dself
__global__ void disassemble_kernel(const int64_t* __restrict__ x, int B, int d, int bin_base, const float* __restrict__ ddyn,
                                   const float* __restrict__ dsta, float* __restrict__ dcomb, float* __restrict__ dEall) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)B * 3 * d) return;
    int f = (int)(i % d); int64_t row = i / d; int r = (int)(row / 3), t = (int)(row % 3);
    if (t == 0) return;                                       // cell table is frozen after stage 1 (off_hook([0]))
    int lb = (int)(x[r * 3 + t] - bin_base);
    float g = dsta[i];
    if (dcomb) dcomb[((size_t)r * 2 + (t - 1)) * d + f] = ddyn[i];
    else g += ddyn[i];                                        // stage 1: dynamic == static
    atomicAdd(dEall + (size_t)lb * d + f, g);
}
__global__ void scatter_rows_kernel(const int64_t* __restrict__ x, int B, int d, int bin_base, const float* __restrict__ dself,
                                    float* __restrict__ dEall) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)B * 2 * d) return;
    int f = (int)(i % d); int64_t row = i / d; int r = (int)(row / 2), t = (int)(row % 2);
    int lb = (int)(x[r * 3 + 1 + t] - bin_base);
    atomicAdd(dEall + (size_t)lb * d + f, dself[i]);
}

// ------------------------------------------------------------------------------------------------ attention
__device__ __forceinline__ void att_row_fwd(float la, float lb, float& wa, float& wb, float* save) {
    float mx = fmaxf(0.f, fmaxf(la, lb));
    float e0 = expf(0.f - mx), ea = expf(la - mx), eb = expf(lb - mx);
    float Z = e0 + ea + eb;
    float pa = ea / Z, pb = eb / Z;
    float den = (pa + pb) + 1e-13f;
    wa = pa / den; wb = pb / den;
    if (save) { save[0] = pa; save[1] = pb; }
}
// thread per (hyperedge, head).  Q,K,V [3B,128] ; ctx [3B,128] ; P [B,8,6] softmax probabilities (pa,pb per row)
__global__ void attn_fwd_kernel(const float* __restrict__ Q, const float* __restrict__ K, const float* __restrict__ V, int B,
                                float* __restrict__ ctx, float* __restrict__ P) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * 8) return;
    int r = i >> 3, h = i & 7;
    const float* q[3]; const float* k[3]; const float* v[3];
    for (int t = 0; t < 3; ++t) {
        q[t] = Q + ((size_t)r * 3 + t) * HSG_HD + h * 16; k[t] = K + ((size_t)r * 3 + t) * HSG_HD + h * 16;
        v[t] = V + ((size_t)r * 3 + t) * HSG_HD + h * 16;
    }
    float s[3][3];
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) {
            float acc = 0.f;
            if (a != b) for (int e = 0; e < 16; ++e) acc = fmaf(q[a][e], k[b][e], acc);
            s[a][b] = acc * 0.25f;
        }
    float w[3][3] = {};
    float* ps = P + (size_t)i * 6;
    att_row_fwd(s[0][1], s[0][2], w[0][1], w[0][2], ps + 0);
    att_row_fwd(s[1][0], s[1][2], w[1][0], w[1][2], ps + 2);
    att_row_fwd(s[2][0], s[2][1], w[2][0], w[2][1], ps + 4);
    for (int t = 0; t < 3; ++t) {
        int a = (t == 0) ? 1 : 0, b = (t == 2) ? 1 : 2;
        float* o = ctx + ((size_t)r * 3 + t) * HSG_HD + h * 16;
        for (int e = 0; e < 16; ++e) o[e] = fmaf(w[t][b], v[b][e], w[t][a] * v[a][e]);
    }
}
// 4 lanes per (hyperedge, head), 4 of the 16 head dimensions per lane: the 32 lanes of a warp cover one full 128-float row per
// token (512 contiguous bytes per load), 36 accumulators per thread instead of 144, 4x the threads of the forward kernel; the
// two 16-element dot products of a row are finished with two shuffles.
__device__ __forceinline__ float dot4(const float4 a, const float4 b) { return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w))); }
__device__ __forceinline__ void axpy4(float4& y, float s, const float4 x) {
    y.x = fmaf(s, x.x, y.x); y.y = fmaf(s, x.y, y.y); y.z = fmaf(s, x.z, y.z); y.w = fmaf(s, x.w, y.w);
}
__global__ void attn_bwd_kernel(const float* __restrict__ Q, const float* __restrict__ K, const float* __restrict__ V,
                                const float* __restrict__ P, const float* __restrict__ dctx, int B, float* __restrict__ dQ,
                                float* __restrict__ dK, float* __restrict__ dV) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = i < B * 32;
    const int ii = valid ? i : B * 32 - 1;               // every lane takes part in the shuffles
    const int r = ii >> 5, h = (ii >> 2) & 7, e4 = ii & 3;
    size_t base[3];
    float4 q[3], k[3], v[3], dc[3];
#pragma unroll
    for (int t = 0; t < 3; ++t) {
        base[t] = ((size_t)r * 3 + t) * HSG_HD + h * 16 + e4 * 4;
        q[t] = *reinterpret_cast<const float4*>(Q + base[t]); k[t] = *reinterpret_cast<const float4*>(K + base[t]);
        v[t] = *reinterpret_cast<const float4*>(V + base[t]); dc[t] = *reinterpret_cast<const float4*>(dctx + base[t]);
    }
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 dq[3] = {z4, z4, z4}, dk[3] = {z4, z4, z4}, dv[3] = {z4, z4, z4};
    const float* ps = P + ((size_t)r * 8 + h) * 6;
#pragma unroll
    for (int t = 0; t < 3; ++t) {
        const int a = (t == 0) ? 1 : 0, b = (t == 2) ? 1 : 2;
        const float pa = ps[t * 2], pb = ps[t * 2 + 1];
        const float den = (pa + pb) + 1e-13f;
        const float wa = pa / den, wb = pb / den;
        float dwa = dot4(dc[t], v[a]), dwb = dot4(dc[t], v[b]);
        dwa += __shfl_xor_sync(0xffffffffu, dwa, 1); dwb += __shfl_xor_sync(0xffffffffu, dwb, 1);
        dwa += __shfl_xor_sync(0xffffffffu, dwa, 2); dwb += __shfl_xor_sync(0xffffffffu, dwb, 2);
        axpy4(dv[a], wa, dc[t]); axpy4(dv[b], wb, dc[t]);
        // w = p / (pa + pb + eps)
        const float common = (dwa * wa + dwb * wb) / den;
        const float dpa = dwa / den - common, dpb = dwb / den - common;
        // p = softmax([0, la, lb])  (the diagonal logit is the constant 0)
        const float dot = pa * dpa + pb * dpb;
        const float dla = pa * (dpa - dot) * 0.25f, dlb = pb * (dpb - dot) * 0.25f;
        axpy4(dq[t], dla, k[a]); axpy4(dq[t], dlb, k[b]);
        axpy4(dk[a], dla, q[t]); axpy4(dk[b], dlb, q[t]);
    }
    if (!valid) return;
#pragma unroll
    for (int t = 0; t < 3; ++t) {
        *reinterpret_cast<float4*>(dQ + base[t]) = dq[t]; *reinterpret_cast<float4*>(dK + base[t]) = dk[t];
        *reinterpret_cast<float4*>(dV + base[t]) = dv[t];
    }
}

// ------------------------------------------------------------------------------------------------ heads and loss
__global__ void sqdiff_fwd_kernel(const float* __restrict__ D, const float* __restrict__ S, int64_t n, float* __restrict__ u) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { float t = D[i] - S[i]; u[i] = t * t; }
}
// dD = 2 (D - S) du ; dS (+)= -dD
__global__ void sqdiff_bwd_kernel(const float* __restrict__ D, const float* __restrict__ S, const float* __restrict__ du, int64_t n,
                                  float* __restrict__ dD, float* __restrict__ dS) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { float g = 2.0f * (D[i] - S[i]) * du[i]; dD[i] = g; dS[i] -= g; }
}
// o[row] = c[row,:] . w1 + b1
__global__ void head_out_kernel(const float* __restrict__ cact, int rows, int hd, const float* __restrict__ w1, const float* __restrict__ b1,
                                float* __restrict__ o) {
    int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= rows) return;
    float acc = b1[0];
    for (int j = 0; j < hd; ++j) acc = fmaf(cact[(size_t)row * hd + j], w1[j], acc);
    o[row] = acc;
}
// distance side channel (Modules.py:729-740): FeedForward([2(C+1)+cf, 4, 1]) with swish
__global__ void xp_fwd_kernel(const int64_t* __restrict__ x, int B, int A, int cf, const float* __restrict__ attr,
                              const float* __restrict__ cellf, int only_model, const float* __restrict__ w0, const float* __restrict__ b0,
                              const float* __restrict__ w1, const float* __restrict__ b1, float* __restrict__ hpre, float* __restrict__ out) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= B) return;
    const int in_w = 2 * A + cf;
    const float* ai = attr + (size_t)x[r * 3 + 1] * A; const float* aj = attr + (size_t)x[r * 3 + 2] * A;
    const float* cfp = cellf + (size_t)x[r * 3] * cf;
    float o = b1[0];
    for (int u = 0; u < 4; ++u) {
        const float* wr = w0 + u * in_w;
        float acc = 0.f;
        for (int t = 0; t < A; ++t) acc = fmaf(wr[t], ai[t], acc);
        for (int t = 0; t < A; ++t) acc = fmaf(wr[A + t], aj[t], acc);
        if (!only_model) for (int t = 0; t < cf; ++t) acc = fmaf(wr[2 * A + t], cfp[t], acc);
        acc += b0[u];
        hpre[r * 4 + u] = acc;
        o = fmaf(w1[u], swish_acc(acc), o);
    }
    out[r] = o;
}
// One block per 256 hyperedges: dh / ids staged in shared memory, then ONE thread per output element (4 x in_w weights, 4 + 4 + 1
// biases / output weights) walks the block's rows -- no contended atomics, one global atomic per output per block.
__global__ void __launch_bounds__(256) xp_bwd_kernel(const int64_t* __restrict__ x, int B, int A, int cf, const float* __restrict__ attr,
                              const float* __restrict__ cellf, int only_model, const float* __restrict__ w1, const float* __restrict__ hpre,
                              const float* __restrict__ dout, float* __restrict__ dw0, float* __restrict__ db0, float* __restrict__ dw1,
                              float* __restrict__ db1) {
    __shared__ float s_dh[256][4], s_gs[256][4], s_g[256];
    __shared__ int s_x[256][3];
    const int in_w = 2 * A + cf;
    const int r0 = blockIdx.x * 256, nr = min(256, B - r0);
    {
        const int r = r0 + threadIdx.x;
        const bool ok = threadIdx.x < nr;
        const float g = ok ? dout[r] : 0.f;
        s_g[threadIdx.x] = g;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const float pre = ok ? hpre[r * 4 + u] : 0.f;
            s_gs[threadIdx.x][u] = g * swish_acc(pre);
            s_dh[threadIdx.x][u] = g * w1[u] * dswish(pre);
        }
#pragma unroll
        for (int q = 0; q < 3; ++q) s_x[threadIdx.x][q] = ok ? (int)x[r * 3 + q] : 0;
    }
    __syncthreads();
    for (int o = threadIdx.x; o < 4 * in_w + 9; o += 256) {
        float acc = 0.f;
        if (o < 4 * in_w) {
            const int u = o / in_w, t = o - u * in_w;
            if (t >= 2 * A && only_model) continue;
            const float* base = (t < 2 * A) ? attr + (t < A ? t : t - A) : cellf + (t - 2 * A);
            const int which = (t < A) ? 1 : (t < 2 * A ? 2 : 0), stride = (t < 2 * A) ? A : cf;
            for (int r = 0; r < nr; ++r) acc = fmaf(s_dh[r][u], __ldg(base + (size_t)s_x[r][which] * stride), acc);
            atomicAdd(dw0 + o, acc);
        } else if (o < 4 * in_w + 4) {
            const int u = o - 4 * in_w;
            for (int r = 0; r < nr; ++r) acc += s_dh[r][u];
            atomicAdd(db0 + u, acc);
        } else if (o < 4 * in_w + 8) {
            const int u = o - 4 * in_w - 4;
            for (int r = 0; r < nr; ++r) acc += s_gs[r][u];
            atomicAdd(dw1 + u, acc);
        } else {
            for (int r = 0; r < nr; ++r) acc += s_g[r];
            atomicAdd(db1, acc);
        }
    }
}
// head = mean over the 3 tokens + side channel
__global__ void pool_kernel(const float* __restrict__ o, const float* __restrict__ dp, int B, float* __restrict__ out) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < B) out[r] = ((o[r * 3] + o[r * 3 + 1]) + o[r * 3 + 2]) / 3.0f + dp[r];
}
// classification loss: mean_r w_r * BCEWithLogits(pred_r, y_r)  (Higashi_wrapper.py:877-878) ; dpred written
__global__ void bce_kernel(const float* __restrict__ pred, const float* __restrict__ y, const float* __restrict__ w, int B,
                           float* __restrict__ loss, float* __restrict__ dpred) {
    __shared__ float red[32];
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    float l = 0.f;
    if (r < B) {
        float p = pred[r], yy = y[r], ww = w[r];
        l = ww * (fmaxf(p, 0.f) - p * yy + log1pf(expf(-fabsf(p)))) / (float)B;
        dpred[r] = ww * (1.0f / (1.0f + expf(-p)) - yy) / (float)B;
    }
    l = warp_sum(l);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = l;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.f;
        v = warp_sum(v);
        if (threadIdx.x == 0) atomicAdd(loss, v);
    }
}
// regression loss: mse(pred, w)  (Higashi_wrapper.py:906-909)
__global__ void mse_kernel(const float* __restrict__ pred, const float* __restrict__ w, int B, float* __restrict__ loss,
                           float* __restrict__ dpred) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= B) return;
    float d = pred[r] - w[r];
    atomicAdd(loss, d * d / (float)B);
    dpred[r] = 2.0f * d / (float)B;
}
// do[row] = dmean[row/3] / 3 ; dc[row, j] = do[row] * w1[j] ; dw1[j] += do * c ; db1 += do
__global__ void head_bwd_kernel(const float* __restrict__ dmean, const float* __restrict__ cact, int rows, int hd,
                                const float* __restrict__ w1, float* __restrict__ dc, float* __restrict__ dw1, float* __restrict__ db1) {
    // 256 rows per block; thread (j = tid % hd, lane group = tid / hd) walks rows so that a warp reads contiguous cact columns:
    // per-thread partial sums, shared-memory combine, one global atomic per output column per block
    extern __shared__ float sh[];                 // [groups][hd + 1]
    const int groups = blockDim.x / hd, j = threadIdx.x % hd, grp = threadIdx.x / hd;
    const int r0 = blockIdx.x * 256, r1 = min(rows, r0 + 256);
    float sw = 0.f, sb = 0.f;
    if (grp < groups) {
        const float wj = w1[j];
        for (int row = r0 + grp; row < r1; row += groups) {
            const float g = dmean[row / 3] / 3.0f;
            dc[(size_t)row * hd + j] = g * wj;
            sw = fmaf(g, cact[(size_t)row * hd + j], sw);
            sb += g;
        }
        sh[grp * (hd + 1) + j] = sw;
        if (j == 0) sh[grp * (hd + 1) + hd] = sb;
    }
    __syncthreads();
    if (threadIdx.x <= hd) {
        float v = 0.f;
        for (int gq = 0; gq < groups; ++gq) v += sh[gq * (hd + 1) + threadIdx.x];
        atomicAdd(threadIdx.x < hd ? dw1 + threadIdx.x : db1, v);
    }
}


// ------------------------------------------------------------------------------------------------ zinb / rank losses
__device__ __forceinline__ float softplus_t(float x) { return x > 20.0f ? x : log1pf(expf(x)); }   // F.softplus(beta=1, threshold=20)
__device__ __forceinline__ float sigmoid_t(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ float digammaf_dev(float x) {
    float r = 0.f;
    while (x < 6.0f) { r -= 1.0f / x; x += 1.0f; }
    float f = 1.0f / (x * x);
    return r + logf(x) - 0.5f / x - f * (1.0f / 12.0f - f * (1.0f / 120.0f - f * (1.0f / 252.0f)));
}
// -log_zinb_positive(x = w, mu = clamp(softplus(mean) * cell2weight[cell]), theta = clamp(softplus(var)), pi = proba).mean()
// (Modules.py:30-80, Higashi_wrapper.py:896-905) and its derivatives with respect to the three heads
__global__ void zinb_kernel(const float* __restrict__ mean, const float* __restrict__ var, const float* __restrict__ proba,
                            const float* __restrict__ w, const int64_t* __restrict__ x, const float* __restrict__ c2w, int B,
                            float* __restrict__ loss, float* __restrict__ dmean, float* __restrict__ dvar, float* __restrict__ dproba,
                            float* __restrict__ pred_out) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= B) return;
    const float eps = 1e-8f;
    float xx = w[r], pi = proba[r];
    float cw = c2w[x[r * 3] - 1];
    float mu_raw = softplus_t(mean[r]) * cw, th_raw = softplus_t(var[r]);
    float mu = fminf(fmaxf(mu_raw, 1e-8f), 1e8f), th = fminf(fmaxf(th_raw, 1e-8f), 1e8f);
    if (pred_out) pred_out[r] = mu;
    float sp_pi = softplus_t(-pi);
    float lte = logf(th + eps), ltme = logf(th + mu + eps);
    float ptl = -pi + th * (lte - ltme);
    float dptl_dth = (lte - ltme) + th * (1.0f / (th + eps) - 1.0f / (th + mu + eps));
    float dptl_dmu = -th / (th + mu + eps);
    float res, d_pi, d_th, d_mu;
    if (xx < eps) {
        res = softplus_t(ptl) - sp_pi;
        float sg = sigmoid_t(ptl);
        d_pi = -sg + sigmoid_t(-pi); d_th = sg * dptl_dth; d_mu = sg * dptl_dmu;
    } else if (xx > eps) {
        res = -sp_pi + ptl + xx * (logf(mu + eps) - ltme) + lgammaf(xx + th) - lgammaf(th) - lgammaf(xx + 1.0f);
        d_pi = sigmoid_t(-pi) - 1.0f;
        d_th = dptl_dth - xx / (th + mu + eps) + digammaf_dev(xx + th) - digammaf_dev(th);
        d_mu = dptl_dmu + xx * (1.0f / (mu + eps) - 1.0f / (th + mu + eps));
    } else { res = 0.f; d_pi = d_th = d_mu = 0.f; }
    const float sc = -1.0f / (float)B;
    atomicAdd(loss, res * sc);
    float mu_live = (mu_raw > 1e-8f && mu_raw < 1e8f) ? 1.f : 0.f, th_live = (th_raw > 1e-8f && th_raw < 1e8f) ? 1.f : 0.f;
    dmean[r] = sc * d_mu * mu_live * sigmoid_t(mean[r]) * cw;
    dvar[r] = sc * d_th * th_live * sigmoid_t(var[r]);
    dproba[r] = sc * d_pi;
}
// rank loss (Higashi_wrapper.py:880-888): BCE-with-logits over all ordered pairs with |w_i - w_j| > thres
__global__ void rank_pairs_kernel(const float* __restrict__ mean, const float* __restrict__ w, int B, float thres,
                                  float* __restrict__ gsum, float* __restrict__ acc /* [0]=loss sum, [1]=pair count */) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    float pi = softplus_t(mean[i]), wi = w[i], g = 0.f, l = 0.f, m = 0.f;
    for (int j = 0; j < B; ++j) {
        float dw = wi - w[j];
        if (fabsf(dw) > thres) {
            float d = pi - softplus_t(mean[j]);
            float lab = dw > 0.f ? 1.f : 0.f;
            l += fmaxf(d, 0.f) - d * lab + log1pf(expf(-fabsf(d)));
            g += sigmoid_t(d) - lab;
            m += 1.f;
        }
    }
    gsum[i] = 2.0f * g;                    // (i,j) and (j,i) contribute the same amount to d/d softplus(mean_i)
    atomicAdd(acc, l); atomicAdd(acc + 1, m);
}
__global__ void rank_finish_kernel(const float* __restrict__ mean, const float* __restrict__ gsum, const float* __restrict__ acc, int B,
                                   float* __restrict__ loss, float* __restrict__ dmean) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    float M = fmaxf(acc[1], 1.f);
    if (i == 0) *loss = acc[0] / M;
    if (i < B) dmean[i] = gsum[i] / M * sigmoid_t(mean[i]);
}

// ------------------------------------------------------------------------------------------------ workspace
struct TrainWs {
    int maxB = 0, max_nnz = 0, max_nc = 0;
    float* buf = nullptr; size_t floats = 0;
    int* indptr = nullptr; int* cols = nullptr; float* vals = nullptr;
    int* rowcnt = nullptr; bool sampled = false;       // batch left by hsg_sample_batch: row r = [indptr[r], indptr[r] + rowcnt[r])
    int64_t* x = nullptr; float* y = nullptr; float* w = nullptr;
    // forward activations (all [rows, *] row-major)
    float *Epre, *Eall, *neigh, *selfb, *combpre, *comb, *dyn, *sta, *mu_d, *rs_d, *mu_s, *rs_s, *qin, *kin, *vin, *Q, *K, *V, *P,
        *ctx, *yv, *sp, *mu_y, *rs_y, *xn, *hpre, *h, *z, *mu_z, *rs_z, *Dn, *mu_p, *rs_p, *Sn, *u, *cpre, *cact, *o, *hp2, *dp2,
        *mean, *loss, *vpre, *vact, *ov, *ppre, *pact, *op, *hp1, *dp1, *hp3, *dp3, *varh, *probah, *pred, *gsum;
    // backward scratch
    float *dvar, *dproba, *dmean, *dc, *du, *dD, *dS, *dz, *dh, *dxn, *dy, *dsp, *dctx, *dQ, *dK, *dV, *dqin, *dkin, *dvin, *ddyn, *dsta, *dcomb,
        *dneigh, *dself, *dEall;
    // last batch
    int B = 0, chrom = 0, nnz = 0, stage = 2, only_model = 0, mode = 0;
    // stage 1: cell branch on-hook
    int dropout = 0; unsigned long long seed = 0;
    float *Xg = nullptr, *cellrows = nullptr, *dcell = nullptr; size_t cell_cap = 0; bool cell_on_hook = false;
};
#include <map>
static std::map<hsg_ctx*, TrainWs*> g_ws;

struct TrainBind {                                  // flat parameter / gradient buffers owned by the host
    float* params = nullptr; float* grads = nullptr; int64_t total = 0;
    std::vector<int64_t> off_slot;                  // [HSG_W_COUNT]  (-1 = not bound)
    std::vector<int64_t> off_pw, off_pb;            // per branch: projection weight [d,in], bias [d]
    std::vector<float*> feat; std::vector<int> in_dim;   // node features on the device, per branch
};
static std::map<hsg_ctx*, TrainBind*> g_bind;
static TrainBind* bind_of(hsg_ctx* c) { auto it = g_bind.find(c); if (it == g_bind.end()) { g_bind[c] = new TrainBind(); return g_bind[c]; } return it->second; }

extern "C" int hsg_train_set_features(hsg_ctx* c, int branch, const float* x_host, int64_t rows, int in_dim) {
    HSG_CHECK(c && c->d > 0 && branch >= 0 && branch <= c->n_chrom && x_host, "bad features");
    HSG_CUDA(cudaSetDevice(c->device));
    TrainBind* b = bind_of(c);
    if ((int)b->feat.size() <= c->n_chrom) { b->feat.assign(c->n_chrom + 1, nullptr); b->in_dim.assign(c->n_chrom + 1, 0); }
    int64_t want = branch == 0 ? c->n_cells : c->bins[branch - 1];
    HSG_CHECK(rows == want, "feature rows do not match the node type");
    if (b->feat[branch]) cudaFree(b->feat[branch]);
    HSG_CUDA(cudaMalloc(&b->feat[branch], (size_t)rows * in_dim * sizeof(float)));
    HSG_CUDA(cudaMemcpy(b->feat[branch], x_host, (size_t)rows * in_dim * sizeof(float), cudaMemcpyHostToDevice));
    b->in_dim[branch] = in_dim;
    return 0;
}

extern "C" int hsg_train_bind(hsg_ctx* c, float* params_dev, float* grads_dev, int64_t total, const int64_t* off_slot,
                              const int64_t* off_proj_w, const int64_t* off_proj_b) {
    HSG_CHECK(c && c->d > 0 && params_dev && grads_dev && off_slot && off_proj_w && off_proj_b, "bad bind arguments");
    TrainBind* b = bind_of(c);
    b->params = params_dev; b->grads = grads_dev; b->total = total;
    b->off_slot.assign(off_slot, off_slot + HSG_W_COUNT);
    b->off_pw.assign(off_proj_w, off_proj_w + c->n_chrom + 1);
    b->off_pb.assign(off_proj_b, off_proj_b + c->n_chrom + 1);
    return 0;
}

static int ws_cell_alloc(hsg_ctx* c, TrainWs* w, int B, int F) {
    size_t need = (size_t)B * F + 2 * (size_t)B * c->d;
    if (w->cell_cap >= need) return 0;
    if (w->Xg) cudaFree(w->Xg);
    HSG_CUDA(cudaMalloc(&w->Xg, need * sizeof(float)));
    w->cellrows = w->Xg + (size_t)B * F; w->dcell = w->cellrows + (size_t)B * c->d; w->cell_cap = need;
    return 0;
}

static int ws_alloc(hsg_ctx* c, int maxB, int max_nnz) {
    TrainWs*& w = g_ws[c];
    if (w && w->maxB >= maxB && w->max_nnz >= max_nnz) return 0;
    if (w) { cudaFree(w->buf); cudaFree(w->indptr); cudaFree(w->rowcnt); cudaFree(w->cols); cudaFree(w->vals); cudaFree(w->x); if (w->Xg) cudaFree(w->Xg); delete w; }
    w = new TrainWs();
    w->maxB = maxB; w->max_nnz = max_nnz;
    int max_nc = 0; for (int v : c->bins) max_nc = std::max(max_nc, v);
    w->max_nc = max_nc;
    const size_t R = (size_t)3 * maxB, NB = (size_t)2 * maxB, d = c->d, hd = c->d / 2;
    size_t total = 0;
    auto take = [&](size_t n) { size_t o = total; total += (n + 63) & ~(size_t)63; return o; };
    std::vector<std::pair<float**, size_t>> plan = {
        {&w->Epre, (size_t)max_nc * d}, {&w->Eall, (size_t)max_nc * d}, {&w->neigh, NB * d}, {&w->selfb, NB * d},
        {&w->combpre, NB * d}, {&w->comb, NB * d}, {&w->dyn, R * d}, {&w->sta, R * d}, {&w->mu_d, R}, {&w->rs_d, R},
        {&w->mu_s, R}, {&w->rs_s, R}, {&w->qin, R * d}, {&w->kin, R * d}, {&w->vin, R * d}, {&w->Q, R * HSG_HD},
        {&w->K, R * HSG_HD}, {&w->V, R * HSG_HD}, {&w->P, (size_t)maxB * 48}, {&w->ctx, R * HSG_HD}, {&w->yv, R * d},
        {&w->sp, R * d}, {&w->mu_y, R}, {&w->rs_y, R}, {&w->xn, R * d}, {&w->hpre, R * d}, {&w->h, R * d}, {&w->z, R * d},
        {&w->mu_z, R}, {&w->rs_z, R}, {&w->Dn, R * d}, {&w->mu_p, R}, {&w->rs_p, R}, {&w->Sn, R * d}, {&w->u, R * d},
        {&w->cpre, R * hd}, {&w->cact, R * hd}, {&w->o, R}, {&w->hp2, (size_t)maxB * 4}, {&w->dp2, (size_t)maxB},
        {&w->mean, (size_t)maxB}, {&w->loss, 64}, {&w->vpre, R * hd}, {&w->vact, R * hd}, {&w->ov, R}, {&w->ppre, R * hd},
        {&w->pact, R * hd}, {&w->op, R}, {&w->hp1, (size_t)maxB * 4}, {&w->dp1, (size_t)maxB}, {&w->hp3, (size_t)maxB * 4},
        {&w->dp3, (size_t)maxB}, {&w->varh, (size_t)maxB}, {&w->probah, (size_t)maxB}, {&w->pred, (size_t)maxB}, {&w->gsum, (size_t)maxB},
        {&w->dvar, (size_t)maxB}, {&w->dproba, (size_t)maxB},
        {&w->dmean, (size_t)maxB}, {&w->dc, R * hd}, {&w->du, R * d}, {&w->dD, R * d}, {&w->dS, R * d}, {&w->dz, R * d},
        {&w->dh, R * d}, {&w->dxn, R * d}, {&w->dy, R * d}, {&w->dsp, R * d}, {&w->dctx, R * HSG_HD}, {&w->dQ, R * HSG_HD},
        {&w->dK, R * HSG_HD}, {&w->dV, R * HSG_HD}, {&w->dqin, R * d}, {&w->dkin, R * d}, {&w->dvin, R * d}, {&w->ddyn, R * d},
        {&w->dsta, R * d}, {&w->dcomb, NB * d}, {&w->dneigh, NB * d}, {&w->dself, NB * d}, {&w->dEall, (size_t)max_nc * d}};
    std::vector<size_t> offs;
    for (auto& p : plan) offs.push_back(take(p.second));
    HSG_CUDA(cudaMalloc(&w->buf, total * sizeof(float)));
    for (size_t i = 0; i < plan.size(); ++i) *plan[i].first = w->buf + offs[i];
    w->floats = total;
    HSG_CUDA(cudaMalloc(&w->indptr, (NB + 1) * sizeof(int)));
    HSG_CUDA(cudaMalloc(&w->rowcnt, (NB + 1) * sizeof(int)));
    HSG_CUDA(cudaMalloc(&w->cols, (size_t)std::max(1, max_nnz) * sizeof(int)));
    HSG_CUDA(cudaMalloc(&w->vals, (size_t)std::max(1, max_nnz) * sizeof(float)));
    HSG_CUDA(cudaMalloc(&w->x, ((size_t)maxB * 3) * sizeof(int64_t) + (size_t)maxB * 2 * sizeof(float)));
    w->y = reinterpret_cast<float*>(w->x + (size_t)maxB * 3);
    w->w = w->y + maxB;
    return 0;
}

#include "sampler.cuh"

struct TrainOpts { float* c2w = nullptr; float rank_thres = 0.f; int dropout = 0; unsigned long long seed = 0; };
static std::map<hsg_ctx*, TrainOpts> g_opts;

extern "C" int hsg_train_options(hsg_ctx* c, const float* cell2weight_host, float rank_thres) {
    HSG_CHECK(c && c->d > 0, "context not initialised");
    HSG_CUDA(cudaSetDevice(c->device));
    TrainOpts& o = g_opts[c];
    o.rank_thres = rank_thres;
    if (cell2weight_host) {
        if (!o.c2w) HSG_CUDA(cudaMalloc(&o.c2w, (size_t)c->n_cells * sizeof(float)));
        HSG_CUDA(cudaMemcpy(o.c2w, cell2weight_host, (size_t)c->n_cells * sizeof(float), cudaMemcpyHostToDevice));
    }
    return 0;
}

// training-mode dropout (Hyper_SAGNN: 0.3 after fc1 / fc2, Modules.py:685,1088-1089; 0.4 inside pff_n1, Modules.py:686,876).
// enabled = 0 gives the eval-mode semantics the parity tests use; `seed` should change every step.
extern "C" int hsg_train_set_dropout(hsg_ctx* c, int enabled, uint64_t seed) {
    HSG_CHECK(c && c->d > 0, "context not initialised");
    g_opts[c].dropout = enabled; g_opts[c].seed = seed;
    return 0;
}

// heads of the last forward batch (mean always; var / proba only after a zinb-mode forward), host outputs, may be NULL
extern "C" int hsg_score_heads(hsg_ctx* c, float* mean_host, float* var_host, float* proba_host, float* pred_host, void* stream) {
    HSG_CHECK(c && g_ws.count(c) && g_ws[c]->B > 0, "no forward batch");
    TrainWs* w = g_ws[c];
    cudaStream_t s = (cudaStream_t)stream;
    if (mean_host) HSG_CUDA(cudaMemcpyAsync(mean_host, w->mean, w->B * sizeof(float), cudaMemcpyDefault, s));
    if (var_host) HSG_CUDA(cudaMemcpyAsync(var_host, w->varh, w->B * sizeof(float), cudaMemcpyDefault, s));
    if (proba_host) HSG_CUDA(cudaMemcpyAsync(proba_host, w->probah, w->B * sizeof(float), cudaMemcpyDefault, s));
    if (pred_host) HSG_CUDA(cudaMemcpyAsync(pred_host, w->pred, w->B * sizeof(float), cudaMemcpyDefault, s));
    HSG_CUDA(cudaStreamSynchronize(s));
    return 0;
}

#define PARAM(slot) (b->params + b->off_slot[slot])
#define GRAD(slot) (b->grads + b->off_slot[slot])

// forward: returns the three heads for the mean head only when the others are not requested (classification / regression use
// `mean`; the var/proba heads of zinb mode are evaluated by the same building blocks -- see hsg_score_fwd docs)
extern "C" int hsg_score_fwd(hsg_ctx* c, int stage, int chrom, const int64_t* x_host, int B, const int32_t* indptr_host,
                             const int32_t* cols_local_host, const float* vals_host, int nnz, const float* y_host,
                             const float* w_host, int loss_mode, int only_model, float* mean_out_host, float* loss_out_host,
                             void* stream) {
    HSG_CHECK(c && c->d > 0, "context not initialised");
    HSG_CUDA(cudaSetDevice(c->device));
    HSG_CHECK(stage == 1 || stage == 2, "stage must be 1 (MultipleEmbedding) or 2 (GraphSAGE on-hook)");
    HSG_CHECK(chrom >= 0 && chrom < c->n_chrom && B > 0, "bad batch");
    // x_host == NULL: run on the batch hsg_sample_batch left in the workspace (B / nnz must be the values it returned)
    const bool dev_batch = (x_host == nullptr);
    if (dev_batch) {
        auto it = g_ws.find(c);
        HSG_CHECK(it != g_ws.end() && it->second && it->second->sampled && it->second->chrom == chrom, "x_host is NULL but there is no sampled batch for this chromosome");
    }
    HSG_CHECK(loss_mode >= 0 && loss_mode <= 3, "loss_mode must be 0..3");
    HSG_CHECK(stage == 1 || dev_batch || (indptr_host && (nnz == 0 || (cols_local_host && vals_host))), "stage 2 needs the neighbour CSR");
    TrainBind* b = bind_of(c);
    HSG_CHECK(b->params != nullptr, "call hsg_train_bind first");
    HSG_CHECK((int)b->feat.size() > chrom + 1 && b->feat[chrom + 1] != nullptr, "call hsg_train_set_features for this chromosome");
    for (int slot : {HSG_W_QS, HSG_W_KS, HSG_W_VS, HSG_W_FC1, HSG_W_FC2, HSG_W_PFF_W0, HSG_W_CLS_W0, HSG_W_XP2_W0, HSG_W_ATTR})
        HSG_CHECK(slot == HSG_W_ATTR ? c->w[slot] != nullptr : b->off_slot[slot] >= 0, "a required parameter is not bound");
    HSG_CHECK(stage == 1 || b->off_slot[HSG_W_SAGE_W] >= 0, "GraphSAGE parameters are not bound");
    HSG_CHECK(c->w[HSG_W_CELL_FEATS] != nullptr, "cell_feats not set");
    if (ws_alloc(c, B, nnz)) return 1;
    TrainWs* w = g_ws[c];
    cudaStream_t s = (cudaStream_t)stream;
    const int d = c->d, hd = d / 2, R = 3 * B, NB = 2 * B, nc = c->bins[chrom], in_c = b->in_dim[chrom + 1];
    const int bin_base = c->n_cells + 1 + c->bin_off[chrom];
    w->B = B; w->chrom = chrom; w->nnz = nnz; w->stage = stage; w->only_model = only_model; w->mode = loss_mode;
    if (!dev_batch) {
        w->sampled = false;
        HSG_CUDA(cudaMemcpyAsync(w->x, x_host, (size_t)B * 3 * sizeof(int64_t), cudaMemcpyHostToDevice, s));
        if (y_host) HSG_CUDA(cudaMemcpyAsync(w->y, y_host, B * sizeof(float), cudaMemcpyHostToDevice, s));
        if (w_host) HSG_CUDA(cudaMemcpyAsync(w->w, w_host, B * sizeof(float), cudaMemcpyHostToDevice, s));
    }
    const bool have_y = dev_batch || y_host != nullptr, have_w = dev_batch || w_host != nullptr;
    if (stage == 2 && !dev_batch) {
        HSG_CUDA(cudaMemcpyAsync(w->indptr, indptr_host, (NB + 1) * sizeof(int), cudaMemcpyHostToDevice, s));
        if (nnz) {
            HSG_CUDA(cudaMemcpyAsync(w->cols, cols_local_host, nnz * sizeof(int), cudaMemcpyHostToDevice, s));
            HSG_CUDA(cudaMemcpyAsync(w->vals, vals_host, nnz * sizeof(float), cudaMemcpyHostToDevice, s));
        }
    }
    // 1. bin embeddings of the whole chromosome: E = swish(X_c W_c^T + b_c)            (Modules.py:235-253, 558)
    const float* Wc = b->params + b->off_pw[chrom + 1]; const float* bc = b->params + b->off_pb[chrom + 1];
    if (LIN_FWD(nc, d, in_c, b->feat[chrom + 1], in_c, Wc, in_c, w->Eall, d, 0.f)) return 1;
    bias_act_kernel<<<nblk((int64_t)nc * d), 256, 0, s>>>(w->Eall, bc, (int64_t)nc * d, d, 1, w->Epre); TR_CHECK_LAUNCH(c);
    // 2. GraphSAGE: neigh = M . E ; comb = swish(nn([neigh | self]))                      (Modules.py:1511-1548)
    if (stage == 2) {
        assemble_kernel<<<nblk((int64_t)R * d), 256, 0, s>>>(w->x, B, d, c->n_cells, bin_base, c->E_cell, w->Eall, nullptr, w->dyn, w->sta, w->selfb);
        TR_CHECK_LAUNCH(c);
        if (d == 64) spmm_fwd_kernel<64><<<(NB + 7) / 8, 256, 0, s>>>(w->indptr, w->sampled ? w->rowcnt : nullptr, w->cols, w->vals, w->Eall, NB, w->neigh);
        else spmm_fwd_kernel<128><<<(NB + 7) / 8, 256, 0, s>>>(w->indptr, w->sampled ? w->rowcnt : nullptr, w->cols, w->vals, w->Eall, NB, w->neigh);
        TR_CHECK_LAUNCH(c);
        const float* G = PARAM(HSG_W_SAGE_W);
        if (LIN_FWD(NB, d, d, w->neigh, d, G, 2 * d, w->comb, d, 0.f)) return 1;
        if (LIN_FWD(NB, d, d, w->selfb, d, G + d, 2 * d, w->comb, d, 1.f)) return 1;
        bias_act_kernel<<<nblk((int64_t)NB * d), 256, 0, s>>>(w->comb, PARAM(HSG_W_SAGE_B), (int64_t)NB * d, d, 1, w->combpre); TR_CHECK_LAUNCH(c);
    }
    const float* cellrows = nullptr;
    if (stage == 1 && b->off_pw[0] >= 0 && !b->feat.empty() && b->feat[0] != nullptr) {
        // cell branch on-hook (stage 1): rows of X_cell for the batch -> linear encoder (Modules.py:550-552, add_activation=False)
        const int F = b->in_dim[0];
        if (ws_cell_alloc(c, w, B, F)) return 1;
        gather_cell_rows_kernel<<<nblk((int64_t)B * F), 256, 0, s>>>(w->x, B, F, b->feat[0], w->Xg); TR_CHECK_LAUNCH(c);
        if (LIN_FWD(B, d, F, w->Xg, F, b->params + b->off_pw[0], F, w->cellrows, d, 0.f)) return 1;
        bias_act_kernel<<<nblk((int64_t)B * d), 256, 0, s>>>(w->cellrows, b->params + b->off_pb[0], (int64_t)B * d, d, 0, nullptr); TR_CHECK_LAUNCH(c);
        cellrows = w->cellrows;
    }
    w->cell_on_hook = cellrows != nullptr;
    assemble_kernel<<<nblk((int64_t)R * d), 256, 0, s>>>(w->x, B, d, c->n_cells, bin_base, c->E_cell, w->Eall,
                                                         stage == 2 ? w->comb : nullptr, w->dyn, w->sta, nullptr, cellrows);
    TR_CHECK_LAUNCH(c);
    // 3. attention block                                                                  (Modules.py:1029-1095)
    if (ln_fwd(c, s, w->dyn, R, PARAM(HSG_W_ALN1_G), PARAM(HSG_W_ALN1_B), w->qin, w->mu_d, w->rs_d)) return 1;
    if (ln_fwd(c, s, w->dyn, R, PARAM(HSG_W_ALN2_G), PARAM(HSG_W_ALN2_B), w->kin, w->mu_d, w->rs_d)) return 1;
    if (ln_fwd(c, s, w->sta, R, PARAM(HSG_W_ALN3_G), PARAM(HSG_W_ALN3_B), w->vin, w->mu_s, w->rs_s)) return 1;
    if (LIN_FWD(R, HSG_HD, d, w->qin, d, PARAM(HSG_W_QS), d, w->Q, HSG_HD, 0.f)) return 1;
    if (LIN_FWD(R, HSG_HD, d, w->kin, d, PARAM(HSG_W_KS), d, w->K, HSG_HD, 0.f)) return 1;
    if (LIN_FWD(R, HSG_HD, d, w->vin, d, PARAM(HSG_W_VS), d, w->V, HSG_HD, 0.f)) return 1;
    attn_fwd_kernel<<<nblk(B * 8, 128), 128, 0, s>>>(w->Q, w->K, w->V, B, w->ctx, w->P); TR_CHECK_LAUNCH(c);
    if (LIN_FWD(R, d, HSG_HD, w->ctx, HSG_HD, PARAM(HSG_W_FC1), HSG_HD, w->yv, d, 0.f)) return 1;
    if (LIN_FWD(R, d, HSG_HD, w->V, HSG_HD, PARAM(HSG_W_FC2), HSG_HD, w->sp, d, 0.f)) return 1;
    const TrainOpts& opt = g_opts[c];
    w->dropout = opt.dropout; w->seed = opt.seed;
    if (opt.dropout) {
        dropout_kernel<<<nblk((int64_t)R * d), 256, 0, s>>>(w->yv, (int64_t)R * d, 0.3f, opt.seed, 1); TR_CHECK_LAUNCH(c);
        dropout_kernel<<<nblk((int64_t)R * d), 256, 0, s>>>(w->sp, (int64_t)R * d, 0.3f, opt.seed, 2); TR_CHECK_LAUNCH(c);
    }
    // 4. pff_n1: LN -> W0 -> swish -> W1 -> + residual                                    (Modules.py:865-889)
    if (ln_fwd(c, s, w->yv, R, PARAM(HSG_W_PFF_LN_G), PARAM(HSG_W_PFF_LN_B), w->xn, w->mu_y, w->rs_y)) return 1;
    if (LIN_FWD(R, d, d, w->xn, d, PARAM(HSG_W_PFF_W0), d, w->h, d, 0.f)) return 1;
    if (opt.dropout) bias_act_dropout_kernel<<<nblk((int64_t)R * d), 256, 0, s>>>(w->h, PARAM(HSG_W_PFF_B0), (int64_t)R * d, d, 1, w->hpre, 0.4f, opt.seed, 3);
    else bias_act_kernel<<<nblk((int64_t)R * d), 256, 0, s>>>(w->h, PARAM(HSG_W_PFF_B0), (int64_t)R * d, d, 1, w->hpre);
    TR_CHECK_LAUNCH(c);
    HSG_CUDA(cudaMemcpyAsync(w->z, w->yv, (size_t)R * d * sizeof(float), cudaMemcpyDeviceToDevice, s));
    if (LIN_FWD(R, d, d, w->h, d, PARAM(HSG_W_PFF_W1), d, w->z, d, 1.f)) return 1;
    bias_act_kernel<<<nblk((int64_t)R * d), 256, 0, s>>>(w->z, PARAM(HSG_W_PFF_B1), (int64_t)R * d, d, 0, nullptr); TR_CHECK_LAUNCH(c);
    // 5. heads                                                                            (Modules.py:752-779)
    if (ln_fwd(c, s, w->z, R, PARAM(HSG_W_LN1_G), PARAM(HSG_W_LN1_B), w->Dn, w->mu_z, w->rs_z)) return 1;
    if (ln_fwd(c, s, w->sp, R, PARAM(HSG_W_LN2_G), PARAM(HSG_W_LN2_B), w->Sn, w->mu_p, w->rs_p)) return 1;
    sqdiff_fwd_kernel<<<nblk((int64_t)R * d), 256, 0, s>>>(w->Dn, w->Sn, (int64_t)R * d, w->u); TR_CHECK_LAUNCH(c);
    if (LIN_FWD(R, hd, d, w->u, d, PARAM(HSG_W_CLS_W0), d, w->cact, hd, 0.f)) return 1;
    bias_act_kernel<<<nblk((int64_t)R * hd), 256, 0, s>>>(w->cact, PARAM(HSG_W_CLS_B0), (int64_t)R * hd, hd, 1, w->cpre); TR_CHECK_LAUNCH(c);
    head_out_kernel<<<nblk(R), 256, 0, s>>>(w->cact, R, hd, PARAM(HSG_W_CLS_W1), PARAM(HSG_W_CLS_B1), w->o); TR_CHECK_LAUNCH(c);
    xp_fwd_kernel<<<nblk(B), 256, 0, s>>>(w->x, B, c->n_chrom + 1, c->cf, c->w[HSG_W_ATTR], c->w[HSG_W_CELL_FEATS], only_model,
                                          PARAM(HSG_W_XP2_W0), PARAM(HSG_W_XP2_B0), PARAM(HSG_W_XP2_W1), PARAM(HSG_W_XP2_B1), w->hp2, w->dp2);
    TR_CHECK_LAUNCH(c);
    pool_kernel<<<nblk(B), 256, 0, s>>>(w->o, w->dp2, B, w->mean); TR_CHECK_LAUNCH(c);
    // var / proba heads on the static branch (zinb mode; Modules.py:760-779)
    if (loss_mode == 2) {
        HSG_CHECK(b->off_slot[HSG_W_VAR_W0] >= 0 && b->off_slot[HSG_W_PRB_W0] >= 0 && b->off_slot[HSG_W_XP_W0] >= 0 &&
                  b->off_slot[HSG_W_XP3_W0] >= 0, "zinb mode needs the var / proba heads bound");
        HSG_CHECK(g_opts[c].c2w != nullptr || !(have_w), "zinb loss needs hsg_train_options(cell2weight)");
        if (LIN_FWD(R, hd, d, w->Sn, d, PARAM(HSG_W_VAR_W0), d, w->vact, hd, 0.f)) return 1;
        bias_act_kernel<<<nblk((int64_t)R * hd), 256, 0, s>>>(w->vact, PARAM(HSG_W_VAR_B0), (int64_t)R * hd, hd, 1, w->vpre); TR_CHECK_LAUNCH(c);
        head_out_kernel<<<nblk(R), 256, 0, s>>>(w->vact, R, hd, PARAM(HSG_W_VAR_W1), PARAM(HSG_W_VAR_B1), w->ov); TR_CHECK_LAUNCH(c);
        if (LIN_FWD(R, hd, d, w->Sn, d, PARAM(HSG_W_PRB_W0), d, w->pact, hd, 0.f)) return 1;
        bias_act_kernel<<<nblk((int64_t)R * hd), 256, 0, s>>>(w->pact, PARAM(HSG_W_PRB_B0), (int64_t)R * hd, hd, 1, w->ppre); TR_CHECK_LAUNCH(c);
        head_out_kernel<<<nblk(R), 256, 0, s>>>(w->pact, R, hd, PARAM(HSG_W_PRB_W1), PARAM(HSG_W_PRB_B1), w->op); TR_CHECK_LAUNCH(c);
        xp_fwd_kernel<<<nblk(B), 256, 0, s>>>(w->x, B, c->n_chrom + 1, c->cf, c->w[HSG_W_ATTR], c->w[HSG_W_CELL_FEATS], only_model,
                                              PARAM(HSG_W_XP3_W0), PARAM(HSG_W_XP3_B0), PARAM(HSG_W_XP3_W1), PARAM(HSG_W_XP3_B1), w->hp3, w->dp3);
        TR_CHECK_LAUNCH(c);
        xp_fwd_kernel<<<nblk(B), 256, 0, s>>>(w->x, B, c->n_chrom + 1, c->cf, c->w[HSG_W_ATTR], c->w[HSG_W_CELL_FEATS], only_model,
                                              PARAM(HSG_W_XP_W0), PARAM(HSG_W_XP_B0), PARAM(HSG_W_XP_W1), PARAM(HSG_W_XP_B1), w->hp1, w->dp1);
        TR_CHECK_LAUNCH(c);
        pool_kernel<<<nblk(B), 256, 0, s>>>(w->ov, w->dp3, B, w->varh); TR_CHECK_LAUNCH(c);
        pool_kernel<<<nblk(B), 256, 0, s>>>(w->op, w->dp1, B, w->probah); TR_CHECK_LAUNCH(c);
    }
    // 6. loss: 0 classification (weighted BCE with logits), 1 regression (MSE against w), 2 zinb, 3 rank
    HSG_CUDA(cudaMemsetAsync(w->loss, 0, 4 * sizeof(float), s));
    if (have_w && (have_y || loss_mode != 0)) {
        if (loss_mode == 0) bce_kernel<<<nblk(B), 256, 0, s>>>(w->mean, w->y, w->w, B, w->loss, w->dmean);
        else if (loss_mode == 1) mse_kernel<<<nblk(B), 256, 0, s>>>(w->mean, w->w, B, w->loss, w->dmean);
        else if (loss_mode == 2)
            zinb_kernel<<<nblk(B), 256, 0, s>>>(w->mean, w->varh, w->probah, w->w, w->x, g_opts[c].c2w, B, w->loss, w->dmean, w->dvar,
                                                w->dproba, w->pred);
        else {
            rank_pairs_kernel<<<nblk(B, 128), 128, 0, s>>>(w->mean, w->w, B, g_opts[c].rank_thres, w->gsum, w->loss + 1);
            TR_CHECK_LAUNCH(c);
            rank_finish_kernel<<<nblk(B), 256, 0, s>>>(w->mean, w->gsum, w->loss + 1, B, w->loss, w->dmean);
        }
        TR_CHECK_LAUNCH(c);
    }
    if (mean_out_host) HSG_CUDA(cudaMemcpyAsync(mean_out_host, w->mean, B * sizeof(float), cudaMemcpyDeviceToHost, s));
    if (loss_out_host) HSG_CUDA(cudaMemcpyAsync(loss_out_host, w->loss, sizeof(float), cudaMemcpyDeviceToHost, s));
    if (mean_out_host || loss_out_host) HSG_CUDA(cudaStreamSynchronize(s));
    return 0;
}

// backward of the last hsg_score_fwd batch: accumulates into the bound flat gradient buffer (the caller zeroes it)
// Upstream gradients of the three heads from an external loss (Hyper_SAGNN.forward used under torch autograd): replaces the
// gradients the built-in loss kernels would have written.  Device pointers, [B] each; dvar / dproba may be NULL (those heads then
// receive no gradient and the backward pass skips them).
extern "C" int hsg_score_set_head_grads(hsg_ctx* c, const float* dmean_dev, const float* dvar_dev, const float* dproba_dev, void* stream) {
    HSG_CHECK(c && dmean_dev, "bad arguments");
    HSG_CUDA(cudaSetDevice(c->device));
    auto it = g_ws.find(c);
    HSG_CHECK(it != g_ws.end() && it->second && it->second->B > 0, "hsg_score_set_head_grads: no forward batch");
    TrainWs* w = it->second;
    cudaStream_t s = (cudaStream_t)stream;
    const size_t nb = (size_t)w->B * sizeof(float);
    HSG_CUDA(cudaMemcpyAsync(w->dmean, dmean_dev, nb, cudaMemcpyDeviceToDevice, s));
    if (dvar_dev || dproba_dev) {
        HSG_CHECK(w->mode == 2, "var / proba gradients need a forward pass that produced those heads (loss_mode 2)");
        if (dvar_dev) HSG_CUDA(cudaMemcpyAsync(w->dvar, dvar_dev, nb, cudaMemcpyDeviceToDevice, s));
        else HSG_CUDA(cudaMemsetAsync(w->dvar, 0, nb, s));
        if (dproba_dev) HSG_CUDA(cudaMemcpyAsync(w->dproba, dproba_dev, nb, cudaMemcpyDeviceToDevice, s));
        else HSG_CUDA(cudaMemsetAsync(w->dproba, 0, nb, s));
    } else if (w->mode == 2) {
        w->mode = 0;                 // the main head only: the backward of the var / proba heads is skipped
    }
    return 0;
}

extern "C" int hsg_score_bwd(hsg_ctx* c, void* stream) {
    HSG_CHECK(c && g_ws.count(c) && g_ws[c]->B > 0, "no forward batch to differentiate");
    HSG_CUDA(cudaSetDevice(c->device));
    TrainWs* w = g_ws[c];
    TrainBind* b = bind_of(c);
    cudaStream_t s = (cudaStream_t)stream;
    const int d = c->d, hd = d / 2, B = w->B, R = 3 * B, NB = 2 * B, chrom = w->chrom, nc = c->bins[chrom], in_c = b->in_dim[chrom + 1];
    const int bin_base = c->n_cells + 1 + c->bin_off[chrom];
    const int64_t nRd = (int64_t)R * d;
    // heads
    xp_bwd_kernel<<<nblk(B), 256, 0, s>>>(w->x, B, c->n_chrom + 1, c->cf, c->w[HSG_W_ATTR], c->w[HSG_W_CELL_FEATS], w->only_model,
                                          PARAM(HSG_W_XP2_W1), w->hp2, w->dmean, GRAD(HSG_W_XP2_W0), GRAD(HSG_W_XP2_B0),
                                          GRAD(HSG_W_XP2_W1), GRAD(HSG_W_XP2_B1));
    TR_CHECK_LAUNCH(c);
    head_bwd_kernel<<<nblk(R), 256, (256 / hd) * (hd + 1) * sizeof(float), s>>>(w->dmean, w->cact, R, hd, PARAM(HSG_W_CLS_W1), w->dc, GRAD(HSG_W_CLS_W1),
                                                                   GRAD(HSG_W_CLS_B1));
    TR_CHECK_LAUNCH(c);
    dswish_colsum_kernel<<<dim3((hd + 63) / 64, COLSUM_SLICES), 256, 0, s>>>(w->dc, w->cpre, R, hd, GRAD(HSG_W_CLS_B0), 0.f, 0ull, 0u); TR_CHECK_LAUNCH(c);
    if (LIN_DW(R, hd, d, w->dc, hd, w->u, d, GRAD(HSG_W_CLS_W0), d)) return 1;
    if (LIN_DX(R, hd, d, w->dc, hd, PARAM(HSG_W_CLS_W0), d, w->du, d, 0.f)) return 1;
    HSG_CUDA(cudaMemsetAsync(w->dS, 0, nRd * sizeof(float), s));
    if (w->mode == 2) {        // zinb: the var and proba heads read the normalised static branch
        xp_bwd_kernel<<<nblk(B), 256, 0, s>>>(w->x, B, c->n_chrom + 1, c->cf, c->w[HSG_W_ATTR], c->w[HSG_W_CELL_FEATS], w->only_model,
                                              PARAM(HSG_W_XP3_W1), w->hp3, w->dvar, GRAD(HSG_W_XP3_W0), GRAD(HSG_W_XP3_B0),
                                              GRAD(HSG_W_XP3_W1), GRAD(HSG_W_XP3_B1));
        TR_CHECK_LAUNCH(c);
        xp_bwd_kernel<<<nblk(B), 256, 0, s>>>(w->x, B, c->n_chrom + 1, c->cf, c->w[HSG_W_ATTR], c->w[HSG_W_CELL_FEATS], w->only_model,
                                              PARAM(HSG_W_XP_W1), w->hp1, w->dproba, GRAD(HSG_W_XP_W0), GRAD(HSG_W_XP_B0),
                                              GRAD(HSG_W_XP_W1), GRAD(HSG_W_XP_B1));
        TR_CHECK_LAUNCH(c);
        const int slots[2][4] = {{HSG_W_VAR_W0, HSG_W_VAR_B0, HSG_W_VAR_W1, HSG_W_VAR_B1}, {HSG_W_PRB_W0, HSG_W_PRB_B0, HSG_W_PRB_W1, HSG_W_PRB_B1}};
        float* dh[2] = {w->dvar, w->dproba}; float* act[2] = {w->vact, w->pact}; float* pre[2] = {w->vpre, w->ppre};
        for (int q = 0; q < 2; ++q) {
            head_bwd_kernel<<<nblk(R), 256, (256 / hd) * (hd + 1) * sizeof(float), s>>>(dh[q], act[q], R, hd, PARAM(slots[q][2]), w->dc, GRAD(slots[q][2]),
                                                                           GRAD(slots[q][3]));
            TR_CHECK_LAUNCH(c);
            dswish_colsum_kernel<<<dim3((hd + 63) / 64, COLSUM_SLICES), 256, 0, s>>>(w->dc, pre[q], R, hd, GRAD(slots[q][1]), 0.f, 0ull, 0u); TR_CHECK_LAUNCH(c);
            if (LIN_DW(R, hd, d, w->dc, hd, w->Sn, d, GRAD(slots[q][0]), d)) return 1;
            if (LIN_DX(R, hd, d, w->dc, hd, PARAM(slots[q][0]), d, w->dS, d, 1.f)) return 1;
        }
    }
    sqdiff_bwd_kernel<<<nblk(nRd), 256, 0, s>>>(w->Dn, w->Sn, w->du, nRd, w->dD, w->dS); TR_CHECK_LAUNCH(c);
    if (ln_bwd(c, s, w->dD, w->z, R, PARAM(HSG_W_LN1_G), w->mu_z, w->rs_z, w->dz, 0, GRAD(HSG_W_LN1_G), GRAD(HSG_W_LN1_B))) return 1;
    if (ln_bwd(c, s, w->dS, w->sp, R, PARAM(HSG_W_LN2_G), w->mu_p, w->rs_p, w->dsp, 0, GRAD(HSG_W_LN2_G), GRAD(HSG_W_LN2_B))) return 1;
    // pff_n1:  z = y + W1 h + b1 ; h = swish(W0 xn + b0) ; xn = LN(y)
    colsum_kernel<<<dim3((d + 63) / 64, COLSUM_SLICES), 256, 0, s>>>(w->dz, R, d, GRAD(HSG_W_PFF_B1)); TR_CHECK_LAUNCH(c);
    if (LIN_DW(R, d, d, w->dz, d, w->h, d, GRAD(HSG_W_PFF_W1), d)) return 1;
    if (LIN_DX(R, d, d, w->dz, d, PARAM(HSG_W_PFF_W1), d, w->dh, d, 0.f)) return 1;
    dswish_colsum_kernel<<<dim3((d + 63) / 64, COLSUM_SLICES), 256, 0, s>>>(w->dh, w->hpre, R, d, GRAD(HSG_W_PFF_B0), w->dropout ? 0.4f : 0.f, w->seed, 3u);
    TR_CHECK_LAUNCH(c);
    if (LIN_DW(R, d, d, w->dh, d, w->xn, d, GRAD(HSG_W_PFF_W0), d)) return 1;
    if (LIN_DX(R, d, d, w->dh, d, PARAM(HSG_W_PFF_W0), d, w->dxn, d, 0.f)) return 1;
    HSG_CUDA(cudaMemcpyAsync(w->dy, w->dz, nRd * sizeof(float), cudaMemcpyDeviceToDevice, s));      // residual branch
    if (ln_bwd(c, s, w->dxn, w->yv, R, PARAM(HSG_W_PFF_LN_G), w->mu_y, w->rs_y, w->dy, 1, GRAD(HSG_W_PFF_LN_G), GRAD(HSG_W_PFF_LN_B))) return 1;
    // attention outputs:  y = fc1 ctx ; sp = fc2 V   (each followed by dropout 0.3 in training mode)
    if (w->dropout) {
        dropout_kernel<<<nblk(nRd), 256, 0, s>>>(w->dy, nRd, 0.3f, w->seed, 1); TR_CHECK_LAUNCH(c);
        dropout_kernel<<<nblk(nRd), 256, 0, s>>>(w->dsp, nRd, 0.3f, w->seed, 2); TR_CHECK_LAUNCH(c);
    }
    if (LIN_DW(R, d, HSG_HD, w->dy, d, w->ctx, HSG_HD, GRAD(HSG_W_FC1), HSG_HD)) return 1;
    if (LIN_DX(R, d, HSG_HD, w->dy, d, PARAM(HSG_W_FC1), HSG_HD, w->dctx, HSG_HD, 0.f)) return 1;
    if (LIN_DW(R, d, HSG_HD, w->dsp, d, w->V, HSG_HD, GRAD(HSG_W_FC2), HSG_HD)) return 1;
    attn_bwd_kernel<<<nblk((int64_t)B * 32, 128), 128, 0, s>>>(w->Q, w->K, w->V, w->P, w->dctx, B, w->dQ, w->dK, w->dV); TR_CHECK_LAUNCH(c);
    if (LIN_DX(R, d, HSG_HD, w->dsp, d, PARAM(HSG_W_FC2), HSG_HD, w->dV, HSG_HD, 1.f)) return 1;
    if (LIN_DW(R, HSG_HD, d, w->dQ, HSG_HD, w->qin, d, GRAD(HSG_W_QS), d)) return 1;
    if (LIN_DW(R, HSG_HD, d, w->dK, HSG_HD, w->kin, d, GRAD(HSG_W_KS), d)) return 1;
    if (LIN_DW(R, HSG_HD, d, w->dV, HSG_HD, w->vin, d, GRAD(HSG_W_VS), d)) return 1;
    if (LIN_DX(R, HSG_HD, d, w->dQ, HSG_HD, PARAM(HSG_W_QS), d, w->dqin, d, 0.f)) return 1;
    if (LIN_DX(R, HSG_HD, d, w->dK, HSG_HD, PARAM(HSG_W_KS), d, w->dkin, d, 0.f)) return 1;
    if (LIN_DX(R, HSG_HD, d, w->dV, HSG_HD, PARAM(HSG_W_VS), d, w->dvin, d, 0.f)) return 1;
    if (ln_bwd(c, s, w->dqin, w->dyn, R, PARAM(HSG_W_ALN1_G), w->mu_d, w->rs_d, w->ddyn, 0, GRAD(HSG_W_ALN1_G), GRAD(HSG_W_ALN1_B))) return 1;
    if (ln_bwd(c, s, w->dkin, w->dyn, R, PARAM(HSG_W_ALN2_G), w->mu_d, w->rs_d, w->ddyn, 1, GRAD(HSG_W_ALN2_G), GRAD(HSG_W_ALN2_B))) return 1;
    if (ln_bwd(c, s, w->dvin, w->sta, R, PARAM(HSG_W_ALN3_G), w->mu_s, w->rs_s, w->dsta, 0, GRAD(HSG_W_ALN3_G), GRAD(HSG_W_ALN3_B))) return 1;
    // node embedding
    HSG_CUDA(cudaMemsetAsync(w->dEall, 0, (size_t)nc * d * sizeof(float), s));
    disassemble_kernel<<<nblk(nRd), 256, 0, s>>>(w->x, B, d, bin_base, w->ddyn, w->dsta, w->stage == 2 ? w->dcomb : nullptr, w->dEall);
    TR_CHECK_LAUNCH(c);
    if (w->stage == 2) {
        const int64_t nNd = (int64_t)NB * d;
        dswish_colsum_kernel<<<dim3((d + 63) / 64, COLSUM_SLICES), 256, 0, s>>>(w->dcomb, w->combpre, NB, d, GRAD(HSG_W_SAGE_B), 0.f, 0ull, 0u); TR_CHECK_LAUNCH(c);
        float* dG = GRAD(HSG_W_SAGE_W); const float* G = PARAM(HSG_W_SAGE_W);
        if (LIN_DW(NB, d, d, w->dcomb, d, w->neigh, d, dG, 2 * d)) return 1;
        if (LIN_DW(NB, d, d, w->dcomb, d, w->selfb, d, dG + d, 2 * d)) return 1;
        if (LIN_DX(NB, d, d, w->dcomb, d, G, 2 * d, w->dneigh, d, 0.f)) return 1;
        if (LIN_DX(NB, d, d, w->dcomb, d, G + d, 2 * d, w->dself, d, 0.f)) return 1;
        scatter_rows_kernel<<<nblk(nNd), 256, 0, s>>>(w->x, B, d, bin_base, w->dself, w->dEall); TR_CHECK_LAUNCH(c);
        if (d == 64) spmm_bwd_kernel<64><<<(NB + 7) / 8, 256, 0, s>>>(w->indptr, w->sampled ? w->rowcnt : nullptr, w->cols, w->vals, w->dneigh, NB, w->dEall);
        else spmm_bwd_kernel<128><<<(NB + 7) / 8, 256, 0, s>>>(w->indptr, w->sampled ? w->rowcnt : nullptr, w->cols, w->vals, w->dneigh, NB, w->dEall);
        TR_CHECK_LAUNCH(c);
    }
    if (w->stage == 1 && w->cell_on_hook) {
        const int F = b->in_dim[0];
        cell_grad_rows_kernel<<<nblk((int64_t)B * d), 256, 0, s>>>(w->ddyn, w->dsta, B, d, w->dcell); TR_CHECK_LAUNCH(c);
        colsum_kernel<<<dim3((d + 63) / 64, COLSUM_SLICES), 256, 0, s>>>(w->dcell, B, d, b->grads + b->off_pb[0]); TR_CHECK_LAUNCH(c);
        if (LIN_DW(B, d, F, w->dcell, d, w->Xg, F, b->grads + b->off_pw[0], F)) return 1;
    }
    // E = swish(X W^T + b)
    dswish_colsum_kernel<<<dim3((d + 63) / 64, COLSUM_SLICES), 256, 0, s>>>(w->dEall, w->Epre, nc, d, b->grads + b->off_pb[chrom + 1], 0.f, 0ull, 0u); TR_CHECK_LAUNCH(c);
    if (LIN_DW(nc, d, in_c, w->dEall, d, b->feat[chrom + 1], in_c, b->grads + b->off_pw[chrom + 1], in_c)) return 1;
    return 0;
}


// ------------------------------------------------------------------------------------------------ stage-1 reconstruction loss
// TiedAutoEncoder.forward(return_recon=True) for the cell branch over ALL cells + F.mse_loss (Higashi_wrapper.py:869-873,
// Modules.py:255-288): enc = X W0^T + b0 ; recon = swish(swish(enc)) . R + rb ; loss = mean((recon - target)^2).
// Gradients go to a scratch block laid out like the flat buffer slots they belong to; hsg_recon_apply adds
// `scale * scratch` to the bound gradient buffer (train_epoch weights the loss by ratio1 after comparing gradient norms).
struct ReconState { float* scratch = nullptr; size_t cap = 0; float* act = nullptr; size_t act_cap = 0; int64_t off_rw = -1, off_rb = -1; int F = 0; };
static std::map<hsg_ctx*, ReconState> g_recon;

extern "C" int hsg_train_bind_recon(hsg_ctx* c, int64_t off_reverse_w, int64_t off_recon_b) {
    HSG_CHECK(c && c->d > 0, "context not initialised");
    g_recon[c].off_rw = off_reverse_w; g_recon[c].off_rb = off_recon_b;
    return 0;
}

extern "C" int hsg_recon_loss(hsg_ctx* c, float* loss_host, float* gnorm_w0_host, void* stream) {
    HSG_CHECK(c && c->d > 0, "context not initialised");
    HSG_CUDA(cudaSetDevice(c->device));
    TrainBind* b = bind_of(c);
    ReconState& rs = g_recon[c];
    HSG_CHECK(b->params && !b->feat.empty() && b->feat[0] && b->off_pw[0] >= 0 && rs.off_rw >= 0 && rs.off_rb >= 0,
              "bind the cell branch (features, weights, reverse weights) first");
    cudaStream_t s = (cudaStream_t)stream;
    const int n = c->n_cells, d = c->d, F = b->in_dim[0];
    rs.F = F;
    size_t need_s = 2 * (size_t)d * F + d + F + 8, need_a = 3 * (size_t)n * d + (size_t)n * F;
    if (rs.cap < need_s) { if (rs.scratch) cudaFree(rs.scratch); HSG_CUDA(cudaMalloc(&rs.scratch, need_s * sizeof(float))); rs.cap = need_s; }
    if (rs.act_cap < need_a) { if (rs.act) cudaFree(rs.act); HSG_CUDA(cudaMalloc(&rs.act, need_a * sizeof(float))); rs.act_cap = need_a; }
    float *gW = rs.scratch, *gR = gW + (size_t)d * F, *gb = gR + (size_t)d * F, *grb = gb + d, *misc = grb + F;
    float *enc = rs.act, *a1 = enc + (size_t)n * d, *a2 = a1 + (size_t)n * d, *rec = a2 + (size_t)n * d;
    HSG_CUDA(cudaMemsetAsync(rs.scratch, 0, need_s * sizeof(float), s));
    const float* W0 = b->params + b->off_pw[0]; const float* b0 = b->params + b->off_pb[0];
    const float* Rw = b->params + rs.off_rw; const float* rb = b->params + rs.off_rb;
    if (LIN_FWD(n, d, F, b->feat[0], F, W0, F, enc, d, 0.f)) return 1;
    bias_act_kernel<<<nblk((int64_t)n * d), 256, 0, s>>>(enc, b0, (int64_t)n * d, d, 0, nullptr); TR_CHECK_LAUNCH(c);
    swish2_fwd_kernel<<<nblk((int64_t)n * d), 256, 0, s>>>(enc, (int64_t)n * d, a1, a2); TR_CHECK_LAUNCH(c);
    // recon[n,F] = a2[n,d] . R[d,F]
    if (gemm(c, s, false, false, n, F, d, a2, d, Rw, F, rec, F, 0.f)) return 1;
    recon_mse_kernel<<<nblk((int64_t)n * F), 256, 0, s>>>(rec, rb, b->feat[0], (int64_t)n * F, F, misc); TR_CHECK_LAUNCH(c);
    // backward: rec now holds d(loss)/d(recon)
    colsum_kernel<<<dim3((F + 63) / 64, COLSUM_SLICES), 256, 0, s>>>(rec, n, F, grb); TR_CHECK_LAUNCH(c);
    if (gemm(c, s, true, false, d, F, n, a2, d, rec, F, gR, F, 1.f)) return 1;              // dR = a2^T drecon
    if (gemm(c, s, false, true, n, d, F, rec, F, Rw, F, a2, d, 0.f)) return 1;              // da2 = drecon R^T (reuse a2)
    swish2_bwd_kernel<<<nblk((int64_t)n * d), 256, 0, s>>>(a2, enc, a1, (int64_t)n * d); TR_CHECK_LAUNCH(c);
    colsum_kernel<<<dim3((d + 63) / 64, COLSUM_SLICES), 256, 0, s>>>(a2, n, d, gb); TR_CHECK_LAUNCH(c);
    if (LIN_DW(n, d, F, a2, d, b->feat[0], F, gW, F)) return 1;
    sumsq_kernel<<<nblk((int64_t)d * F), 256, 0, s>>>(gW, (int64_t)d * F, misc + 1); TR_CHECK_LAUNCH(c);
    float h[2] = {0, 0};
    HSG_CUDA(cudaMemcpyAsync(h, misc, 2 * sizeof(float), cudaMemcpyDeviceToHost, s));
    HSG_CUDA(cudaStreamSynchronize(s));
    if (loss_host) *loss_host = h[0];
    if (gnorm_w0_host) *gnorm_w0_host = sqrtf(h[1]);
    return 0;
}

extern "C" int hsg_recon_apply(hsg_ctx* c, float scale, void* stream) {
    HSG_CHECK(c && g_recon.count(c) && g_recon[c].scratch, "call hsg_recon_loss first");
    TrainBind* b = bind_of(c);
    ReconState& rs = g_recon[c];
    cudaStream_t s = (cudaStream_t)stream;
    const int d = c->d, F = rs.F;
    float *gW = rs.scratch, *gR = gW + (size_t)d * F, *gb = gR + (size_t)d * F, *grb = gb + d;
    axpy_scaled_kernel<<<nblk((int64_t)d * F), 256, 0, s>>>(b->grads + b->off_pw[0], gW, scale, (int64_t)d * F); TR_CHECK_LAUNCH(c);
    axpy_scaled_kernel<<<nblk((int64_t)d * F), 256, 0, s>>>(b->grads + rs.off_rw, gR, scale, (int64_t)d * F); TR_CHECK_LAUNCH(c);
    axpy_scaled_kernel<<<nblk(d), 256, 0, s>>>(b->grads + b->off_pb[0], gb, scale, d); TR_CHECK_LAUNCH(c);
    axpy_scaled_kernel<<<nblk(F), 256, 0, s>>>(b->grads + rs.off_rb, grb, scale, F); TR_CHECK_LAUNCH(c);
    return 0;
}

// ------------------------------------------------------------------------------------------------ teardown
// Called by hsg_destroy: all per-context training state is keyed by the context pointer, and a later context may be given the
// same address -- stale entries would hand it buffers sized for another model.
void train_release(hsg_ctx* c) {
    auto w = g_ws.find(c);
    if (w != g_ws.end()) {
        if (w->second) {
            cudaFree(w->second->buf); cudaFree(w->second->indptr); cudaFree(w->second->rowcnt); cudaFree(w->second->cols);
            cudaFree(w->second->vals); cudaFree(w->second->x);
            if (w->second->Xg) cudaFree(w->second->Xg);
            delete w->second;
        }
        g_ws.erase(w);
    }
    auto b = g_bind.find(c);
    if (b != g_bind.end()) {
        if (b->second) { for (float* f : b->second->feat) if (f) cudaFree(f); delete b->second; }
        g_bind.erase(b);
    }
    auto o = g_opts.find(c);
    if (o != g_opts.end()) { if (o->second.c2w) cudaFree(o->second.c2w); g_opts.erase(o); }
    auto r = g_recon.find(c);
    if (r != g_recon.end()) { if (r->second.scratch) cudaFree(r->second.scratch); if (r->second.act) cudaFree(r->second.act); g_recon.erase(r); }
    auto s = g_sampler.find(c);
    if (s != g_sampler.end()) {
        SamplerWs* sw = s->second;
        if (sw) {
            cudaFree(sw->pos); cudaFree(sw->pos_w); cudaFree(sw->neg); cudaFree(sw->flag); cudaFree(sw->scan); cudaFree(sw->x_all);
            cudaFree(sw->y_all); cudaFree(sw->w_all); cudaFree(sw->cand_cnt); cudaFree(sw->cand_off); cudaFree(sw->s_col);
            cudaFree(sw->s_val); cudaFree(sw->s_sum); cudaFree(sw->s_first); cudaFree(sw->counts); cudaFree(sw->draws);
            if (sw->h_counts) cudaFreeHost(sw->h_counts);
            if (sw->sized) cudaEventDestroy(sw->sized);
            if (sw->consumed) cudaEventDestroy(sw->consumed);
            if (sw->side) cudaStreamDestroy(sw->side);
            delete sw;
        }
        g_sampler.erase(s);
    }
}
