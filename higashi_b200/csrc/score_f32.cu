// score_f32.cu -- K2 (fp32 mode): the fused Hyper-SAGNN scorer over a tile of triplets.
//
// Per triplet (cell, b_i, b_j) with tokens (cell, b_i, b_j)  [SURVEY.md 3.3a; Modules.py:726-783, 1029-1095, 865-889]:
//   s_h[i,j] = Q_h[i].K_h[j] / 4 ;  a = softmax_j(s*m)*m / (sum + 1e-13), m = 1 - I3     (Modules.py:948-953)
//   ctx[i]   = concat_h sum_j a_h[i,j] V_h[j]
//   y[i]     = fc1 . ctx[i]
//   z[i]     = y[i] + W1 . swish(W0 . LN_pff(y[i]) + b0) + b1                            (pff_n1)
//   o[i]     = c1 . swish(C0 . (LN_1(z[i]) - S[tok_i])^2 + cb0) + cb1                     (pff_classifier)
//   score    = act(mean_i o[i] + pair_bias)
// Everything stays in shared memory / registers; the only HBM traffic is the 4-byte score.
//
// fp32 everywhere (parity bar 1e-5 absolute).  32 triplets (96 token rows) per block of 256 threads;
// activations are staged transposed ([feature][row]) so the register-tiled GEMMs read rows with
// 128-bit shared loads; weights (transposed once per hsg_prepare) come through L1/L2.
#include "common.cuh"

#define SC_T 32
#define SC_R 96
#define SC_LDA 100

template <int CT>
__device__ __forceinline__ void sc_tile_gemm(const float* __restrict__ sA, int K, int row0,
                                             const float* __restrict__ Wt, int ldw, int col0, float (&acc)[4][CT]) {
#pragma unroll 4
    for (int k = 0; k < K; ++k) {
        float4 a = *reinterpret_cast<const float4*>(sA + k * SC_LDA + row0);
        float wv[CT];
#pragma unroll
        for (int q = 0; q < CT; q += 4) {
            float4 w4 = __ldg(reinterpret_cast<const float4*>(Wt + (size_t)k * ldw + col0 + q));
            wv[q] = w4.x; wv[q + 1] = w4.y; wv[q + 2] = w4.z; wv[q + 3] = w4.w;
        }
#pragma unroll
        for (int q = 0; q < CT; ++q) {
            acc[0][q] = fmaf(a.x, wv[q], acc[0][q]);
            acc[1][q] = fmaf(a.y, wv[q], acc[1][q]);
            acc[2][q] = fmaf(a.z, wv[q], acc[2][q]);
            acc[3][q] = fmaf(a.w, wv[q], acc[3][q]);
        }
    }
}

__device__ __forceinline__ float dot16(const float* __restrict__ a, const float* __restrict__ b) {
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        float4 x = __ldg(reinterpret_cast<const float4*>(a) + q);
        float4 y = __ldg(reinterpret_cast<const float4*>(b) + q);
        s = fmaf(x.x, y.x, s); s = fmaf(x.y, y.y, s); s = fmaf(x.z, y.z, s); s = fmaf(x.w, y.w, s);
    }
    return s;
}

// masked 3-token softmax row: logits (0 on the diagonal, la, lb) -> weights of the two off-diagonal keys
__device__ __forceinline__ void masked_row(float la, float lb, float& wa, float& wb) {
    float mx = fmaxf(0.f, fmaxf(la, lb));
    float e0 = expf(0.f - mx), ea = expf(la - mx), eb = expf(lb - mx);
    float Z = e0 + ea + eb;
    float pa = ea / Z, pb = eb / Z;
    float den = (pa + pb) + 1e-13f;
    wa = pa / den; wb = pb / den;
}

struct ScoreParams {
    // token tables
    const float* cellQ; const float* cellK; const float* cellV; const float* cellS;
    const float* binV; const float* binS; const float* QK;
    // weights
    const float* fc1T; const float* pff_g; const float* pff_b; const float* w0T; const float* b0;
    const float* w1T; const float* b1; const float* ln1_g; const float* ln1_b;
    const float* c0T; const float* cb0; const float* cw1; const float* cb1;
    // pair bias inputs for the explicit mode
    const float* attr; const float* xw0; const float* xb0; const float* xw1; const float* xb1;
    int n_cells, n_bins, n_chrom, cf, act;
    // implicit enumeration
    const int2* pairs; const float* pair_bias; int64_t P; int cell_start; int tiles_per_cell;
    // explicit triplets
    const int64_t* x; int64_t B;
    float* out;
};

__device__ __forceinline__ float pair_bias_dev(const ScoreParams& p, int bi, int bj) {
    const int A = p.n_chrom + 1, in_w = 2 * A + p.cf;
    const float* ai = p.attr + (int64_t)(p.n_cells + 1 + bi) * A;
    const float* aj = p.attr + (int64_t)(p.n_cells + 1 + bj) * A;
    float o = p.xb1[0];
    for (int u = 0; u < 4; ++u) {
        float acc = 0.f;
        const float* wr = p.xw0 + u * in_w;
        for (int t = 0; t < A; ++t) acc = fmaf(wr[t], ai[t], acc);
        for (int t = 0; t < A; ++t) acc = fmaf(wr[A + t], aj[t], acc);
        acc += p.xb0[u];
        o = fmaf(p.xw1[u], swish_acc(acc), o);
    }
    return o;
}

template <int D, bool EXPLICIT>
__global__ void __launch_bounds__(256) score_f32_kernel(ScoreParams p) {
    extern __shared__ __align__(16) float sm[];
    float* ctxT = sm;                               // [128][LDA]   ctx, later h (D rows) and c (D/2 rows)
    float* yT = ctxT + HSG_HD * SC_LDA;             // [D][LDA]     y, then z
    float* xT = yT + D * SC_LDA;                    // [D][LDA]     LN(y), later (LN1(z)-S)^2
    __shared__ int s_cell[SC_T], s_bi[SC_T], s_bj[SC_T], s_slot[SC_T];
    __shared__ float s_bias[SC_T];
    __shared__ float s_o[SC_R];
    __shared__ long long s_out[SC_T];
    const int tid = threadIdx.x;

    if (tid < SC_T) {
        int cell = 0, bi = 0, bj = 0, slot = 0; float bias = 0.f; long long oidx = -1;
        if (EXPLICIT) {
            int64_t i = (int64_t)blockIdx.x * SC_T + tid;
            if (i < p.B) {
                cell = (int)p.x[i * 3] - 1;
                bi = (int)p.x[i * 3 + 1] - p.n_cells - 1;
                bj = (int)p.x[i * 3 + 2] - p.n_cells - 1;
                bias = pair_bias_dev(p, bi, bj);
                oidx = i;
            }
        } else {
            int cb = blockIdx.x / p.tiles_per_cell, tile = blockIdx.x - cb * p.tiles_per_cell;
            int64_t pi = (int64_t)tile * SC_T + tid;
            cell = p.cell_start + cb; slot = cb;
            if (pi < p.P) {
                int2 pr = p.pairs[pi];
                bi = pr.x; bj = pr.y; bias = p.pair_bias[pi];
                oidx = (int64_t)cb * p.P + pi;
            }
        }
        s_cell[tid] = cell; s_bi[tid] = bi; s_bj[tid] = bj; s_slot[tid] = slot; s_bias[tid] = bias; s_out[tid] = oidx;
    }
    __syncthreads();

    // ---- attention: one (triplet, head) per thread ------------------------------------------------
    {
        const int r = tid & 31, h = tid >> 5;
        const int cell = s_cell[r], bi = s_bi[r], bj = s_bj[r], slot = s_slot[r];
        const float* q0 = p.cellQ + (int64_t)cell * HSG_HD + h * 16;
        const float* k0 = p.cellK + (int64_t)cell * HSG_HD + h * 16;
        const float* v0 = p.cellV + (int64_t)cell * HSG_HD + h * 16;
        const float* qk1 = p.QK + ((int64_t)slot * p.n_bins + bi) * 2 * HSG_HD + h * 16;
        const float* qk2 = p.QK + ((int64_t)slot * p.n_bins + bj) * 2 * HSG_HD + h * 16;
        const float* v1 = p.binV + (int64_t)bi * HSG_HD + h * 16;
        const float* v2 = p.binV + (int64_t)bj * HSG_HD + h * 16;
        const float* q1 = qk1; const float* k1 = qk1 + HSG_HD;
        const float* q2 = qk2; const float* k2 = qk2 + HSG_HD;
        float w01, w02, w10, w12, w20, w21;
        masked_row(dot16(q0, k1) * 0.25f, dot16(q0, k2) * 0.25f, w01, w02);
        masked_row(dot16(q1, k0) * 0.25f, dot16(q1, k2) * 0.25f, w10, w12);
        masked_row(dot16(q2, k0) * 0.25f, dot16(q2, k1) * 0.25f, w20, w21);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float4 a = __ldg(reinterpret_cast<const float4*>(v0) + q);
            float4 b = __ldg(reinterpret_cast<const float4*>(v1) + q);
            float4 c = __ldg(reinterpret_cast<const float4*>(v2) + q);
            const float va[4] = {a.x, a.y, a.z, a.w}, vb[4] = {b.x, b.y, b.z, b.w}, vc[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float* dst = ctxT + (h * 16 + q * 4 + e) * SC_LDA + r;
                // torch.bmm(attn, v): zero-weight diagonal term first, then keys in order
                dst[0] = fmaf(w02, vc[e], w01 * vb[e]);
                dst[32] = fmaf(w12, vc[e], w10 * va[e]);
                dst[64] = fmaf(w21, vb[e], w20 * va[e]);
            }
        }
    }
    __syncthreads();

    // ---- y = ctx . fc1^T ---------------------------------------------------------------------------
    constexpr int NRG = SC_R / 4;
    for (int t = tid; t < NRG * (D / 4); t += 256) {
        int rg = t % NRG, cg = t / NRG;
        float acc[4][4] = {};
        sc_tile_gemm<4>(ctxT, HSG_HD, rg * 4, p.fc1T, D, cg * 4, acc);
#pragma unroll
        for (int q = 0; q < 4; ++q)
            *reinterpret_cast<float4*>(yT + (cg * 4 + q) * SC_LDA + rg * 4) = make_float4(acc[0][q], acc[1][q], acc[2][q], acc[3][q]);
    }
    __syncthreads();

    // ---- pff_n1 layer norm (per row) -----------------------------------------------------------------
    if (tid < SC_R) {
        float s = 0.f;
        for (int f = 0; f < D; ++f) s += yT[f * SC_LDA + tid];
        float mean = s / (float)D, v = 0.f;
        for (int f = 0; f < D; ++f) { float dd = yT[f * SC_LDA + tid] - mean; v = fmaf(dd, dd, v); }
        float rstd = 1.0f / sqrtf(v / (float)D + HSG_LN_EPS);
        for (int f = 0; f < D; ++f) xT[f * SC_LDA + tid] = fmaf((yT[f * SC_LDA + tid] - mean) * rstd, p.pff_g[f], p.pff_b[f]);
    }
    __syncthreads();

    // ---- h = swish(x . W0^T + b0)  -> ctxT rows [0,D) ------------------------------------------------
    for (int t = tid; t < NRG * (D / 4); t += 256) {
        int rg = t % NRG, cg = t / NRG;
        float acc[4][4] = {};
        sc_tile_gemm<4>(xT, D, rg * 4, p.w0T, D, cg * 4, acc);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float bb = p.b0[cg * 4 + q];
            *reinterpret_cast<float4*>(ctxT + (cg * 4 + q) * SC_LDA + rg * 4) =
                make_float4(swish_acc(acc[0][q] + bb), swish_acc(acc[1][q] + bb), swish_acc(acc[2][q] + bb), swish_acc(acc[3][q] + bb));
        }
    }
    __syncthreads();

    // ---- z = h . W1^T + b1 + y  (in place over y) ----------------------------------------------------
    for (int t = tid; t < NRG * (D / 4); t += 256) {
        int rg = t % NRG, cg = t / NRG;
        float acc[4][4] = {};
        sc_tile_gemm<4>(ctxT, D, rg * 4, p.w1T, D, cg * 4, acc);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float bb = p.b1[cg * 4 + q];
            float4* dst = reinterpret_cast<float4*>(yT + (cg * 4 + q) * SC_LDA + rg * 4);
            float4 y = *dst;
            *dst = make_float4((acc[0][q] + bb) + y.x, (acc[1][q] + bb) + y.y, (acc[2][q] + bb) + y.z, (acc[3][q] + bb) + y.w);
        }
    }
    __syncthreads();

    // ---- u = (layer_norm1(z) - S[token])^2 -------------------------------------------------------------
    if (tid < SC_R) {
        const int r = tid & 31, tpos = tid >> 5;
        const float* Srow = (tpos == 0) ? p.cellS + (int64_t)s_cell[r] * D
                                        : p.binS + (int64_t)(tpos == 1 ? s_bi[r] : s_bj[r]) * D;
        float s = 0.f;
        for (int f = 0; f < D; ++f) s += yT[f * SC_LDA + tid];
        float mean = s / (float)D, v = 0.f;
        for (int f = 0; f < D; ++f) { float dd = yT[f * SC_LDA + tid] - mean; v = fmaf(dd, dd, v); }
        float rstd = 1.0f / sqrtf(v / (float)D + HSG_LN_EPS);
        for (int f = 0; f < D; ++f) {
            float dn = fmaf((yT[f * SC_LDA + tid] - mean) * rstd, p.ln1_g[f], p.ln1_b[f]);
            float df = dn - __ldg(Srow + f);
            xT[f * SC_LDA + tid] = df * df;
        }
    }
    __syncthreads();

    // ---- c = swish(u . C0^T + cb0)  [96][D/2] -> ctxT rows [0,D/2) ------------------------------------
    for (int t = tid; t < NRG * (D / 8); t += 256) {
        int rg = t % NRG, cg = t / NRG;
        float acc[4][4] = {};
        sc_tile_gemm<4>(xT, D, rg * 4, p.c0T, D / 2, cg * 4, acc);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float bb = p.cb0[cg * 4 + q];
            *reinterpret_cast<float4*>(ctxT + (cg * 4 + q) * SC_LDA + rg * 4) =
                make_float4(swish_acc(acc[0][q] + bb), swish_acc(acc[1][q] + bb), swish_acc(acc[2][q] + bb), swish_acc(acc[3][q] + bb));
        }
    }
    __syncthreads();

    // ---- o = c . w1 + b1 ; mean over the 3 tokens ; + pair bias ; activation ---------------------------
    if (tid < SC_R) {
        float o = 0.f;
        for (int j = 0; j < D / 2; ++j) o = fmaf(ctxT[j * SC_LDA + tid], p.cw1[j], o);
        s_o[tid] = o + p.cb1[0];
    }
    __syncthreads();
    if (tid < SC_T && s_out[tid] >= 0) {
        float m = ((s_o[tid] + s_o[32 + tid]) + s_o[64 + tid]) / 3.0f + s_bias[tid];
        p.out[s_out[tid]] = apply_act(m, p.act);
    }
}

static void fill_params(hsg_ctx* c, ScoreParams& p) {
    float** w = c->w;
    p.cellQ = c->cellQ; p.cellK = c->cellK; p.cellV = c->cellV; p.cellS = c->cellS;
    p.binV = c->binV; p.binS = c->binS; p.QK = c->QK;
    p.fc1T = c->dw.fc1T; p.pff_g = w[HSG_W_PFF_LN_G]; p.pff_b = w[HSG_W_PFF_LN_B];
    p.w0T = c->dw.w0T; p.b0 = w[HSG_W_PFF_B0]; p.w1T = c->dw.w1T; p.b1 = w[HSG_W_PFF_B1];
    p.ln1_g = w[HSG_W_LN1_G]; p.ln1_b = w[HSG_W_LN1_B];
    p.c0T = c->dw.c0T; p.cb0 = w[HSG_W_CLS_B0]; p.cw1 = w[HSG_W_CLS_W1]; p.cb1 = w[HSG_W_CLS_B1];
    p.attr = w[HSG_W_ATTR]; p.xw0 = w[HSG_W_XP2_W0]; p.xb0 = w[HSG_W_XP2_B0]; p.xw1 = w[HSG_W_XP2_W1]; p.xb1 = w[HSG_W_XP2_B1];
    p.n_cells = c->n_cells; p.n_bins = c->n_bins; p.n_chrom = c->n_chrom; p.cf = c->cf; p.act = c->act;
    p.pairs = c->pairs; p.pair_bias = c->pair_bias; p.P = c->P; p.cell_start = 0; p.tiles_per_cell = 0;
    p.x = nullptr; p.B = 0; p.out = nullptr;
}

template <int D, bool EXPLICIT>
static int launch_score(hsg_ctx* c, const ScoreParams& p, unsigned grid, cudaStream_t s) {
    size_t smem = (size_t)(HSG_HD + 2 * D) * SC_LDA * sizeof(float);
    HSG_CUDA(cudaFuncSetAttribute(score_f32_kernel<D, EXPLICIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    score_f32_kernel<D, EXPLICIT><<<grid, 256, smem, s>>>(p);
    c->launches++;
    HSG_CUDA(cudaGetLastError());
    return 0;
}

int launch_score_pairs_f32(hsg_ctx* c, int cell_start, int n_batch, float* out, cudaStream_t s) {
    if (c->P == 0 || n_batch == 0) return 0;
    ScoreParams p; fill_params(c, p);
    p.cell_start = cell_start; p.out = out;
    p.tiles_per_cell = (int)((c->P + SC_T - 1) / SC_T);
    int64_t grid = (int64_t)n_batch * p.tiles_per_cell;
    HSG_CHECK(grid < 2147483647LL, "score grid too large; lower the cell batch");
    return c->d == 64 ? launch_score<64, false>(c, p, (unsigned)grid, s) : launch_score<128, false>(c, p, (unsigned)grid, s);
}

int launch_score_triplets_f32(hsg_ctx* c, const int64_t* x, int64_t B, float* out, cudaStream_t s) {
    if (B == 0) return 0;
    ScoreParams p; fill_params(c, p);
    p.x = x; p.B = B; p.out = out;
    unsigned grid = (unsigned)((B + SC_T - 1) / SC_T);
    return c->d == 64 ? launch_score<64, true>(c, p, grid, s) : launch_score<128, true>(c, p, grid, s);
}
