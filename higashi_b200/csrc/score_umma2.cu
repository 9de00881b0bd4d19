// score_umma2.cu -- K2, fast tensor-core variant, second generation: row-per-thread attention over CHUNK-MAJOR token tables.
//
// Same math and the same MMA chain as score_umma.cu (Modules.py:726-783, 1029-1095, 865-889, 936-972):
//   [y_c | g] = ctx . [fc1c ; W0' fc1c]^T   (M128 N128 K128)      h = swish(rstd(y_c) g + b0')   -> TMEM A operand
//   z_c       = y_c + h . W1c^T + b1c       (M128 N64  K64+16)    u = (z_c rstd + S')^2          -> TMEM A operand
//   c         = u . C0'^T + cb0/2           (M128 N32  K64+16)    o = cw1 . swish(c) + cb1
// What changed against the first generation is everything in front of MMA 1.  There a warp walked its 32 rows four at a time with
// lane = (row, head), because the Q / K / V tables were row-major: 155 SASS instructions per 8 rows, 42 % of the kernel's
// instructions, most of them address arithmetic, pair look-ups and one softmax per lane (ncu source page of round 1).
// Here the tables are CHUNK-MAJOR inside groups of 32 bins -- table[bin >> 5][chunk c][bin & 31] is one 16-byte chunk (8 halves), so
// consecutive bins are consecutive addresses and the chunks of one bin sit at the compile-time stride of 512 bytes (every load of
// a row is base + immediate: no per-load address arithmetic) -- and thread r of a warpgroup owns triplet r for the attention too:
//   * pairs are enumerated row by row (generate_binpair order), so the 32 rows of a warp are one or two runs of consecutive
//     bin_j under a fixed bin_i: a bin_j load is a fully coalesced 512-byte warp access, a bin_i load is a broadcast;
//   * a thread owns all 8 heads of its row: address arithmetic, the pair, the masked softmax bookkeeping are paid once per row
//     instead of once per (row, head) lane, and every load of a token is independent of every other (32 to 36 LDG.128 in flight
//     per thread instead of 6);
//   * the ctx chunks go straight into the K-major SWIZZLE_128B A tile in natural K order (conflict-free 128-bit stores: the 8
//     rows of a swizzle atom hit 8 distinct 16-byte slots), so fc1's image needs no column permutation.
// Rounding points are the ones of the first generation (half2 products of Q.K accumulated per 16-byte chunk, fp32 logit sum,
// fp32 softmax, half2 ctx), so the parity budget is unchanged.
#include "umma_common.cuh"

#include <algorithm>
#include <cstdlib>

struct Umma2Params {
    // G = HSG_GRP = 32 bins per group, n_grp = ceil(n_bins / 32); a "chunk" is 16 bytes
    const uint4* QKc;        // [slot][n_grp][32 chunks][G]: c < 16 -> Q (x log2(e)/4) halves 8c..8c+7 ; 16 + c -> K halves 8c..8c+7
    const float4* XSc;       // [slot][n_grp][4 quads][G]: quads 0,1 = qcK[0..7] (Q_cell . K_b) ; 2,3 = kcQ[0..7] (Q_b . K_cell), log2 units
    const uint4* binVc;      // [n_grp][16 chunks][G]   V of the bins
    const uint4* cellV;      // [n_cells][16 chunks]    V of the cells
    const float4* binSb;     // [n_grp][16 quads][G]    (beta - S) / gamma of layer_norm1, 4 features per quad
    const float4* cellSb;    // [n_cells][16 quads]
    const unsigned char* wimg;
    float cst[CST_COUNT];
    const int2* pairs; const float* pair_bias; long long P;
    int n_bins, n_grp, n_cells, cell_start, n_batch, tiles_per_cell, act;
    int nwg;                 // warpgroups per CTA (blockDim.x / 128)
    int exp_flags;           // experiments: bit 0 = skip the cell-V loads (wrong results; data-pipe sensitivity test)
    int quad_cells;          // 1: the four warpgroups of a CTA take one pair tile for four consecutive cells; 0: four consecutive tiles
    float* out; __half* out16;
};

// element offset of (bin b, chunk 0) inside a grouped table with CH chunks per bin; chunk c is at + c * HSG_GRP
template <int CH>
__device__ __forceinline__ uint32_t grp_off(uint32_t b) { return (b >> 5) * (uint32_t)(CH * HSG_GRP) + (b & 31u); }

// one head of a logit: 16 products in two half2 chains (one per 16-byte chunk), summed in fp32 -- the rounding of score_umma.cu
__device__ __forceinline__ float head_dot(const uint4 q0, const uint4 q1, const uint4 k0, const uint4 k1) {
    __half2 accA = __hmul2(u2h(q0.x), u2h(k0.x)), accB = __hmul2(u2h(q1.x), u2h(k1.x));
    accA = __hfma2(u2h(q0.y), u2h(k0.y), accA); accB = __hfma2(u2h(q1.y), u2h(k1.y), accB);
    accA = __hfma2(u2h(q0.z), u2h(k0.z), accA); accB = __hfma2(u2h(q1.z), u2h(k1.z), accB);
    accA = __hfma2(u2h(q0.w), u2h(k0.w), accA); accB = __hfma2(u2h(q1.w), u2h(k1.w), accB);
    const float2 fA = __half22float2(accA), fB = __half22float2(accB);
    return (fA.x + fA.y) + (fB.x + fB.y);
}

// ctx chunk = wa * Va + wb * Vb in half2 (product first, then fma: the order of score_umma.cu)
__device__ __forceinline__ void ctx_chunk(uint32_t addr, const __half2 wa2, const __half2 wb2, const uint4 a, const uint4 b) {
    sts128(addr,
           h2u(__hfma2(wb2, u2h(b.x), __hmul2(wa2, u2h(a.x)))), h2u(__hfma2(wb2, u2h(b.y), __hmul2(wa2, u2h(a.y)))),
           h2u(__hfma2(wb2, u2h(b.z), __hmul2(wa2, u2h(a.z)))), h2u(__hfma2(wb2, u2h(b.w), __hmul2(wa2, u2h(a.w)))));
}

// Attention of token position T (0 cell, 1 bin_i, 2 bin_j) for ONE row: row T attends to the two other tokens
//   T = 0: (a, b) = (bin_i, bin_j), logits from the cross-term table only
//   T = 1: (a, b) = (cell, bin_j),  l_a = Q_bi . K_cell (table), l_b = Q_bi . K_bj
//   T = 2: (a, b) = (cell, bin_i),  l_a = Q_bj . K_cell (table), l_b = Q_bj . K_bi
// Part 1: the 8 masked-softmax weight pairs (w_a, w_b) of the row, packed as half2 (the rounding the ctx product uses anyway).
template <int T>
__device__ __forceinline__ void attn_weights(const uint4* __restrict__ QKs, const float4* __restrict__ XSs,
                                             const uint32_t bi, const uint32_t bj, uint32_t (&w)[8]) {
    constexpr uint32_t G = HSG_GRP;
    float la[8], lb[8];
    {
        const float4* xa = XSs + grp_off<4>((T == 2) ? bj : bi) + (T == 0 ? 0u : 2u * G);      // the token whose table row holds l_a
        const float4 x0 = __ldg(xa), x1 = __ldg(xa + G);
        la[0] = x0.x; la[1] = x0.y; la[2] = x0.z; la[3] = x0.w; la[4] = x1.x; la[5] = x1.y; la[6] = x1.z; la[7] = x1.w;
    }
    if (T == 0) {
        const float4* xb = XSs + grp_off<4>(bj);
        const float4 x0 = __ldg(xb), x1 = __ldg(xb + G);
        lb[0] = x0.x; lb[1] = x0.y; lb[2] = x0.z; lb[3] = x0.w; lb[4] = x1.x; lb[5] = x1.y; lb[6] = x1.z; lb[7] = x1.w;
    } else {
        const uint4* q = QKs + grp_off<32>((T == 1) ? bi : bj);
        const uint4* k = QKs + grp_off<32>((T == 1) ? bj : bi) + 16u * G;
#pragma unroll
        for (int h = 0; h < 8; ++h)
            lb[h] = head_dot(__ldg(q + (2 * h) * G), __ldg(q + (2 * h + 1) * G), __ldg(k + (2 * h) * G), __ldg(k + (2 * h + 1) * G));
    }
    // masked 3-token softmax per head (masked_row_fast, with the rare exact branch taken once per row instead of once per head)
    float wa[8], wb[8];
    float lowest = fmaxf(la[0], lb[0]);
#pragma unroll
    for (int h = 0; h < 8; ++h) {
        wa[h] = rcp_approx(1.0f + ex2_approx(lb[h] - la[h]));
        wb[h] = 1.0f - wa[h];
        lowest = fminf(lowest, fmaxf(la[h], lb[h]));
    }
    if (lowest < -20.f) {
#pragma unroll
        for (int h = 0; h < 8; ++h)
            if (fmaxf(la[h], lb[h]) < -20.f) masked_row_exact(la[h], lb[h], wa[h], wb[h]);
    }
#pragma unroll
    for (int h = 0; h < 8; ++h) w[h] = pack_h2(wa[h], wb[h]);
}

// Part 2: ctx = w_a V_a + w_b V_b, written as the 16 chunks of the row into the warpgroup's A tile.
// a_row_x = address of the row in the A tile with (row & 7) ALSO in address bits 4..6: xor with the chunk index gives the
// SWIZZLE_128B slot of the chunk.
template <int T>
__device__ __forceinline__ void attn_ctx(const uint4* __restrict__ binVc, const uint4* __restrict__ cellVrow,
                                         const uint32_t bi, const uint32_t bj, const uint32_t a_row_x, const uint32_t (&w)[8]) {
    constexpr uint32_t G = HSG_GRP;
    const uint4* va = (T == 0) ? binVc + grp_off<16>(bi) : cellVrow;
    constexpr uint32_t sa = (T == 0) ? G : 1u;
    const uint4* vb = binVc + grp_off<16>((T == 2) ? bi : bj);
#pragma unroll
    for (int c = 0; c < 16; ++c) {
        const __half2 wa2 = __low2half2(u2h(w[c >> 1])), wb2 = __high2half2(u2h(w[c >> 1]));
        uint32_t addr;
        asm volatile("xor.b32 %0, %1, %2;" : "=r"(addr) : "r"(a_row_x), "r"((uint32_t)((c & 7) << 4)));
        ctx_chunk(addr + (uint32_t)(c >> 3) * 16384u, wa2, wb2, __ldg(va + c * sa), __ldg(vb + c * G));
    }
}

// ---- the three row-wise epilogues of a token (thread r <-> row r <-> TMEM lane r), packed fp32x2 math -----------------------
// after MMA 1: rstd of y_c (columns [0, 64) of t_y), h = swish(rstd g + b0') -> fp16 A operand over the consumed g columns, plus
// the constant K columns 64 / 65 = 1 that pick up the bias rows of the W1 / C0 images
template <bool TANH2>
__device__ __forceinline__ void epi_h(const uint32_t t_y, const uint32_t t_h, const uint32_t t_lane, const float* cst) {
    float2 rs;
    {
        uint32_t r0[32], r1[32];
        tmem_ld32(t_y + t_lane, r0);
        tmem_ld32(t_y + t_lane + 32, r1);
        tmem_ld_wait();
        float2 vA = make_float2(0.f, 0.f), vB = vA, vC = vA, vD = vA;
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
            float2 a = f2(r0[2 * i], r0[2 * i + 1]), b = f2(r0[2 * i + 2], r0[2 * i + 3]);
            float2 c2 = f2(r1[2 * i], r1[2 * i + 1]), d2 = f2(r1[2 * i + 2], r1[2 * i + 3]);
            vA = __ffma2_rn(a, a, vA); vB = __ffma2_rn(b, b, vB); vC = __ffma2_rn(c2, c2, vC); vD = __ffma2_rn(d2, d2, vD);
        }
        vA = __fadd2_rn(__fadd2_rn(vA, vB), __fadd2_rn(vC, vD));
        const float rstd = rsqrtf((vA.x + vA.y) * (1.0f / 64.0f) + HSG_LN_EPS);
        rs = make_float2(rstd, rstd);
    }
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        uint32_t r[32];
        tmem_ld32(t_h + t_lane + half * 32, r);
        tmem_ld_wait();
        uint32_t hp[16];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const float* bq = cst + CST_B0F + half * 32 + c * 8;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float2 w = __ffma2_rn(f2(r[c * 8 + 2 * q], r[c * 8 + 2 * q + 1]), rs, make_float2(bq[2 * q], bq[2 * q + 1]));
                if constexpr (TANH2) {
                    const __half2 wh = __floats2half2_rn(w.x, w.y);
                    uint32_t th;
                    asm("tanh.approx.f16x2 %0, %1;" : "=r"(th) : "r"(h2u(wh)));
                    hp[c * 4 + q] = h2u(__hfma2(wh, u2h(th), wh));
                } else {
                    float2 th = make_float2(fast_tanh(w.x), fast_tanh(w.y));
                    float2 hh = __ffma2_rn(w, th, w);
                    hp[c * 4 + q] = pack_h2(hh.x, hh.y);
                }
            }
        }
        tmem_st16(t_h + t_lane + half * 16, hp);
    }
    const uint32_t one[8] = {0x3C003C00u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
    tmem_st8(t_h + t_lane + 32, one);
}

// the S' row of a token, requested before the wait on MMA 2 (volatile: the compiler must not sink the loads below the wait)
__device__ __forceinline__ void load_s_row(const Umma2Params& p, const int t, const int cell, const int2 pr, float4 (&sreg)[16]) {
    if (t == 0) {
        const float4* Srow = p.cellSb + (size_t)cell * 16u;
#pragma unroll
        for (int c = 0; c < 16; ++c) sreg[c] = ldg_nc_v4_volatile(Srow + c);
    } else {
        const float4* Scol = p.binSb + grp_off<16>((uint32_t)(t == 1 ? pr.x : pr.y));
#pragma unroll
        for (int c = 0; c < 16; ++c) sreg[c] = ldg_nc_v4_volatile(Scol + c * HSG_GRP);
    }
}

// after MMA 2: z_c (centred by construction) -> rstd, u' = (z_c rstd + S')^2 -> fp16 A operand over the consumed h columns
__device__ __forceinline__ void epi_u(const uint32_t t_y, const uint32_t t_h, const uint32_t t_lane, const float4 (&sreg)[16]) {
    float2 rs;
    {
        float2 vA = make_float2(0.f, 0.f), vB = vA, vC = vA, vD = vA;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            uint32_t r[32];
            tmem_ld32(t_y + t_lane + half * 32, r);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
                float2 a = f2(r[2 * i], r[2 * i + 1]), b = f2(r[2 * i + 2], r[2 * i + 3]);
                float2 c2 = f2(r[2 * i + 4], r[2 * i + 5]), d2 = f2(r[2 * i + 6], r[2 * i + 7]);
                vA = __ffma2_rn(a, a, vA); vB = __ffma2_rn(b, b, vB); vC = __ffma2_rn(c2, c2, vC); vD = __ffma2_rn(d2, d2, vD);
            }
        }
        vA = __fadd2_rn(__fadd2_rn(vA, vB), __fadd2_rn(vC, vD));
        const float rstd = rsqrtf((vA.x + vA.y) * (1.0f / 64.0f) + HSG_LN_EPS);
        rs = make_float2(rstd, rstd);
    }
#pragma unroll
    for (int qt = 0; qt < 4; ++qt) {
        uint32_t r[16], up[8];
        tmem_ld16(t_y + t_lane + qt * 16, r);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const float4 s0 = sreg[qt * 4 + 2 * c], s1 = sreg[qt * 4 + 2 * c + 1];
            const float svv[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float2 u = __ffma2_rn(f2(r[c * 8 + 2 * q], r[c * 8 + 2 * q + 1]), rs, make_float2(svv[2 * q], svv[2 * q + 1]));
                u = __fmul2_rn(u, u);
                up[c * 4 + q] = pack_h2(u.x, u.y);
            }
        }
        tmem_st8(t_h + t_lane + qt * 8, up);
    }
}

// after MMA 3: o_t = cw1 . swish(c) + cb1
__device__ __forceinline__ float epi_o(const uint32_t t_y, const uint32_t t_lane, const float* cst) {
    uint32_t r[32];
    tmem_ld32(t_y + t_lane, r);
    tmem_ld_wait();
    float2 oA = make_float2(0.f, 0.f), oB = oA;
#pragma unroll
    for (int i = 0; i < 16; i += 2) {
        const float* wq = cst + CST_CW1 + 2 * i;
        float2 w0 = f2(r[2 * i], r[2 * i + 1]), w1 = f2(r[2 * i + 2], r[2 * i + 3]);
        float2 h0 = __ffma2_rn(w0, make_float2(fast_tanh(w0.x), fast_tanh(w0.y)), w0);
        float2 h1 = __ffma2_rn(w1, make_float2(fast_tanh(w1.x), fast_tanh(w1.y)), w1);
        oA = __ffma2_rn(h0, make_float2(wq[0], wq[1]), oA);
        oB = __ffma2_rn(h1, make_float2(wq[2], wq[3]), oB);
    }
    return ((oA.x + oA.y) + (oB.x + oB.y)) + cst[CST_CB1];
}

// one 128-pair tile of one cell as a warpgroup sees it
struct TileRow {
    int cb; long long pi; bool valid; int2 pr; float bias;
};
__device__ __forceinline__ TileRow decode_tile(const Umma2Params& p, const long long q, const int wg, const int wtid) {
    TileRow t;
    int pt; bool tile_ok;
    if (p.quad_cells) {
        const int cg = (int)(q / p.tiles_per_cell);
        pt = (int)(q - (long long)cg * p.tiles_per_cell);
        tile_ok = cg * p.nwg + wg < p.n_batch;
        t.cb = tile_ok ? cg * p.nwg + wg : p.n_batch - 1;
    } else {
        const long long n_tiles = (long long)p.n_batch * p.tiles_per_cell;
        tile_ok = q * p.nwg + wg < n_tiles;
        const long long tile = tile_ok ? q * p.nwg + wg : n_tiles - 1;
        t.cb = (int)(tile / p.tiles_per_cell); pt = (int)(tile - (long long)t.cb * p.tiles_per_cell);
    }
    t.pi = (long long)pt * 128 + wtid;
    t.valid = tile_ok && t.pi < p.P;
    t.pr = __ldg(p.pairs + (t.pi < p.P ? t.pi : (p.P - 1)));
    t.bias = t.valid ? __ldg(p.pair_bias + t.pi) : 0.f;
    return t;
}
__device__ __forceinline__ void store_score(const Umma2Params& p, const TileRow& t, const float osum) {
    if (!t.valid) return;
    const float m = osum * (1.0f / 3.0f) + t.bias;
    const float sc = (p.act == HSG_ACT_SIGMOID) ? __frcp_rn(1.0f + __expf(-m))
                     : (p.act == HSG_ACT_SOFTPLUS) ? (m > 20.f ? m : log1pf(__expf(m))) : m;
    if (p.out16) p.out16[(long long)t.cb * p.P + t.pi] = __float2half_rn(sc);
    else p.out[(long long)t.cb * p.P + t.pi] = sc;
}

// weights (PART & 1) and / or ctx (PART & 2) of token T of tile row `tr`
template <int T, int PART>
__device__ __forceinline__ void attn_part(const Umma2Params& p, const TileRow& tr, const uint32_t a_row_x, uint32_t (&w)[8]) {
    if (PART & 1)
        attn_weights<T>(p.QKc + (size_t)tr.cb * (size_t)(p.n_grp * 32 * HSG_GRP), p.XSc + (size_t)tr.cb * (size_t)(p.n_grp * 4 * HSG_GRP),
                        (uint32_t)tr.pr.x, (uint32_t)tr.pr.y, w);
    if (PART & 2) {
        if ((p.exp_flags & 1) && T != 0) attn_ctx<0>(p.binVc, p.cellV, (uint32_t)tr.pr.y, (uint32_t)tr.pr.y, a_row_x, w);   // experiment
        else attn_ctx<T>(p.binVc, p.cellV + (size_t)(p.cell_start + tr.cb) * 16u, (uint32_t)tr.pr.x, (uint32_t)tr.pr.y, a_row_x, w);
    }
}

__device__ __forceinline__ void named_bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void named_bar_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ uint32_t tmem_base_of(const unsigned char* smem, int off) { return *reinterpret_cast<const uint32_t*>(smem + off); }
// operand ready -> MMA group: with issue warps the warpgroup only arrives (operand writes fenced first), otherwise its thread 0 issues
#define REQ_MMA1()                                                                                                        \
    if constexpr (ISS) { fence_async_smem(); tc_fence_before(); named_bar_arrive(5 + wg, 160); }                          \
    else { ISSUE_BEGIN() mma_group<8, 128>(a_base, sbase + SMF_B_FC1, 128, t_y, 0u); ISSUE_END() }
#define REQ_MMA2()                                                                                                        \
    if constexpr (ISS) { tmem_st_wait(); tc_fence_before(); named_bar_arrive(5 + wg, 160); }                              \
    else { ISSUE_BEGIN_TS() mma_group_ts<4, 64>(t_h, sbase + SMF_B_W1, 64, t_y, 1u);                                       \
           umma_f16_ts(t_y, t_h + 32, umma_desc(sbase + SMF_X_W1B), umma_idesc_f16(128, 64), 1u); ISSUE_END() }
#define REQ_MMA3()                                                                                                        \
    if constexpr (ISS) { tmem_st_wait(); tc_fence_before(); named_bar_arrive(5 + wg, 160); }                              \
    else { ISSUE_BEGIN_TS() mma_group_ts<4, 32>(t_h, sbase + SMF_B_C0, 32, t_y, 0u);                                       \
           umma_f16_ts(t_y, t_h + 32, umma_desc(sbase + SMF_X_C0B), umma_idesc_f16(128, 32), 1u); ISSUE_END() }

// TANH2: packed tanh.approx.f16x2 in the feed-forward swish (MUFU work of h: 64 -> 32 per row and token)
// PIPE : software pipeline over the tokens of a warpgroup's own chain.  0: attention, MMA 1, epilogues strictly one after the
//        other (every tensor-core round trip and every load burst is exposed).  1: the softmax weights of the NEXT token (Q / K /
//        cross-term loads, 8 head dots) are computed while MMA 1 of the current token is in flight, its ctx tile (V loads, 16
//        chunk products) is built while MMA 3 is in flight (the A tile is free once MMA 1 has completed) -- across tile
//        boundaries too; the warpgroup carries only the 8 packed weight pairs from one part to the other.
// ISS  : dedicated MMA-issue warps.  0: thread 0 of a warpgroup issues its MMAs after a warpgroup barrier (its warp runs ~300
//        uniform-datapath instructions per token while the three sibling warps wait, and every issue point is a barrier).
//        1: a fifth warpgroup of four issue warps, warp w serving warpgroup w: the 128 epilogue threads only ARRIVE on a named
//        barrier (no wait) once their operand is written and go on; the issue warp completes the barrier, issues the MMA group
//        and commits to the warpgroup's mbarrier.  Register budget by setmaxnreg out of the 640 x 96 registers of the launch: 4 x 128 threads x 112 + 128 x 24.
template <bool TANH2, int PIPE, bool ISS>
__global__ void __launch_bounds__(ISS ? UM_THREADS + 128 : UM_THREADS, 1) score_umma2_kernel(Umma2Params p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, wg = tid >> 7, wtid = tid & 127, warp_in_wg = (tid >> 5) & 3;
    const uint32_t sraw = smem_u32(smem_raw);
    const uint32_t sbase = (sraw + 1023u) & ~1023u;          // SWIZZLE_128B operand tiles need 1024-byte aligned bases
    unsigned char* smem = smem_raw + (sbase - sraw);
    const float* cst = p.cst;
    const uint32_t bar_w = sbase + SM_BAR_OF(0);
    const uint32_t bar_mma = sbase + SM_BAR_OF(0) + 8 + wg * 8;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + SM_TMEM_OF(0));

    if (tid == 0) {
        mbar_init(bar_w, 1);
        for (int i = 0; i < UM_WG; ++i) mbar_init(sbase + SM_BAR_OF(0) + 8 + i * 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (tid == 0) {
        mbar_expect_tx(bar_w, SMF_WIMG_BYTES);
        tma_bulk_g2s(sbase, p.wimg, SMF_WIMG_BYTES, bar_w);
    }
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t t_lane = ((uint32_t)(warp_in_wg * 32)) << 16;
    const uint32_t t_y = tmem_base + (uint32_t)(wg * 128);
    const uint32_t t_h = t_y + 64;
    const uint32_t a_base = sbase + SM_A0_OF(0) + wg * SM_A_BYTES;
    const uint32_t a_row_x = a_base + (uint32_t)(wtid >> 3) * 1024u + (uint32_t)(wtid & 7) * (128u + 16u);
    mbar_wait(bar_w, 0);

    const long long n_tiles = (long long)p.n_batch * p.tiles_per_cell;
    if constexpr (ISS) {
        if (wg == UM_WG) {
            // ---------------------------------------------------------------- issue warps: warp w serves warpgroup w
            asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");
            // warp-uniform values go through a shuffle so that the compiler keeps them (and the MMA descriptors derived from them) in
            // uniform registers: a UTCHMMA then needs no per-instruction R2UR / election sequence
            const int sw = __shfl_sync(0xffffffffu, warp_in_wg, 0);                      // the warpgroup this warp serves
            const uint32_t s_ty = __shfl_sync(0xffffffffu, tmem_base, 0) + (uint32_t)(sw * 128), s_th = s_ty + 64;
            const uint32_t s_a = sbase + SM_A0_OF(0) + sw * SM_A_BYTES, s_bar = sbase + SM_BAR_OF(0) + 8 + sw * 8;
            uint32_t lead_p;
            asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(lead_p));
            const bool lead = lead_p != 0;
            const long long nq = p.quad_cells ? (long long)((p.n_batch + p.nwg - 1) / p.nwg) * p.tiles_per_cell : (n_tiles + p.nwg - 1) / p.nwg;
            long long n_tok = 0;
            for (long long q = blockIdx.x; q < nq; q += gridDim.x) n_tok += 3;            // tokens this CTA scores per warpgroup
            for (long long i = 0; i < n_tok; ++i) {
                named_bar_sync(5 + sw, 160); tc_fence_after();                            // ctx tile complete in shared memory
                if (lead) { mma_group<8, 128>(s_a, sbase + SMF_B_FC1, 128, s_ty, 0u); umma_commit(s_bar); }
                __syncwarp();
                named_bar_sync(5 + sw, 160); tc_fence_after();                            // h complete in TMEM
                if (lead) {
                    mma_group_ts<4, 64>(s_th, sbase + SMF_B_W1, 64, s_ty, 1u);
                    umma_f16_ts(s_ty, s_th + 32, umma_desc(sbase + SMF_X_W1B), umma_idesc_f16(128, 64), 1u);
                    umma_commit(s_bar);
                }
                __syncwarp();
                named_bar_sync(5 + sw, 160); tc_fence_after();                            // u' complete in TMEM
                if (lead) {
                    mma_group_ts<4, 32>(s_th, sbase + SMF_B_C0, 32, s_ty, 0u);
                    umma_f16_ts(s_ty, s_th + 32, umma_desc(sbase + SMF_X_C0B), umma_idesc_f16(128, 32), 1u);
                    umma_commit(s_bar);
                }
                __syncwarp();
            }
            tc_fence_before();
            __syncthreads();
            return;
        }
        asm volatile("setmaxnreg.inc.sync.aligned.u32 112;");
    }
    uint32_t phase = 0;
    // quad_cells: the four warpgroups of a CTA score the SAME 128-pair tile for four consecutive cells, so everything
    // cell-invariant a tile reads (pairs, pair bias, V and S rows of its bins: about half of the operand bytes) comes from L2 once
    // per CTA and the siblings find it in L1.  The loop bound depends on the block only: control flow stays warp-uniform; a
    // warpgroup whose tile / cell runs past the end recomputes the last one with its stores masked.
    const long long n_quads = p.quad_cells ? (long long)((p.n_batch + p.nwg - 1) / p.nwg) * p.tiles_per_cell : (n_tiles + p.nwg - 1) / p.nwg;

    if constexpr (PIPE == 0) {
        for (long long q = blockIdx.x; q < n_quads; q += gridDim.x) {
            const TileRow tr = decode_tile(p, q, wg, wtid);
            const int cell = p.cell_start + tr.cb;
            float osum = 0.f;
#pragma unroll 1
            for (int t = 0; t < 3; ++t) {
                uint32_t w[8];
                if (t == 0) attn_part<0, 3>(p, tr, a_row_x, w);
                else if (t == 1) attn_part<1, 3>(p, tr, a_row_x, w);
                else attn_part<2, 3>(p, tr, a_row_x, w);
                REQ_MMA1()                                      // MMA 1: [y_c | g] = ctx . [fc1c ; W0' fc1c]^T
                mbar_wait(bar_mma, phase); phase ^= 1;
                tc_fence_after();
                epi_h<TANH2>(t_y, t_h, t_lane, cst);
                REQ_MMA2()                                      // MMA 2: z_c = y_c + h . W1c^T + b1c
                float4 sreg[16];
                load_s_row(p, t, cell, tr.pr, sreg);
                mbar_wait(bar_mma, phase); phase ^= 1;
                tc_fence_after();
                epi_u(t_y, t_h, t_lane, sreg);
                REQ_MMA3()                                      // MMA 3: c = u' . C0'^T + cb0 / 2
                mbar_wait(bar_mma, phase); phase ^= 1;
                tc_fence_after();
                osum += epi_o(t_y, t_lane, cst);
                tc_fence_before();      // TMEM reads done before the next token's MMAs overwrite the columns
            }
            store_score(p, tr, osum);
        }
    } else {
        uint32_t w[8];
        TileRow cur = decode_tile(p, blockIdx.x < n_quads ? (long long)blockIdx.x : 0, wg, wtid);
        if ((long long)blockIdx.x < n_quads) {
            attn_part<0, 3>(p, cur, a_row_x, w);
            REQ_MMA1()
        }
        for (long long q = blockIdx.x; q < n_quads; q += gridDim.x) {
            const bool has_next = q + gridDim.x < n_quads;                                  // block-uniform
            const TileRow nxt = decode_tile(p, has_next ? q + gridDim.x : q, wg, wtid);     // pair / bias loads of the next tile, early
            const int cell = p.cell_start + cur.cb;
            float osum = 0.f;
#pragma unroll 1
            for (int t = 0; t < 3; ++t) {
                // MMA 1 of token t is in flight: softmax weights of the next token
                if (PIPE == 1) {
                    if (t == 0) attn_part<1, 1>(p, cur, a_row_x, w);
                    else if (t == 1) attn_part<2, 1>(p, cur, a_row_x, w);
                    else if (has_next) attn_part<0, 1>(p, nxt, a_row_x, w);
                }
                mbar_wait(bar_mma, phase); phase ^= 1;
                tc_fence_after();
                epi_h<TANH2>(t_y, t_h, t_lane, cst);
                REQ_MMA2()
                float4 sreg[16];
                load_s_row(p, t, cell, cur.pr, sreg);
                mbar_wait(bar_mma, phase); phase ^= 1;
                tc_fence_after();
                epi_u(t_y, t_h, t_lane, sreg);
                REQ_MMA3()
                // MMA 3 is in flight and MMA 1 of this token has completed: the A tile is free -> ctx tile of the next token
                if (PIPE == 1) {
                    if (t == 0) attn_part<1, 2>(p, cur, a_row_x, w);
                    else if (t == 1) attn_part<2, 2>(p, cur, a_row_x, w);
                    else if (has_next) attn_part<0, 2>(p, nxt, a_row_x, w);
                } else {
                    if (t == 0) attn_part<1, 3>(p, cur, a_row_x, w);
                    else if (t == 1) attn_part<2, 3>(p, cur, a_row_x, w);
                    else if (has_next) attn_part<0, 3>(p, nxt, a_row_x, w);
                }
                mbar_wait(bar_mma, phase); phase ^= 1;
                tc_fence_after();
                osum += epi_o(t_y, t_lane, cst);
                if (t < 2 || has_next) {
                    REQ_MMA1()                                  // orders the TMEM reads above and the A-tile writes before MMA 1
                }
            }
            store_score(p, cur, osum);
            cur = nxt;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (tid < 32) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------
// Third generation: the ctx operand of MMA 1 lives in TENSOR MEMORY as well.
// ncu on the kernel above: the L1 / shared-memory data pipe is 84 % busy (62 % LSU + 22 % tensor-core operand reads); of the 58.7
// wavefronts an SM spends per triplet, 8 are the 128-bit shared-memory stores of the ctx tile and 6 the tensor core reading it
// back.  Here thread r writes its ctx row with tcgen05.st (lane r, 64 columns = K 128 in fp16 pairs) and MMA 1 runs in TS mode:
// no A tile in shared memory at all -- the CTA asks for 58 KB instead of 190 KB, which also leaves ~190 KB of L1 for the token
// tables.  Tensor memory then holds 3 accumulator blocks of 128 columns (THREE epilogue warpgroups) plus TWO ctx slots of 64
// columns which the warpgroups take in turn (a slot is held from the ctx build to the completion of MMA 1; the softmax weights are
// computed before the slot is taken).  512 threads: 3 x 128 epilogue threads at 160 registers, 4 issue warps at 24 (setmaxnreg).
// ---------------------------------------------------------------------------------------------
#define U3_WG 3
#define U3_SM_BAR SMF_WIMG_BYTES                 // 8 mbarriers
#define U3_SM_TMEM (U3_SM_BAR + 64)
#define U3_SM_SLOT (U3_SM_TMEM + 16)             // int owner[2], int slot_of[U3_WG]
#define U3_SM_TOTAL (U3_SM_SLOT + 32)

// ctx = w_a V_a + w_b V_b of one row -> the 64 TMEM columns of an A slot (column j = K elements 2j, 2j + 1)
template <int T>
__device__ __forceinline__ void attn_ctx_tmem(const uint4* __restrict__ binVc, const uint4* __restrict__ cellVrow,
                                              const uint32_t bi, const uint32_t bj, const uint32_t t_a, const uint32_t (&w)[8]) {
    constexpr uint32_t G = HSG_GRP;
    const uint4* va = (T == 0) ? binVc + grp_off<16>(bi) : cellVrow;
    constexpr uint32_t sa = (T == 0) ? G : 1u;
    const uint4* vb = binVc + grp_off<16>((T == 2) ? bi : bj);
#pragma unroll
    for (int c4 = 0; c4 < 4; ++c4) {
        uint32_t o[16];
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
            const int c = c4 * 4 + cc;
            const __half2 wa2 = __low2half2(u2h(w[c >> 1])), wb2 = __high2half2(u2h(w[c >> 1]));
            const uint4 a = __ldg(va + c * sa), b = __ldg(vb + c * G);
            o[cc * 4 + 0] = h2u(__hfma2(wb2, u2h(b.x), __hmul2(wa2, u2h(a.x))));
            o[cc * 4 + 1] = h2u(__hfma2(wb2, u2h(b.y), __hmul2(wa2, u2h(a.y))));
            o[cc * 4 + 2] = h2u(__hfma2(wb2, u2h(b.z), __hmul2(wa2, u2h(a.z))));
            o[cc * 4 + 3] = h2u(__hfma2(wb2, u2h(b.w), __hmul2(wa2, u2h(a.w))));
        }
        tmem_st16(t_a + (uint32_t)c4 * 16u, o);
    }
}

template <bool TANH2, bool PIPE>
__global__ void __launch_bounds__(512, 1) score_umma3_kernel(Umma2Params p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, wg = tid >> 7, wtid = tid & 127, warp_in_wg = (tid >> 5) & 3;
    const uint32_t sraw = smem_u32(smem_raw);
    const uint32_t sbase = (sraw + 1023u) & ~1023u;
    unsigned char* smem = smem_raw + (sbase - sraw);
    const float* cst = p.cst;
    const uint32_t bar_w = sbase + U3_SM_BAR;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + U3_SM_TMEM);
    volatile int* slot_owner = reinterpret_cast<volatile int*>(smem + U3_SM_SLOT);       // [2]: -1 free, else the owning warpgroup
    volatile int* slot_of = slot_owner + 2;                                             // [U3_WG]: the slot a warpgroup holds

    if (tid == 0) {
        mbar_init(bar_w, 1);
        for (int i = 0; i < U3_WG; ++i) mbar_init(sbase + U3_SM_BAR + 8 + i * 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        slot_owner[0] = -1; slot_owner[1] = -1;
    }
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (tid == 0) {
        mbar_expect_tx(bar_w, SMF_WIMG_BYTES);
        tma_bulk_g2s(sbase, p.wimg, SMF_WIMG_BYTES, bar_w);
    }
    const uint32_t tmem_base = *tmem_slot;
    mbar_wait(bar_w, 0);
    const long long n_tiles = (long long)p.n_batch * p.tiles_per_cell;
    const long long n_quads = (n_tiles + U3_WG - 1) / U3_WG;

    if (wg == U3_WG) {
        // ---------------------------------------------------------------- issue warps: warp w < 3 serves warpgroup w
        asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");
        const int sw = __shfl_sync(0xffffffffu, warp_in_wg, 0);
        if (sw < U3_WG) {
            const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
            const uint32_t s_ty = tb + (uint32_t)(sw * 128), s_th = s_ty + 64, s_bar = sbase + U3_SM_BAR + 8 + sw * 8;
            uint32_t lead_p;
            asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(lead_p));
            const bool lead = lead_p != 0;
            long long n_tok = 0;
            for (long long q = blockIdx.x; q < n_quads; q += gridDim.x) n_tok += 3;
            for (long long i = 0; i < n_tok; ++i) {
                named_bar_sync(5 + sw, 160); tc_fence_after();                            // ctx complete in the A slot
                const uint32_t s_a = tb + 384u + 64u * (uint32_t)__shfl_sync(0xffffffffu, slot_of[sw], 0);
                if (lead) { mma_group_ts<8, 128>(s_a, sbase + SMF_B_FC1, 128, s_ty, 0u); umma_commit(s_bar); }
                __syncwarp();
                named_bar_sync(5 + sw, 160); tc_fence_after();                            // h complete in TMEM
                if (lead) {
                    mma_group_ts<4, 64>(s_th, sbase + SMF_B_W1, 64, s_ty, 1u);
                    umma_f16_ts(s_ty, s_th + 32, umma_desc(sbase + SMF_X_W1B), umma_idesc_f16(128, 64), 1u);
                    umma_commit(s_bar);
                }
                __syncwarp();
                named_bar_sync(5 + sw, 160); tc_fence_after();                            // u' complete in TMEM
                if (lead) {
                    mma_group_ts<4, 32>(s_th, sbase + SMF_B_C0, 32, s_ty, 0u);
                    umma_f16_ts(s_ty, s_th + 32, umma_desc(sbase + SMF_X_C0B), umma_idesc_f16(128, 32), 1u);
                    umma_commit(s_bar);
                }
                __syncwarp();
            }
        }
        tc_fence_before();
        __syncthreads();
        return;
    }
    asm volatile("setmaxnreg.inc.sync.aligned.u32 160;");
    const uint32_t bar_mma = sbase + U3_SM_BAR + 8 + wg * 8;
    const uint32_t t_lane = ((uint32_t)(warp_in_wg * 32)) << 16;
    const uint32_t t_y = tmem_base + (uint32_t)(wg * 128), t_h = t_y + 64;
    uint32_t phase = 0;
    auto decode = [&](long long q) {
        TileRow tr;
        const bool tile_ok = q * U3_WG + wg < n_tiles;
        const long long tile = tile_ok ? q * U3_WG + wg : n_tiles - 1;
        tr.cb = (int)(tile / p.tiles_per_cell);
        const int pt = (int)(tile - (long long)tr.cb * p.tiles_per_cell);
        tr.pi = (long long)pt * 128 + wtid;
        tr.valid = tile_ok && tr.pi < p.P;
        tr.pr = __ldg(p.pairs + (tr.pi < p.P ? tr.pi : (p.P - 1)));
        tr.bias = tr.valid ? __ldg(p.pair_bias + tr.pi) : 0.f;
        return tr;
    };
    // softmax weights of token tk of tile row tr (no ctx slot needed)
    auto weights = [&](const TileRow& tr, int tk, uint32_t (&w)[8]) {
        const uint4* QKs = p.QKc + (size_t)tr.cb * (size_t)(p.n_grp * 32 * HSG_GRP);
        const float4* XSs = p.XSc + (size_t)tr.cb * (size_t)(p.n_grp * 4 * HSG_GRP);
        if (tk == 0) attn_weights<0>(QKs, XSs, (uint32_t)tr.pr.x, (uint32_t)tr.pr.y, w);
        else if (tk == 1) attn_weights<1>(QKs, XSs, (uint32_t)tr.pr.x, (uint32_t)tr.pr.y, w);
        else attn_weights<2>(QKs, XSs, (uint32_t)tr.pr.x, (uint32_t)tr.pr.y, w);
    };
    // take a ctx slot (the elected thread spins on the two owner words; 3 warpgroups hold a slot for about a quarter of their time
    // each), build the ctx rows of token tk in it and request MMA 1
    auto ctx_and_request = [&](const TileRow& tr, int tk, const uint32_t (&w)[8]) {
        if (wtid == 0) {
            int sl = -1;
            while (sl < 0) {
                if (atomicCAS((int*)&slot_owner[0], -1, wg) == -1) sl = 0;
                else if (atomicCAS((int*)&slot_owner[1], -1, wg) == -1) sl = 1;
                else __nanosleep(64);
            }
            slot_of[wg] = sl;
        }
        wg_bar(wg);
        tc_fence_after();                    // the previous owner's MMA 1 (its reads of the slot) completed before it released
        const uint32_t t_a = tmem_base + 384u + 64u * (uint32_t)slot_of[wg] + t_lane;
        const uint4* cellVrow = p.cellV + (size_t)(p.cell_start + tr.cb) * 16u;
        if (tk == 0) attn_ctx_tmem<0>(p.binVc, cellVrow, (uint32_t)tr.pr.x, (uint32_t)tr.pr.y, t_a, w);
        else if (tk == 1) attn_ctx_tmem<1>(p.binVc, cellVrow, (uint32_t)tr.pr.x, (uint32_t)tr.pr.y, t_a, w);
        else attn_ctx_tmem<2>(p.binVc, cellVrow, (uint32_t)tr.pr.x, (uint32_t)tr.pr.y, t_a, w);
    };
    auto request = [&]() { tmem_st_wait(); tc_fence_before(); named_bar_arrive(5 + wg, 160); };
    auto wait_mma = [&]() { mbar_wait(bar_mma, phase); phase ^= 1; tc_fence_after(); };
    auto release_slot = [&]() { if (wtid == 0) { tc_fence_before(); __threadfence_block(); slot_owner[slot_of[wg]] = -1; } };

    if constexpr (!PIPE) {
        for (long long q = blockIdx.x; q < n_quads; q += gridDim.x) {
            const TileRow tr = decode(q);
            const int cell = p.cell_start + tr.cb;
            float osum = 0.f;
#pragma unroll 1
            for (int t = 0; t < 3; ++t) {
                uint32_t w[8];
                weights(tr, t, w);
                ctx_and_request(tr, t, w);
                request();                                                     // MMA 1
                wait_mma();
                release_slot();                                                // MMA 1 has read the slot
                epi_h<TANH2>(t_y, t_h, t_lane, cst);
                request();                                                     // MMA 2
                float4 sreg[16];
                load_s_row(p, t, cell, tr.pr, sreg);
                wait_mma();
                epi_u(t_y, t_h, t_lane, sreg);
                request();                                                     // MMA 3
                wait_mma();
                osum += epi_o(t_y, t_lane, cst);
                tc_fence_before();
            }
            store_score(p, tr, osum);
        }
    } else {
        // software pipeline over the warpgroup's own chain: the softmax weights of the NEXT token are computed while MMA 1 of the
        // current one is in flight, its ctx rows are built (in whichever slot is free) while MMA 3 is in flight; across tiles too
        uint32_t w[8];
        TileRow cur = decode(blockIdx.x);
        weights(cur, 0, w);
        ctx_and_request(cur, 0, w);
        request();
        for (long long q = blockIdx.x; q < n_quads; q += gridDim.x) {
            const bool has_next = q + gridDim.x < n_quads;                     // block-uniform
            const TileRow nxt = decode(has_next ? q + gridDim.x : q);
            const int cell = p.cell_start + cur.cb;
            float osum = 0.f;
#pragma unroll 1
            for (int t = 0; t < 3; ++t) {
                const bool more = t < 2 || has_next;
                if (more) weights(t < 2 ? cur : nxt, t < 2 ? t + 1 : 0, w);    // MMA 1 of token t is in flight
                wait_mma();
                release_slot();
                epi_h<TANH2>(t_y, t_h, t_lane, cst);
                request();                                                     // MMA 2
                float4 sreg[16];
                load_s_row(p, t, cell, cur.pr, sreg);
                wait_mma();
                epi_u(t_y, t_h, t_lane, sreg);
                request();                                                     // MMA 3
                if (more) ctx_and_request(t < 2 ? cur : nxt, t < 2 ? t + 1 : 0, w);     // MMA 3 is in flight
                wait_mma();
                osum += epi_o(t_y, t_lane, cst);
                if (more) request();                                           // MMA 1 of the next token (orders the TMEM reads above)
                else tc_fence_before();
            }
            store_score(p, cur, osum);
            cur = nxt;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (tid < 32) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
// fp32 static tables -> the 16-bit / folded copies of this kernel.  grouped != 0 (bins): V16 [n_grp][16 chunks][32] chunks and
// Sb [n_grp][16 quads][32] quads; grouped == 0 (cells): both row-major.  Sb = (beta - S) / gamma of layer_norm1.
__global__ void convert_static2_kernel(const float* __restrict__ V, const float* __restrict__ S, const float* __restrict__ ln1_b,
                                       const float* __restrict__ ln1_g, long long rows, int d, int grouped,
                                       __half* __restrict__ V16, float* __restrict__ Sb) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < rows * HSG_HD) {
        const long long r = i / HSG_HD; const int col = (int)(i - r * HSG_HD);
        const long long dst = grouped ? (((r >> 5) * 16 + (col >> 3)) * HSG_GRP + (r & 31)) * 8 + (col & 7) : i;
        V16[dst] = __float2half_rn(V[i]);
    }
    if (i < rows * d) {
        const long long r = i / d; const int f = (int)(i - r * d);
        const long long dst = grouped ? (((r >> 5) * 16 + (f >> 2)) * HSG_GRP + (r & 31)) * 4 + (f & 3) : i;
        Sb[dst] = (ln1_b[f] - S[i]) / ln1_g[f];
    }
}

int launch_convert_static2(hsg_ctx* c, const float* V, const float* S, long long rows, int grouped, void* V16, float* Sb, cudaStream_t s) {
    const long long n = rows * HSG_HD;
    convert_static2_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(V, S, c->w[HSG_W_LN1_B], c->w[HSG_W_LN1_G], rows, c->d, grouped,
                                                                        (__half*)V16, Sb);
    c->launches++;
    HSG_CUDA(cudaGetLastError());
    return 0;
}

int launch_score_pairs_umma2(hsg_ctx* c, int cell_start, int n_batch, float* out, void* out16, cudaStream_t s) {
    if (c->P == 0 || n_batch == 0) return 0;
    Umma2Params p;
    p.QKc = (const uint4*)c->QK16; p.XSc = (const float4*)c->XS; p.binVc = (const uint4*)c->binV16; p.cellV = (const uint4*)c->cellV16;
    p.binSb = (const float4*)c->binSb; p.cellSb = (const float4*)c->cellSb; p.wimg = c->wimg;
    memcpy(p.cst, c->h_cst, sizeof(p.cst));
    p.pairs = c->pairs; p.pair_bias = c->pair_bias; p.P = c->P;
    p.n_bins = c->n_bins; p.n_grp = (c->n_bins + HSG_GRP - 1) / HSG_GRP; p.n_cells = c->n_cells; p.cell_start = cell_start; p.n_batch = n_batch;
    p.tiles_per_cell = (int)((c->P + 127) / 128); p.act = c->act; p.out = out; p.out16 = (__half*)out16;
    static const bool env_quad = getenv("HSG_K2_QUAD") && getenv("HSG_K2_QUAD")[0] == '1';      // measured slower: off by default
    static const bool env_tanh2 = getenv("HSG_K2_TANH2") && getenv("HSG_K2_TANH2")[0] == '1';
    // generation 3 (ctx operand in tensor memory, 3 epilogue warpgroups, pipelined tokens) is the default; HSG_K2_GEN=2 selects the
    // shared-memory ctx kernel above (4 warpgroups), kept for A/B measurements
    static const bool env_gen3 = !(getenv("HSG_K2_GEN") && getenv("HSG_K2_GEN")[0] == '2');
    // token pipeline: generation 3 default 1 (split pipeline); generation 2 default 0 (serial: measured fastest there, the pipelined
    // variants spill at its 112-register budget), 2 = whole attention under MMA 3
    static const int env_pipe = getenv("HSG_K2_PIPE") ? (getenv("HSG_K2_PIPE")[0] - '0') % 3 : (env_gen3 ? 1 : 0);
    static const bool env_iss = !(getenv("HSG_K2_ISS") && getenv("HSG_K2_ISS")[0] == '0');
    static const int env_nwg = getenv("HSG_K2_NWG") ? std::max(1, std::min(UM_WG, atoi(getenv("HSG_K2_NWG")))) : UM_WG;
    static const int env_exp = getenv("HSG_K2_EXP") ? atoi(getenv("HSG_K2_EXP")) : 0;
    p.quad_cells = env_quad ? 1 : 0; p.nwg = env_nwg; p.exp_flags = env_exp;
    const long long quads = p.quad_cells ? (long long)((n_batch + env_nwg - 1) / env_nwg) * p.tiles_per_cell
                                         : ((long long)n_batch * p.tiles_per_cell + env_nwg - 1) / env_nwg;
    const int grid = (int)(quads < c->n_sm ? quads : c->n_sm);
    static bool attr_set[64] = {};
    if (!attr_set[c->device & 63]) {
        const int sm_bytes = SM_PAIRS_OF(0) + 1024;
#define K2ATTR(T2, PP, II) HSG_CUDA(cudaFuncSetAttribute(score_umma2_kernel<T2, PP, II>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm_bytes))
        K2ATTR(false, 0, false); K2ATTR(false, 1, false); K2ATTR(false, 2, false); K2ATTR(true, 0, false); K2ATTR(true, 1, false); K2ATTR(true, 2, false);
        K2ATTR(false, 0, true); K2ATTR(false, 1, true); K2ATTR(false, 2, true); K2ATTR(true, 0, true); K2ATTR(true, 1, true); K2ATTR(true, 2, true);
#undef K2ATTR
        attr_set[c->device & 63] = true;
    }
    if (env_gen3) {
        const long long q3 = ((long long)n_batch * p.tiles_per_cell + U3_WG - 1) / U3_WG;
        const int g3 = (int)(q3 < c->n_sm ? q3 : c->n_sm);
        static bool attr3[64] = {};
        if (!attr3[c->device & 63]) {
            HSG_CUDA(cudaFuncSetAttribute(score_umma3_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, U3_SM_TOTAL + 1024));
            HSG_CUDA(cudaFuncSetAttribute(score_umma3_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, U3_SM_TOTAL + 1024));
            HSG_CUDA(cudaFuncSetAttribute(score_umma3_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, U3_SM_TOTAL + 1024));
            HSG_CUDA(cudaFuncSetAttribute(score_umma3_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, U3_SM_TOTAL + 1024));
            attr3[c->device & 63] = true;
        }
        p.quad_cells = 0; p.nwg = U3_WG;
        const size_t sm3 = U3_SM_TOTAL + 1024;
        if (env_pipe) { if (env_tanh2) score_umma3_kernel<true, true><<<g3, 512, sm3, s>>>(p); else score_umma3_kernel<false, true><<<g3, 512, sm3, s>>>(p); }
        else { if (env_tanh2) score_umma3_kernel<true, false><<<g3, 512, sm3, s>>>(p); else score_umma3_kernel<false, false><<<g3, 512, sm3, s>>>(p); }
        c->launches++;
        HSG_CUDA(cudaGetLastError());
        return 0;
    }
    const size_t smem = SM_PAIRS_OF(0) + 1024;
    if (env_iss) p.nwg = UM_WG;
#define K2V(T2, PP, II) score_umma2_kernel<T2, PP, II><<<grid, (II) ? UM_THREADS + 128 : p.nwg * 128, smem, s>>>(p)
    const int variant = (env_iss ? 6 : 0) + (env_tanh2 ? 3 : 0) + env_pipe;
    switch (variant) {
        case 0: K2V(false, 0, false); break;  case 1: K2V(false, 1, false); break;  case 2: K2V(false, 2, false); break;
        case 3: K2V(true, 0, false); break;   case 4: K2V(true, 1, false); break;   case 5: K2V(true, 2, false); break;
        case 6: K2V(false, 0, true); break;   case 7: K2V(false, 1, true); break;   case 8: K2V(false, 2, true); break;
        case 9: K2V(true, 0, true); break;    case 10: K2V(true, 1, true); break;   default: K2V(true, 2, true); break;
    }
#undef K2V
    c->launches++;
    HSG_CUDA(cudaGetLastError());
    return 0;
}
