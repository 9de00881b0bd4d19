// score_rows.cuh -- row-per-thread building blocks of the fast tensor-core scorer (score_umma2.cu: generations 2 and 3;
// score_umma4.cu: generation 4): parameter block, grouped-table addressing, the attention of one row, the three row epilogues.
#pragma once
#include "umma_common.cuh"

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
// HSG_EPI_PIPE bits (B200, C2, K2 ms per 225 M triplets; plain code 48.67):
//   1 (default)  h epilogue: the next tcgen05.ld is in flight while the registers of the previous one are consumed (a tcgen05.ld
//                x32 of a warp occupies the 64 B/clk tensor-memory read port for 64 cycles)                              48.24
//   2            u epilogue keeps z in registers instead of reading it twice (z + S' = 128 live registers: spills)      50.7
//   4            u epilogue two-pass with pipelined loads (with bit 1)                                                   48.6
//   8            L1 prefetch of the S' row before the h epilogue (with bit 1)                                            49.1
//   (-DHSG_K2_TANH2: packed tanh.approx.f16x2 in the h epilogue, with bit 1: 48.41)
#ifndef HSG_EPI_PIPE
#define HSG_EPI_PIPE 1
#endif
template <bool TANH2>
__device__ __forceinline__ void epi_h_half(const uint32_t (&r)[32], const float2 rs, const int half, const uint32_t t_h, const uint32_t t_lane,
                                           const float* cst) {
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

template <bool TANH2>
__device__ __forceinline__ void epi_h(const uint32_t t_y, const uint32_t t_h, const uint32_t t_lane, const float* cst) {
#if HSG_EPI_PIPE & 1
    {
        uint32_t r0[32], r1[32], g0[32];
        tmem_ld32(t_y + t_lane, r0);
        tmem_ld32(t_y + t_lane + 32, r1);
        tmem_ld_wait();
        tmem_ld32(t_h + t_lane, g0);                       // first half of g: in flight during the variance
        float2 vA = make_float2(0.f, 0.f), vB = vA, vC = vA, vD = vA;
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
            float2 a = f2(r0[2 * i], r0[2 * i + 1]), b = f2(r0[2 * i + 2], r0[2 * i + 3]);
            float2 c2 = f2(r1[2 * i], r1[2 * i + 1]), d2 = f2(r1[2 * i + 2], r1[2 * i + 3]);
            vA = __ffma2_rn(a, a, vA); vB = __ffma2_rn(b, b, vB); vC = __ffma2_rn(c2, c2, vC); vD = __ffma2_rn(d2, d2, vD);
        }
        vA = __fadd2_rn(__fadd2_rn(vA, vB), __fadd2_rn(vC, vD));
        const float rstd = rsqrtf((vA.x + vA.y) * (1.0f / 64.0f) + HSG_LN_EPS);
        const float2 rs = make_float2(rstd, rstd);
        tmem_ld_wait();
        tmem_ld32(t_h + t_lane + 32, r0);                  // second half of g: in flight during the first half of h
        epi_h_half<TANH2>(g0, rs, 0, t_h, t_lane, cst);    // (h columns [0, 16) of the slot: not the columns being read)
        tmem_ld_wait();
        epi_h_half<TANH2>(r0, rs, 1, t_h, t_lane, cst);
        const uint32_t one[8] = {0x3C003C00u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
        tmem_st8(t_h + t_lane + 32, one);
        return;
    }
#endif
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

// HSG_EPI_PIPE & 8: the S' row is pulled into L1 before the h epilogue, so that the loads issued behind MMA 2 are L1 hits
__device__ __forceinline__ void prefetch_s_row(const Umma2Params& p, const int t, const int cell, const int2 pr) {
    if (t == 0) {
        const float4* Srow = p.cellSb + (size_t)cell * 16u;
#pragma unroll
        for (int c = 0; c < 16; c += 8) asm volatile("prefetch.global.L1 [%0];" ::"l"(Srow + c));
    } else {
        const float4* Scol = p.binSb + grp_off<16>((uint32_t)(t == 1 ? pr.x : pr.y));
#pragma unroll
        for (int c = 0; c < 16; ++c) asm volatile("prefetch.global.L1 [%0];" ::"l"(Scol + c * HSG_GRP));
    }
}

// after MMA 2: z_c (centred by construction) -> rstd, u' = (z_c rstd + S')^2 -> fp16 A operand over the consumed h columns
__device__ __forceinline__ void epi_u(const uint32_t t_y, const uint32_t t_h, const uint32_t t_lane, const float4 (&sreg)[16]) {
#if HSG_EPI_PIPE & 2
    {
        uint32_t za[32], zb[32];
        tmem_ld32(t_y + t_lane, za);
        tmem_ld_wait();
        tmem_ld32(t_y + t_lane + 32, zb);                  // second half of z: in flight during the first half of the variance
        float2 vA = make_float2(0.f, 0.f), vB = vA, vC = vA, vD = vA;
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
            float2 a = f2(za[2 * i], za[2 * i + 1]), b = f2(za[2 * i + 2], za[2 * i + 3]);
            float2 c2 = f2(za[2 * i + 4], za[2 * i + 5]), d2 = f2(za[2 * i + 6], za[2 * i + 7]);
            vA = __ffma2_rn(a, a, vA); vB = __ffma2_rn(b, b, vB); vC = __ffma2_rn(c2, c2, vC); vD = __ffma2_rn(d2, d2, vD);
        }
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
            float2 a = f2(zb[2 * i], zb[2 * i + 1]), b = f2(zb[2 * i + 2], zb[2 * i + 3]);
            float2 c2 = f2(zb[2 * i + 4], zb[2 * i + 5]), d2 = f2(zb[2 * i + 6], zb[2 * i + 7]);
            vA = __ffma2_rn(a, a, vA); vB = __ffma2_rn(b, b, vB); vC = __ffma2_rn(c2, c2, vC); vD = __ffma2_rn(d2, d2, vD);
        }
        vA = __fadd2_rn(__fadd2_rn(vA, vB), __fadd2_rn(vC, vD));
        const float rstd = rsqrtf((vA.x + vA.y) * (1.0f / 64.0f) + HSG_LN_EPS);
        const float2 rs = make_float2(rstd, rstd);
#pragma unroll
        for (int qt = 0; qt < 4; ++qt) {                  // z stays in registers: no second read of the accumulator
            uint32_t up[8];
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const float4 s0 = sreg[qt * 4 + 2 * c], s1 = sreg[qt * 4 + 2 * c + 1];
                const float svv[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int col = (qt & 1) * 16 + c * 8 + 2 * q;
                    const uint32_t z0 = qt < 2 ? za[col] : zb[col], z1 = qt < 2 ? za[col + 1] : zb[col + 1];
                    float2 u = __ffma2_rn(f2(z0, z1), rs, make_float2(svv[2 * q], svv[2 * q + 1]));
                    u = __fmul2_rn(u, u);
                    up[c * 4 + q] = pack_h2(u.x, u.y);
                }
            }
            tmem_st8(t_h + t_lane + qt * 8, up);
        }
        return;
    }
#endif
#if HSG_EPI_PIPE & 4
    {
        // two passes over z as in the plain version (registers), every tcgen05.ld issued before the previous block is consumed
        uint32_t za[32], zb[32];
        tmem_ld32(t_y + t_lane, za);
        tmem_ld_wait();
        tmem_ld32(t_y + t_lane + 32, zb);
        float2 vA = make_float2(0.f, 0.f), vB = vA, vC = vA, vD = vA;
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
            float2 a = f2(za[2 * i], za[2 * i + 1]), b = f2(za[2 * i + 2], za[2 * i + 3]);
            float2 c2 = f2(za[2 * i + 4], za[2 * i + 5]), d2 = f2(za[2 * i + 6], za[2 * i + 7]);
            vA = __ffma2_rn(a, a, vA); vB = __ffma2_rn(b, b, vB); vC = __ffma2_rn(c2, c2, vC); vD = __ffma2_rn(d2, d2, vD);
        }
        tmem_ld_wait();
        uint32_t q0[16], q1[16];
        tmem_ld16(t_y + t_lane, q0);                       // first 16 columns again: in flight during the second half of the variance
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
            float2 a = f2(zb[2 * i], zb[2 * i + 1]), b = f2(zb[2 * i + 2], zb[2 * i + 3]);
            float2 c2 = f2(zb[2 * i + 4], zb[2 * i + 5]), d2 = f2(zb[2 * i + 6], zb[2 * i + 7]);
            vA = __ffma2_rn(a, a, vA); vB = __ffma2_rn(b, b, vB); vC = __ffma2_rn(c2, c2, vC); vD = __ffma2_rn(d2, d2, vD);
        }
        vA = __fadd2_rn(__fadd2_rn(vA, vB), __fadd2_rn(vC, vD));
        const float rstd = rsqrtf((vA.x + vA.y) * (1.0f / 64.0f) + HSG_LN_EPS);
        const float2 rs = make_float2(rstd, rstd);
        auto u_block = [&](const uint32_t (&r)[16], const int qt) {
            uint32_t up[8];
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
        };
        tmem_ld_wait();
        tmem_ld16(t_y + t_lane + 16, q1);
        u_block(q0, 0);
        tmem_ld_wait();
        tmem_ld16(t_y + t_lane + 32, q0);
        u_block(q1, 1);
        tmem_ld_wait();
        tmem_ld16(t_y + t_lane + 48, q1);
        u_block(q0, 2);
        tmem_ld_wait();
        u_block(q1, 3);
        return;
    }
#endif
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
