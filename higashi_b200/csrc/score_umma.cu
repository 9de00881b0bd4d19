// score_umma.cu -- K2 (tensor-core mode): the fused Hyper-SAGNN scorer on tcgen05 / TMEM (sm_100a).
//
// Same math as score_f32.cu (Modules.py:726-783, 1029-1095, 865-889), restructured for
// Blackwell:
//   * one CTA per SM (persistent), 4 warpgroups of 128 threads; a warpgroup owns a tile of 128 triplets,
//     thread r <-> triplet r <-> TMEM lane r, and walks the 3 tokens of its triplets one after the other;
//   * per token row-tile the dense contractions run on the 5th-gen tensor cores in THREE MMA round trips
//       [y_c | g] = ctx . [fc1c ; W0' fc1c]^T   (M128 N128 K128; centred fc1, LayerNorm . W0 folded in: see the body)
//       z_c       = y_c + h . W1c^T             (M128 N64 K64, accumulated onto y_c in TMEM; A operand h in TMEM)
//       c         = u . C0^T                    (M128 N32 K64; A operand u = (LN1(z) - S)^2 in TMEM)
//     issued by one elected thread of the warpgroup (tcgen05.mma, fp16 operands, fp32 accumulation in TMEM), completion
//     tracked with tcgen05.commit -> mbarrier, accumulators read back with tcgen05.ld;
//   * the ctx operand is written by the warpgroup itself straight into the canonical K-major SWIZZLE_128B shared-memory
//     layout; h and u go back to TMEM with tcgen05.st and are consumed as TMEM A operands (no shared-memory round trip);
//     the B operands (weights, pre-swizzled on the host) are staged once per CTA with one TMA bulk copy (cp.async.bulk);
//   * the 3-token masked attention stays in registers (packed half2 math, software-pipelined table loads); the only HBM
//     traffic of the kernel is the 4-byte score per triplet, token tables are read through L1/L2;
//   * template flag HI: high-accuracy variant (fp32 tables and attention, hi + lo weight / operand splits) for softplus heads.
// Operands are fp16, not bf16: with bf16 operands the measured score error is 1.5e-3..2e-3 (above the 1e-3
// parity bar), with fp16 it is <= 5e-4 at identical tensor throughput (scripts/bf16_emulation.py, scripts/tail_error.py).
#include <algorithm>
#include <cmath>

#include "umma_common.cuh"

// 1: build the ctx tile of token t + 1 between the issue and the wait of token t's classifier MMA.  Measured on the B200
// (C2, 12 steps): 54.3-54.7 ms against 53.76 ms without -- the four warpgroups already fill each other's MMA waits -- so it
// stays off.  Also measured and dropped: MMA 1 as two N = 64 halves with separate commits (variance pass under the g half):
// 56.2 ms against 53.6 ms.
#ifndef K2_OVERLAP_ATT
#define K2_OVERLAP_ATT 0
#endif

template <bool HI>
__global__ void __launch_bounds__(UM_THREADS, 1) score_umma_kernel(UmmaParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, wg = tid >> 7, wtid = tid & 127, warp_in_wg = (tid >> 5) & 3, lane = tid & 31;
    // SWIZZLE_128B operand tiles need 1024-byte aligned bases: round the dynamic segment up by hand
    const uint32_t sraw = smem_u32(smem_raw);
    const uint32_t sbase = (sraw + 1023u) & ~1023u;
    unsigned char* smem = smem_raw + (sbase - sraw);
    const float* cst = p.cst;
    const uint32_t bar_w = sbase + SM_BAR_OF(HI);                 // weights landed
    const uint32_t bar_mma = sbase + SM_BAR_OF(HI) + 8 + wg * 8;  // this warpgroup's MMA completion
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + SM_TMEM_OF(HI));

    if (tid == 0) {
        mbar_init(bar_w, 1);
        for (int i = 0; i < UM_WG; ++i) mbar_init(sbase + SM_BAR_OF(HI) + 8 + i * 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 32) {     // warp 0 allocates all 512 TMEM columns (one CTA per SM)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (tid == 0) {
        mbar_expect_tx(bar_w, HI ? SMH_WIMG_BYTES : SMF_WIMG_BYTES);
        tma_bulk_g2s(sbase, p.wimg, HI ? SMH_WIMG_BYTES : SMF_WIMG_BYTES, bar_w);
    }
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t t_lane = ((uint32_t)(warp_in_wg * 32)) << 16;
    const uint32_t t_y = tmem_base + (uint32_t)(wg * 128);          // columns [0,64) of this warpgroup: y, then z
    const uint32_t t_h = t_y + 64;                                  // columns [64,128): h, then c
    const uint32_t a_base = sbase + SM_A0_OF(HI) + wg * SM_A_BYTES;
    mbar_wait(bar_w, 0);

    uint32_t phase = 0;
    const long long n_tiles = (long long)p.n_batch * p.tiles_per_cell;
    // The loop bound depends on the block only (not on the warpgroup), so the whole body is warp-uniform control flow for the
    // compiler; a warpgroup whose tile index runs past the end recomputes the last tile with its stores masked.
    for (long long tile0 = (long long)blockIdx.x * UM_WG; tile0 < n_tiles; tile0 += (long long)gridDim.x * UM_WG) {
        const bool tile_ok = tile0 + wg < n_tiles;
        const long long tile = tile_ok ? tile0 + wg : n_tiles - 1;
        const int cb = (int)(tile / p.tiles_per_cell), pt = (int)(tile - (long long)cb * p.tiles_per_cell);
        const long long pi = (long long)pt * 128 + wtid;
        const bool valid = tile_ok && pi < p.P;
        const int2 pr = p.pairs[pi < p.P ? pi : (p.P - 1)];
        const float bias = valid ? p.pair_bias[pi] : 0.f;
        int2* s_pr_wg = reinterpret_cast<int2*>(smem + SM_PAIRS_OF(HI)) + wg * 128;    // rows of this warpgroup's tile
        __syncwarp();
        s_pr_wg[wtid] = pr;                                                      // a warp only reads its own 32 rows
        __syncwarp();
        const int cell = p.cell_start + cb;
        float osum = 0.f;

#pragma unroll 1
        for (int t = 0; t < 3; ++t) {
            // ------------------------------------------------------------------ attention -> ctx (A tile, K = 128)
            // Row t attends to the two other tokens: (a, b) = (bi, bj) | (cell, bj) | (cell, bi).
            // Lane mapping for this phase: lane = (row-in-group rho = lane>>3, head h = lane&7); a warp walks its
            // own 32 rows four at a time, so the 8 lanes of a row read one contiguous 256-byte table row
            // (coalesced) instead of every lane streaming through a private row.
            if constexpr (HI) {
                // high-accuracy variant: fp32 Q/K/V tables, fp32 dot products and ctx, one fp16 rounding of ctx
                const int h = lane & 7, rsub = lane >> 3;
                const float4* QK4 = reinterpret_cast<const float4*>(p.QK32);
                const float4* BV4 = reinterpret_cast<const float4*>(p.binV32);
                // rows are stored [quad e][head h][4 floats] (HSG_HI_POS): the 8 lanes of a row read 128 contiguous bytes per load
                const float4* CV4 = reinterpret_cast<const float4*>(p.cellV32) + (uint32_t)cell * 32u + h;
                float4 cv[4];
                if (t != 0) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) cv[e] = __ldg(CV4 + e * 8);
                }
                const uint32_t tokbase = (uint32_t)cb * (uint32_t)p.n_bins;
#pragma unroll 2
                for (int it = 0; it < 8; ++it) {
                    const int rw = it * 4 + rsub;
                    const uint32_t bi_r = (uint32_t)__shfl_sync(0xffffffffu, pr.x, rw), bj_r = (uint32_t)__shfl_sync(0xffffffffu, pr.y, rw);
                    const uint32_t ti = tokbase + bi_r, tj = tokbase + bj_r;
                    float la, lb;
                    const float4 *va, *vb;
                    if (t == 0) {
                        la = __ldg(p.XS + ti * 16u + HSG_XS_POS(h)); lb = __ldg(p.XS + tj * 16u + HSG_XS_POS(h));
                        va = BV4 + bi_r * 32u + h; vb = BV4 + bj_r * 32u + h;
                    } else {
                        const uint32_t tq = (t == 1) ? ti : tj, tk = (t == 1) ? tj : ti;
                        la = __ldg(p.XS + tq * 16u + 8 + HSG_XS_POS(h));
                        const float4* q = QK4 + tq * 64u + h; const float4* k = QK4 + tk * 64u + 32 + h;
                        float2 acc = make_float2(0.f, 0.f), acc2 = acc;
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            float4 qv = __ldg(q + e * 8), kv = __ldg(k + e * 8);
                            acc = __ffma2_rn(make_float2(qv.x, qv.y), make_float2(kv.x, kv.y), acc);
                            acc2 = __ffma2_rn(make_float2(qv.z, qv.w), make_float2(kv.z, kv.w), acc2);
                        }
                        lb = ((acc.x + acc.y) + (acc2.x + acc2.y)) * HSG_L2E_QUARTER;
                        va = nullptr; vb = BV4 + ((t == 1) ? bj_r : bi_r) * 32u + h;
                    }
                    float wa, wb;
                    masked_row_fast(la, lb, wa, wb);
                    const float2 wa2 = make_float2(wa, wa), wb2 = make_float2(wb, wb);
                    const int arow = warp_in_wg * 32 + rw;
                    uint32_t pk[8];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        float4 av = (t == 0) ? __ldg(va + e * 8) : cv[e], bv = __ldg(vb + e * 8);
                        float2 c0 = __ffma2_rn(wb2, make_float2(bv.x, bv.y), __fmul2_rn(wa2, make_float2(av.x, av.y)));
                        float2 c1 = __ffma2_rn(wb2, make_float2(bv.z, bv.w), __fmul2_rn(wa2, make_float2(av.z, av.w)));
                        pk[e * 2] = pack_h2(c0.x, c0.y); pk[e * 2 + 1] = pack_h2(c1.x, c1.y);
                    }
                    // head h = chunk positions h and 8 + h (HSG_CHUNK_POS): the 8 lanes of a row hit 8 distinct swizzle slots
                    sts128(a_chunk_addr(a_base, arow, h), pk[0], pk[1], pk[2], pk[3]);
                    sts128(a_chunk_addr(a_base, arow, 8 + h), pk[4], pk[5], pk[6], pk[7]);
                }
            } else if (!K2_OVERLAP_ATT || t == 0) {
                // with K2_OVERLAP_ATT only the first token's attention runs here; the ctx tile of token t + 1 is then built while
                // MMA 3 of token t is in flight (the A tile is free as soon as MMA 1 of token t has completed), see below
                const uint32_t tokbase = (uint32_t)cb * (uint32_t)p.n_bins;
                if (t == 0) attention_fast<0, 8>(p.QK16, p.XS, p.binV16, p.cellV16, s_pr_wg, lane, warp_in_wg * 32, a_base, tokbase, cell);
                else if (t == 1) attention_fast<1, 8>(p.QK16, p.XS, p.binV16, p.cellV16, s_pr_wg, lane, warp_in_wg * 32, a_base, tokbase, cell);
                else attention_fast<2, 8>(p.QK16, p.XS, p.binV16, p.cellV16, s_pr_wg, lane, warp_in_wg * 32, a_base, tokbase, cell);
            }
            // ------------------------------------------------------------------ MMA 1: [y_c | g] = ctx . [fc1c ; W0' fc1c]^T
            // ONE N=128 GEMM gives y_c = ctx . fc1c^T (columns 0..63: y with its row mean already removed, because the mean over
            // the outputs was subtracted from fc1 on the host) and g = ctx . (W0' fc1c)^T (columns 64..127).
            // LayerNorm(y) . W0'^T = rstd(y_c) * g, so the LN -> fp16 -> shared memory -> MMA round trip of the feed-forward
            // input disappears: the epilogue only needs the row variance of y_c.  High-accuracy variant: the weight image is
            // split into fp16 hi + lo parts, the lo part is one more accumulating MMA group.
            ISSUE_BEGIN()
            mma_group<8, 128>(a_base, sbase + SMF_B_FC1, 128, t_y, 0u);
            if constexpr (HI) mma_group<8, 128>(a_base, sbase + SMH_B_FC1_LO, 128, t_y, 1u);
            ISSUE_END()
            mbar_wait(bar_mma, phase); phase ^= 1;
            tc_fence_after();

            // The row-wise epilogues use the packed fp32x2 instructions of sm_100 (FADD2/FMUL2/FFMA2): two columns per issue slot.
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
            // ------------------------------------------------------------------ h = swish(rstd * g + b0')  -> TMEM A operand
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
                        float2 th = make_float2(fast_tanh(w.x), fast_tanh(w.y));
                        float2 hh = __ffma2_rn(w, th, w);
                        hp[c * 4 + q] = pack_h2(hh.x, hh.y);
                    }
                }
                // h goes back to TMEM as the fp16 A operand of the next MMA (two K elements per column): it overwrites the g
                // columns this thread has just consumed -- no shared-memory round trip, no proxy fence
                tmem_st16(t_h + t_lane + half * 16, hp);
            }
            if constexpr (!HI) {
                // K columns 64 / 65 of the A operand = 1 (the bias columns of the W1 / C0 images); g is consumed, MMA 1 of the
                // next token overwrites them again
                const uint32_t one[8] = {0x3C003C00u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
                tmem_st8(t_h + t_lane + 32, one);
            }
            // ------------------------------------------------------------------ MMA 2: z_c = y_c + h . W1c^T (residual for free)
            ISSUE_BEGIN_TS()
            mma_group_ts<4, 64>(t_h, sbase + SMF_B_W1, 64, t_y, 1u);
            if constexpr (!HI) umma_f16_ts(t_y, t_h + 32, umma_desc(sbase + SMF_X_W1B), umma_idesc_f16(128, 64), 1u);     // + b1c
            ISSUE_END()
            // S rows of this token (feature blocks of 4): requested BEFORE waiting for the MMA so that their L2 latency hides
            // behind the tensor-core round trip (volatile loads: the compiler must not sink them below the wait)
            float4 sreg[16];
            {
                const float4* Scol = reinterpret_cast<const float4*>((t == 0) ? p.cellSb : p.binSb) + ((t == 0) ? cell : (t == 1 ? pr.x : pr.y));
                const long long sstr = (t == 0) ? p.n_cells : p.n_bins;
#pragma unroll
                for (int c = 0; c < 16; ++c) sreg[c] = ldg_nc_v4_volatile(Scol + c * sstr);
            }
            mbar_wait(bar_mma, phase); phase ^= 1;
            tc_fence_after();

            // ------------------------------------------------------------------ u = (LN1(z) - S)^2  -> TMEM A operand
            // z arrives centred (y_c + h . W1c^T + b1c has zero row mean by construction).  Two passes over TMEM keep the register
            // footprint at 16 columns + the 64 prefetched S values: pass 1 = row variance, pass 2 = u.
            // Fast variant: b1c was added by the tensor core (bias K columns) and the LayerNorm gain is folded away
            // (S table = (beta - S) / gamma, gamma^2 in the columns of C0): u' = (z rstd + S')^2 -- FFMA2 + FMUL2 + pack per pair.
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
                        if constexpr (HI) {
                            const float* bq = cst + CST_B1C + half * 32 + 2 * i;
                            a = __fadd2_rn(a, make_float2(bq[0], bq[1])); b = __fadd2_rn(b, make_float2(bq[2], bq[3]));
                            c2 = __fadd2_rn(c2, make_float2(bq[4], bq[5])); d2 = __fadd2_rn(d2, make_float2(bq[6], bq[7]));
                        }
                        vA = __ffma2_rn(a, a, vA); vB = __ffma2_rn(b, b, vB); vC = __ffma2_rn(c2, c2, vC); vD = __ffma2_rn(d2, d2, vD);
                    }
                }
                vA = __fadd2_rn(__fadd2_rn(vA, vB), __fadd2_rn(vC, vD));
                const float rstd = rsqrtf((vA.x + vA.y) * (1.0f / 64.0f) + HSG_LN_EPS);
                rs = make_float2(rstd, rstd);
            }
#pragma unroll
            for (int qt = 0; qt < 4; ++qt) {                 // 16 columns at a time next to the 64 S values
                uint32_t r[16], up[8], ul[8];
                tmem_ld16(t_y + t_lane + qt * 16, r);
                tmem_ld_wait();
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const float* bq = cst + CST_B1C + qt * 16 + c * 8;
                    const float* gq = cst + CST_G1 + qt * 16 + c * 8;
                    const float4 s0 = sreg[qt * 4 + 2 * c], s1 = sreg[qt * 4 + 2 * c + 1];
                    const float svv[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        float2 u;
                        if constexpr (HI) {
                            float2 x = __fadd2_rn(f2(r[c * 8 + 2 * q], r[c * 8 + 2 * q + 1]), make_float2(bq[2 * q], bq[2 * q + 1]));
                            float2 tz = __fmul2_rn(x, rs);
                            u = __ffma2_rn(tz, make_float2(gq[2 * q], gq[2 * q + 1]), make_float2(svv[2 * q], svv[2 * q + 1]));
                        } else {
                            u = __ffma2_rn(f2(r[c * 8 + 2 * q], r[c * 8 + 2 * q + 1]), rs, make_float2(svv[2 * q], svv[2 * q + 1]));
                        }
                        u = __fmul2_rn(u, u);
                        up[c * 4 + q] = pack_h2(u.x, u.y);
                        if constexpr (HI) {             // u = hi + lo, both fp16: the classifier GEMM sees ~22 bits of u
                            float2 hf = __half22float2(u2h(up[c * 4 + q]));
                            float2 lo = __fadd2_rn(u, make_float2(-hf.x, -hf.y));
                            ul[c * 4 + q] = pack_h2(lo.x, lo.y);
                        }
                    }
                }
                tmem_st8(t_h + t_lane + qt * 8, up);         // u as the fp16 A operand of the classifier GEMM (h is consumed)
                if constexpr (HI) tmem_st8(t_h + t_lane + 32 + qt * 8, ul);
            }
            // ------------------------------------------------------------------ MMA 3: c = u . C0^T  (first 32 columns of the consumed z)
            ISSUE_BEGIN_TS()
            mma_group_ts<4, 32>(t_h, sbase + SMF_B_C0, 32, t_y, 0u);                                 // u_hi . C0_hi
            if constexpr (HI) {
                mma_group_ts<4, 32>(t_h + 32, sbase + SMF_B_C0, 32, t_y, 1u);                        // u_lo . C0_hi
                mma_group_ts<4, 32>(t_h, sbase + SMH_B_C0_LO, 32, t_y, 1u);                          // u_hi . C0_lo
            } else {
                umma_f16_ts(t_y, t_h + 32, umma_desc(sbase + SMF_X_C0B), umma_idesc_f16(128, 32), 1u);  // + cb0 / 2
            }
            ISSUE_END()
            if constexpr (!HI && K2_OVERLAP_ATT) {        // next token's attention overlaps the classifier MMA round trip
                const uint32_t tokbase = (uint32_t)cb * (uint32_t)p.n_bins;
                if (t == 0) attention_fast<1, 8>(p.QK16, p.XS, p.binV16, p.cellV16, s_pr_wg, lane, warp_in_wg * 32, a_base, tokbase, cell);
                else if (t == 1) attention_fast<2, 8>(p.QK16, p.XS, p.binV16, p.cellV16, s_pr_wg, lane, warp_in_wg * 32, a_base, tokbase, cell);
            }
            mbar_wait(bar_mma, phase); phase ^= 1;
            tc_fence_after();

            // ------------------------------------------------------------------ o_t = c1 . swish(c + cb0) + cb1
            {
                uint32_t r[32];
                tmem_ld32(t_y + t_lane, r);
                tmem_ld_wait();
                float2 oA = make_float2(0.f, 0.f), oB = oA;
#pragma unroll
                for (int i = 0; i < 16; i += 2) {
                    const float* bq = cst + CST_CB0 + 2 * i; const float* wq = cst + CST_CW1 + 2 * i;
                    float2 w0 = f2(r[2 * i], r[2 * i + 1]), w1 = f2(r[2 * i + 2], r[2 * i + 3]);
                    if constexpr (HI) { w0 = __fadd2_rn(w0, make_float2(bq[0], bq[1])); w1 = __fadd2_rn(w1, make_float2(bq[2], bq[3])); }
                    float2 h0 = __ffma2_rn(w0, make_float2(fast_tanh(w0.x), fast_tanh(w0.y)), w0);
                    float2 h1 = __ffma2_rn(w1, make_float2(fast_tanh(w1.x), fast_tanh(w1.y)), w1);
                    oA = __ffma2_rn(h0, make_float2(wq[0], wq[1]), oA);
                    oB = __ffma2_rn(h1, make_float2(wq[2], wq[3]), oB);
                }
                osum += ((oA.x + oA.y) + (oB.x + oB.y)) + cst[CST_CB1];
            }
            tc_fence_before();      // TMEM reads done before the next token's MMAs overwrite the columns
        }
        if (valid) {
            float m = osum * (1.0f / 3.0f) + bias;
            float sc = (p.act == HSG_ACT_SIGMOID) ? __frcp_rn(1.0f + __expf(-m))
                       : (p.act == HSG_ACT_SOFTPLUS) ? (m > 20.f ? m : log1pf(__expf(m))) : m;
            p.out[(long long)cb * p.P + pi] = sc;
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

int umma_build_weights(const float* fc1, const float* w0, const float* b0, const float* pff_g, const float* pff_b,
                       const float* w1, const float* b1, const float* ln1_g, const float* c0, const float* cb0,
                       const float* cw1, const float* cb1, int hi, int natural_k, unsigned char* wimg_host, float* cst_host,
                       float* comb_out) {
    const int d = UM_D;
    std::vector<float> w0f((size_t)d * d);
    for (int n = 0; n < d; ++n) {
        double acc = b0[n];
        for (int k = 0; k < d; ++k) {
            w0f[(size_t)n * d + k] = 0.5f * w0[(size_t)n * d + k] * pff_g[k];   // LN gain and the swish 1/2 folded into W0
            acc += (double)w0[(size_t)n * d + k] * pff_b[k];                // and the LN bias into b0
        }
        cst_host[CST_B0F + n] = (float)(0.5 * acc);
    }
    memset(wimg_host, 0, SMH_WIMG_BYTES);
    for (int i = 0; i < d; ++i) cst_host[CST_G1 + i] = ln1_g[i];
    for (int i = 0; i < d / 2; ++i) { cst_host[CST_CB0 + i] = 0.5f * cb0[i]; cst_host[CST_CW1 + i] = cw1[i]; }
    cst_host[CST_CB1] = cb1[0];
    // every LayerNorm input is produced centred, and LN_pff . W0' is folded into fc1
    std::vector<double> fc1c((size_t)d * HSG_HD);
    for (int k = 0; k < HSG_HD; ++k) {
        double m = 0;
        for (int n = 0; n < d; ++n) m += fc1[(size_t)n * HSG_HD + k];
        m /= d;
        for (int n = 0; n < d; ++n) fc1c[(size_t)n * HSG_HD + k] = fc1[(size_t)n * HSG_HD + k] - m;
    }
    std::vector<float> comb((size_t)2 * d * HSG_HD), lo((size_t)2 * d * HSG_HD);
    for (int n = 0; n < d; ++n)
        for (int k = 0; k < HSG_HD; ++k) {
            comb[(size_t)n * HSG_HD + k] = (float)fc1c[(size_t)n * HSG_HD + k];
            double acc = 0;
            for (int j = 0; j < d; ++j) acc += (double)w0f[(size_t)n * d + j] * fc1c[(size_t)j * HSG_HD + k];
            comb[(size_t)(d + n) * HSG_HD + k] = (float)acc;
        }
    if (comb_out) memcpy(comb_out, comb.data(), comb.size() * sizeof(float));      // F = [fc1c ; W0' fc1c], natural K order (score_umma4.cu)
    if (!natural_k) {   // the ctx A tile arrives with its 16-byte chunks in HSG_CHUNK_POS order: permute the K columns of the image to match (score_umma2.cu keeps the natural order)
        std::vector<float> perm(comb.size());
        for (int n = 0; n < 2 * d; ++n)
            for (int k = 0; k < HSG_HD; ++k) perm[(size_t)n * HSG_HD + HSG_COL_POS(k)] = comb[(size_t)n * HSG_HD + k];
        comb.swap(perm);
    }
    swizzle_image(comb.data(), 2 * d, HSG_HD, reinterpret_cast<__half*>(wimg_host + SMF_B_FC1));
    for (size_t i = 0; i < comb.size(); ++i) lo[i] = comb[i] - __half2float(__float2half_rn(comb[i]));      // written below (hi only)
    std::vector<float> w1c((size_t)d * d);
    for (int k = 0; k < d; ++k) {
        double m = 0;
        for (int n = 0; n < d; ++n) m += w1[(size_t)n * d + k];
        m /= d;
        for (int n = 0; n < d; ++n) w1c[(size_t)n * d + k] = (float)(w1[(size_t)n * d + k] - m);
    }
    swizzle_image(w1c.data(), d, d, reinterpret_cast<__half*>(wimg_host + SMF_B_W1));
    double mb = 0;
    for (int i = 0; i < d; ++i) mb += b1[i];
    mb /= d;
    for (int i = 0; i < d; ++i) cst_host[CST_B1C + i] = (float)(b1[i] - mb);
    // fast variant: u' = u / gamma^2 (the S table is divided by gamma, convert_static_kernel), so gamma^2 goes into C0's columns
    std::vector<float> c0h((size_t)(d / 2) * d);
    for (int n = 0; n < d / 2; ++n)
        for (int k = 0; k < d; ++k) c0h[(size_t)n * d + k] = 0.5f * c0[(size_t)n * d + k] * (hi ? 1.0f : ln1_g[k] * ln1_g[k]);
    swizzle_image(c0h.data(), d / 2, d, reinterpret_cast<__half*>(wimg_host + SMF_B_C0));
    if (hi) {
        swizzle_image(lo.data(), 2 * d, HSG_HD, reinterpret_cast<__half*>(wimg_host + SMH_B_FC1_LO));
        for (size_t i = 0; i < c0h.size(); ++i) lo[i] = c0h[i] - __half2float(__float2half_rn(c0h[i]));
        swizzle_image(lo.data(), d / 2, d, reinterpret_cast<__half*>(wimg_host + SMH_B_C0_LO));
    } else {
        // bias atoms: K column 0 = fp16(b), column 1 = fp16(b - fp16(b)); the A operand carries (1, 1, 0, ...) in these columns
        std::vector<float> wb((size_t)d * d, 0.f), cb((size_t)(d / 2) * d, 0.f);
        for (int n = 0; n < d; ++n) {
            const float b = cst_host[CST_B1C + n], bh = __half2float(__float2half_rn(b));
            wb[(size_t)n * d] = bh; wb[(size_t)n * d + 1] = b - bh;
        }
        for (int n = 0; n < d / 2; ++n) {
            const float b = 0.5f * cb0[n], bh = __half2float(__float2half_rn(b));
            cb[(size_t)n * d] = bh; cb[(size_t)n * d + 1] = b - bh;
        }
        swizzle_image(wb.data(), d, d, reinterpret_cast<__half*>(wimg_host + SMF_X_W1B));
        swizzle_image(cb.data(), d / 2, d, reinterpret_cast<__half*>(wimg_host + SMF_X_C0B));
    }
    return 0;
}

// the fast variant folds the layer_norm1 gain into the S table and C0: it needs |gamma| bounded away from 0
int umma_gain_foldable(const float* ln1_g) {
    for (int i = 0; i < UM_D; ++i) if (!(fabsf(ln1_g[i]) >= 0.05f)) return 0;
    return 1;
}

// fp16 range analysis of a model (host, once per hsg_prepare).  Every 16-bit operand of the tensor-core scorer is bounded from
// the weights (a LayerNorm output obeys |x_hat_k| <= sqrt(d - 1)) and from the model's own S tables:
//   Q, K rows (fast variant: the half2 chains of a head hold sums of 4 products), V rows / ctx, h = swish(.) and the squared
//   difference u -- fast variant u' = (z rstd + (beta - S) / gamma)^2, high-accuracy variant u = (z rstd gamma + beta - S)^2.
// returns 0: the fast variant is safe; 1: take the high-accuracy variant; 2: take the fp32 kernels.
int umma_range_level(const float* wq, const float* wk, const float* wv, const float* g1, const float* b1, const float* g2, const float* b2,
                     const float* g3, const float* b3, const float* w0, const float* b0, const float* pff_g, const float* pff_b,
                     const float* ln1_g, const float* ln1_b, const float* S, long long s_rows, int d) {
    const double L = sqrt((double)d - 1.0), LIM = 6.0e4, ULIM = 240.0;      // 240^2 = 57600 < 65504
    auto row_bound = [&](const float* w, const float* g, const float* b, int rows) {
        double worst = 0;
        for (int n = 0; n < rows; ++n) {
            double a = 0, c0 = 0;
            for (int k = 0; k < d; ++k) { a += fabs((double)w[(size_t)n * d + k] * g[k]); c0 += (double)w[(size_t)n * d + k] * b[k]; }
            worst = std::max(worst, a * L + fabs(c0));
        }
        return worst;
    };
    const double qmax = row_bound(wq, g1, b1, HSG_HD) * HSG_L2E_QUARTER, kmax = row_bound(wk, g2, b2, HSG_HD), vmax = row_bound(wv, g3, b3, HSG_HD);
    double hmax = 0;
    for (int n = 0; n < d; ++n) {
        double a = 0, c0 = b0[n];
        for (int k = 0; k < d; ++k) { a += fabs(0.5 * (double)w0[(size_t)n * d + k] * pff_g[k]); c0 += (double)w0[(size_t)n * d + k] * pff_b[k]; }
        hmax = std::max(hmax, a * L + fabs(0.5 * c0));
    }
    double u_fast = 0, u_hi = 0;
    for (int f = 0; f < d; ++f) {
        double dmax = 0;
        for (long long r = 0; r < s_rows; ++r) dmax = std::max(dmax, fabs((double)ln1_b[f] - (double)S[r * d + f]));
        const double g = fabs((double)ln1_g[f]);
        u_hi = std::max(u_hi, dmax + L * g);
        u_fast = std::max(u_fast, (g > 0 ? dmax / g : 1e30) + L);
    }
    const bool finite_ok = std::isfinite(qmax) && std::isfinite(kmax) && std::isfinite(vmax) && std::isfinite(hmax) && std::isfinite(u_hi);
    if (!finite_ok || vmax > LIM || hmax > LIM || u_hi > ULIM) return 2;
    if (qmax > LIM || kmax > LIM || 4.0 * qmax * kmax > LIM || u_fast > ULIM) return 1;
    return 0;
}

int umma_wimg_bytes() { return SMH_WIMG_BYTES; }
int umma_cst_count() { return CST_COUNT; }

// fp32 static tables -> 16-bit / offset copies used by the tensor-core kernel
// Sb = beta(layer_norm1) - S, divided by gamma(layer_norm1) when ln1_g is given (fast variant, see umma_build_weights)
__global__ void convert_static_kernel(const float* __restrict__ V, const float* __restrict__ S, const float* __restrict__ ln1_b,
                                      const float* __restrict__ ln1_g, long long rows, int d, __half* __restrict__ V16,
                                      float* __restrict__ Sb, float* __restrict__ V32p) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < rows * HSG_HD) {
        { const int cc = (int)(i & (HSG_HD - 1)); V16[i - cc + HSG_COL_POS(cc)] = __float2half_rn(V[i]); }     // head-per-lane chunk order
        if (V32p) { const int cc = (int)(i & (HSG_HD - 1)); V32p[i - cc + HSG_HI_POS(cc)] = V[i]; }
    }
    if (i < rows * d) { long long r = i / d; int f = (int)(i - r * d); Sb[((long long)(f >> 2) * rows + r) * 4 + (f & 3)] = ln1_g ? (ln1_b[f] - S[i]) / ln1_g[f] : ln1_b[f] - S[i]; }
}

int launch_convert_static(hsg_ctx* c, const float* V, const float* S, long long rows, void* V16, float* Sb, float* V32p, cudaStream_t s) {
    long long n = rows * HSG_HD;
    convert_static_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(V, S, c->w[HSG_W_LN1_B], c->umma_hi ? nullptr : c->w[HSG_W_LN1_G], rows, c->d,
                                                                       (__half*)V16, Sb, V32p);
    c->launches++;
    HSG_CUDA(cudaGetLastError());
    return 0;
}

int launch_score_pairs_umma(hsg_ctx* c, int cell_start, int n_batch, float* out, cudaStream_t s) {
    if (c->P == 0 || n_batch == 0) return 0;
    UmmaParams p;
    p.QK16 = (const __half*)c->QK16; p.XS = c->XS; p.binV16 = (const __half*)c->binV16; p.binSb = c->binSb;
    p.cellV16 = (const __half*)c->cellV16; p.cellSb = c->cellSb; p.wimg = c->wimg;
    p.QK32 = c->QK; p.binV32 = c->binV32p; p.cellV32 = c->cellV32p;
    memcpy(p.cst, c->h_cst, sizeof(p.cst));
    p.pairs = c->pairs; p.pair_bias = c->pair_bias; p.P = c->P;
    p.n_bins = c->n_bins; p.n_cells = c->n_cells; p.cell_start = cell_start; p.n_batch = n_batch;
    p.tiles_per_cell = (int)((c->P + 127) / 128); p.act = c->act; p.out = out;
    long long n_tiles = (long long)n_batch * p.tiles_per_cell;
    long long quads = (n_tiles + UM_WG - 1) / UM_WG;
    int grid = (int)(quads < c->n_sm ? quads : c->n_sm);
    static bool attr_set[64] = {};               // per device: the attribute belongs to the device's function instance
    if (!attr_set[c->device & 63]) {
        HSG_CUDA(cudaFuncSetAttribute(score_umma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL_OF(0) + 1024));
        HSG_CUDA(cudaFuncSetAttribute(score_umma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL_OF(1) + 1024));
        attr_set[c->device & 63] = true;
    }
    if (c->umma_hi) score_umma_kernel<true><<<grid, UM_THREADS, SM_TOTAL_OF(1) + 1024, s>>>(p);
    else score_umma_kernel<false><<<grid, UM_THREADS, SM_TOTAL_OF(0) + 1024, s>>>(p);
    c->launches++;
    HSG_CUDA(cudaGetLastError());
    return 0;
}
