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

#include "score_rows.cuh"

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
// HYB: every warpgroup OWNS a ctx slot -- warpgroups 0 and 1 the two 64-column TMEM slots, warpgroup 2 a 32 KB K-major SWIZZLE_128B
// tile in shared memory (MMA 1 of that warpgroup runs in SS mode).  ncu on the dynamic variant: 7.3 % of all samples (10 % of the
// epilogue warps' time) sat in the warpgroup barrier behind the slot acquisition -- three warpgroups were taking turns on two slots.
#define U3H_SM_A (SMF_WIMG_BYTES + 1024)
#define U3H_SM_TOTAL (U3H_SM_A + SM_A_BYTES)

__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
// L1 prefetch of 16 consecutive chunks (first chunk c0) of the grouped-table rows this warp is about to read: consecutive lanes
// hold consecutive bins, so the rows of lanes 0 / 8 / 16 / 24 name the (up to) four 128-byte lines per chunk; lane l asks for the
// line of chunk c0 + (l >> 2) and c0 + 8 + (l >> 2) of the row of lane 8 (l & 3): two instructions per table and token.
template <int CH>
__device__ __forceinline__ void prefetch_rows(const uint4* __restrict__ table, const uint32_t b, const int lane, const int c0) {
    const uint32_t bs = __shfl_sync(0xffffffffu, b, (lane & 3) * 8);
    const uint4* q = table + grp_off<CH>(bs) + (uint32_t)(c0 + (lane >> 2)) * HSG_GRP;
    prefetch_l1(q); prefetch_l1(q + 8 * HSG_GRP);
}

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

template <bool TANH2, bool PIPE, bool HYB>
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
                if constexpr (HYB) {
                    if (lead) {
                        if (sw == 2) mma_group<8, 128>(sbase + U3H_SM_A, sbase + SMF_B_FC1, 128, s_ty, 0u);
                        else mma_group_ts<8, 128>(tb + 384u + 64u * (uint32_t)sw, sbase + SMF_B_FC1, 128, s_ty, 0u);
                        umma_commit(s_bar);
                    }
                } else {
                    const uint32_t s_a = tb + 384u + 64u * (uint32_t)__shfl_sync(0xffffffffu, slot_of[sw], 0);
                    if (lead) { mma_group_ts<8, 128>(s_a, sbase + SMF_B_FC1, 128, s_ty, 0u); umma_commit(s_bar); }
                }
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
    const uint32_t a_row_x = sbase + U3H_SM_A + (uint32_t)(wtid >> 3) * 1024u + (uint32_t)(wtid & 7) * (128u + 16u);
    auto ctx_and_request = [&](const TileRow& tr, int tk, const uint32_t (&w)[8]) {
        if constexpr (HYB) {
            const uint4* cellVrow = p.cellV + (size_t)(p.cell_start + tr.cb) * 16u;
            if (wg == 2) {
                if (tk == 0) attn_ctx<0>(p.binVc, cellVrow, (uint32_t)tr.pr.x, (uint32_t)tr.pr.y, a_row_x, w);
                else if (tk == 1) attn_ctx<1>(p.binVc, cellVrow, (uint32_t)tr.pr.x, (uint32_t)tr.pr.y, a_row_x, w);
                else attn_ctx<2>(p.binVc, cellVrow, (uint32_t)tr.pr.x, (uint32_t)tr.pr.y, a_row_x, w);
                fence_async_smem();
            } else {
                const uint32_t t_a = tmem_base + 384u + 64u * (uint32_t)wg + t_lane;
                if (tk == 0) attn_ctx_tmem<0>(p.binVc, cellVrow, (uint32_t)tr.pr.x, (uint32_t)tr.pr.y, t_a, w);
                else if (tk == 1) attn_ctx_tmem<1>(p.binVc, cellVrow, (uint32_t)tr.pr.x, (uint32_t)tr.pr.y, t_a, w);
                else attn_ctx_tmem<2>(p.binVc, cellVrow, (uint32_t)tr.pr.x, (uint32_t)tr.pr.y, t_a, w);
            }
            return;
        }
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
    auto release_slot = [&]() { if constexpr (!HYB) { if (wtid == 0) { tc_fence_before(); __threadfence_block(); slot_owner[slot_of[wg]] = -1; } } };
    // L1 prefetch of the bin_j rows the NEXT phases will read (exp_flags bit 1 switches it off): called once per token, right
    // after the S' loads and before the wait on MMA 2, i.e. about one MMA round trip + one epilogue ahead of the first use
    const bool do_pf = !(p.exp_flags & 2);
    const int lane = tid & 31;
    auto prefetch_ahead = [&](const TileRow& cur, const TileRow& nxt, int t, bool has_next) {
        if (!do_pf) return;
        const size_t qk_slot = (size_t)(p.n_grp * 32 * HSG_GRP);
        if (t == 0) {
            prefetch_rows<32>(p.QKc + (size_t)cur.cb * qk_slot, (uint32_t)cur.pr.y, lane, 0);        // Q[bin_j] for the weights of token 2
        } else if (t == 1) {
            prefetch_rows<16>(reinterpret_cast<const uint4*>(p.binSb), (uint32_t)cur.pr.y, lane, 0);  // S'[bin_j] of token 2
        } else if (has_next) {
            prefetch_rows<16>(p.binVc, (uint32_t)nxt.pr.y, lane, 0);                                  // V[bin_j] for the ctx of tokens 0 / 1
            prefetch_rows<32>(p.QKc + (size_t)nxt.cb * qk_slot, (uint32_t)nxt.pr.y, lane, 16);       // K[bin_j] for the weights of token 1
        }
    };

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
                prefetch_ahead(cur, nxt, t, has_next);
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
            HSG_CUDA(cudaFuncSetAttribute(score_umma3_kernel<false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, U3_SM_TOTAL + 1024));
            HSG_CUDA(cudaFuncSetAttribute(score_umma3_kernel<false, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, U3_SM_TOTAL + 1024));
            HSG_CUDA(cudaFuncSetAttribute(score_umma3_kernel<false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, U3H_SM_TOTAL + 1024));
            HSG_CUDA(cudaFuncSetAttribute(score_umma3_kernel<true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, U3H_SM_TOTAL + 1024));
            attr3[c->device & 63] = true;
        }
        p.quad_cells = 0; p.nwg = U3_WG;
        // HSG_K2_SLOTS=dyn: two TMEM ctx slots taken in turn by the three warpgroups (the first form of this kernel); default: owned slots
        static const bool env_hyb = !(getenv("HSG_K2_SLOTS") && getenv("HSG_K2_SLOTS")[0] == 'd');
        if (env_pipe && env_hyb) {
            if (env_tanh2) score_umma3_kernel<true, true, true><<<g3, 512, U3H_SM_TOTAL + 1024, s>>>(p);
            else score_umma3_kernel<false, true, true><<<g3, 512, U3H_SM_TOTAL + 1024, s>>>(p);
        } else if (env_pipe) score_umma3_kernel<false, true, false><<<g3, 512, U3_SM_TOTAL + 1024, s>>>(p);
        else score_umma3_kernel<false, false, false><<<g3, 512, U3_SM_TOTAL + 1024, s>>>(p);
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
