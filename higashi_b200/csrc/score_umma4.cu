// score_umma4.cu -- K2, fast tensor-core variant, FOURTH generation: the warp-uniform half of the attention context rides on the
// tensor core.
//
// Same math as score_umma2.cu (Modules.py:726-783, 1029-1095, 865-889, 936-972).  What changed is MMA 1.
// ncu on generation 3 (profiles/r02_k2_gen3_ncu.txt): the L1 data pipe is the busiest unit of the SM (55 % LSU wavefronts + 13 %
// tensor-core operand reads), and experiments that hide latency without taking load off it do not pay (L1 prefetch of the bin_j
// rows: +2 % time; a third ctx slot: +1 %; packed tanh: 0).  A row reads 224 x 16 bytes per triplet, and a 128-bit load costs four
// data-pipe cycles per warp whatever its address pattern -- also when all 32 lanes read the SAME row.  128 of the 224 loads are
// such warp-uniform rows: pairs are enumerated row by row (generate_binpair order), so a tile of consecutive pairs is a few runs
// ("segments") of consecutive bin_j under one bin_i, and the cell is the same for the whole tile.
//
// The context of a token is ctx = w_a (.) V_a + w_b (.) V_b with one softmax weight per head, and MMA 1 is linear in it:
//     [y_c | g] = ctx . F^T = sum_h  w_a[h] (V_a[h,:] . F_h^T) + w_b[h] (V_b[h,:] . F_h^T),          F = [fc1c ; W0' fc1c]
// P[x][n][h] = V_x[h,:] . F[n, 16h .. 16h+15] does not depend on the cell or on the pair: it is a table over bins and cells, built
// once per model (8 halves per output n = 2 KB per bin).  For a warp-uniform token x the contribution is W . P[x]^T with W the
// [128 rows x 8 heads] matrix of softmax weights -- a K = 8 slab of an MMA whose B operand is P[x] itself.  So:
//   * a tile carries at most THREE segments (tile table built once per pair set: `build_tiles4`); the warpgroup keeps a 16 KB
//     K-major SWIZZLE_128B tile in shared memory whose K columns 0..31 are the A operand (8 weights of the cell slot, 8 of each
//     segment slot: a row writes its weights into its own segment's slot and zeros elsewhere) and whose K columns 32..63 are the B
//     operand (P[cell], P[bin_i of segment 0 / 1 / 2]: one 16-byte chunk per thread and slot, once per tile);
//   * token 0 (cell; attends bin_i, bin_j):  MMA 1 = (w_b (.) V[bin_j]) . F^T  (A in tensor memory, K = 128)  +  slots (K = 32)
//     token 1 (bin_i; attends cell, bin_j):  the same with the cell slot carrying w_a
//     token 2 (bin_j; attends cell, bin_i):  the slots only -- no ctx rows, no V loads, no tensor-memory slot, K = 32 instead of 128;
//   * per triplet: 152 instead of 224 128-bit loads, 128 instead of 384 packed half2 operations in the ctx build, 8 instead of 12
//     tcgen05.st of ctx chunks.
// Rounding: P is rounded to fp16 once (from fp32 V and fp32 F) where generation 3 rounded V, the products and the ctx sums; the
// varying half keeps generation 3's rounding points.  y_c stays centred up to that rounding (the columns of F are centred in fp32).
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "score_rows.cuh"

#define U4_WG 3
// ablation switches (HSG_K2_EXP bits, HSG_K2_NWG) are compiled in only with -DHSG_K2_ABLATE=1: as run-time tests they cost the default
// kernel 14 % (measured); results are WRONG with any bit set.  bit 0 no softmax weights, 1 no ctx rows, 2 no S' loads, 3 / 4 / 5 no
// h / u / o epilogue.  Round 2, 600 cells of C2, K2 ms: all 33.2 | 2 warpgroups 46.0 | 1 warpgroup 74.8 | -weights 26.7 | -ctx 30.6 |
// -S' 31.6 | -h 29.3 | -u 30.7 | -o 30.6 | -all three epilogues 24.9 | MMA chain + barriers only 12.8.
#ifndef HSG_K2_ABLATE
#define HSG_K2_ABLATE 0
#endif
#if HSG_K2_ABLATE
#define ABL(bit) (p.exp_flags & (bit))
#else
#define ABL(bit) false
#endif
#define U4_SM_BAR SMF_WIMG_BYTES                 // 8 mbarriers
#define U4_SM_TMEM (U4_SM_BAR + 64)
#define U4_SM_SLOT (U4_SM_TMEM + 16)             // int owner[2], int slot_of[U4_WG]
#define U4_SM_AW (SMF_WIMG_BYTES + 1024)         // U4_WG x 16 KB: weights (K 0..31) and P images (K 32..63) of the current tile
#define U4_AW_BYTES 16384
#define U4_SM_TOTAL (U4_SM_AW + U4_WG * U4_AW_BYTES)

struct Umma4Params {
    Umma2Params b;
    const int4* tiles;       // [n_tiles_cell][2]: (p0, cnt, s1, s2), (bin_i of segment 0, 1, 2, -)
    int n_tiles_cell;
    const uint4* binP;       // [n_bins][128 n] : 8 halves = heads 0..7 of P[bin][n]
    const uint4* cellP;      // [n_cells][128 n]
};

struct TileRow4 {
    int cb; long long pi; bool valid; int2 pr; float bias;
    int seg;                 // segment of this row inside its tile (0..2)
    int bi0, bi1, bi2;       // bin_i of the tile's segments (unused ones repeat segment 0)
    // seg == 3 for every row of a DENSE tile (more than three segments: the short rows at the end of a chromosome): no segment
    // slot matches, both bin terms of every token go through the varying ctx operand (all three tokens take a slot)
    __device__ __forceinline__ bool dense() const { return seg == 3; }
};

// softmax weights of a row as two chunks of 8 halves: WA = w_a[0..7], WB = w_b[0..7] (same arithmetic as attn_weights)
template <int T>
__device__ __forceinline__ void attn_weights4(const uint4* __restrict__ QKs, const float4* __restrict__ XSs, const uint32_t bi,
                                              const uint32_t bj, uint32_t (&WA)[4], uint32_t (&WB)[4]) {
    uint32_t w[8];
    attn_weights<T>(QKs, XSs, bi, bj, w);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const __half2 lo = u2h(w[2 * i]), hi = u2h(w[2 * i + 1]);        // (wa, wb) of heads 2i, 2i + 1
        WA[i] = h2u(__lows2half2(lo, hi)); WB[i] = h2u(__highs2half2(lo, hi));
    }
}

// ctx_var = w_b (.) V[bin_j] -> the 64 TMEM columns of an A slot
__device__ __forceinline__ void ctx_var_tmem(const uint4* __restrict__ binVc, const uint32_t bj, const uint32_t t_a, const uint32_t (&WB)[4]) {
    const uint4* vb = binVc + grp_off<16>(bj);
#pragma unroll
    for (int c4 = 0; c4 < 4; ++c4) {
        uint32_t o[16];
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
            const int c = c4 * 4 + cc, h = c >> 1;
            const __half2 wb2 = (h & 1) ? __high2half2(u2h(WB[h >> 1])) : __low2half2(u2h(WB[h >> 1]));
            const uint4 b = __ldg(vb + c * HSG_GRP);
            o[cc * 4 + 0] = h2u(__hmul2(wb2, u2h(b.x))); o[cc * 4 + 1] = h2u(__hmul2(wb2, u2h(b.y)));
            o[cc * 4 + 2] = h2u(__hmul2(wb2, u2h(b.z))); o[cc * 4 + 3] = h2u(__hmul2(wb2, u2h(b.w)));
        }
        tmem_st16(t_a + (uint32_t)c4 * 16u, o);
    }
}

// dense tiles: ctx_var = w_a (.) V[x] + w_b (.) V[y]  (token 0: x = bin_i, y = bin_j) -> the 64 TMEM columns of an A slot
// (not inlined: it runs for one token of a few per cent of the tiles, and the main loop is already at the size where more inlined
//  code costs the common path -- measured: inlining it slowed the tiles that never call it by 4 %)
__device__ __noinline__ void ctx_var2_tmem(const uint4* __restrict__ binVc, const uint32_t bx, const uint32_t by, const uint32_t t_a,
                                           const uint4 wa4, const uint4 wb4) {
    const uint32_t WA[4] = {wa4.x, wa4.y, wa4.z, wa4.w}, WB[4] = {wb4.x, wb4.y, wb4.z, wb4.w};
    const uint4* va = binVc + grp_off<16>(bx);
    const uint4* vb = binVc + grp_off<16>(by);
#pragma unroll
    for (int c4 = 0; c4 < 4; ++c4) {
        uint32_t o[16];
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
            const int c = c4 * 4 + cc, h = c >> 1;
            const __half2 wa2 = (h & 1) ? __high2half2(u2h(WA[h >> 1])) : __low2half2(u2h(WA[h >> 1]));
            const __half2 wb2 = (h & 1) ? __high2half2(u2h(WB[h >> 1])) : __low2half2(u2h(WB[h >> 1]));
            const uint4 a = __ldg(va + c * HSG_GRP), b = __ldg(vb + c * HSG_GRP);
            o[cc * 4 + 0] = h2u(__hfma2(wb2, u2h(b.x), __hmul2(wa2, u2h(a.x)))); o[cc * 4 + 1] = h2u(__hfma2(wb2, u2h(b.y), __hmul2(wa2, u2h(a.y))));
            o[cc * 4 + 2] = h2u(__hfma2(wb2, u2h(b.z), __hmul2(wa2, u2h(a.z)))); o[cc * 4 + 3] = h2u(__hfma2(wb2, u2h(b.w), __hmul2(wa2, u2h(a.w))));
        }
        tmem_st16(t_a + (uint32_t)c4 * 16u, o);
    }
}

// P [rows][128 n][8 heads] fp16 from V [rows][128] fp32 and F [128 n][128 k] fp32
__global__ void build_ptable_kernel(const float* __restrict__ V, const float* __restrict__ F, long long rows, __half* __restrict__ P) {
    __shared__ float v[HSG_HD];
    const long long x = blockIdx.x;
    if (x >= rows) return;
    v[threadIdx.x] = V[x * HSG_HD + threadIdx.x];
    __syncthreads();
    const int n = threadIdx.x;
    const float4* f = reinterpret_cast<const float4*>(F + (size_t)n * HSG_HD);
    uint32_t o[4];
    float acc[8];
#pragma unroll
    for (int h = 0; h < 8; ++h) {
        float a = 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float4 w = __ldg(f + h * 4 + q);
            a = fmaf(v[h * 16 + q * 4 + 0], w.x, a); a = fmaf(v[h * 16 + q * 4 + 1], w.y, a);
            a = fmaf(v[h * 16 + q * 4 + 2], w.z, a); a = fmaf(v[h * 16 + q * 4 + 3], w.w, a);
        }
        acc[h] = a;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) o[i] = pack_h2(acc[2 * i], acc[2 * i + 1]);
    reinterpret_cast<uint4*>(P)[x * 128 + n] = make_uint4(o[0], o[1], o[2], o[3]);
}

// DENSE: the tile table may hold dense tiles (see build_tiles4).  A second instantiation rather than a run-time flag: the mere
// presence of the dense-tile code costs the plain tiles ~3 % (B200: C3 45.1 -> 46.7 ms per 250 cells), so pair sets whose
// chromosome tails are a small share of the tiles (C3, C5) keep the kernel without it.
template <bool DENSE>
__global__ void __launch_bounds__(512, 1) score_umma4_kernel(Umma4Params pp) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const Umma2Params& p = pp.b;
    const int tid = threadIdx.x, wg = tid >> 7, wtid = tid & 127, warp_in_wg = (tid >> 5) & 3;
    const uint32_t sraw = smem_u32(smem_raw);
    const uint32_t sbase = (sraw + 1023u) & ~1023u;
    unsigned char* smem = smem_raw + (sbase - sraw);
    const float* cst = p.cst;
    const uint32_t bar_w = sbase + U4_SM_BAR;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + U4_SM_TMEM);
    volatile int* slot_owner = reinterpret_cast<volatile int*>(smem + U4_SM_SLOT);       // [2]: -1 free, else the owning warpgroup
    volatile int* slot_of = slot_owner + 2;                                             // [U4_WG]: the slot a warpgroup holds

    if (tid == 0) {
        mbar_init(bar_w, 1);
        for (int i = 0; i < U4_WG; ++i) mbar_init(sbase + U4_SM_BAR + 8 + i * 8, 1);
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
    const long long n_tiles = (long long)p.n_batch * pp.n_tiles_cell;
#if HSG_K2_ABLATE
    const int nwg = p.nwg;                                                     // active warpgroups (fewer than U4_WG only in experiments)
#else
    constexpr int nwg = U4_WG;
#endif
    const long long n_quads = (n_tiles + nwg - 1) / nwg;

    if (wg == U4_WG) {
        // ---------------------------------------------------------------- issue warps: warp w < 3 serves warpgroup w
        asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");
        const int sw = __shfl_sync(0xffffffffu, warp_in_wg, 0);
        if (sw < nwg) {
            const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
            const uint32_t s_ty = tb + (uint32_t)(sw * 128), s_th = s_ty + 64, s_bar = sbase + U4_SM_BAR + 8 + sw * 8;
            const uint32_t s_aw = sbase + U4_SM_AW + (uint32_t)sw * U4_AW_BYTES;
            constexpr uint32_t idesc1 = umma_idesc_f16(128, 128);
            uint32_t lead_p;
            asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(lead_p));
            const bool lead = lead_p != 0;
            long long n_tile_own = 0;
            for (long long q = blockIdx.x; q < n_quads; q += gridDim.x) ++n_tile_own;
            for (long long i = 0; i < n_tile_own; ++i) {
                // dense tile (flag in the tile table): token 2 has a K = 128 part too
                const long long tile_i = ((long long)blockIdx.x + i * gridDim.x) * nwg + sw;
                const long long tile_c = tile_i < n_tiles ? tile_i : n_tiles - 1;           // (a warpgroup past the end re-runs the last tile, unstored)
                int dense_i = 0;
                if constexpr (DENSE) dense_i = __ldg(reinterpret_cast<const int*>(pp.tiles + 2 * (int)(tile_c % pp.n_tiles_cell) + 1) + 3);
#pragma unroll 1
                for (int t = 0; t < 3; ++t) {
                    named_bar_sync(5 + sw, 160); tc_fence_after();                        // ctx_var / weight slots / P images complete
                    const bool has_var = t != 2 || (DENSE && dense_i != 0);                        // (warp-uniform)
                    if (has_var) {
                        const uint32_t s_a = tb + 384u + 64u * (uint32_t)__shfl_sync(0xffffffffu, slot_of[sw], 0);
                        if (lead) mma_group_ts<8, 128>(s_a, sbase + SMF_B_FC1, 128, s_ty, 0u);
                    }
                    if (lead) {
                        umma_f16(s_ty, umma_desc(s_aw), umma_desc(s_aw + 64u), idesc1, has_var ? 1u : 0u);
                        umma_f16(s_ty, umma_desc(s_aw + 32u), umma_desc(s_aw + 96u), idesc1, 1u);
                        umma_commit(s_bar);
                    }
                    __syncwarp();
                    named_bar_sync(5 + sw, 160); tc_fence_after();                        // h complete in TMEM
                    if (lead) {
                        mma_group_ts<4, 64>(s_th, sbase + SMF_B_W1, 64, s_ty, 1u);
                        umma_f16_ts(s_ty, s_th + 32, umma_desc(sbase + SMF_X_W1B), umma_idesc_f16(128, 64), 1u);
                        umma_commit(s_bar);
                    }
                    __syncwarp();
                    named_bar_sync(5 + sw, 160); tc_fence_after();                        // u' complete in TMEM
                    if (lead) {
                        mma_group_ts<4, 32>(s_th, sbase + SMF_B_C0, 32, s_ty, 0u);
                        umma_f16_ts(s_ty, s_th + 32, umma_desc(sbase + SMF_X_C0B), umma_idesc_f16(128, 32), 1u);
                        umma_commit(s_bar);
                    }
                    __syncwarp();
                }
            }
        }
        tc_fence_before();
        __syncthreads();
        return;
    }
    asm volatile("setmaxnreg.inc.sync.aligned.u32 160;");
    const uint32_t bar_mma = sbase + U4_SM_BAR + 8 + wg * 8;
    const uint32_t t_lane = ((uint32_t)(warp_in_wg * 32)) << 16;
    const uint32_t t_y = tmem_base + (uint32_t)(wg * 128), t_h = t_y + 64;
    // this thread's row of the weight / P tile with (row & 7) also in address bits 4..6: xor with (chunk << 4) = swizzled chunk slot
    const uint32_t aw_row_x = sbase + U4_SM_AW + (uint32_t)wg * U4_AW_BYTES + (uint32_t)(wtid >> 3) * 1024u + (uint32_t)(wtid & 7) * (128u + 16u);
    uint32_t phase = 0;
    auto decode = [&](long long q) {
        TileRow4 tr;
        const bool tile_ok = q * nwg + wg < n_tiles;
        const long long tile = tile_ok ? q * nwg + wg : n_tiles - 1;
        tr.cb = (int)(tile / pp.n_tiles_cell);
        const int pt = (int)(tile - (long long)tr.cb * pp.n_tiles_cell);
        const int4 ta = __ldg(pp.tiles + 2 * pt), tb4 = __ldg(pp.tiles + 2 * pt + 1);
        tr.pi = (long long)ta.x + wtid;
        tr.valid = tile_ok && wtid < ta.y;
        tr.pr = __ldg(p.pairs + (wtid < ta.y ? tr.pi : (long long)ta.x + ta.y - 1));
        tr.bias = tr.valid ? __ldg(p.pair_bias + tr.pi) : 0.f;
        tr.seg = (wtid >= ta.z ? 1 : 0) + (wtid >= ta.w ? 1 : 0);
        if constexpr (DENSE) tr.seg += tb4.w;                                   // dense tile: s1 = s2 = 0 and flag 1 -> 3
        HSG_DASSERT(pt >= 0 && pt < pp.n_tiles_cell && ta.y >= 1 && ta.y <= 128 && (long long)ta.x + ta.y <= p.P);
        HSG_DASSERT(tr.pr.x >= 0 && tr.pr.x < p.n_bins && tr.pr.y >= 0 && tr.pr.y < p.n_bins);
        HSG_DASSERT(tb4.x >= 0 && tb4.x < p.n_bins && tb4.y >= 0 && tb4.y < p.n_bins && tb4.z >= 0 && tb4.z < p.n_bins);
        tr.bi0 = tb4.x; tr.bi1 = tb4.y; tr.bi2 = tb4.z;
        return tr;
    };
    auto weights = [&](const TileRow4& tr, int tk, uint32_t (&WA)[4], uint32_t (&WB)[4]) {
        const uint4* QKs = p.QKc + (size_t)tr.cb * (size_t)(p.n_grp * 32 * HSG_GRP);
        const float4* XSs = p.XSc + (size_t)tr.cb * (size_t)(p.n_grp * 4 * HSG_GRP);
        if (tk == 0) attn_weights4<0>(QKs, XSs, (uint32_t)tr.pr.x, (uint32_t)tr.pr.y, WA, WB);
        else if (tk == 1) attn_weights4<1>(QKs, XSs, (uint32_t)tr.pr.x, (uint32_t)tr.pr.y, WA, WB);
        else attn_weights4<2>(QKs, XSs, (uint32_t)tr.pr.x, (uint32_t)tr.pr.y, WA, WB);
    };
    auto sts_chunk = [&](int chunk, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
        sts128(aw_row_x ^ ((uint32_t)chunk << 4), a, b, c, d);
    };
    // operands of MMA 1 of token tk of tile row tr: (tk == 0: the P images of the tile,) the weight slots, and for tk != 2 the
    // varying half of the ctx in a tensor-memory slot taken for the purpose
    auto ctx_build = [&](const TileRow4& tr, int tk, const uint32_t (&WA)[4], const uint32_t (&WB)[4]) {
        if (tk == 0) {
            // thread n copies the 8 halves P[x][n][0..7] of the four slot tokens into row n, chunks 4..7
            const uint4 pc = __ldg(pp.cellP + (size_t)(p.cell_start + tr.cb) * 128u + wtid);
            const uint4 p0 = __ldg(pp.binP + (size_t)tr.bi0 * 128u + wtid), p1 = __ldg(pp.binP + (size_t)tr.bi1 * 128u + wtid),
                        p2 = __ldg(pp.binP + (size_t)tr.bi2 * 128u + wtid);
            sts_chunk(4, pc.x, pc.y, pc.z, pc.w); sts_chunk(5, p0.x, p0.y, p0.z, p0.w);
            sts_chunk(6, p1.x, p1.y, p1.z, p1.w); sts_chunk(7, p2.x, p2.y, p2.z, p2.w);
        }
        // weight slots: chunk 0 = cell token, chunk 1 + s = bin_i of segment s
        {
            const bool c_on = tk != 0;                                   // tokens 1 and 2 attend to the cell with w_a
            sts_chunk(0, c_on ? WA[0] : 0u, c_on ? WA[1] : 0u, c_on ? WA[2] : 0u, c_on ? WA[3] : 0u);
            uint32_t s0 = 0u, s1 = 0u, s2 = 0u, s3 = 0u;                 // token 0 attends to bin_i with w_a, token 2 with w_b
            if (tk == 0) { s0 = WA[0]; s1 = WA[1]; s2 = WA[2]; s3 = WA[3]; }
            else if (tk == 2) { s0 = WB[0]; s1 = WB[1]; s2 = WB[2]; s3 = WB[3]; }
#pragma unroll
            for (int s = 0; s < 3; ++s) {
                const bool on = tr.seg == s;
                sts_chunk(1 + s, on ? s0 : 0u, on ? s1 : 0u, on ? s2 : 0u, on ? s3 : 0u);
            }
        }
        fence_async_smem();
        if (tk != 2 || (DENSE && tr.dense())) {                     // (warpgroup-uniform)
            if (wtid == 0) {
                int sl = -1;
                while (sl < 0) {
                    if (atomicCAS((int*)&slot_owner[0], -1, wg) == -1) sl = 0;
                    else if (atomicCAS((int*)&slot_owner[1], -1, wg) == -1) sl = 1;
                    else __nanosleep(64);
                }
                HSG_DASSERT(sl == 0 || sl == 1);          // tensor-memory columns 384 + 64 sl .. + 63 of the 512 allocated
                slot_of[wg] = sl;
            }
            wg_bar(wg);
            tc_fence_after();                // the previous owner's MMA 1 (its reads of the slot) completed before it released
            const uint32_t t_slot = tmem_base + 384u + 64u * (uint32_t)slot_of[wg] + t_lane;
            // dense tile: token 0 attends bin_i (w_a) and bin_j (w_b) through the operand; tokens 1 / 2 the cell through its slot and
            // bin_j / bin_i here
            if constexpr (DENSE) {
                if (tr.dense() && tk == 0)
                    ctx_var2_tmem(p.binVc, (uint32_t)tr.pr.x, (uint32_t)tr.pr.y, t_slot, make_uint4(WA[0], WA[1], WA[2], WA[3]), make_uint4(WB[0], WB[1], WB[2], WB[3]));
                else if (!ABL(2)) ctx_var_tmem(p.binVc, (uint32_t)((tr.dense() && tk == 2) ? tr.pr.x : tr.pr.y), t_slot, WB);
            } else {
                if (!ABL(2)) ctx_var_tmem(p.binVc, (uint32_t)tr.pr.y, t_slot, WB);
            }
        }
    };
    auto request = [&]() { tmem_st_wait(); tc_fence_before(); named_bar_arrive(5 + wg, 160); };
    auto wait_mma = [&]() { mbar_wait(bar_mma, phase); phase ^= 1; tc_fence_after(); };
    auto release_slot = [&]() { if (wtid == 0) { tc_fence_before(); __threadfence_block(); slot_owner[slot_of[wg]] = -1; } };

    // software pipeline over the warpgroup's own chain (generation 3's): the softmax weights of the NEXT token are computed while MMA 1
    // of the current one is in flight, its MMA 1 operands are built while MMA 3 is in flight; across tiles too
    if ((long long)blockIdx.x < n_quads && wg < nwg) {
        uint32_t WA[4], WB[4];
        TileRow4 cur = decode(blockIdx.x);
        weights(cur, 0, WA, WB);
        ctx_build(cur, 0, WA, WB);
        request();
        for (long long q = blockIdx.x; q < n_quads; q += gridDim.x) {
            const bool has_next = q + gridDim.x < n_quads;                     // block-uniform
            const TileRow4 nxt = decode(has_next ? q + gridDim.x : q);
            const int cell = p.cell_start + cur.cb;
            float osum = 0.f;
#pragma unroll 1
            for (int t = 0; t < 3; ++t) {
                const bool more = t < 2 || has_next;
                if (more && !ABL(1)) weights(t < 2 ? cur : nxt, t < 2 ? t + 1 : 0, WA, WB);       // MMA 1 of token t is in flight
                wait_mma();
                if (t != 2 || (DENSE && cur.dense())) release_slot();                       // MMA 1 has read the slot
#if HSG_EPI_PIPE & 8
                prefetch_s_row(p, t, cell, cur.pr);
#endif
#ifdef HSG_K2_TANH2
                if (!ABL(8)) epi_h<true>(t_y, t_h, t_lane, cst);
#else
                if (!ABL(8)) epi_h<false>(t_y, t_h, t_lane, cst);
#endif
                request();                                                     // MMA 2
                float4 sreg[16];
                if (!ABL(4)) load_s_row(p, t, cell, cur.pr, sreg);
                else {
#pragma unroll
                    for (int i = 0; i < 16; ++i) sreg[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
                wait_mma();
                if (!ABL(16)) epi_u(t_y, t_h, t_lane, sreg);
                request();                                                     // MMA 3
                if (more) ctx_build(t < 2 ? cur : nxt, t < 2 ? t + 1 : 0, WA, WB);      // MMA 3 is in flight
                wait_mma();
                if (!ABL(32)) osum += epi_o(t_y, t_lane, cst);
                if (more) request();                                           // MMA 1 of the next token (orders the TMEM reads above)
                else tc_fence_before();
            }
            if (cur.valid) {
                TileRow tr; tr.cb = cur.cb; tr.pi = cur.pi; tr.valid = true; tr.pr = cur.pr; tr.bias = cur.bias;
                store_score(p, tr, osum);
            }
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
// Tile table of a pair list in generate_binpair order: greedy runs of at most 128 consecutive pairs with at most three distinct
// bin_i (a new segment starts where bin_i changes).  Entry = (p0, cnt, s1, s2), (bin_i of segment 0, 1, 2, dense); rows [0, s1) are
// segment 0, [s1, s2) segment 1, [s2, cnt) segment 2; unused segments start at 128 and repeat bin_i of segment 0.
// The rows at the end of a chromosome are short (row i of n bins has n - i - 1 pairs): three of them fill a fraction of a tile (C2:
// 1904 tiles per cell for 1758 full ones, 8 % of the rows idle).  With `allow_dense`, a run whose three-segment prefix would leave
// more than 16 rows idle becomes a DENSE tile instead: 128 pairs whatever their bin_i, flagged in the last field; the kernel then
// sends both bin terms of every token through the varying ctx operand (no segment slots).
void build_tiles4(const int2* pairs, long long P, std::vector<int>& out, bool allow_dense) {
    out.clear();
    long long p0 = 0;
    while (p0 < P) {
        int cnt = 0, nseg = 0, s[3] = {128, 128, 128}, bi[3] = {pairs[p0].x, pairs[p0].x, pairs[p0].x};
        int last = -1;
        while (p0 + cnt < P && cnt < 128) {
            const int b = pairs[p0 + cnt].x;
            if (b != last) {
                if (nseg == 3) break;
                s[nseg] = cnt; bi[nseg] = b; ++nseg; last = b;
            }
            ++cnt;
        }
        int dense = 0;
        if (allow_dense && cnt < 112 && p0 + cnt < P) {             // cut short by the segment limit, not by the end of the list
            cnt = (int)std::min<long long>(128, P - p0);
            s[1] = s[2] = 0; bi[1] = bi[2] = bi[0];            // every row gets segment index 1 + 1 + flag = 3 in the kernel
            dense = 1;
        }
        const int e[8] = {(int)p0, cnt, s[1], s[2], bi[0], bi[1], bi[2], dense};
        out.insert(out.end(), e, e + 8);
        p0 += cnt;
    }
}

int launch_build_ptable(hsg_ctx* c, const float* V, const float* F, long long rows, void* P, cudaStream_t s) {
    if (rows == 0) return 0;
    build_ptable_kernel<<<(unsigned)rows, 128, 0, s>>>(V, F, rows, (__half*)P);
    c->launches++;
    HSG_CUDA(cudaGetLastError());
    return 0;
}

int launch_score_pairs_umma4(hsg_ctx* c, int cell_start, int n_batch, float* out, void* out16, cudaStream_t s) {
    if (c->P == 0 || n_batch == 0) return 0;
    Umma4Params pp;
    Umma2Params& p = pp.b;
    p.QKc = (const uint4*)c->QK16; p.XSc = (const float4*)c->XS; p.binVc = (const uint4*)c->binV16; p.cellV = (const uint4*)c->cellV16;
    p.binSb = (const float4*)c->binSb; p.cellSb = (const float4*)c->cellSb; p.wimg = c->wimg;
    memcpy(p.cst, c->h_cst, sizeof(p.cst));
    p.pairs = c->pairs; p.pair_bias = c->pair_bias; p.P = c->P;
    p.n_bins = c->n_bins; p.n_grp = (c->n_bins + HSG_GRP - 1) / HSG_GRP; p.n_cells = c->n_cells; p.cell_start = cell_start; p.n_batch = n_batch;
    p.tiles_per_cell = c->n_tiles4; p.act = c->act; p.out = out; p.out16 = (__half*)out16;
    static const int env_nwg = getenv("HSG_K2_NWG") ? std::max(1, std::min(U4_WG, atoi(getenv("HSG_K2_NWG")))) : U4_WG;
    static const int env_exp = getenv("HSG_K2_EXP") ? atoi(getenv("HSG_K2_EXP")) : 0;
    p.nwg = env_nwg; p.exp_flags = env_exp; p.quad_cells = 0;
    pp.tiles = (const int4*)c->tiles4; pp.n_tiles_cell = c->n_tiles4; pp.binP = (const uint4*)c->binP; pp.cellP = (const uint4*)c->cellP;
    const long long q4 = ((long long)n_batch * c->n_tiles4 + env_nwg - 1) / env_nwg;
    const int g4 = (int)(q4 < c->n_sm ? q4 : c->n_sm);
    static bool attr4[64] = {};
    if (!attr4[c->device & 63]) {
        HSG_CUDA(cudaFuncSetAttribute(score_umma4_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, U4_SM_TOTAL + 1024));
        HSG_CUDA(cudaFuncSetAttribute(score_umma4_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, U4_SM_TOTAL + 1024));
        attr4[c->device & 63] = true;
    }
    if (c->k2_dense_tiles) score_umma4_kernel<true><<<g4, 512, U4_SM_TOTAL + 1024, s>>>(pp);
    else score_umma4_kernel<false><<<g4, 512, U4_SM_TOTAL + 1024, s>>>(pp);
    c->launches++;
    HSG_CUDA(cudaGetLastError());
    return 0;
}
