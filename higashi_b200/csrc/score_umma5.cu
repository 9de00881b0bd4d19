// score_umma5.cu -- K2, fast tensor-core variant, FIFTH generation: generation 4's MMA chain with TWO threads per triplet row.
// EXPERIMENT, NOT THE DEFAULT (HSG_K2_GEN=5): parity-green, but K2 takes 62.3 ms per 225 M triplets against generation 4's 48.7 --
// the two exchanges per token couple eight warps, every warp repeats the tile bookkeeping, the mbarrier polls and the waits, and at
// 80 registers the epilogues spill.  What the hypothesis below missed: the chain of a row tile is a SUM of unit round trips (tensor
// pipe ~28 %, tensor-memory reads at 64 B/clk per quadrant ~23 %, XU 37 %, L1 data pipe ~50 %, issue 42 %) that more threads per row
// do not shorten; only more row tiles in flight would overlap them, and tensor memory (512 columns) holds three.
//
// Ablations on generation 4 (score_umma4.cu, -DHSG_K2_ABLATE=1; 600 cells of C2): one warpgroup alone 74.8 ms, two 46.0, three
// 33.2 -- throughput follows the number of independent row chains almost linearly, and no pipe is the bound (L1 data pipe, XU,
// tensor pipe all below 60 %).  A warp of generation 4 runs ~1000 instructions per token at ~5 cycles each (its own dependency
// chains: tcgen05.ld -> FFMA2 -> MUFU -> pack -> tcgen05.st, L2 round trips of the table rows), and tensor memory caps the chains
// at three (3 x 128 accumulator columns + 2 x 64 ctx columns = 512).  More rows cannot be in flight; more THREADS per row can:
// tensor-memory lanes are tied to warp index mod 4, so warps q and q + 4 of a 256-thread group see the same 32 rows, and every
// phase of the chain splits by COLUMNS:
//   softmax weights  heads 4 hs .. 4 hs + 3           (half hs = 0 / 1; 16 instead of 32 row loads, 4 instead of 8 head dots)
//   ctx rows         chunks 8 hs .. 8 hs + 7          -> columns 32 hs .. 32 hs + 31 of the tensor-memory slot
//   weight slots     bytes 8 hs .. 8 hs + 7 of each 16-byte slot chunk;  P images: slots 2 hs, 2 hs + 1
//   h epilogue       sum of squares over y columns 32 hs ..; g columns 64 + 32 hs .. -> h columns 16 hs ..
//   u epilogue       sum of squares over z columns 32 hs ..; S' quads 8 hs ..        -> u' columns 16 hs ..
//   o epilogue       c columns 16 hs .. 16 hs + 15, partial dot with the classifier row
// The two partial sums of squares per token and the partial score per tile cross between the halves through 4 bytes of shared
// memory per row and a 256-thread named barrier (double-buffered: the barrier of exchange i + 1 orders the reads of exchange i
// before the writes of exchange i + 2).  896 threads per CTA: 3 groups x 256 epilogue threads at 80 registers + 4 issue warps at 24.
// Arithmetic, rounding points and the order of every per-thread sum are generation 4's, except that a 64-term sum of squares is
// now (sum over the own 32 columns) + (sum over the other 32) -- both halves add the same two numbers, so they agree bit for bit.
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "score_rows.cuh"

#define U5_GRP 3
#define U5_THREADS (U5_GRP * 256 + 128)
#define U5_SM_BAR SMF_WIMG_BYTES                 // 8 mbarriers
#define U5_SM_TMEM (U5_SM_BAR + 64)
#define U5_SM_SLOT (U5_SM_TMEM + 16)             // int owner[2], int slot_of[U5_GRP]
#define U5_SM_XBUF (SMF_WIMG_BYTES + 128)        // float [U5_GRP][2 parities][2 halves][128 rows]
#define U5_XBUF_BYTES (2 * 2 * 128 * 4)
#define U5_SM_AW 64512                           // 63 KB: U5_GRP x 16 KB weight-slot / P-image tiles (1024-byte aligned)
#define U5_AW_BYTES 16384
#define U5_SM_TOTAL (U5_SM_AW + U5_GRP * U5_AW_BYTES)
static_assert(U5_SM_XBUF + U5_GRP * U5_XBUF_BYTES <= U5_SM_AW, "exchange buffers overlap the slot tiles");

struct Umma5Params {
    Umma2Params b;
    const int4* tiles;       // [n_tiles_cell][2]: (p0, cnt, s1, s2), (bin_i of segment 0, 1, 2, -)   (build_tiles4)
    int n_tiles_cell;
    const uint4* binP;       // [n_bins][128 n] : 8 halves = heads 0..7 of P[bin][n]
    const uint4* cellP;      // [n_cells][128 n]
};

struct TileRow5 {
    int cb, pt, pi; bool valid; int2 pr; int seg;
};

__device__ __forceinline__ void sts64(uint32_t addr, uint32_t a, uint32_t b) {
    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ void group_bar(int g) { asm volatile("bar.sync %0, 256;" ::"r"(g + 1) : "memory"); }

// softmax weights of heads 4 hs .. 4 hs + 3 of a row: WA = (w_a[0], w_a[1]), (w_a[2], w_a[3]) as half2, WB likewise
template <int T>
__device__ __forceinline__ void attn_weights_half(const uint4* __restrict__ QKs, const float4* __restrict__ XSs, const uint32_t bi,
                                                  const uint32_t bj, const int hs, uint32_t (&WA)[2], uint32_t (&WB)[2]) {
    constexpr uint32_t G = HSG_GRP;
    float la[4], lb[4];
    {
        const float4 x = __ldg(XSs + grp_off<4>((T == 2) ? bj : bi) + (uint32_t)((T == 0 ? 0 : 2) + hs) * G);
        la[0] = x.x; la[1] = x.y; la[2] = x.z; la[3] = x.w;
    }
    if (T == 0) {
        const float4 x = __ldg(XSs + grp_off<4>(bj) + (uint32_t)hs * G);
        lb[0] = x.x; lb[1] = x.y; lb[2] = x.z; lb[3] = x.w;
    } else {
        const uint4* q = QKs + grp_off<32>((T == 1) ? bi : bj) + (uint32_t)(8 * hs) * G;
        const uint4* k = QKs + grp_off<32>((T == 1) ? bj : bi) + (uint32_t)(16 + 8 * hs) * G;
#pragma unroll
        for (int h = 0; h < 4; ++h)
            lb[h] = head_dot(__ldg(q + (2 * h) * G), __ldg(q + (2 * h + 1) * G), __ldg(k + (2 * h) * G), __ldg(k + (2 * h + 1) * G));
    }
    float wa[4], wb[4];
#pragma unroll
    for (int h = 0; h < 4; ++h) {
        wa[h] = rcp_approx(1.0f + ex2_approx(lb[h] - la[h]));
        wb[h] = 1.0f - wa[h];
        if (fmaxf(la[h], lb[h]) < -20.f) masked_row_exact(la[h], lb[h], wa[h], wb[h]);
    }
    WA[0] = pack_h2(wa[0], wa[1]); WA[1] = pack_h2(wa[2], wa[3]);
    WB[0] = pack_h2(wb[0], wb[1]); WB[1] = pack_h2(wb[2], wb[3]);
}

// chunks 8 hs .. 8 hs + 7 of ctx_var = w_b (.) V[bin_j] -> columns 32 hs .. 32 hs + 31 of the A slot
__device__ __forceinline__ void ctx_var_half(const uint4* __restrict__ binVc, const uint32_t bj, const int hs, const uint32_t t_a,
                                             const uint32_t (&WB)[2]) {
    const uint4* vb = binVc + grp_off<16>(bj) + (uint32_t)(8 * hs) * HSG_GRP;
#pragma unroll
    for (int c4 = 0; c4 < 2; ++c4) {
        uint32_t o[16];
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
            const int c = c4 * 4 + cc, lh = c >> 1;                    // local chunk, local head
            const __half2 wb2 = (lh & 1) ? __high2half2(u2h(WB[lh >> 1])) : __low2half2(u2h(WB[lh >> 1]));
            const uint4 b = __ldg(vb + c * HSG_GRP);
            o[cc * 4 + 0] = h2u(__hmul2(wb2, u2h(b.x))); o[cc * 4 + 1] = h2u(__hmul2(wb2, u2h(b.y)));
            o[cc * 4 + 2] = h2u(__hmul2(wb2, u2h(b.z))); o[cc * 4 + 3] = h2u(__hmul2(wb2, u2h(b.w)));
        }
        tmem_st16(t_a + (uint32_t)(32 * hs + 16 * c4), o);
    }
}

__device__ __forceinline__ float sumsq32(const uint32_t (&r)[32]) {
    float2 vA = make_float2(0.f, 0.f), vB = vA, vC = vA, vD = vA;
#pragma unroll
    for (int i = 0; i < 16; i += 4) {
        float2 a = f2(r[2 * i], r[2 * i + 1]), b = f2(r[2 * i + 2], r[2 * i + 3]);
        float2 c2 = f2(r[2 * i + 4], r[2 * i + 5]), d2 = f2(r[2 * i + 6], r[2 * i + 7]);
        vA = __ffma2_rn(a, a, vA); vB = __ffma2_rn(b, b, vB); vC = __ffma2_rn(c2, c2, vC); vD = __ffma2_rn(d2, d2, vD);
    }
    vA = __fadd2_rn(__fadd2_rn(vA, vB), __fadd2_rn(vC, vD));
    return vA.x + vA.y;
}

__global__ void __launch_bounds__(U5_THREADS, 1) score_umma5_kernel(Umma5Params pp) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const Umma2Params& p = pp.b;
    const int tid = threadIdx.x;
    const uint32_t sraw = smem_u32(smem_raw);
    const uint32_t sbase = (sraw + 1023u) & ~1023u;
    unsigned char* smem = smem_raw + (sbase - sraw);
    const float* cst = p.cst;
    const uint32_t bar_w = sbase + U5_SM_BAR;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + U5_SM_TMEM);
    volatile int* slot_owner = reinterpret_cast<volatile int*>(smem + U5_SM_SLOT);       // [2]: -1 free, else the owning group
    volatile int* slot_of = slot_owner + 2;                                             // [U5_GRP]: the slot a group holds

    if (tid == 0) {
        mbar_init(bar_w, 1);
        for (int i = 0; i < U5_GRP; ++i) mbar_init(sbase + U5_SM_BAR + 8 + i * 8, 1);
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
    const long long n_quads = (n_tiles + U5_GRP - 1) / U5_GRP;

    if (tid >= U5_GRP * 256) {
        // ---------------------------------------------------------------- issue warps: warp w < 3 serves group w
        asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");
        const int sw = __shfl_sync(0xffffffffu, (tid - U5_GRP * 256) >> 5, 0);
        if (sw < U5_GRP) {
            const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
            const uint32_t s_ty = tb + (uint32_t)(sw * 128), s_th = s_ty + 64, s_bar = sbase + U5_SM_BAR + 8 + sw * 8;
            const uint32_t s_aw = sbase + U5_SM_AW + (uint32_t)sw * U5_AW_BYTES;
            constexpr uint32_t idesc1 = umma_idesc_f16(128, 128);
            uint32_t lead_p;
            asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(lead_p));
            const bool lead = lead_p != 0;
            long long n_tile_own = 0;
            for (long long q = blockIdx.x; q < n_quads; q += gridDim.x) ++n_tile_own;
            for (long long i = 0; i < n_tile_own; ++i) {
#pragma unroll 1
                for (int t = 0; t < 3; ++t) {
                    named_bar_sync(5 + sw, 288); tc_fence_after();                        // ctx rows / weight slots / P images complete
                    if (t != 2) {
                        const uint32_t s_a = tb + 384u + 64u * (uint32_t)__shfl_sync(0xffffffffu, slot_of[sw], 0);
                        if (lead) mma_group_ts<8, 128>(s_a, sbase + SMF_B_FC1, 128, s_ty, 0u);
                    }
                    if (lead) {
                        umma_f16(s_ty, umma_desc(s_aw), umma_desc(s_aw + 64u), idesc1, (t != 2) ? 1u : 0u);
                        umma_f16(s_ty, umma_desc(s_aw + 32u), umma_desc(s_aw + 96u), idesc1, 1u);
                        umma_commit(s_bar);
                    }
                    __syncwarp();
                    named_bar_sync(5 + sw, 288); tc_fence_after();                        // h complete in TMEM
                    if (lead) {
                        mma_group_ts<4, 64>(s_th, sbase + SMF_B_W1, 64, s_ty, 1u);
                        umma_f16_ts(s_ty, s_th + 32, umma_desc(sbase + SMF_X_W1B), umma_idesc_f16(128, 64), 1u);
                        umma_commit(s_bar);
                    }
                    __syncwarp();
                    named_bar_sync(5 + sw, 288); tc_fence_after();                        // u' complete in TMEM
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
    asm volatile("setmaxnreg.inc.sync.aligned.u32 80;");
    const int g = tid >> 8, gt = tid & 255, hs = gt >> 7, row = gt & 127, quad = (gt >> 5) & 3;
    const uint32_t bar_mma = sbase + U5_SM_BAR + 8 + g * 8;
    const uint32_t t_lane = ((uint32_t)(quad * 32)) << 16;
    const uint32_t t_y = tmem_base + (uint32_t)(g * 128), t_h = t_y + 64;
    // this thread's row of the group's weight / P tile with (row & 7) also in address bits 4..6: xor with (chunk << 4) = swizzled slot
    const uint32_t aw_row_x = sbase + U5_SM_AW + (uint32_t)g * U5_AW_BYTES + (uint32_t)(row >> 3) * 1024u + (uint32_t)(row & 7) * (128u + 16u);
    volatile float* xb = reinterpret_cast<volatile float*>(smem + U5_SM_XBUF + g * U5_XBUF_BYTES);
    uint32_t phase = 0;
    int xpar = 0;
    // partial sums cross between the two halves of a row: returns the OTHER half's value
    auto exchange = [&](float v) {
        xb[xpar * 256 + hs * 128 + row] = v;
        group_bar(g);
        const float o = xb[xpar * 256 + (1 - hs) * 128 + row];
        xpar ^= 1;
        return o;
    };
    auto decode = [&](long long q) {
        TileRow5 tr;
        const bool tile_ok = q * U5_GRP + g < n_tiles;
        const long long tile = tile_ok ? q * U5_GRP + g : n_tiles - 1;
        tr.cb = (int)(tile / pp.n_tiles_cell);
        tr.pt = (int)(tile - (long long)tr.cb * pp.n_tiles_cell);
        const int4 ta = __ldg(pp.tiles + 2 * tr.pt);
        tr.pi = ta.x + row;
        tr.valid = tile_ok && row < ta.y;
        tr.pr = __ldg(p.pairs + (row < ta.y ? tr.pi : ta.x + ta.y - 1));
        tr.seg = (row >= ta.z ? 1 : 0) + (row >= ta.w ? 1 : 0);
        return tr;
    };
    auto weights = [&](const TileRow5& tr, int tk, uint32_t (&WA)[2], uint32_t (&WB)[2]) {
        const uint4* QKs = p.QKc + (size_t)tr.cb * (size_t)(p.n_grp * 32 * HSG_GRP);
        const float4* XSs = p.XSc + (size_t)tr.cb * (size_t)(p.n_grp * 4 * HSG_GRP);
        if (tk == 0) attn_weights_half<0>(QKs, XSs, (uint32_t)tr.pr.x, (uint32_t)tr.pr.y, hs, WA, WB);
        else if (tk == 1) attn_weights_half<1>(QKs, XSs, (uint32_t)tr.pr.x, (uint32_t)tr.pr.y, hs, WA, WB);
        else attn_weights_half<2>(QKs, XSs, (uint32_t)tr.pr.x, (uint32_t)tr.pr.y, hs, WA, WB);
    };
    // operands of MMA 1 of token tk of tile row tr (this thread's half of them)
    auto ctx_build = [&](const TileRow5& tr, int tk, const uint32_t (&WA)[2], const uint32_t (&WB)[2]) {
        if (tk == 0) {
            // P images of the tile: row n of the tile gets P[x][n][0..7] of slot token x in chunk 4 + slot; this half copies slots 2 hs, 2 hs + 1
            const int4 tb4 = __ldg(pp.tiles + 2 * tr.pt + 1);
            const uint4* s0 = hs ? pp.binP + (size_t)tb4.y * 128u : pp.cellP + (size_t)(p.cell_start + tr.cb) * 128u;
            const uint4* s1 = pp.binP + (size_t)(hs ? tb4.z : tb4.x) * 128u;
            const uint4 a = __ldg(s0 + row), b = __ldg(s1 + row);
            sts128(aw_row_x ^ ((uint32_t)(4 + 2 * hs) << 4), a.x, a.y, a.z, a.w);
            sts128(aw_row_x ^ ((uint32_t)(5 + 2 * hs) << 4), b.x, b.y, b.z, b.w);
        }
        // weight slots: chunk 0 = cell token, chunk 1 + s = bin_i of segment s; this half owns bytes 8 hs .. 8 hs + 7 of every chunk
        {
            const uint32_t ho = (uint32_t)(8 * hs);
            const bool c_on = tk != 0;                                   // tokens 1 and 2 attend to the cell with w_a
            sts64((aw_row_x ^ 0u) + ho, c_on ? WA[0] : 0u, c_on ? WA[1] : 0u);
            uint32_t s0 = 0u, s1 = 0u;                                   // token 0 attends to bin_i with w_a, token 2 with w_b
            if (tk == 0) { s0 = WA[0]; s1 = WA[1]; }
            else if (tk == 2) { s0 = WB[0]; s1 = WB[1]; }
#pragma unroll
            for (int s = 0; s < 3; ++s) {
                const bool on = tr.seg == s;
                sts64((aw_row_x ^ ((uint32_t)(1 + s) << 4)) + ho, on ? s0 : 0u, on ? s1 : 0u);
            }
        }
        fence_async_smem();
        if (tk != 2) {
            if (gt == 0) {
                int sl = -1;
                while (sl < 0) {
                    if (atomicCAS((int*)&slot_owner[0], -1, g) == -1) sl = 0;
                    else if (atomicCAS((int*)&slot_owner[1], -1, g) == -1) sl = 1;
                    else __nanosleep(64);
                }
                slot_of[g] = sl;
            }
            group_bar(g);
            tc_fence_after();                // the previous owner's MMA 1 (its reads of the slot) completed before it released
            ctx_var_half(p.binVc, (uint32_t)tr.pr.y, hs, tmem_base + 384u + 64u * (uint32_t)slot_of[g] + t_lane, WB);
        }
    };
    auto request = [&]() { tmem_st_wait(); tc_fence_before(); named_bar_arrive(5 + g, 288); };
    auto wait_mma = [&]() { mbar_wait(bar_mma, phase); phase ^= 1; tc_fence_after(); };
    auto release_slot = [&]() { if (gt == 0) { tc_fence_before(); __threadfence_block(); slot_owner[slot_of[g]] = -1; } };

    // h epilogue, this half: rstd of y_c from both halves' sums of squares; h = swish(rstd g + b0') for g columns 32 hs .. 32 hs + 31
    auto epi_h5 = [&]() {
        uint32_t r[32];
        tmem_ld32(t_y + t_lane + 32 * hs, r);
        tmem_ld_wait();
        const float own = sumsq32(r);
        tmem_ld32(t_h + t_lane + 32 * hs, r);              // the g columns: in flight across the exchange
        const float tot = hs ? exchange(own) + own : own + exchange(own);       // columns 0..31 first in both halves
        const float rstd = rsqrtf(tot * (1.0f / 64.0f) + HSG_LN_EPS);
        const float2 rs = make_float2(rstd, rstd);
        tmem_ld_wait();
        uint32_t hp[16];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const float* bq = cst + CST_B0F + 32 * hs + c * 8;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float2 w = __ffma2_rn(f2(r[c * 8 + 2 * q], r[c * 8 + 2 * q + 1]), rs, make_float2(bq[2 * q], bq[2 * q + 1]));
                float2 th = make_float2(fast_tanh(w.x), fast_tanh(w.y));
                float2 hh = __ffma2_rn(w, th, w);
                hp[c * 4 + q] = pack_h2(hh.x, hh.y);
            }
        }
        tmem_st16(t_h + t_lane + 16 * hs, hp);
        if (hs == 0) {
            const uint32_t one[8] = {0x3C003C00u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
            tmem_st8(t_h + t_lane + 32, one);
        }
    };
    // the S' quads 8 hs .. 8 hs + 7 of a token's row, requested before the wait on MMA 2
    auto load_s_half = [&](int t, int cell, int2 pr, float4 (&sreg)[8]) {
        if (t == 0) {
            const float4* Srow = p.cellSb + (size_t)cell * 16u + 8 * hs;
#pragma unroll
            for (int c = 0; c < 8; ++c) sreg[c] = ldg_nc_v4_volatile(Srow + c);
        } else {
            const float4* Scol = p.binSb + grp_off<16>((uint32_t)(t == 1 ? pr.x : pr.y)) + (uint32_t)(8 * hs) * HSG_GRP;
#pragma unroll
            for (int c = 0; c < 8; ++c) sreg[c] = ldg_nc_v4_volatile(Scol + c * HSG_GRP);
        }
    };
    // u epilogue, this half: rstd of z_c, u' = (z_c rstd + S')^2 for features 32 hs .. 32 hs + 31 -> u' columns 16 hs .. 16 hs + 15
    auto epi_u5 = [&](const float4 (&sreg)[8]) {
        float2 rs;
        {
            uint32_t r[32];
            tmem_ld32(t_y + t_lane + 32 * hs, r);
            tmem_ld_wait();
            const float own = sumsq32(r);
            const float tot = hs ? exchange(own) + own : own + exchange(own);
            const float rstd = rsqrtf(tot * (1.0f / 64.0f) + HSG_LN_EPS);
            rs = make_float2(rstd, rstd);
        }
#pragma unroll
        for (int qt = 0; qt < 2; ++qt) {
            uint32_t r[16], up[8];
            tmem_ld16(t_y + t_lane + 32 * hs + qt * 16, r);
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
            tmem_st8(t_h + t_lane + 16 * hs + qt * 8, up);
        }
    };
    // o epilogue, this half: partial cw1 . swish(c) over c columns 16 hs .. 16 hs + 15
    auto epi_o5 = [&]() {
        uint32_t r[16];
        tmem_ld16(t_y + t_lane + 16 * hs, r);
        tmem_ld_wait();
        float2 oA = make_float2(0.f, 0.f), oB = oA;
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
            const float* wq = cst + CST_CW1 + 16 * hs + 2 * i;
            float2 w0 = f2(r[2 * i], r[2 * i + 1]), w1 = f2(r[2 * i + 2], r[2 * i + 3]);
            float2 h0 = __ffma2_rn(w0, make_float2(fast_tanh(w0.x), fast_tanh(w0.y)), w0);
            float2 h1 = __ffma2_rn(w1, make_float2(fast_tanh(w1.x), fast_tanh(w1.y)), w1);
            oA = __ffma2_rn(h0, make_float2(wq[0], wq[1]), oA);
            oB = __ffma2_rn(h1, make_float2(wq[2], wq[3]), oB);
        }
        return (oA.x + oA.y) + (oB.x + oB.y);
    };

    // software pipeline over the group's own chain (generation 3's): the softmax weights of the NEXT token are computed while MMA 1 of
    // the current one is in flight, its MMA 1 operands are built while MMA 3 is in flight; across tiles too
    if ((long long)blockIdx.x < n_quads) {
        uint32_t WA[2], WB[2];
        TileRow5 cur = decode(blockIdx.x), nxt = cur;
        weights(cur, 0, WA, WB);
        ctx_build(cur, 0, WA, WB);
        request();
        for (long long q = blockIdx.x; q < n_quads; q += gridDim.x) {
            const bool has_next = q + gridDim.x < n_quads;                     // block-uniform
            const int cell = p.cell_start + cur.cb;
            float osum = 0.f;
#pragma unroll 1
            for (int t = 0; t < 3; ++t) {
                const bool more = t < 2 || has_next;
                if (more) weights(t < 2 ? cur : nxt, t < 2 ? t + 1 : 0, WA, WB);       // MMA 1 of token t is in flight
                wait_mma();
                if (t != 2) release_slot();                                    // MMA 1 has read the slot
                epi_h5();
                request();                                                     // MMA 2
                float4 sreg[8];
                load_s_half(t, cell, cur.pr, sreg);
                if (t == 1 && has_next) nxt = decode(q + gridDim.x);           // pair loads of the next tile, one token ahead of their use
                wait_mma();
                epi_u5(sreg);
                request();                                                     // MMA 3
                if (more) ctx_build(t < 2 ? cur : nxt, t < 2 ? t + 1 : 0, WA, WB);      // MMA 3 is in flight
                wait_mma();
                osum += epi_o5();
                if (more) request();                                           // MMA 1 of the next token (orders the TMEM reads above)
                else tc_fence_before();
            }
            // the halves add their partial scores in the same order (columns 0..15 first)
            const float oth = exchange(osum);
            if (hs == 0 && cur.valid) {
                const float m = ((osum + oth) + 3.0f * cst[CST_CB1]) * (1.0f / 3.0f) + __ldg(p.pair_bias + cur.pi);
                const float sc = (p.act == HSG_ACT_SIGMOID) ? __frcp_rn(1.0f + __expf(-m))
                                 : (p.act == HSG_ACT_SOFTPLUS) ? (m > 20.f ? m : log1pf(__expf(m))) : m;
                if (p.out16) p.out16[(long long)cur.cb * p.P + cur.pi] = __float2half_rn(sc);
                else p.out[(long long)cur.cb * p.P + cur.pi] = sc;
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

int launch_score_pairs_umma5(hsg_ctx* c, int cell_start, int n_batch, float* out, void* out16, cudaStream_t s) {
    if (c->P == 0 || n_batch == 0) return 0;
    Umma5Params pp;
    Umma2Params& p = pp.b;
    p.QKc = (const uint4*)c->QK16; p.XSc = (const float4*)c->XS; p.binVc = (const uint4*)c->binV16; p.cellV = (const uint4*)c->cellV16;
    p.binSb = (const float4*)c->binSb; p.cellSb = (const float4*)c->cellSb; p.wimg = c->wimg;
    memcpy(p.cst, c->h_cst, sizeof(p.cst));
    p.pairs = c->pairs; p.pair_bias = c->pair_bias; p.P = c->P;
    p.n_bins = c->n_bins; p.n_grp = (c->n_bins + HSG_GRP - 1) / HSG_GRP; p.n_cells = c->n_cells; p.cell_start = cell_start; p.n_batch = n_batch;
    p.tiles_per_cell = c->n_tiles4; p.act = c->act; p.out = out; p.out16 = (__half*)out16;
    p.nwg = U5_GRP; p.exp_flags = 0; p.quad_cells = 0;
    pp.tiles = (const int4*)c->tiles4; pp.n_tiles_cell = c->n_tiles4; pp.binP = (const uint4*)c->binP; pp.cellP = (const uint4*)c->cellP;
    const long long q5 = ((long long)n_batch * c->n_tiles4 + U5_GRP - 1) / U5_GRP;
    const int g5 = (int)(q5 < c->n_sm ? q5 : c->n_sm);
    static bool attr5[64] = {};
    if (!attr5[c->device & 63]) {
        HSG_CUDA(cudaFuncSetAttribute(score_umma5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, U5_SM_TOTAL + 1024));
        attr5[c->device & 63] = true;
    }
    score_umma5_kernel<<<g5, U5_THREADS, U5_SM_TOTAL + 1024, s>>>(pp);
    c->launches++;
    HSG_CUDA(cudaGetLastError());
    return 0;
}
