// gather_stream.cu -- K1a as a STREAM: the neighbour rows of a (cell, bin range) work item are contiguous in the packed CSR
// (row = cell * n_bins + g), so a block fetches them with bulk asynchronous copies (cp.async.bulk + mbarrier, the TMA engine)
// instead of chasing nbr -> row_ptr -> col / val per token.
//
// Same arithmetic as the other K1a kernels (Impute.py:78-121 prep_one with local_transfer_range = 0; Modules.py:1364-1371
// forward_GCN): A = sum_n w_n A_n in neighbour order (fp32), x pdf(0), log1p, L1 row normalisation, neigh = A_hat . E_bin.
//
//   * work item = (chromosome c, cell of the batch, tile of TB consecutive bins of c).  Items are numbered chromosome-major and
//     every block takes ONE contiguous range of them (equal ranges: one wave, no tail), so a block works on one chromosome
//     (two or three when its range straddles a boundary) and keeps that chromosome's E table [n_c x 64] fp32 in shared memory
//     (one bulk copy per chromosome change).
//   * a PRODUCER warp runs ahead of the consumers through a ring of NS stages (one consumer pass of tokens each; as many stages
//     as shared memory holds, <= 12):
//       hop A (item i + NS):     the k+1 row_ptr slices [row0, row0 + nt] of the neighbour cells -> shared memory (k+1 bulk copies);
//       hop B (item i + NS / 2): with the extents of hop A in shared memory, the k+1 col and val segments [lo, hi) of the item
//                                (16-byte aligned outwards) -> shared memory (2 (k+1) bulk copies);
//     the consumers merge item i from shared memory meanwhile.  Three dependent global round trips per TOKEN in the older
//     kernels become two per ITEM (62 tokens), several items deep (measured: with one item of lookahead the consumers spent
//     most of their time polling the full barrier -- they finish an item in about a microsecond).
//   * CONSUMER warps take the tokens of the item two at a time (16 lanes per token while both hold at most 16 entries, else one
//     token per warp).  Sparse token (at most 32 entries in its k+1 rows): one entry per
//     lane in row-major order, duplicates found with match.any, the owner (first occurrence) adds its peers' values in lane
//     order = neighbour order.  Longer token: a dense per-warp accumulator in shared memory, rows added one after the other,
//     then one atomicExch per entry hands each column's sum to its first visitor (and clears the accumulator).  Either way the
//     candidates land in a per-warp list in first-touch order; the gather reads four candidates at a time from the E table in
//     shared memory (8 lanes x 8 columns, two conflict-free 128-bit loads per 8 FMAs) like the hash / chromosome kernels.
//   * items whose segments exceed the stage capacity and tokens with more distinct columns than the candidate list holds go to
//     the overflow list that the generic kernel finishes.
//
// MEASURED (B200, round 2) -- opt-in, NOT the default: parity-green (tests/test_gpu_parity.py::test_gather_stream_*), but C2 K1a takes
// 2.16 - 2.39 ms per 1000 cells against 1.93 for the register kernel (C1 0.26 vs 0.21; C3, where only 8 consumer warps fit next to
// the 128 KB E table, 5.4 vs 3.4 for the hash-merge kernel).  ncu (profiles/r02_k1_stream_ncu.txt): DRAM 1.7 % of peak, issue
// 0.64 per scheduler, short scoreboard 30 % -- with the three global hops gone, the per-token dependency chain (extents -> prefix
// -> entry -> match.any -> log1p -> norm -> gather) is what bounds the kernel, at ~220 warp instructions per token; one item of
// lookahead or nine made no difference.  What would: four tokens per warp for rows of <= 8 entries and a densified-tile x E_c MMA
// for the long ones.
#include <algorithm>
#include <cmath>
#include <cstdlib>

#include "umma_common.cuh"

#define GST_MAX_CHROM 64
#define GST_MAX_NS 16
#define GST_MAX_K1 8
#define GST_MAX_GRID 256

struct GatherStreamParams {
    const int64_t* row_ptr; const int32_t* col; const float* val; const int32_t* nbr; const float* nbr_w;
    const float* E_bin; const int* bin_off; float* neigh; int* tok_list; int* tok_count;
    int k1, weighted, n_bins, cell_start, n_batch, n_chrom, max_nc;
    int TB, CAP, LCAP, n_warps, n_items, NS, AH;     // NS ring stages; hop A runs AH items ahead of hop B
    int item_first[GST_MAX_CHROM + 1];        // items [item_first[c], item_first[c + 1]) belong to chromosome c
    int tiles[GST_MAX_CHROM];                 // bin tiles per cell on chromosome c
    int blk_first[GST_MAX_GRID + 1];          // block b works on items [blk_first[b], blk_first[b + 1]): equal TOKEN counts
};

struct GstLayout {
    uint32_t e_bytes, bar_off, stage_off, stage_bytes, meta_bytes, rp_bytes, col_off, val_off, warp_off, warp_bytes, acc_bytes, total;
};
// meta of a stage: int64 sbase[8] | float w[8] | int d[8] | int flag
__host__ __device__ inline GstLayout gst_layout(int max_nc, int k1, int TB, int CAP, int LCAP, int n_warps, int NS) {
    GstLayout L;
    L.e_bytes = (uint32_t)max_nc * 256u;
    L.bar_off = L.e_bytes;
    L.stage_off = L.bar_off + 512u;
    L.meta_bytes = 144u;
    L.rp_bytes = (uint32_t)k1 * (uint32_t)(TB + 4) * 8u;
    L.col_off = L.meta_bytes + L.rp_bytes;
    L.val_off = L.col_off + (uint32_t)CAP * 4u;
    L.stage_bytes = L.val_off + (uint32_t)CAP * 4u;
    L.warp_off = L.stage_off + (uint32_t)NS * L.stage_bytes;
    L.acc_bytes = (((uint32_t)max_nc + 3u) & ~3u) * 4u;
    L.warp_bytes = L.acc_bytes + (uint32_t)LCAP * 8u;
    L.total = L.warp_off + (uint32_t)n_warps * L.warp_bytes;
    return L;
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// wait with a watchdog (2 s of wall clock): a protocol error aborts the kernel (launch failure) instead of hanging the device
__device__ __forceinline__ void mbar_wait_wd(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    unsigned long long t0 = 0;
    unsigned polls = 0;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity), "r"(1000000u) : "memory");
        if (!ok && (++polls & 255u) == 0u) {
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (t0 == 0) t0 = now;
            else if (now - t0 > 2000000000ull) __trap();
        }
    } while (!ok);
}

struct GstItem { int c, cb, l0, nt, off, nc; };
__device__ __forceinline__ GstItem gst_decode(const GatherStreamParams& p, int item, int& c_hint) {
    int c = c_hint;
    while (item >= p.item_first[c + 1]) ++c;
    c_hint = c;
    GstItem it;
    it.c = c;
    const int r = item - p.item_first[c], tl = p.tiles[c];
    it.cb = r / tl;
    it.l0 = (r - it.cb * tl) * p.TB;
    it.off = p.bin_off[c];
    it.nc = p.bin_off[c + 1] - it.off;
    it.nt = min(p.TB, it.nc - it.l0);
    return it;
}

__global__ void __launch_bounds__(1024, 1) neigh_gather_stream_kernel(const __grid_constant__ GatherStreamParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    const GstLayout L = gst_layout(p.max_nc, p.k1, p.TB, p.CAP, p.LCAP, p.n_warps, p.NS);
    const int NS = p.NS;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t sbase = smem_u32(smem);
    const uint32_t bar_e = sbase + L.bar_off;
    auto bar_a = [&](int s) { return sbase + L.bar_off + 8u + (uint32_t)s * 8u; };
    auto bar_b = [&](int s) { return sbase + L.bar_off + 8u + (uint32_t)(NS + s) * 8u; };
    auto bar_empty = [&](int s) { return sbase + L.bar_off + 8u + (uint32_t)(2 * NS + s) * 8u; };
    if (threadIdx.x == 0) {
        mbar_init(bar_e, 1);
        for (int s = 0; s < NS; ++s) { mbar_init(bar_a(s), 1); mbar_init(bar_b(s), 1); mbar_init(bar_empty(s), (uint32_t)p.n_warps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int i0 = p.blk_first[blockIdx.x], i1 = p.blk_first[blockIdx.x + 1];
    const int n_my = i1 - i0;
    if (n_my <= 0) return;

    if (warp == p.n_warps) {
        // ------------------------------------------------------------------ producer warp (lane n < k1 <-> neighbour row n)
        int c_hint_a = 0, c_hint_b = 0;
        int nb_cell = -1, nb_src = 0;                    // neighbour id / weight of lane n, reloaded when the cell changes
        float nb_w = 1.0f;
        auto stage_ptr = [&](int s) { return smem + L.stage_off + (uint32_t)s * L.stage_bytes; };
        auto hop_a = [&](int j) {                       // row_ptr slices of item i0 + j
            const int s = j % NS, use = j / NS;
            if (use > 0) mbar_wait_wd(bar_empty(s), (uint32_t)(use - 1) & 1u);
            const GstItem it = gst_decode(p, i0 + j, c_hint_a);
            unsigned char* st = stage_ptr(s);
            uint32_t bytes = 0;
            int64_t row0 = 0;
            int d = 0;
            const int cell = p.cell_start + it.cb;
            if (cell != nb_cell) {
                nb_cell = cell; nb_src = cell; nb_w = 1.0f;
                if (p.weighted && lane < p.k1) { nb_src = __ldg(p.nbr + (int64_t)cell * p.k1 + lane); nb_w = __ldg(p.nbr_w + (int64_t)cell * p.k1 + lane); }
            }
            if (lane < p.k1) {
                row0 = (int64_t)nb_src * p.n_bins + it.off + it.l0;
                d = (int)(row0 & 1);
                bytes = (uint32_t)((d + it.nt + 2) & ~1) * 8u;
                reinterpret_cast<float*>(st + 64)[lane] = nb_w;
                reinterpret_cast<int*>(st + 96)[lane] = d;
            }
            uint32_t tot = bytes;
#pragma unroll
            for (int o = 1; o < GST_MAX_K1; o <<= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
            if (lane == 0) mbar_expect_tx(bar_a(s), tot);
            __syncwarp();
            if (lane < p.k1)
                tma_bulk_g2s(smem_u32(st + L.meta_bytes) + (uint32_t)lane * (uint32_t)(p.TB + 4) * 8u, p.row_ptr + (row0 - d), bytes, bar_a(s));
        };
        auto hop_b = [&](int j) {                       // col / val segments of item i0 + j
            const int s = j % NS, use = j / NS;
            mbar_wait_wd(bar_a(s), (uint32_t)use & 1u);
            const GstItem it = gst_decode(p, i0 + j, c_hint_b);
            unsigned char* st = stage_ptr(s);
            int64_t a_lo = 0;
            int cnt = 0;
            if (lane < p.k1) {
                const int64_t* rp = reinterpret_cast<const int64_t*>(st + L.meta_bytes) + (size_t)lane * (size_t)(p.TB + 4);
                const int d = reinterpret_cast<const int*>(st + 96)[lane];
                const int64_t lo = rp[d], hi = rp[d + it.nt];
                a_lo = lo & ~(int64_t)3;
                cnt = hi > lo ? (int)(((hi + 3) & ~(int64_t)3) - a_lo) : 0;
            }
            int end = cnt;                               // inclusive prefix over lanes 0..7
#pragma unroll
            for (int o = 1; o < GST_MAX_K1; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, end, o, GST_MAX_K1);
                if ((lane & (GST_MAX_K1 - 1)) >= o) end += t;
            }
            const int total = __shfl_sync(0xffffffffu, end, GST_MAX_K1 - 1);
            const int base = end - cnt;
            const bool over = total > p.CAP;
            HSG_DASSERT(cnt >= 0 && base >= 0 && (over || base + cnt <= p.CAP));
            if (lane < p.k1) reinterpret_cast<int64_t*>(st)[lane] = (int64_t)base - a_lo;          // smem index of entry e = e + sbase
            if (lane == 0) reinterpret_cast<int*>(st + 128)[0] = over ? 1 : 0;
            __syncwarp();
            if (lane == 0) mbar_expect_tx(bar_b(s), over ? 0u : (uint32_t)total * 8u);
            __syncwarp();
            if (!over && cnt > 0) {
                tma_bulk_g2s(smem_u32(st + L.col_off) + (uint32_t)base * 4u, p.col + a_lo, (uint32_t)cnt * 4u, bar_b(s));
                tma_bulk_g2s(smem_u32(st + L.val_off) + (uint32_t)base * 4u, p.val + a_lo, (uint32_t)cnt * 4u, bar_b(s));
            }
        };
        for (int j = 0; j < p.AH && j < n_my; ++j) hop_a(j);
        for (int j = 0; j < n_my; ++j) {
            if (j + p.AH < n_my) hop_a(j + p.AH);
            hop_b(j);
        }
        return;
    }

    // ---------------------------------------------------------------------- consumer warps
    float* Es = reinterpret_cast<float*>(smem);
    unsigned char* wbase = smem + L.warp_off + (uint32_t)warp * L.warp_bytes;
    float* acc = reinterpret_cast<float*>(wbase);
    int* lc = reinterpret_cast<int*>(wbase + L.acc_bytes);
    float* lv = reinterpret_cast<float*>(wbase + L.acc_bytes + (uint32_t)p.LCAP * 4u);
    for (int i = lane; i < (int)(L.acc_bytes / 4u); i += 32) acc[i] = 0.f;
    __syncwarp();
    const int grp = lane >> 3, gl = lane & 7;
    const int n_cons = p.n_warps * 32;
    int cur_c = -1, c_hint = 0;
    uint32_t e_phase = 0;
    for (int j = 0; j < n_my; ++j) {
        const int s = j % NS, use = j / NS;
        const GstItem it = gst_decode(p, i0 + j, c_hint);
        if (it.c != cur_c) {                               // (block-uniform) new chromosome: its E table -> shared memory
            asm volatile("bar.sync 1, %0;" ::"r"(n_cons) : "memory");
            if (warp == 0 && lane == 0) {
                mbar_expect_tx(bar_e, (uint32_t)it.nc * 256u);
                tma_bulk_g2s(sbase, p.E_bin + (int64_t)it.off * 64, (uint32_t)it.nc * 256u, bar_e);
            }
            mbar_wait_wd(bar_e, e_phase);
            e_phase ^= 1u;
            cur_c = it.c;
        }
        mbar_wait_wd(bar_b(s), (uint32_t)use & 1u);
        const unsigned char* st = smem + L.stage_off + (uint32_t)s * L.stage_bytes;
        const int over_item = reinterpret_cast<const int*>(st + 128)[0];
        const int* scol = reinterpret_cast<const int*>(st + L.col_off);
        const float* sval = reinterpret_cast<const float*>(st + L.val_off);
        // one token with the whole warp (tokens with more than 16 entries, or a small candidate list)
        auto slow_token = [&](const int t) {
            const int tok = it.cb * p.n_bins + it.off + it.l0 + t;
            // ---- extents of the k1 rows of this token (lane n < k1)
            int len = 0, idx0 = 0;
            float w = 1.0f;
            if (lane < p.k1) {
                const int64_t* rp = reinterpret_cast<const int64_t*>(st + L.meta_bytes) + (size_t)lane * (size_t)(p.TB + 4);
                const int d = reinterpret_cast<const int*>(st + 96)[lane];
                const int64_t e_lo = rp[d + t], e_hi = rp[d + t + 1];
                len = (int)(e_hi - e_lo);
                idx0 = (int)(e_lo + reinterpret_cast<const int64_t*>(st)[lane]);
                w = reinterpret_cast<const float*>(st + 64)[lane];
            }
            int end = len;
#pragma unroll
            for (int o = 1; o < GST_MAX_K1; o <<= 1) {
                const int x = __shfl_up_sync(0xffffffffu, end, o, GST_MAX_K1);
                if ((lane & (GST_MAX_K1 - 1)) >= o) end += x;
            }
            const int m = __shfl_sync(0xffffffffu, end, p.k1 - 1);
            float4* dst = reinterpret_cast<float4*>(p.neigh + (int64_t)tok * 64);
            if (m == 0) {                                  // empty rows: A_hat row = 0
                if (lane < 16) dst[lane] = make_float4(0.f, 0.f, 0.f, 0.f);
                return;
            }
            int acount = 0;
            float part = 0.f;
            if (m <= 32) {
                // ---- one entry per lane in row-major order
                int r = 0;
                for (int q = 0; q + 1 < p.k1; ++q) r += (lane >= __shfl_sync(0xffffffffu, end, q)) ? 1 : 0;
                const int start_r = __shfl_sync(0xffffffffu, end - len, r);
                const int idx = __shfl_sync(0xffffffffu, idx0, r) + (lane - start_r);
                const float wr = __shfl_sync(0xffffffffu, w, r);
                const bool valid = lane < m;
                const int ccol = valid ? scol[idx] : -1 - lane;
                float v = valid ? sval[idx] : 0.f;
                if (p.weighted) v = __fmul_rn(wr, v);
                const unsigned peers = __match_any_sync(0xffffffffu, ccol);
                const bool own = valid && (__ffs(peers) - 1 == lane);
                unsigned mm = peers & (peers - 1);                   // peers after the first occurrence, in lane (= neighbour) order
                float sum = __shfl_sync(0xffffffffu, v, __ffs(peers) - 1);
                for (int q = 1; q < p.k1; ++q) {
                    const int src = mm ? __ffs(mm) - 1 : lane;
                    const float x = __shfl_sync(0xffffffffu, v, src);
                    if (mm) sum = __fadd_rn(sum, x);
                    mm &= mm - 1;
                }
                const float tv = own ? log1pf(__fmul_rn(sum, HSG_PDF0)) : 0.f;      // x pdf(0), log1p (Impute.py:19, 91)
                const unsigned ob = __ballot_sync(0xffffffffu, own);
                const int pos = __popc(ob & ((1u << lane) - 1u));
                if (own && pos < p.LCAP) { lc[pos] = ccol; lv[pos] = tv; }
                acount = __popc(ob);
                part = fabsf(tv);
            } else {
                // ---- dense accumulator, rows one after the other (neighbour order), then first-visitor ownership
                for (int n = 0; n < p.k1; ++n) {
                    const int ln = __shfl_sync(0xffffffffu, len, n), i0n = __shfl_sync(0xffffffffu, idx0, n);
                    const float wn = __shfl_sync(0xffffffffu, w, n);
                    for (int e = lane; e < ln; e += 32) {
                        HSG_DASSERT(i0n + e >= 0 && i0n + e < p.CAP);
                        const int ccol = scol[i0n + e];
                        const float v = sval[i0n + e];
                        HSG_DASSERT(ccol >= 0 && ccol < it.nc);
                        acc[ccol] = __fadd_rn(acc[ccol], p.weighted ? __fmul_rn(wn, v) : v);
                    }
                    __syncwarp();
                }
                for (int n = 0; n < p.k1; ++n) {
                    const int ln = __shfl_sync(0xffffffffu, len, n), i0n = __shfl_sync(0xffffffffu, idx0, n);
                    for (int e0 = 0; e0 < ln; e0 += 32) {
                        const int e = e0 + lane;
                        int ccol = 0;
                        float sum = 0.f;
                        if (e < ln) { ccol = scol[i0n + e]; sum = atomicExch(&acc[ccol], 0.f); }
                        const bool own = sum != 0.f;
                        const float tv = own ? log1pf(__fmul_rn(sum, HSG_PDF0)) : 0.f;
                        const unsigned ob = __ballot_sync(0xffffffffu, own);
                        const int pos = acount + __popc(ob & ((1u << lane) - 1u));
                        if (own && pos < p.LCAP) { lc[pos] = ccol; lv[pos] = tv; }
                        acount += __popc(ob);
                        part += fabsf(tv);
                    }
                }
            }
            if (acount > p.LCAP) {                         // more distinct columns than the list holds: the generic kernel finishes it
                if (lane == 0) p.tok_list[atomicAdd(p.tok_count, 1)] = tok;
                __syncwarp();
                return;
            }
            const float total = warp_sum(part);
            __syncwarp();
            if (total > 0.f)
                for (int i = lane; i < acount; i += 32) lv[i] = __fdiv_rn(lv[i], total);           // L1 normalisation (Impute.py:92)
            __syncwarp();
            // ---- neigh = sum_t A_hat[b, t] E_c[t, :] from shared memory: group g takes candidates g, g + 4, ...; lane gl owns
            //      columns 4 gl .. 4 gl + 3 and 32 + 4 gl .. 32 + 4 gl + 3 (each 128-bit load covers 128 contiguous bytes per group)
            float4 o0 = make_float4(0.f, 0.f, 0.f, 0.f), o1 = o0;
            const float4* E4 = reinterpret_cast<const float4*>(Es) + gl;
#pragma unroll 2
            for (int i = grp; i < acount; i += 4) {
                const int ci = lc[i];
                const float a = lv[i];
                const float4 x0 = E4[ci * 16], x1 = E4[ci * 16 + 8];
                o0.x = fmaf(a, x0.x, o0.x); o0.y = fmaf(a, x0.y, o0.y); o0.z = fmaf(a, x0.z, o0.z); o0.w = fmaf(a, x0.w, o0.w);
                o1.x = fmaf(a, x1.x, o1.x); o1.y = fmaf(a, x1.y, o1.y); o1.z = fmaf(a, x1.z, o1.z); o1.w = fmaf(a, x1.w, o1.w);
            }
#pragma unroll
            for (int sh = 8; sh <= 16; sh <<= 1) {
                o0.x += __shfl_xor_sync(0xffffffffu, o0.x, sh); o0.y += __shfl_xor_sync(0xffffffffu, o0.y, sh);
                o0.z += __shfl_xor_sync(0xffffffffu, o0.z, sh); o0.w += __shfl_xor_sync(0xffffffffu, o0.w, sh);
                o1.x += __shfl_xor_sync(0xffffffffu, o1.x, sh); o1.y += __shfl_xor_sync(0xffffffffu, o1.y, sh);
                o1.z += __shfl_xor_sync(0xffffffffu, o1.z, sh); o1.w += __shfl_xor_sync(0xffffffffu, o1.w, sh);
            }
            if (grp == 0) dst[gl] = o0;
            else if (grp == 1) dst[8 + gl] = o1;
            __syncwarp();
        };

        if (over_item) {
            for (int t = warp; t < it.nt; t += p.n_warps)
                if (lane == 0) p.tok_list[atomicAdd(p.tok_count, 1)] = it.cb * p.n_bins + it.off + it.l0 + t;
        } else {
            // TWO tokens per pass, 16 lanes each (the sparse regime: a token's k1 rows hold a handful of entries): every
            // instruction below serves both tokens; a pair with a token of more than 16 entries takes slow_token() twice
            const int half = lane >> 4, le = lane & 15, hg = le >> 3;
            for (int tp = 2 * warp; tp < it.nt; tp += 2 * p.n_warps) {
                const int t = tp + half;
                const bool tvalid = t < it.nt;
                int len = 0, idx0 = 0;
                float w = 1.0f;
                if (le < p.k1 && tvalid) {
                    const int64_t* rp = reinterpret_cast<const int64_t*>(st + L.meta_bytes) + (size_t)le * (size_t)(p.TB + 4);
                    const int d = reinterpret_cast<const int*>(st + 96)[le];
                    const int64_t e_lo = rp[d + t], e_hi = rp[d + t + 1];
                    len = (int)(e_hi - e_lo);
                    idx0 = (int)(e_lo + reinterpret_cast<const int64_t*>(st)[le]);
                    w = reinterpret_cast<const float*>(st + 64)[le];
                }
                int end = len;
#pragma unroll
                for (int o = 1; o < GST_MAX_K1; o <<= 1) {
                    const int x = __shfl_up_sync(0xffffffffu, end, o, GST_MAX_K1);
                    if ((lane & (GST_MAX_K1 - 1)) >= o) end += x;
                }
                const int m = __shfl_sync(0xffffffffu, end, p.k1 - 1, 16);                 // entries of this half's token
                if (__any_sync(0xffffffffu, m > 16) || p.LCAP < 32) {
                    slow_token(tp);
                    if (tp + 1 < it.nt) slow_token(tp + 1);
                    continue;
                }
                int r = 0;
                for (int q = 0; q + 1 < p.k1; ++q) r += (le >= __shfl_sync(0xffffffffu, end, q, 16)) ? 1 : 0;
                const int start_r = __shfl_sync(0xffffffffu, end - len, r, 16);
                const int idx = __shfl_sync(0xffffffffu, idx0, r, 16) + (le - start_r);
                const float wr = __shfl_sync(0xffffffffu, w, r, 16);
                const bool valid = le < m;
                HSG_DASSERT(!valid || (idx >= 0 && idx < p.CAP));
                const int ccol = valid ? scol[idx] : 0;
                HSG_DASSERT(ccol >= 0 && ccol < it.nc);
                float v = valid ? sval[idx] : 0.f;
                if (p.weighted) v = __fmul_rn(wr, v);
                const unsigned peers = __match_any_sync(0xffffffffu, valid ? (ccol | (half << 24)) : (-1 - lane));
                const int first = __ffs(peers) - 1;
                const bool own = valid && first == lane;
                unsigned mm = peers & (peers - 1);                   // peers after the first occurrence, in lane (= neighbour) order
                float sum = __shfl_sync(0xffffffffu, v, first);
                for (int q = 1; q < p.k1; ++q) {
                    const int src = mm ? __ffs(mm) - 1 : lane;
                    const float x = __shfl_sync(0xffffffffu, v, src);
                    if (mm) sum = __fadd_rn(sum, x);
                    mm &= mm - 1;
                }
                float tv = own ? log1pf(__fmul_rn(sum, HSG_PDF0)) : 0.f;              // x pdf(0), log1p (Impute.py:19, 91)
                float total = fabsf(tv);
#pragma unroll
                for (int o = 8; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
                if (total > 0.f) tv = __fdiv_rn(tv, total);                          // L1 normalisation (Impute.py:92)
                const unsigned ob = (__ballot_sync(0xffffffffu, own) >> (half * 16)) & 0xffffu;
                if (own) { const int pos = half * 16 + __popc(ob & ((1u << le) - 1u)); lc[pos] = ccol; lv[pos] = tv; }
                const int acount = __popc(ob);
                const int amax = max(acount, __shfl_xor_sync(0xffffffffu, acount, 16));
                __syncwarp();
                // ---- gather: the two 8-lane groups of a half take the token's candidates alternately; lane gl owns columns
                //      4 gl .. 4 gl + 3 and 32 + 4 gl .. 32 + 4 gl + 3
                float4 o0 = make_float4(0.f, 0.f, 0.f, 0.f), o1 = o0;
                const float4* E4 = reinterpret_cast<const float4*>(Es) + gl;
                for (int i = hg; i < amax; i += 2) {
                    if (i < acount) {
                        const int ci = lc[half * 16 + i];
                        const float a = lv[half * 16 + i];
                        const float4 x0 = E4[ci * 16], x1 = E4[ci * 16 + 8];
                        o0.x = fmaf(a, x0.x, o0.x); o0.y = fmaf(a, x0.y, o0.y); o0.z = fmaf(a, x0.z, o0.z); o0.w = fmaf(a, x0.w, o0.w);
                        o1.x = fmaf(a, x1.x, o1.x); o1.y = fmaf(a, x1.y, o1.y); o1.z = fmaf(a, x1.z, o1.z); o1.w = fmaf(a, x1.w, o1.w);
                    }
                }
                o0.x += __shfl_xor_sync(0xffffffffu, o0.x, 8); o0.y += __shfl_xor_sync(0xffffffffu, o0.y, 8);
                o0.z += __shfl_xor_sync(0xffffffffu, o0.z, 8); o0.w += __shfl_xor_sync(0xffffffffu, o0.w, 8);
                o1.x += __shfl_xor_sync(0xffffffffu, o1.x, 8); o1.y += __shfl_xor_sync(0xffffffffu, o1.y, 8);
                o1.z += __shfl_xor_sync(0xffffffffu, o1.z, 8); o1.w += __shfl_xor_sync(0xffffffffu, o1.w, 8);
                if (tvalid) {
                    float4* dst = reinterpret_cast<float4*>(p.neigh + (int64_t)(it.cb * p.n_bins + it.off + it.l0 + t) * 64);
                    if (hg == 0) dst[gl] = o0;
                    else dst[8 + gl] = o1;
                }
                __syncwarp();
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_empty(s));
    }
}

// returns 0 and sets *done when the stream kernel was launched (tokens it could not finish are on c->ovf_list, *n_list_tokens
// capacity already allocated by the caller)
int try_launch_gather_stream(hsg_ctx* c, int cell_start, int n_batch, int max_nc, int* cnt, cudaStream_t s, bool* done) {
    *done = false;
    if (c->d != 64 || c->ltr != 0 || c->k1 > GST_MAX_K1 || c->n_chrom > GST_MAX_CHROM || c->n_rows <= 0) return 0;
    if ((int64_t)n_batch * c->n_bins >= (1LL << 31)) return 0;
    const double avg_row = (double)c->nnz / (double)c->n_rows;
    const double avg_cand = (double)c->k1 * avg_row;
    const uint32_t budget = 226u * 1024u;
    GatherStreamParams p;
    int best_tb = 0;
    // consumer warps: as many as fit (the merge of a token is a chain of dependent shared-memory / shuffle steps: warps, not
    // issue slots, are what a scheduler runs out of); tile: the largest whose stages fit next to the E table
    // (31 + the producer = 1024 threads; fewer than 16 -- long chromosomes with dense rows, whose E table leaves little room
    // -- and the hash-merge kernel with its 48 warps per SM is faster: measured on C3, 5.4 ms with 8 warps against 3.4 ms)
    // ring depth: the consumers finish an item of the sparse regime in about a microsecond, a hop takes a few -- so the ring is
    // as deep as shared memory allows (<= 12 stages of ONE consumer pass each), hop A running half a ring ahead of hop B
    static const int nw_opts[2] = {31, 16};
    for (int o = 0; o < 2 && !best_tb; ++o) {
        const int nw = nw_opts[o];
        const int lcap = c->gather_stream_lcap > 0 ? c->gather_stream_lcap : (int)std::min(512.0, std::max(64.0, 32.0 * std::ceil(4.0 * avg_cand / 32.0)));
        const int tb = 2 * nw;
        double worst = 0.0;                               // expected entries of an item on the densest chromosome
        for (int j = 0; j < c->n_chrom; ++j) {
            const double a = (size_t)j < c->chrom_row_avg.size() ? c->chrom_row_avg[(size_t)j] : avg_row;
            worst = std::max(worst, a * c->k1 * std::min(tb, c->bins[j]));
        }
        const int cap = c->gather_stream_cap > 0 ? (c->gather_stream_cap & ~3) : ((int)(2.0 * worst) + 128 + 3) & ~3;
        for (int ns = 12; ns >= 4; --ns) {
            const GstLayout L = gst_layout(max_nc, c->k1, tb, cap, lcap, nw, ns);
            if (L.total <= budget) { best_tb = tb; p.TB = tb; p.CAP = cap; p.LCAP = lcap; p.n_warps = nw; p.NS = ns; p.AH = ns / 2; break; }
        }
    }
    if (!best_tb) return 0;
    p.row_ptr = c->row_ptr; p.col = c->col; p.val = c->val; p.nbr = c->nbr; p.nbr_w = c->nbr_w; p.E_bin = c->E_bin;
    p.bin_off = c->d_bin_off; p.neigh = c->neigh; p.tok_list = c->ovf_list; p.tok_count = cnt;
    p.k1 = c->k1; p.weighted = c->weighted ? 1 : 0; p.n_bins = c->n_bins; p.cell_start = cell_start; p.n_batch = n_batch;
    p.n_chrom = c->n_chrom; p.max_nc = max_nc;
    long long first = 0;
    for (int j = 0; j < c->n_chrom; ++j) {
        p.item_first[j] = (int)first;
        p.tiles[j] = (c->bins[j] + p.TB - 1) / p.TB;
        first += (long long)n_batch * p.tiles[j];
    }
    if (first >= (1LL << 31) || first == 0) return 0;
    for (int j = c->n_chrom; j <= GST_MAX_CHROM; ++j) p.item_first[j] = (int)first;
    for (int j = c->n_chrom; j < GST_MAX_CHROM; ++j) p.tiles[j] = 1;
    p.n_items = (int)first;
    const GstLayout L = gst_layout(max_nc, c->k1, p.TB, p.CAP, p.LCAP, p.n_warps, p.NS);
    static bool attr_set[64] = {};
    if (!attr_set[c->device & 63]) {
        HSG_CUDA(cudaFuncSetAttribute(neigh_gather_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set[c->device & 63] = true;
    }
    // one block per SM; block b starts at the item that holds token b * n_tok / grid (items differ in size at chromosome ends)
    const int grid = (int)std::min<long long>(std::min<long long>(first, c->n_sm), GST_MAX_GRID);
    const long long n_tok = (long long)n_batch * c->n_bins;
    for (int b = 0; b <= grid; ++b) {
        if (b == grid) { p.blk_first[b] = (int)first; break; }
        const long long target = n_tok * b / grid;
        int j = 0;
        while (j + 1 < c->n_chrom && (long long)n_batch * c->bin_off[j + 1] <= target) ++j;
        const long long rem = target - (long long)n_batch * c->bin_off[j];
        const int cb = (int)(rem / c->bins[j]), l = (int)(rem % c->bins[j]);
        p.blk_first[b] = p.item_first[j] + cb * p.tiles[j] + l / p.TB;
    }
    neigh_gather_stream_kernel<<<grid, (p.n_warps + 1) * 32, L.total, s>>>(p);
    c->launches++;
    HSG_CUDA(cudaGetLastError());
    *done = true;
    return 0;
}
