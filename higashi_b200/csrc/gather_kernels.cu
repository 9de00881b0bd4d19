// gather_kernels.cu -- K1: node embedding of every (cell, bin) token of a batch of cells.
//
//   K1a neigh_gather_kernel  : weighted merge of the k+1 neighbour cells' CSR rows (+ optional row-shift
//                              smoothing), x pdf(0), log1p, L1 row normalisation, then the sparse gather
//                              A_hat[b,:] . E_bin   (Impute.py:78-121 prep_one + Modules.py:1364-1371 forward_GCN)
//   K1b token_project_kernel : dyn = swish(nn([neigh | E_self]))  (Modules.py:1480), Q = w_qs.LN1(dyn),
//                              K = w_ks.LN2(dyn)                   (Modules.py:1035-1044)
#include <cuda_fp16.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "common.cuh"

// ---------------------------------------------------------------------------------------------
// K1a.  One warp per (cell, bin) token.  Per-warp scratch in shared memory:
//   acc[max_nc] (+ tmp[max_nc] when local_transfer_range > 0), touched-column lists and bit masks.
// Values of one CSR row have unique columns, so the lanes of a warp never collide on acc[].
// ---------------------------------------------------------------------------------------------
struct GatherSmem {
    float* acc; float* tmp; unsigned* amask; unsigned* tmask; unsigned short* alist; unsigned short* tlist;
};

__host__ __device__ inline size_t gather_warp_bytes(int max_nc, int ltr) {
    size_t words = (max_nc + 31) / 32;
    size_t b = (size_t)max_nc * 4 + words * 4 + (((size_t)max_nc * 2 + 3) & ~(size_t)3);
    if (ltr > 0) b *= 2;
    return (b + 15) & ~(size_t)15;
}

__device__ __forceinline__ float norm_pdf_dev(float x) { return expf(-0.5f * x * x) * HSG_PDF0; }

// accumulate  T[col] (+)= w * val  over one CSR row, recording first-touched columns
__device__ __forceinline__ void merge_row(const int32_t* __restrict__ col, const float* __restrict__ val,
                                          int64_t lo, int64_t hi, float w, bool scale, float* T, unsigned* mask,
                                          unsigned short* list, int& count, int lane) {
    for (int64_t e0 = lo; e0 < hi; e0 += 32) {
        int64_t e = e0 + lane;
        bool valid = e < hi;
        bool is_new = false;
        int cidx = 0;
        if (valid) {
            cidx = __ldg(col + e);
            HSG_DASSERT(cidx >= 0 && (cidx >> 5) < 2048);          // chromosome-local column inside the per-warp accumulator
            float v = __ldg(val + e);
            float prod = scale ? __fmul_rn(w, v) : v;
            T[cidx] = __fadd_rn(T[cidx], prod);
            unsigned bit = 1u << (cidx & 31);
            unsigned old = atomicOr(&mask[cidx >> 5], bit);
            is_new = !(old & bit);
        }
        unsigned b = __ballot_sync(0xffffffffu, is_new);
        if (is_new) list[count + __popc(b & ((1u << lane) - 1u))] = (unsigned short)cidx;
        count += __popc(b);
    }
    __syncwarp();
}

template <int D>
__global__ void neigh_gather_kernel(const int64_t* __restrict__ row_ptr, const int32_t* __restrict__ col,
                                    const float* __restrict__ val, const int32_t* __restrict__ nbr,
                                    const float* __restrict__ nbr_w, int k1, int weighted,
                                    const float* __restrict__ E_bin, const int* __restrict__ bin_chrom,
                                    const int* __restrict__ bin_off, int n_bins, int cell_start,
                                    int n_batch_cells, int ltr, int max_nc, float* __restrict__ neigh,
                                    const int* __restrict__ tok_list, const int* __restrict__ tok_count) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const size_t wbytes = gather_warp_bytes(max_nc, ltr);
    unsigned char* base = smem_raw + (size_t)warp * wbytes;
    const int words = (max_nc + 31) / 32;
    const size_t lbytes = (((size_t)max_nc * 2 + 3) & ~(size_t)3);
    GatherSmem S;
    S.acc = (float*)base;
    S.amask = (unsigned*)(base + (size_t)max_nc * 4);
    S.alist = (unsigned short*)(base + (size_t)max_nc * 4 + (size_t)words * 4);
    unsigned char* b2 = base + (size_t)max_nc * 4 + (size_t)words * 4 + lbytes;
    S.tmp = (float*)b2;
    S.tmask = (unsigned*)(b2 + (size_t)max_nc * 4);
    S.tlist = (unsigned short*)(b2 + (size_t)max_nc * 4 + (size_t)words * 4);
    // list mode: only the tokens the sub-warp kernel could not hold in registers (tok_list / tok_count written by it)
    const int64_t n_tok = tok_list ? (int64_t)*tok_count : (int64_t)n_batch_cells * n_bins;
    if ((int64_t)blockIdx.x * wpb + warp >= n_tok) return;          // nothing for this warp (the usual case in list mode)
    for (int i = lane; i < max_nc; i += 32) { S.acc[i] = 0.f; if (ltr > 0) S.tmp[i] = 0.f; }
    for (int i = lane; i < words; i += 32) { S.amask[i] = 0u; if (ltr > 0) S.tmask[i] = 0u; }
    __syncwarp();

    for (int64_t ti = (int64_t)blockIdx.x * wpb + warp; ti < n_tok; ti += (int64_t)gridDim.x * wpb) {
        const int64_t tok = tok_list ? (int64_t)tok_list[ti] : ti;
        HSG_DASSERT(tok >= 0 && tok < (int64_t)n_batch_cells * n_bins);
        const int cb = (int)(tok / n_bins), g = (int)(tok - (int64_t)cb * n_bins);
        const int cell = cell_start + cb;
        const int c = bin_chrom[g], off = bin_off[c], nc = bin_off[c + 1] - off, l = g - off;
        int acount = 0;
        if (ltr == 0) {
            for (int nb = 0; nb < k1; ++nb) {
                int src = weighted ? nbr[(int64_t)cell * k1 + nb] : cell;
                float w = weighted ? nbr_w[(int64_t)cell * k1 + nb] : 1.0f;
                int64_t row = (int64_t)src * n_bins + off + l;
                merge_row(col, val, row_ptr[row], row_ptr[row + 1], w, weighted != 0, S.acc, S.amask, S.alist,
                          acount, lane);
            }
            for (int i = lane; i < acount; i += 32) {           // adj * norm.pdf(0)   (Impute.py:19)
                int ci = S.alist[i];
                S.acc[ci] = __fmul_rn(S.acc[ci], HSG_PDF0);
            }
            __syncwarp();
        } else {
            // adj = M*pdf(0) + sum_i (M[row+i+1] + M[row-i-1]) * pdf((i+1)/range), rows clamped (Impute.py:16-26)
            for (int s = -2 * ltr; s <= 2 * ltr; ++s) {
                int as = s < 0 ? -s : s;
                float pdf = (s == 0) ? HSG_PDF0 : norm_pdf_dev((float)as / (float)ltr);
                int sl = min(max(l + s, 0), nc - 1);
                int tcount = 0;
                for (int nb = 0; nb < k1; ++nb) {
                    int src = weighted ? nbr[(int64_t)cell * k1 + nb] : cell;
                    float w = weighted ? nbr_w[(int64_t)cell * k1 + nb] : 1.0f;
                    int64_t row = (int64_t)src * n_bins + off + sl;
                    merge_row(col, val, row_ptr[row], row_ptr[row + 1], w, weighted != 0, S.tmp, S.tmask, S.tlist,
                              tcount, lane);
                }
                for (int i0 = 0; i0 < tcount; i0 += 32) {
                    int i = i0 + lane;
                    bool valid = i < tcount, is_new = false;
                    int ci = 0;
                    if (valid) {
                        ci = S.tlist[i];
                        S.acc[ci] = __fadd_rn(S.acc[ci], __fmul_rn(S.tmp[ci], pdf));
                        S.tmp[ci] = 0.f;
                        unsigned bit = 1u << (ci & 31);
                        unsigned old = atomicOr(&S.amask[ci >> 5], bit);
                        is_new = !(old & bit);
                        S.tmask[ci >> 5] = 0u;
                    }
                    unsigned b = __ballot_sync(0xffffffffu, is_new);
                    if (is_new) S.alist[acount + __popc(b & ((1u << lane) - 1u))] = (unsigned short)ci;
                    acount += __popc(b);
                }
                __syncwarp();
            }
        }
        // log1p + L1 normalisation (Impute.py:91-92)
        float part = 0.f;
        for (int i = lane; i < acount; i += 32) {
            int ci = S.alist[i];
            float v = log1pf(S.acc[ci]);
            S.acc[ci] = v;
            part += fabsf(v);
        }
        float total = warp_sum(part);
        __syncwarp();
        // sparse gather: neigh = sum_t A_hat[b,t] * E_bin[t,:]   (lanes split the d columns)
        constexpr int VPL = D / 32;
        float out[VPL];
#pragma unroll
        for (int q = 0; q < VPL; ++q) out[q] = 0.f;
        const float* Ec = E_bin + (int64_t)off * D;
        for (int i = 0; i < acount; ++i) {
            int ci = S.alist[i];
            float a = (total > 0.f) ? __fdiv_rn(S.acc[ci], total) : S.acc[ci];
            const float* er = Ec + (int64_t)ci * D + lane * VPL;
            if (VPL == 2) {
                float2 ev = __ldg(reinterpret_cast<const float2*>(er));
                out[0] = fmaf(a, ev.x, out[0]); out[1] = fmaf(a, ev.y, out[1]);
            } else {
                float4 ev = __ldg(reinterpret_cast<const float4*>(er));
                out[0] = fmaf(a, ev.x, out[0]); out[1] = fmaf(a, ev.y, out[1]);
                out[2 % VPL] = fmaf(a, ev.z, out[2 % VPL]); out[3 % VPL] = fmaf(a, ev.w, out[3 % VPL]);
            }
        }
        float* dst = neigh + tok * D + lane * VPL;
        if (VPL == 2) *reinterpret_cast<float2*>(dst) = make_float2(out[0], out[1]);
        else *reinterpret_cast<float4*>(dst) = make_float4(out[0], out[1], out[2 % VPL], out[3 % VPL]);
        __syncwarp();
        for (int i = lane; i < acount; i += 32) {            // reset scratch for the next token
            int ci = S.alist[i];
            S.acc[ci] = 0.f;
            S.amask[ci >> 5] = 0u;
        }
        __syncwarp();
    }
}


// ---------------------------------------------------------------------------------------------
// K1a, sparse rows (the scHi-C regime: a handful of contacts per row).  EIGHT lanes per (cell, bin) token, four tokens per
// warp, everything in registers:
//   * lanes 0..k1-1 of a group fetch the k+1 neighbour rows' extents together (one dependent round trip instead of k+1),
//     then push their entries, already scaled by the neighbour weight, into a 32-entry staging strip in row order;
//   * duplicates across neighbour rows are merged with an all-pairs compare over shuffles, every sum taken in row order
//     (same order as the reference's sequential sparse adds, deterministic);
//   * x pdf(0), log1p, L1 normalisation and the gather of the E rows (each lane owns 8 of the 64 output columns: two 128-bit
//     loads per entry, one contiguous 256-byte row per group).
// Tokens with more than 8 SLOTS merged candidates go to an overflow list that the generic warp-per-token kernel finishes.
// ---------------------------------------------------------------------------------------------
// SLOTS x 8 candidates per token are held in registers (SLOTS slots per lane): 4 (32 candidates) for the sparsest data, 8 / 12
// (64 / 96) when the merged rows are longer -- the all-pairs dedup costs SLOTS^2, so the launcher picks the smallest that fits
#define GS_WPB 8

template <int SLOTS, int S>
__device__ __forceinline__ void gs_dedup(const int (&col)[SLOTS], const float (&val)[SLOTS], int gl, float (&sum)[SLOTS],
                                         bool (&dup)[SLOTS]) {
#pragma unroll
    for (int s = 0; s < SLOTS; ++s) { sum[s] = 0.f; dup[s] = false; }
#pragma unroll
    for (int s2 = 0; s2 < S; ++s2) {
#pragma unroll
        for (int l2 = 0; l2 < 8; ++l2) {
            const int c2 = __shfl_sync(0xffffffffu, col[s2], l2, 8);
            const float v2 = __shfl_sync(0xffffffffu, val[s2], l2, 8);
            const int e2 = s2 * 8 + l2;
#pragma unroll
            for (int s = 0; s < S; ++s) {
                const int e = s * 8 + gl;
                const bool same = (c2 == col[s]);
                if (same && e2 >= e) sum[s] = __fadd_rn(sum[s], v2);     // owner accumulates in row order (itself first)
                if (same && e2 < e) dup[s] = true;
            }
        }
    }
}

// compile-time dispatch on the number of slots actually in use by this warp's four tokens
template <int SLOTS, int S>
__device__ __forceinline__ void gs_dedup_dispatch(int nmax, const int (&col)[SLOTS], const float (&val)[SLOTS], int gl,
                                                  float (&sum)[SLOTS], bool (&dup)[SLOTS]) {
    if constexpr (S >= SLOTS) {
        gs_dedup<SLOTS, SLOTS>(col, val, gl, sum, dup);
    } else {
        if (nmax <= S * 8) gs_dedup<SLOTS, S>(col, val, gl, sum, dup);
        else gs_dedup_dispatch<SLOTS, S + 1>(nmax, col, val, gl, sum, dup);
    }
}

template <int SLOTS, int D>
__global__ void __launch_bounds__(GS_WPB * 32) neigh_gather_sparse_kernel(
    const int64_t* __restrict__ row_ptr, const int32_t* __restrict__ col, const float* __restrict__ val,
    const int32_t* __restrict__ nbr, const float* __restrict__ nbr_w, int k1, int weighted, const float* __restrict__ E_bin,
    const int* __restrict__ bin_chrom, const int* __restrict__ bin_off, int n_bins, int cell_start, int n_tok,
    float* __restrict__ neigh, int* __restrict__ tok_list, int* __restrict__ tok_count) {
    constexpr int GS_CAP = SLOTS * 8;
    constexpr int GS_SLOTS = SLOTS;
    __shared__ int s_col[GS_WPB][4][GS_CAP];
    __shared__ float s_val[GS_WPB][4][GS_CAP];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, gl = lane & 7, grp = lane >> 3;
    const int n_quads = (n_tok + 3) >> 2;
    for (int quad = blockIdx.x * GS_WPB + warp; quad < n_quads; quad += gridDim.x * GS_WPB) {
        const int tok = quad * 4 + grp;
        const bool tvalid = tok < n_tok;
        const int tk = tvalid ? tok : n_tok - 1;
        const int cb = tk / n_bins, g = tk - cb * n_bins;
        const int cell = cell_start + cb;
        const int off = bin_off[bin_chrom[g]];
        // ---- extents of the k1 neighbour rows (lane gl < k1 owns row gl)
        int64_t lo = 0;
        int len = 0;
        float w = 1.0f;
        if (gl < k1) {
            int src = cell;
            if (weighted) { src = __ldg(nbr + (int64_t)cell * k1 + gl); w = __ldg(nbr_w + (int64_t)cell * k1 + gl); }
            const int64_t row = (int64_t)src * n_bins + g;
            lo = __ldg(row_ptr + row);
            len = (int)(__ldg(row_ptr + row + 1) - lo);
        }
        int start = len;                         // inclusive prefix over the 8 lanes of the group
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, start, o, 8);
            if (gl >= o) start += t;
        }
        const int n = __shfl_sync(0xffffffffu, start, 7, 8);
        start -= len;
        const bool over = n > GS_CAP;
        if (over && gl == 0 && tvalid) { const int slot = atomicAdd(tok_count, 1); HSG_DASSERT(slot >= 0 && slot < n_tok); tok_list[slot] = tok; }
        // ---- push the (weighted) entries into the staging strip in row order
        if (!over) {
            for (int i = 0; i < len; ++i) {
                const float v = __ldg(val + lo + i);
                s_col[warp][grp][start + i] = __ldg(col + lo + i);
                s_val[warp][grp][start + i] = weighted ? __fmul_rn(w, v) : v;
            }
        }
        __syncwarp();
        const int n_eff = over ? 0 : n;
        int nmax = n_eff;
        nmax = max(nmax, __shfl_xor_sync(0xffffffffu, nmax, 8));
        nmax = max(nmax, __shfl_xor_sync(0xffffffffu, nmax, 16));
        int ec[GS_SLOTS];
        float ev[GS_SLOTS];
#pragma unroll
        for (int s = 0; s < GS_SLOTS; ++s) {
            const int e = s * 8 + gl;
            const bool ok = e < n_eff;
            ec[s] = ok ? s_col[warp][grp][e] : -1 - e;           // distinct negative keys: never equal to a real column
            ev[s] = ok ? s_val[warp][grp][e] : 0.f;
        }
        __syncwarp();
        float sum[GS_SLOTS];
        bool dup[GS_SLOTS];
        gs_dedup_dispatch<SLOTS, 1>(nmax, ec, ev, gl, sum, dup);
        // ---- x pdf(0), log1p, L1 normalisation (Impute.py:19, 91-92)
        float part = 0.f;
#pragma unroll
        for (int s = 0; s < GS_SLOTS; ++s) {
            const bool own = (s * 8 + gl < n_eff) && !dup[s];
            const float v = own ? log1pf(__fmul_rn(sum[s], HSG_PDF0)) : 0.f;
            sum[s] = v;
            part += fabsf(v);
        }
        part += __shfl_xor_sync(0xffffffffu, part, 1, 8);
        part += __shfl_xor_sync(0xffffffffu, part, 2, 8);
        part += __shfl_xor_sync(0xffffffffu, part, 4, 8);
        if (part > 0.f) {
#pragma unroll
            for (int s = 0; s < GS_SLOTS; ++s) sum[s] = __fdiv_rn(sum[s], part);
        }
        // ---- neigh = sum_t A_hat[b,t] * E_bin[t,:] : lane gl owns columns [8 gl, 8 gl + 8) of every 64-column block
        constexpr int NB = D / 64;
        float4 o0[NB], o1[NB];
#pragma unroll
        for (int b = 0; b < NB; ++b) { o0[b] = make_float4(0.f, 0.f, 0.f, 0.f); o1[b] = o0[b]; }
        const float4* Eg = reinterpret_cast<const float4*>(E_bin + (int64_t)off * D) + gl * 2;
#pragma unroll
        for (int s = 0; s < GS_SLOTS; ++s) {
            if (s * 8 < nmax) {
#pragma unroll
                for (int l2 = 0; l2 < 8; ++l2) {
                    const float a = __shfl_sync(0xffffffffu, sum[s], l2, 8);
                    const int c = __shfl_sync(0xffffffffu, ec[s], l2, 8);
                    if (a != 0.f) {
#pragma unroll
                        for (int b = 0; b < NB; ++b) {
                            const float4 x0 = __ldg(Eg + (int64_t)c * (D / 4) + b * 16), x1 = __ldg(Eg + (int64_t)c * (D / 4) + b * 16 + 1);
                            o0[b].x = fmaf(a, x0.x, o0[b].x); o0[b].y = fmaf(a, x0.y, o0[b].y); o0[b].z = fmaf(a, x0.z, o0[b].z); o0[b].w = fmaf(a, x0.w, o0[b].w);
                            o1[b].x = fmaf(a, x1.x, o1[b].x); o1[b].y = fmaf(a, x1.y, o1[b].y); o1[b].z = fmaf(a, x1.z, o1[b].z); o1[b].w = fmaf(a, x1.w, o1[b].w);
                        }
                    }
                }
            }
        }
        if (tvalid && !over) {
            float4* dst = reinterpret_cast<float4*>(neigh + (int64_t)tok * D) + gl * 2;
#pragma unroll
            for (int b = 0; b < NB; ++b) { dst[b * 16] = o0[b]; dst[b * 16 + 1] = o1[b]; }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K1a, dense short-chromosome regime (C3 / C4 shapes: tens of merged candidates per row, chromosomes of a few hundred bins).
// The generic kernel above spends its time on the E-row gather: ~60 candidates x 256 B per token out of L2 (22 GB of L2
// traffic per 250 cells of C3) with one warp iteration per candidate.  Here a block works on ONE chromosome:
//   * the chromosome's E table [n_c x 64] fp32 (<= 128 KB) is staged in shared memory once per block, the gather reads it from there;
//   * the k+1 row extents of a token are fetched together (lanes 0..k) one token ahead, and the first 32 entries of every
//     neighbour row are requested before any of them is accumulated: one exposed round trip per token (col / val) instead of
//     2 (k+1) dependent ones;
//   * the gather runs 4 candidates at a time: 8 lanes per candidate, 8 columns per lane (two 128-bit shared-memory loads per
//     8 FMAs), the 4 partial sums are combined with shuffles at the end.
// The merge itself is the generic kernel's (row by row in neighbour order, fp32, bit-identical A_hat values); only the order of
// the 64-column dot-product sum changes (4 interleaved partial sums), well inside the fp32-mode tolerance.
// Blocks are assigned to chromosomes in proportion to n_c (n_c + 400) (rows x (fixed per-token cost + candidates per row)).
// ---------------------------------------------------------------------------------------------
#define GC_WARPS 24
#define GC_MAX_CHROM 64
struct GatherChromParams {
    const int64_t* row_ptr; const int32_t* col; const float* val; const int32_t* nbr; const float* nbr_w;
    const float* E_bin; const int* bin_off; float* neigh;
    int k1, weighted, n_bins, cell_start, n_batch, max_nc, n_chrom;
    int blk_first[GC_MAX_CHROM + 1];       // blocks [blk_first[c], blk_first[c + 1]) work on chromosome c
};

__global__ void __launch_bounds__(GC_WARPS * 32, 1) neigh_gather_chrom_kernel(GatherChromParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int c = 0;
    while (c + 1 < p.n_chrom && (int)blockIdx.x >= p.blk_first[c + 1]) ++c;
    const int jb = blockIdx.x - p.blk_first[c], nb_c = p.blk_first[c + 1] - p.blk_first[c];
    const int off = p.bin_off[c], nc = p.bin_off[c + 1] - off;
    // shared memory: E_c [max_nc][64] fp32, then the per-warp merge scratch
    float* Es = reinterpret_cast<float*>(smem_raw);
    const size_t wbytes = gather_warp_bytes(p.max_nc, 0);
    unsigned char* base = smem_raw + (size_t)p.max_nc * 64 * sizeof(float) + (size_t)warp * wbytes;
    const int words = (p.max_nc + 31) / 32;
    float* acc = (float*)base;
    unsigned* amask = (unsigned*)(base + (size_t)p.max_nc * 4);
    unsigned short* alist = (unsigned short*)(base + (size_t)p.max_nc * 4 + (size_t)words * 4);
    {
        const float4* src = reinterpret_cast<const float4*>(p.E_bin + (int64_t)off * 64);
        float4* dst = reinterpret_cast<float4*>(Es);
        for (int i = threadIdx.x; i < nc * 16; i += blockDim.x) dst[i] = __ldg(src + i);
    }
    for (int i = lane; i < p.max_nc; i += 32) acc[i] = 0.f;
    for (int i = lane; i < words; i += 32) amask[i] = 0u;
    __syncthreads();

    const int grp = lane >> 3, gl = lane & 7;
    const int64_t n_tok = (int64_t)p.n_batch * nc;
    // lanes 0..k: neighbour id, weight and row extent of a token, all in flight together; the extents of the warp's NEXT token
    // are requested at the top of the current token's work, so only the col / val round trip is exposed per token
    auto fetch_extents = [&](int64_t t, int64_t& lo, int64_t& hi, float& w) {
        lo = 0; hi = 0; w = 1.0f;
        if (lane < p.k1 && t < n_tok) {
            const int cb = (int)(t / nc), l = (int)(t - (int64_t)cb * nc);
            const int cell = p.cell_start + cb;
            const int src = p.weighted ? __ldg(p.nbr + (int64_t)cell * p.k1 + lane) : cell;
            if (p.weighted) w = __ldg(p.nbr_w + (int64_t)cell * p.k1 + lane);
            const int64_t row = (int64_t)src * p.n_bins + off + l;
            lo = __ldg(p.row_ptr + row); hi = __ldg(p.row_ptr + row + 1);
        }
    };
    const int64_t t_first = (int64_t)jb * GC_WARPS + warp, t_step = (int64_t)nb_c * GC_WARPS;
    int64_t lo_n, hi_n;
    float w_n;
    fetch_extents(t_first, lo_n, hi_n, w_n);
    for (int64_t t = t_first; t < n_tok; t += t_step) {
        const int cb = (int)(t / nc), l = (int)(t - (int64_t)cb * nc);
        const int64_t lo_l = lo_n, hi_l = hi_n;
        const float w_l = w_n;
        fetch_extents(t + t_step, lo_n, hi_n, w_n);
        // first 32 entries of every neighbour row requested before any accumulation (k+1 <= 8 rows per pass)
        int acount = 0;
        for (int nb0 = 0; nb0 < p.k1; nb0 += 8) {
            int cc[8]; float vv[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int64_t lo = __shfl_sync(0xffffffffu, lo_l, (nb0 + q) & 31), hi = __shfl_sync(0xffffffffu, hi_l, (nb0 + q) & 31);
                cc[q] = -1; vv[q] = 0.f;
                if (nb0 + q < p.k1 && lo + lane < hi) { cc[q] = __ldg(p.col + lo + lane); vv[q] = __ldg(p.val + lo + lane); }
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                if (nb0 + q >= p.k1) break;                       // warp-uniform
                const float w = __shfl_sync(0xffffffffu, w_l, nb0 + q);
                const int64_t lo = __shfl_sync(0xffffffffu, lo_l, nb0 + q), hi = __shfl_sync(0xffffffffu, hi_l, nb0 + q);
                bool is_new = false;
                const int cidx = cc[q];
                if (cidx >= 0) {
                    const float prod = p.weighted ? __fmul_rn(w, vv[q]) : vv[q];
                    acc[cidx] = __fadd_rn(acc[cidx], prod);
                    const unsigned bit = 1u << (cidx & 31);
                    const unsigned old = atomicOr(&amask[cidx >> 5], bit);
                    is_new = !(old & bit);
                }
                const unsigned b = __ballot_sync(0xffffffffu, is_new);
                if (is_new) alist[acount + __popc(b & ((1u << lane) - 1u))] = (unsigned short)cidx;
                acount += __popc(b);
                __syncwarp();
                if (hi - lo > 32) merge_row(p.col, p.val, lo + 32, hi, w, p.weighted != 0, acc, amask, alist, acount, lane);   // long row: the rest
            }
        }
        // x pdf(0), log1p, L1 norm (Impute.py:19, 91-92)
        float part = 0.f;
        for (int i = lane; i < acount; i += 32) {
            const int ci = alist[i];
            const float v = log1pf(__fmul_rn(acc[ci], HSG_PDF0));
            acc[ci] = v;
            part += fabsf(v);
        }
        const float total = warp_sum(part);
        const float inv_total = (total > 0.f) ? __frcp_rn(total) : 1.0f;      // one reciprocal per row (1 ulp from the division)
        __syncwarp();
        // gather from shared memory: group g takes candidates g, g + 4, ...; lane gl of a group owns columns 8 gl .. 8 gl + 7
        float4 o0 = make_float4(0.f, 0.f, 0.f, 0.f), o1 = o0;
#pragma unroll 2
        for (int i = grp; i < acount; i += 4) {
            const int ci = alist[i];
            const float a = acc[ci] * inv_total;
            const float4* er = reinterpret_cast<const float4*>(Es + ci * 64 + gl * 8);
            const float4 x0 = er[0], x1 = er[1];
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
        if (grp == 0) {
            float4* dst = reinterpret_cast<float4*>(p.neigh + ((int64_t)cb * p.n_bins + off + l) * 64 + gl * 8);
            dst[0] = o0; dst[1] = o1;
        }
        __syncwarp();
        for (int i = lane; i < acount; i += 32) {            // reset scratch for the next token
            const int ci = alist[i];
            acc[ci] = 0.f;
            amask[ci >> 5] = 0u;
        }
        __syncwarp();
    }
}

// returns 1 if the chromosome-block kernel was launched (it then covers every token of the batch)
static int try_launch_gather_chrom(hsg_ctx* c, int cell_start, int n_batch, int max_nc, cudaStream_t s, bool* done) {
    *done = false;
    const size_t smem = (size_t)max_nc * 64 * sizeof(float) + (size_t)GC_WARPS * gather_warp_bytes(max_nc, 0);
    if (c->d != 64 || c->ltr != 0 || c->n_chrom > GC_MAX_CHROM || smem > 220 * 1024 || c->k1 > 32) return 0;
    GatherChromParams p;
    p.row_ptr = c->row_ptr; p.col = c->col; p.val = c->val; p.nbr = c->nbr; p.nbr_w = c->nbr_w; p.E_bin = c->E_bin;
    p.bin_off = c->d_bin_off; p.neigh = c->neigh; p.k1 = c->k1; p.weighted = c->weighted ? 1 : 0; p.n_bins = c->n_bins;
    p.cell_start = cell_start; p.n_batch = n_batch; p.max_nc = max_nc; p.n_chrom = c->n_chrom;
    // blocks per chromosome in proportion to n_c (n_c + 400): rows x (fixed cost of a token -- three dependent global round
    // trips -- plus the candidates of a row, which grow with n_c), at least one, about 2 waves in total
    auto weight = [](int b) { return (double)b * ((double)b + 400.0); };
    double tot = 0;
    for (int b : c->bins) tot += weight(b);
    const int target = 2 * c->n_sm;
    int first = 0;
    for (int j = 0; j < c->n_chrom; ++j) {
        p.blk_first[j] = first;
        const int64_t toks = (int64_t)n_batch * c->bins[j];
        int nb = (int)(target * weight(c->bins[j]) / tot + 0.5);
        nb = (int)std::max<int64_t>(1, std::min<int64_t>(nb, (toks + GC_WARPS - 1) / GC_WARPS));
        first += nb;
    }
    p.blk_first[c->n_chrom] = first;
    static bool attr_set[64] = {};               // per device: the attribute belongs to the device's function instance
    if (!attr_set[c->device & 63]) {
        HSG_CUDA(cudaFuncSetAttribute(neigh_gather_chrom_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        attr_set[c->device & 63] = true;
    }
    neigh_gather_chrom_kernel<<<first, GC_WARPS * 32, smem, s>>>(p);
    c->launches++;
    HSG_CUDA(cudaGetLastError());
    *done = true;
    return 0;
}

// ---------------------------------------------------------------------------------------------
// K1a, rows with tens of merged candidates (C3, C4, C5).  For long chromosomes (C5: 4864 bins at 50 kb) the chromosome's E
// table (1.2 MB) does not fit in shared memory and a dense per-warp accumulator (19 KB) leaves the generic kernel ~10 warps per
// SM; the register kernel pays SLOTS^2 shuffles for its all-pairs dedup.  Here a warp merges the k+1 rows of its token into a
// 256-slot HASH table in shared memory (3.5 KB per warp): one atomicCAS on the key claims / finds the slot of a column, the value
// is added in row order (columns are unique inside a row, rows are processed one after the other -> the same fp32 sums as the
// generic kernel), then the table is compacted, x pdf(0), log1p, L1 norm, and the E rows are gathered from L2 four candidates at
// a time (8 lanes x 8 columns, two 128-bit loads per 8 FMAs).  Row extents are fetched one token ahead, the first 32 entries of
// all rows are requested before any is inserted.  Tokens whose probe sequence exceeds GH_PROBES (table nearly full) go to the
// overflow list that the generic kernel finishes.
// ---------------------------------------------------------------------------------------------
#define GH_WARPS 8
#define GH_SLOTS 256
#define GH_PROBES 48
struct GatherHashParams {
    const int64_t* row_ptr; const int32_t* col; const float* val; const int32_t* nbr; const float* nbr_w;
    const float* E_bin; const int* bin_chrom; const int* bin_off; float* neigh; int* tok_list; int* tok_count;
    int k1, weighted, n_bins, cell_start;
    int64_t n_tok;
    const int* in_list; const int* in_count;      // list mode: only the tokens in_list[0 .. *in_count) (the register kernel's overflow)
};

__global__ void __launch_bounds__(GH_WARPS * 32) neigh_gather_hash_kernel(GatherHashParams p) {
    __shared__ int s_key[GH_WARPS][GH_SLOTS];
    __shared__ float s_val[GH_WARPS][GH_SLOTS];
    __shared__ float s_cval[GH_WARPS][GH_SLOTS];
    __shared__ unsigned short s_ccol[GH_WARPS][GH_SLOTS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int* key = s_key[warp]; float* hv = s_val[warp]; float* cval = s_cval[warp]; unsigned short* ccol = s_ccol[warp];
    for (int i = lane; i < GH_SLOTS; i += 32) { key[i] = -1; hv[i] = 0.f; }
    __syncwarp();
    const int grp = lane >> 3, gl = lane & 7;

    const int64_t n_work = p.in_list ? (int64_t)*p.in_count : p.n_tok;
    if ((int64_t)blockIdx.x * GH_WARPS + warp >= n_work) return;
    auto token_of = [&](int64_t ti) -> int64_t { return p.in_list ? (int64_t)__ldg(p.in_list + ti) : ti; };
    auto fetch_extents = [&](int64_t ti, int64_t& lo, int64_t& hi, float& w) {
        lo = 0; hi = 0; w = 1.0f;
        if (lane < p.k1 && ti < n_work) {
            const int64_t t = token_of(ti);
            const int cb = (int)(t / p.n_bins), g = (int)(t - (int64_t)cb * p.n_bins);
            const int cell = p.cell_start + cb;
            const int src = p.weighted ? __ldg(p.nbr + (int64_t)cell * p.k1 + lane) : cell;
            if (p.weighted) w = __ldg(p.nbr_w + (int64_t)cell * p.k1 + lane);
            const int64_t row = (int64_t)src * p.n_bins + g;
            lo = __ldg(p.row_ptr + row); hi = __ldg(p.row_ptr + row + 1);
        }
    };
    // insert one (column, weighted value) of the current row; false = probe limit hit
    auto insert = [&](int cidx, float prod) -> bool {
        unsigned h = ((unsigned)cidx * 0x9E3779B1u) >> 24;
        for (int pr = 0; pr < GH_PROBES; ++pr) {
            const int slot = (int)((h + pr) & (GH_SLOTS - 1));
            HSG_DASSERT(cidx >= 0 && cidx < 65536 && slot >= 0 && slot < GH_SLOTS);
            const int old = atomicCAS(&key[slot], -1, cidx);
            if (old == -1 || old == cidx) { hv[slot] = __fadd_rn(hv[slot], prod); return true; }
        }
        return false;
    };

    const int64_t t_first = (int64_t)blockIdx.x * GH_WARPS + warp, t_step = (int64_t)gridDim.x * GH_WARPS;
    int64_t lo_n, hi_n;
    float w_n;
    fetch_extents(t_first, lo_n, hi_n, w_n);
    for (int64_t ti = t_first; ti < n_work; ti += t_step) {
        const int64_t t = token_of(ti);
        const int g = (int)(t % p.n_bins);
        const int off = p.bin_off[p.bin_chrom[g]];
        const int64_t lo_l = lo_n, hi_l = hi_n;
        const float w_l = w_n;
        fetch_extents(ti + t_step, lo_n, hi_n, w_n);
        bool fail = false;
        for (int nb0 = 0; nb0 < p.k1; nb0 += 8) {
            int cc[8]; float vv[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int64_t lo = __shfl_sync(0xffffffffu, lo_l, (nb0 + q) & 31), hi = __shfl_sync(0xffffffffu, hi_l, (nb0 + q) & 31);
                cc[q] = -1; vv[q] = 0.f;
                if (nb0 + q < p.k1 && lo + lane < hi) { cc[q] = __ldg(p.col + lo + lane); vv[q] = __ldg(p.val + lo + lane); }
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                if (nb0 + q >= p.k1) break;                       // warp-uniform
                const float w = __shfl_sync(0xffffffffu, w_l, nb0 + q);
                const int64_t lo = __shfl_sync(0xffffffffu, lo_l, nb0 + q), hi = __shfl_sync(0xffffffffu, hi_l, nb0 + q);
                if (cc[q] >= 0 && !insert(cc[q], p.weighted ? __fmul_rn(w, vv[q]) : vv[q])) fail = true;
                __syncwarp();
                for (int64_t e0 = lo + 32; e0 < hi; e0 += 32) {   // long row: the rest, 32 entries at a time
                    const int64_t e = e0 + lane;
                    if (e < hi) {
                        const float v = __ldg(p.val + e);
                        if (!insert(__ldg(p.col + e), p.weighted ? __fmul_rn(w, v) : v)) fail = true;
                    }
                    __syncwarp();
                }
            }
        }
        fail = __any_sync(0xffffffffu, fail);
        // compact the table (and reset it for the next token)
        int acount = 0;
        float part = 0.f;
#pragma unroll
        for (int s0 = 0; s0 < GH_SLOTS; s0 += 32) {
            const int sidx = s0 + lane;
            const int kcol = key[sidx];
            const bool used = kcol >= 0;
            const unsigned b = __ballot_sync(0xffffffffu, used);
            if (used) {
                const float v = log1pf(__fmul_rn(hv[sidx], HSG_PDF0));       // x pdf(0), log1p (Impute.py:19, 91)
                const int pos = acount + __popc(b & ((1u << lane) - 1u));
                HSG_DASSERT(pos >= 0 && pos < GH_SLOTS);
                ccol[pos] = (unsigned short)kcol; cval[pos] = v;
                part += fabsf(v);
                key[sidx] = -1; hv[sidx] = 0.f;
            }
            acount += __popc(b);
        }
        const float total = warp_sum(part);
        __syncwarp();
        if (fail) {                                             // table (nearly) full: the generic kernel finishes this token
            if (lane == 0) p.tok_list[atomicAdd(p.tok_count, 1)] = (int)t;
            continue;
        }
        const float inv_total = (total > 0.f) ? __frcp_rn(total) : 1.0f;
        const float4* Ec = reinterpret_cast<const float4*>(p.E_bin + (int64_t)off * 64) + gl * 2;
        float4 o0 = make_float4(0.f, 0.f, 0.f, 0.f), o1 = o0;
#pragma unroll 2
        for (int i = grp; i < acount; i += 4) {
            const int ci = ccol[i];
            const float a = cval[i] * inv_total;
            const float4 x0 = __ldg(Ec + (int64_t)ci * 16), x1 = __ldg(Ec + (int64_t)ci * 16 + 1);
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
        if (grp == 0) {
            float4* dst = reinterpret_cast<float4*>(p.neigh + t * 64 + gl * 8);
            dst[0] = o0; dst[1] = o1;
        }
        __syncwarp();
    }
}

int launch_gather(hsg_ctx* c, const int*, int cell_start, int n_batch, cudaStream_t s) {
    int max_nc = 0;
    for (int b : c->bins) max_nc = b > max_nc ? b : max_nc;
    HSG_CHECK(max_nc < 65536, "chromosomes with >= 65536 bins are not supported by the gather kernel");
    size_t wbytes = gather_warp_bytes(max_nc, c->ltr);
    int wpb = (int)((200 * 1024) / wbytes);
    if (wpb > 16) wpb = 16;
    HSG_CHECK(wpb >= 1, "chromosome too long for the gather scratch");
    size_t smem = wbytes * wpb;
    int64_t n_tok = (int64_t)n_batch * c->n_bins;
    int64_t blocks = (n_tok + wpb - 1) / wpb;
    int64_t cap = 148LL * 8;
    unsigned grid = (unsigned)(blocks < cap ? blocks : cap);
    const int* list = nullptr;
    const int* count = nullptr;
    // sparse-row fast path (d = 64, no row-shift smoothing, at most 8 neighbour rows): the generic kernel then only finishes the
    // tokens on the overflow list
    // Which kernel: the register kernel wins for sparse rows (<= 16 merged candidates on average).  For longer rows its
    // all-pairs dedup (SLOTS^2) loses to the generic kernel's dense accumulator -- unless the chromosomes are so long that the
    // accumulator (4 B x max_nc per warp) leaves the generic kernel with a handful of warps per SM (measured, C3 / C4 / C5 shapes).
    const double avg_cand = (double)c->k1 * (double)c->nnz / (double)std::max<int64_t>(c->n_rows, 1);
    const bool sparse_ok = (c->d == 64 || c->d == 128) && c->ltr == 0 && c->k1 <= 8 && n_tok < (1LL << 31) && c->gather_sparse;
    // long rows (> 16 merged candidates): the hash-merge kernel wherever the fast path applies (measured: C5 7.28 -> 2.54 ms against
    // the 96-candidate register kernel, C3 3.53 -> 3.38 ms against the chromosome-block kernel); with it switched off the old rule
    // stands (register kernel only for chromosomes too long for a dense per-warp accumulator)
    // (d_model = 128: only the register kernel is templated on the width; longer rows stay on the generic kernel there)
    const bool to_sparse = sparse_ok && (c->d == 64 ? (avg_cand <= 16.0 || max_nc > 1536 || c->gather_hash || c->gather_stream) : avg_cand <= 16.0);
    if (!to_sparse && c->gather_chrom) {        // dense rows, short chromosomes: one chromosome per block, E_c in shared memory
        bool done = false;
        if (try_launch_gather_chrom(c, cell_start, n_batch, max_nc, s, &done)) return 1;
        if (done) return 0;
    }
    if (to_sparse) {
        if (c->ovf_cap < n_tok + 1) {
            if (c->ovf_list) cudaFree(c->ovf_list);
            HSG_CUDA(cudaMalloc((void**)&c->ovf_list, (size_t)(2 * n_tok + 2) * sizeof(int)));      // two lists, each followed by its counter
            c->ovf_cap = n_tok + 1;
        }
        const bool second_list = false;      // (list mode of the hash-merge kernel: see the note at the register-kernel launch below)
        int* cnt = c->ovf_list + n_tok;
        HSG_CUDA(cudaMemsetAsync(cnt, 0, sizeof(int), s));
        int64_t quads = (n_tok + 3) / 4;
        int64_t gb = (quads + GS_WPB - 1) / GS_WPB;
        int64_t gcap = (int64_t)c->n_sm * 8;
        // register capacity from the average number of merged candidates per token (k+1 rows of nnz / n_rows entries each)
        const double avg = avg_cand;
        const unsigned g = (unsigned)(gb < gcap ? gb : gcap);
        // the streaming kernel (gather_stream.cu: bulk-copied CSR segments, E table of the chromosome in shared memory) whenever
        // its stages fit next to the chromosome's E table; it leaves what it cannot hold on the same overflow list
        bool streamed = false;
        if (c->gather_stream && c->d == 64 && try_launch_gather_stream(c, cell_start, n_batch, max_nc, cnt, s, &streamed)) return 1;
#define GS_LAUNCH(SL) neigh_gather_sparse_kernel<SL, 64><<<g, GS_WPB * 32, 0, s>>>(                                                    \
            c->row_ptr, c->col, c->val, c->nbr, c->nbr_w, c->k1, c->weighted ? 1 : 0, c->E_bin, c->d_bin_chrom, c->d_bin_off,      \
            c->n_bins, cell_start, (int)n_tok, c->neigh, c->ovf_list, cnt)
        if (streamed) {
        } else if (c->d == 128) {
            neigh_gather_sparse_kernel<4, 128><<<g, GS_WPB * 32, 0, s>>>(
                c->row_ptr, c->col, c->val, c->nbr, c->nbr_w, c->k1, c->weighted ? 1 : 0, c->E_bin, c->d_bin_chrom, c->d_bin_off,
                c->n_bins, cell_start, (int)n_tok, c->neigh, c->ovf_list, cnt);
        } else if (avg > 16.0 && c->gather_hash && max_nc < 65536) {
            // long chromosome, tens of candidates per token: shared-memory hash merge instead of the SLOTS^2 register dedup
            GatherHashParams hp;
            hp.row_ptr = c->row_ptr; hp.col = c->col; hp.val = c->val; hp.nbr = c->nbr; hp.nbr_w = c->nbr_w; hp.E_bin = c->E_bin;
            hp.bin_chrom = c->d_bin_chrom; hp.bin_off = c->d_bin_off; hp.neigh = c->neigh; hp.tok_list = c->ovf_list; hp.tok_count = cnt;
            hp.k1 = c->k1; hp.weighted = c->weighted ? 1 : 0; hp.n_bins = c->n_bins; hp.cell_start = cell_start; hp.n_tok = n_tok;
            hp.in_list = nullptr; hp.in_count = nullptr;
            const int64_t hb = (n_tok + GH_WARPS - 1) / GH_WARPS;
            neigh_gather_hash_kernel<<<(unsigned)std::min<int64_t>(hb, (int64_t)c->n_sm * 6), GH_WARPS * 32, 0, s>>>(hp);
        }
        // (tried: finishing the register kernel's overflow list with the hash-merge kernel in list mode before the generic kernel --
        //  C2 K1a 1.94 -> 2.06 ms per 1000 cells: the list is almost always empty and a third launch costs more than it saves)
        else if (avg <= 16.0) GS_LAUNCH(4);
        else if (avg <= 36.0) GS_LAUNCH(8);
        else GS_LAUNCH(12);
#undef GS_LAUNCH
        if (!streamed) c->launches++;
        HSG_CUDA(cudaGetLastError());
        list = second_list ? c->ovf_list + n_tok + 1 : c->ovf_list; count = second_list ? c->ovf_list + 2 * n_tok + 1 : cnt;
        grid = (unsigned)std::min<int64_t>(grid, (int64_t)c->n_sm * 2);
    }
    if (c->d == 64) {
        HSG_CUDA(cudaFuncSetAttribute(neigh_gather_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        neigh_gather_kernel<64><<<grid, wpb * 32, smem, s>>>(c->row_ptr, c->col, c->val, c->nbr, c->nbr_w, c->k1,
                                                             c->weighted ? 1 : 0, c->E_bin, c->d_bin_chrom, c->d_bin_off,
                                                             c->n_bins, cell_start, n_batch, c->ltr, max_nc, c->neigh, list, count);
    } else {
        HSG_CUDA(cudaFuncSetAttribute(neigh_gather_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        neigh_gather_kernel<128><<<grid, wpb * 32, smem, s>>>(c->row_ptr, c->col, c->val, c->nbr, c->nbr_w, c->k1,
                                                              c->weighted ? 1 : 0, c->E_bin, c->d_bin_chrom, c->d_bin_off,
                                                              c->n_bins, cell_start, n_batch, c->ltr, max_nc, c->neigh, list, count);
    }
    c->launches++;
    HSG_CUDA(cudaGetLastError());
    static const bool dbg = getenv("HSG_GATHER_DEBUG") != nullptr;
    if (dbg && count) {                              // diagnostics: how many tokens the fast path left to the generic kernel
        int h = 0;
        HSG_CUDA(cudaStreamSynchronize(s));
        HSG_CUDA(cudaMemcpy(&h, count, sizeof(int), cudaMemcpyDeviceToHost));
        fprintf(stderr, "[hsg] K1a overflow tokens: %d of %lld\n", h, (long long)n_tok);
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------
// K1b.  32 tokens per block, 256 threads; register-tiled fp32 GEMMs with the activations staged
// transposed in shared memory ([feature][token]) and the (transposed) weights read through L1.
// ---------------------------------------------------------------------------------------------
#define TP_TOK 32
#define TP_LDA 36        // padded token stride (multiple of 4 for 128-bit loads)

// acc[4][CT] += sum_k A^T[k][row0..row0+3] * Wt[k][col0..col0+CT-1]
template <int CT>
__device__ __forceinline__ void tile_gemm(const float* __restrict__ sA, int K, int row0,
                                          const float* __restrict__ Wt, int ldw, int col0, float (&acc)[4][CT]) {
#pragma unroll 4
    for (int k = 0; k < K; ++k) {
        float4 a = *reinterpret_cast<const float4*>(sA + k * TP_LDA + row0);
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

template <int D, int MODE>   // MODE 0: fp32 Q/K ; 1: fp16 Q/K + cell cross terms ; 2: fp32 Q/K + cell cross terms
__global__ void __launch_bounds__(256) token_project_kernel(
    const float* __restrict__ neigh, const float* __restrict__ E_bin, int n_bins, int64_t n_tok,
    const float* __restrict__ sageT, const float* __restrict__ sage_b,
    const float* __restrict__ g1, const float* __restrict__ b1, const float* __restrict__ g2,
    const float* __restrict__ b2, const float* __restrict__ wqT, const float* __restrict__ wkT,
    float* __restrict__ QK, __half* __restrict__ QK16, float* __restrict__ XS,
    const float* __restrict__ cellQ, const float* __restrict__ cellK, int cell_start, int chunk_major) {
    extern __shared__ __align__(16) float sm[];
    float* xT = sm;                              // [2D][LDA]
    float* dT = xT + 2 * D * TP_LDA;             // [D][LDA]   dyn, then LN1 input
    float* kT = dT + D * TP_LDA;                 // [D][LDA]   LN2 input
    float* qkT = kT + D * TP_LDA;                // [256][LDA] Q (rows 0..127) and K (rows 128..255), MODE != 0 only
    constexpr bool OUT16 = (MODE == 1);
    const int tid = threadIdx.x;
    const int64_t tok0 = (int64_t)blockIdx.x * TP_TOK;
    // stage [neigh | E_self] transposed
    for (int i = tid; i < TP_TOK * 2 * D; i += 256) {
        int tk = i / (2 * D), f = i - tk * 2 * D;
        int64_t tok = tok0 + tk;
        float v = 0.f;
        if (tok < n_tok) {
            if (f < D) v = neigh[tok * D + f];
            else v = E_bin[(int64_t)(tok % n_bins) * D + (f - D)];
        }
        xT[f * TP_LDA + tk] = v;
    }
    __syncthreads();
    // GEMM1: dyn = swish(x . sageT + b)   [32][D]
    {
        constexpr int NCG = D / 4;
        for (int t = tid; t < (TP_TOK / 4) * NCG; t += 256) {
            int rg = t % (TP_TOK / 4), cg = t / (TP_TOK / 4);
            float acc[4][4] = {};
            tile_gemm<4>(xT, 2 * D, rg * 4, sageT, D, cg * 4, acc);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float bb = sage_b[cg * 4 + q];
                *reinterpret_cast<float4*>(dT + (cg * 4 + q) * TP_LDA + rg * 4) =
                    make_float4(swish_acc(acc[0][q] + bb), swish_acc(acc[1][q] + bb), swish_acc(acc[2][q] + bb), swish_acc(acc[3][q] + bb));
            }
        }
    }
    __syncthreads();
    // LayerNorm statistics per token (shared by layer_norm1 and layer_norm2), 8 threads per token
    {
        int tk = tid >> 3, sub = tid & 7;
        float s = 0.f;
        for (int f = sub; f < D; f += 8) s += dT[f * TP_LDA + tk];
        s += __shfl_xor_sync(0xffffffffu, s, 1); s += __shfl_xor_sync(0xffffffffu, s, 2); s += __shfl_xor_sync(0xffffffffu, s, 4);
        float mean = s / (float)D;
        float v = 0.f;
        for (int f = sub; f < D; f += 8) { float dd = dT[f * TP_LDA + tk] - mean; v = fmaf(dd, dd, v); }
        v += __shfl_xor_sync(0xffffffffu, v, 1); v += __shfl_xor_sync(0xffffffffu, v, 2); v += __shfl_xor_sync(0xffffffffu, v, 4);
        float rstd = 1.0f / sqrtf(v / (float)D + HSG_LN_EPS);
        for (int f = sub; f < D; f += 8) {
            float xh = (dT[f * TP_LDA + tk] - mean) * rstd;
            dT[f * TP_LDA + tk] = fmaf(xh, g1[f], b1[f]);
            kT[f * TP_LDA + tk] = fmaf(xh, g2[f], b2[f]);
        }
    }
    __syncthreads();
    // GEMM2: Q = LN1(dyn) . wqT, K = LN2(dyn) . wkT   [32][128] each
    for (int t = tid; t < 2 * (TP_TOK / 4) * (HSG_HD / 4); t += 256) {
        int which = t / ((TP_TOK / 4) * (HSG_HD / 4));
        int u = t - which * ((TP_TOK / 4) * (HSG_HD / 4));
        int rg = u % (TP_TOK / 4), cg = u / (TP_TOK / 4);
        float acc[4][4] = {};
        tile_gemm<4>(which ? kT : dT, D, rg * 4, which ? wkT : wqT, HSG_HD, cg * 4, acc);
        if (MODE != 0) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
                *reinterpret_cast<float4*>(qkT + (which * HSG_HD + cg * 4 + q) * TP_LDA + rg * 4) =
                    make_float4(acc[0][q], acc[1][q], acc[2][q], acc[3][q]);
        }
        if (!OUT16) {
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                int64_t tok = tok0 + rg * 4 + r;
                // MODE 2 (high-accuracy scorer): quads in (quad e, head h) order, see HSG_HI_POS
                const int qpos = (MODE == 2) ? ((cg & 3) * 8 + (cg >> 2)) * 4 : cg * 4;
                if (tok < n_tok)
                    *reinterpret_cast<float4*>(QK + (tok * 2 + which) * HSG_HD + qpos) =
                        make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
            }
        }
    }
    if (MODE != 0) {
        __syncthreads();
        // 16-bit rows: thread -> (token, 8-column chunk); lanes walk tokens so the smem reads are conflict-free
        for (int i = tid; OUT16 && i < TP_TOK * 32; i += 256) {
            int tk = i & 31, ch = i >> 5;
            int64_t tok = tok0 + tk;
            if (tok >= n_tok) continue;
            const float* src = qkT + (ch * 8) * TP_LDA + tk;
            const float sc = (ch < 16) ? HSG_L2E_QUARTER : 1.0f;        // Q rows carry log2(e)/4: the scorer's logits are in log2 units
            __half2 h0 = __floats2half2_rn(sc * src[0], sc * src[TP_LDA]), h1 = __floats2half2_rn(sc * src[2 * TP_LDA], sc * src[3 * TP_LDA]);
            __half2 h2 = __floats2half2_rn(sc * src[4 * TP_LDA], sc * src[5 * TP_LDA]), h3 = __floats2half2_rn(sc * src[6 * TP_LDA], sc * src[7 * TP_LDA]);
            uint4 pk;
            pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
            pk.z = *reinterpret_cast<uint32_t*>(&h2); pk.w = *reinterpret_cast<uint32_t*>(&h3);
            if (chunk_major)      // score_umma2.cu: table[slot][bin >> 5][chunk][bin & 31], natural chunk order
                reinterpret_cast<uint4*>(QK16)[(((size_t)(tok / n_bins) * (size_t)((n_bins + HSG_GRP - 1) / HSG_GRP) + (size_t)((tok % n_bins) >> 5)) * 32u + (size_t)ch) * HSG_GRP + (size_t)((tok % n_bins) & 31)] = pk;
            else
            *reinterpret_cast<uint4*>(QK16 + tok * 256 + (ch & 16) * 8 + HSG_CHUNK_POS(ch & 15) * 8) = pk;      // head-per-lane chunk order
        }
        // cross terms with the cell token: qcK[h] = Q_cell,h . K_b,h / 4 ; kcQ[h] = Q_b,h . K_cell,h / 4
        {
            int tk = tid & 31, h = tid >> 5;
            int64_t tok = tok0 + tk;
            if (tok < n_tok) {
                int cell = cell_start + (int)(tok / n_bins);
                const float* qc = cellQ + (int64_t)cell * HSG_HD + h * 16;
                const float* kc = cellK + (int64_t)cell * HSG_HD + h * 16;
                float a = 0.f, b = 0.f;
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                    a = fmaf(__ldg(qc + e), qkT[(HSG_HD + h * 16 + e) * TP_LDA + tk], a);
                    b = fmaf(qkT[(h * 16 + e) * TP_LDA + tk], __ldg(kc + e), b);
                }
                if (chunk_major) {        // XS[slot][bin >> 5][quad][bin & 31][4]: quads 0, 1 = qcK, quads 2, 3 = kcQ
                    const size_t sl = (size_t)(tok / n_bins), g = (size_t)(tok % n_bins), ng = (size_t)((n_bins + HSG_GRP - 1) / HSG_GRP);
                    float* xg = XS + (((sl * ng + (g >> 5)) * 4) * HSG_GRP + (g & 31)) * 4;      // quad q of this bin at + q * HSG_GRP * 4 floats
                    xg[(size_t)(h >> 2) * HSG_GRP * 4 + (h & 3)] = a * HSG_L2E_QUARTER;
                    xg[(size_t)(2 + (h >> 2)) * HSG_GRP * 4 + (h & 3)] = b * HSG_L2E_QUARTER;
                } else {
                XS[tok * 16 + HSG_XS_POS(h)] = a * HSG_L2E_QUARTER;
                XS[tok * 16 + 8 + HSG_XS_POS(h)] = b * HSG_L2E_QUARTER;
                }
            }
        }
    }
}

template <int D, int MODE>
static int launch_tp(hsg_ctx* c, int cell_start, int64_t n_tok, cudaStream_t s) {
    unsigned grid = (unsigned)((n_tok + TP_TOK - 1) / TP_TOK);
    float** w = c->w;
    size_t smem = (size_t)(4 * D + (MODE != 0 ? 2 * HSG_HD : 0)) * TP_LDA * sizeof(float);
    HSG_CUDA(cudaFuncSetAttribute(token_project_kernel<D, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    token_project_kernel<D, MODE><<<grid, 256, smem, s>>>(c->neigh, c->E_bin, c->n_bins, n_tok, c->dw.sageT, w[HSG_W_SAGE_B],
                                                           w[HSG_W_ALN1_G], w[HSG_W_ALN1_B], w[HSG_W_ALN2_G], w[HSG_W_ALN2_B],
                                                           c->dw.wqT, c->dw.wkT, c->QK, (__half*)c->QK16, c->XS, c->cellQ, c->cellK,
                                                           cell_start, (MODE == 1 && c->k2v2) ? 1 : 0);
    c->launches++;
    HSG_CUDA(cudaGetLastError());
    return 0;
}

int launch_token_project(hsg_ctx* c, int cell_start, int n_batch, int mode, cudaStream_t s) {
    int64_t n_tok = (int64_t)n_batch * c->n_bins;
    if (c->d == 64) {
        if (mode == 1) return launch_tp<64, 1>(c, cell_start, n_tok, s);
        if (mode == 2) return launch_tp<64, 2>(c, cell_start, n_tok, s);
        return launch_tp<64, 0>(c, cell_start, n_tok, s);
    }
    if (mode == 1) return launch_tp<128, 1>(c, cell_start, n_tok, s);
    if (mode == 2) return launch_tp<128, 2>(c, cell_start, n_tok, s);
    return launch_tp<128, 0>(c, cell_start, n_tok, s);
}
