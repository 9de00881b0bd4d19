// umma_common.cuh -- tcgen05 / TMEM / mbarrier helpers, shared-memory map and parameter block shared by the tensor-core
// kernels (score_umma.cu: the fused scorer; project_umma.cu: the token projection).
#pragma once
#include <cuda_fp16.h>

#include <cstring>
#include <vector>

#include "common.cuh"

#define UM_WG 4
#define UM_THREADS (UM_WG * 128)
#define UM_D 64

// shared memory map (bytes).  Weight images (fp16, K-major SWIZZLE_128B, built on the host):
//   [y_c | g] = ctx . [fc1c ; W0' fc1c]^T in ONE N=128 GEMM (see score_umma.cu), then W1c and C0; the high-accuracy variant
//   appends the low halves (w - fp16(w)) of the first and the last image.
#define SMF_B_FC1 0           // [128 x 128], 2 K-atoms of 16 KB: rows 0..63 fc1c, rows 64..127 W0' fc1c
#define SMF_B_W1 32768        // [64 x 64]  W1 with its output mean removed
#define SMF_B_C0 40960        // [32 x 64]  C0 / 2
// fast variant only: one more K atom each for W1 and C0 whose first two K columns hold the bias (fp16 hi + lo); the A operand
// supplies the matching constant columns (1, 1, 0, ...), so z_c and c leave the tensor core with their biases added
#define SMF_X_W1B 45056       // [64 x 64]  column 0 / 1 = hi / lo of b1 - mean(b1)
#define SMF_X_C0B 53248       // [32 x 64]  column 0 / 1 = hi / lo of cb0 / 2
#define SMF_WIMG_BYTES 57344
#define SMH_B_FC1_LO 45056    // [128 x 128] low halves
#define SMH_B_C0_LO 77824     // [32 x 64]   low halves
#define SMH_WIMG_BYTES 81920
// everything behind the weight images is placed relative to the end of the variant's images (HI = 1: high-accuracy variant), so
// the fast variant asks for ~190 KB instead of ~214 KB and the SM keeps a 60 KB L1 (196 KB carve-out) instead of 28 KB
#define SM_A0_OF(HI) ((HI) ? SMH_WIMG_BYTES : SMF_WIMG_BYTES)       // 4 x 32 KB ctx tiles (one per warpgroup)
#define SM_A_BYTES 32768
#define CST_B0F 0
#define CST_G1 128
#define CST_CB0 192
#define CST_CW1 224
#define CST_CB1 256
#define CST_B1C 264           // b1 - mean(b1): every LayerNorm input is produced already centred
#define CST_COUNT 328
#define SM_CONST_OF(HI) (SM_A0_OF(HI) + UM_WG * SM_A_BYTES)         // fp32 constants (reserved)
#define SM_BAR_OF(HI) (SM_CONST_OF(HI) + CST_COUNT * 4)            // 8 mbarriers
#define SM_TMEM_OF(HI) (SM_BAR_OF(HI) + 64)
#define SM_PAIRS_OF(HI) (SM_TMEM_OF(HI) + 16)                      // int2 [UM_WG * 128]: (bin_i, bin_j) of every row of the current tiles
#define SM_TOTAL_OF(HI) (SM_PAIRS_OF(HI) + UM_WG * 128 * 8)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity), "r"(0x989680u) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void wg_bar(int wg) { asm volatile("bar.sync %0, 128;" ::"r"(wg + 1) : "memory"); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
// start>>4 [0,14) | LBO>>4 [16,30) (unused for swizzled K-major, 1) | SBO>>4 [32,46) = 1024 B between
// 8-row groups | version=1 [46,48) | layout_type=2 (SWIZZLE_128B) [61,64)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor kind::f16: D=f32 (1<<4), A=B=f16 (0<<7, 0<<10), K-major both, N>>3 at 17, M>>4 at 24
__host__ __device__ constexpr uint32_t umma_idesc_f16(int M, int N) {
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
// A operand from TMEM (fp16, lane = row, two K elements per 32-bit column), B from shared memory
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

#define TMEM_LD_ARGS16(r, o) "=r"(r[o + 0]), "=r"(r[o + 1]), "=r"(r[o + 2]), "=r"(r[o + 3]), "=r"(r[o + 4]), "=r"(r[o + 5]),   \
    "=r"(r[o + 6]), "=r"(r[o + 7]), "=r"(r[o + 8]), "=r"(r[o + 9]), "=r"(r[o + 10]), "=r"(r[o + 11]), "=r"(r[o + 12]),         \
    "=r"(r[o + 13]), "=r"(r[o + 14]), "=r"(r[o + 15])

// 32 lanes x 32 columns -> 32 registers per thread
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : TMEM_LD_ARGS16(r, 0), TMEM_LD_ARGS16(r, 16) : "r"(taddr) : "memory");
}
// 32 lanes x 16 columns
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : TMEM_LD_ARGS16(r, 0) : "r"(taddr) : "memory");
}
// registers -> TMEM (thread r writes lane r): 16 / 8 consecutive 32-bit columns.  Used to hand an fp16 A operand (two K
// elements per column, K-major) to the tensor core without the shared-memory round trip.
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
                   "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ float fast_tanh(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// swish(x) = x * sigmoid(x) = x * (0.5 + 0.5 tanh(x/2)).  The GEMMs that feed a swish carry the factor 1/2 in
// their (pre-scaled) weights and biases, so the epilogue sees w = x/2 and swish(x) = w + w tanh(w): MUFU + FFMA.
__device__ __forceinline__ float swish_half(float w) { return fmaf(w, fast_tanh(w), w); }

// row r of a [128 x K] fp16 A tile, 16-byte chunk `c` (8 halves), K-major SWIZZLE_128B
__device__ __forceinline__ uint32_t a_chunk_addr(uint32_t a_base, int r, int c) {
    return a_base + (uint32_t)(c >> 3) * 16384u + (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u +
           (uint32_t)(((c & 7) ^ (r & 7)) << 4);
}
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// 16-byte read-only global load that the compiler may not move (used to start long-latency loads before an mbarrier wait)
__device__ __forceinline__ float4 ldg_nc_v4_volatile(const float4* p) {
    float4 v;
    asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float2 f2(uint32_t a, uint32_t b) { return make_float2(__uint_as_float(a), __uint_as_float(b)); }
__device__ __forceinline__ __half2 u2h(uint32_t u) { return *reinterpret_cast<__half2*>(&u); }
__device__ __forceinline__ uint32_t h2u(__half2 h) { return *reinterpret_cast<uint32_t*>(&h); }

struct UmmaParams {
    // 16-bit token tables
    const __half* QK16;      // [slot][n_bins][256]  Q then K
    const float* XS;         // [slot][n_bins][16]   qcK[8] = Q_cell.K_b, kcQ[8] = Q_b.K_cell (x log2(e)/4), pair order HSG_XS_POS
    const __half* binV16;    // [n_bins][128]
    const float* binSb;      // [16][n_bins][4] feature blocks: beta(layer_norm1) - layer_norm2(fc2 V)
    const __half* cellV16;   // [n_cells][128]
    // high-accuracy variant: fp32 token tables (Q/K per (cell, bin), V per bin / cell)
    const float* QK32;       // [slot][n_bins][2][128]
    const float* binV32;     // [n_bins][128]
    const float* cellV32;    // [n_cells][128]
    const float* cellSb;     // [16][n_cells][4]
    const unsigned char* wimg;   // SM_WIMG_BYTES pre-swizzled fp16 weight images
    float cst[CST_COUNT];        // fp32 constants, by value: they become constant-bank operands of the epilogue math
    const int2* pairs; const float* pair_bias; long long P;
    int n_bins, n_cells, cell_start, n_batch, tiles_per_cell, act;
    float* out;
};

__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
// masked 3-token softmax row (Modules.py:948-953) with the diagonal logit forced to 0:
// w_a = p_a / (p_a + p_b + 1e-13), p = softmax([0, la, lb])  ==  e_a / (e_a + e_b + 1e-13 (e_0 + e_a + e_b)).
// The logits arrive in log2 units (K1b folds log2(e)/4 into the Q rows and the cell cross terms).
__device__ __forceinline__ void masked_row_exact(float la, float lb, float& wa, float& wb) {
    float mx = fmaxf(0.f, fmaxf(la, lb));
    float e0 = ex2_approx(-mx), ea = ex2_approx(la - mx), eb = ex2_approx(lb - mx);
    float den = fmaf(1e-13f, e0 + ea + eb, ea + eb);
    float inv = rcp_approx(den);
    wa = ea * inv; wb = eb * inv;
}
// The 1e-13 term changes the result by less than 1e-7 relative unless both logits are below -20 (2^-20 ~ 1e-6 >> 1e-13), so the
// common case is a two-way softmax: w_a = 1 / (1 + 2^(lb - la)), w_b = 1 - w_a  (2 MUFU instead of 4).
__device__ __forceinline__ void masked_row_fast(float la, float lb, float& wa, float& wb) {
    wa = rcp_approx(1.0f + ex2_approx(lb - la));
    wb = 1.0f - wa;
    if (fmaxf(la, lb) < -20.f) masked_row_exact(la, lb, wa, wb);
}

struct UmmaParams;
// Attention of one token row-tile (fast variant), T = token position (0 cell, 1 bin_i, 2 bin_j): row T attends to the two
// other tokens.  Lane mapping: lane = (row-in-group rho = lane>>3, head j = lane&7); a warp walks its own 32 rows four at a
// time.  The 16-bit Q / K / V rows are stored with their 16-byte chunks in HSG_CHUNK_POS order (position j = first half,
// position 8 + j = second half of head j), so lane j reads positions j and 8 + j: the 8 lanes of a row cover two fully used
// 128-byte lines per load AND every lane owns one complete head -- its logit needs no cross-lane reduction and the softmax
// is computed once per (row, head).  The ctx chunks go to the same positions of the A tile (8 lanes of a row = 8 distinct
// swizzle slots, conflict-free); fc1's K columns are permuted to match.  s_pr: the (bin_i, bin_j) pairs of the warpgroup's
// 128 rows in shared memory; the calling warp handles rows row0 .. row0 + 4 NIT - 1.
__device__ __forceinline__ uint4 ldg_nc_u4_volatile(const uint4* p) {
    uint4 v;
    asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ float ldg_nc_f32_volatile(const float* p) {
    float v;
    asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
// everything one lane needs from the token tables for one row of one token position
struct AttRow {
    uint4 qA, qB, kA, kB;     // T != 0: this lane's head of Q[tq] and K[tk] ;  T == 0: V[bi] in qA/qB
    uint4 bA, bB;             // V of the "b" token
    float la, lb;             // logits that come from the cross-term table
};
template <int T>
__device__ __forceinline__ void att_load(AttRow& r, int rw, const int2* s_pr, const uint4* QKj, const float* XSj, const uint4* BVj,
                                         uint32_t tokbase) {
    const int2 pr = s_pr[rw];
    const uint32_t ti = tokbase + (uint32_t)pr.x, tj = tokbase + (uint32_t)pr.y;
    if (T == 0) {
        r.la = ldg_nc_f32_volatile(XSj + ti * 16u); r.lb = ldg_nc_f32_volatile(XSj + tj * 16u);
        r.qA = ldg_nc_u4_volatile(BVj + (uint32_t)pr.x * 16u); r.qB = ldg_nc_u4_volatile(BVj + (uint32_t)pr.x * 16u + 8);
        r.bA = ldg_nc_u4_volatile(BVj + (uint32_t)pr.y * 16u); r.bB = ldg_nc_u4_volatile(BVj + (uint32_t)pr.y * 16u + 8);
    } else {
        const uint32_t tq = (T == 1) ? ti : tj, tk = (T == 1) ? tj : ti, bv = (uint32_t)((T == 1) ? pr.y : pr.x);
        r.la = ldg_nc_f32_volatile(XSj + tq * 16u);
        r.qA = ldg_nc_u4_volatile(QKj + tq * 32u); r.qB = ldg_nc_u4_volatile(QKj + tq * 32u + 8);
        r.kA = ldg_nc_u4_volatile(QKj + tk * 32u + 16); r.kB = ldg_nc_u4_volatile(QKj + tk * 32u + 24);
        r.bA = ldg_nc_u4_volatile(BVj + bv * 16u); r.bB = ldg_nc_u4_volatile(BVj + bv * 16u + 8);
    }
}
// st_addr: shared-memory address of chunk position j of this row in the A tile; position 8 + j is one K atom (16 KB) further
template <int T>
__device__ __forceinline__ void att_compute(const AttRow& r, uint4 vcA, uint4 vcB, uint32_t st_addr) {
    float la = r.la, lb;
    uint4 aA, aB;
    if (T == 0) {
        lb = r.lb; aA = r.qA; aB = r.qB;
    } else {
        aA = vcA; aB = vcB;
        __half2 accA = __hmul2(u2h(r.qA.x), u2h(r.kA.x)), accB = __hmul2(u2h(r.qB.x), u2h(r.kB.x));
        accA = __hfma2(u2h(r.qA.y), u2h(r.kA.y), accA); accB = __hfma2(u2h(r.qB.y), u2h(r.kB.y), accB);
        accA = __hfma2(u2h(r.qA.z), u2h(r.kA.z), accA); accB = __hfma2(u2h(r.qB.z), u2h(r.kB.z), accB);
        accA = __hfma2(u2h(r.qA.w), u2h(r.kA.w), accA); accB = __hfma2(u2h(r.qB.w), u2h(r.kB.w), accB);
        const float2 fA = __half22float2(accA), fB = __half22float2(accB);
        lb = (fA.x + fA.y) + (fB.x + fB.y);      // the Q rows carry log2(e) / 4
    }
    float wa, wb;
    masked_row_fast(la, lb, wa, wb);
    const uint4 bA = r.bA, bB = r.bB;
    const __half2 wa2 = __float2half2_rn(wa), wb2 = __float2half2_rn(wb);
    sts128(st_addr,
           h2u(__hfma2(wb2, u2h(bA.x), __hmul2(wa2, u2h(aA.x)))), h2u(__hfma2(wb2, u2h(bA.y), __hmul2(wa2, u2h(aA.y)))),
           h2u(__hfma2(wb2, u2h(bA.z), __hmul2(wa2, u2h(aA.z)))), h2u(__hfma2(wb2, u2h(bA.w), __hmul2(wa2, u2h(aA.w)))));
    sts128(st_addr + 16384u,
           h2u(__hfma2(wb2, u2h(bB.x), __hmul2(wa2, u2h(aB.x)))), h2u(__hfma2(wb2, u2h(bB.y), __hmul2(wa2, u2h(aB.y)))),
           h2u(__hfma2(wb2, u2h(bB.z), __hmul2(wa2, u2h(aB.z)))), h2u(__hfma2(wb2, u2h(bB.w), __hmul2(wa2, u2h(aB.w)))));
}
// Software-pipelined: the table rows of row group it + 1 are requested (volatile loads, the compiler keeps the order) before the
// math of row group it, so one L2 round trip overlaps one group's compute instead of stalling every loop trip.
template <int T, int NIT>
__device__ __forceinline__ void attention_fast(const __half* __restrict__ QK16, const float* __restrict__ XS,
                                               const __half* __restrict__ binV16, const __half* __restrict__ cellV16,
                                               const int2* s_pr, int lane, int row0, uint32_t a_base, uint32_t tokbase, int cell) {
    const int j = lane & 7, rsub = lane >> 3;
    const uint4* QKj = reinterpret_cast<const uint4*>(QK16) + j;
    const uint4* BVj = reinterpret_cast<const uint4*>(binV16) + j;
    const float* XSj = XS + j + (T == 0 ? 0 : 8);
    uint4 vcA = make_uint4(0, 0, 0, 0), vcB = vcA;
    if (T != 0) {
        const uint4* vc = reinterpret_cast<const uint4*>(cellV16) + (uint32_t)cell * 16u + j;
        vcA = __ldg(vc); vcB = __ldg(vc + 8);
    }
    // row0 is a multiple of 32 and the rows of an even / odd group are rsub / rsub + 4 (mod 8): the swizzled store address is a
    // per-lane constant plus 1024 bytes per pair of groups
    uint32_t stA = a_chunk_addr(a_base, row0 + rsub, j), stB = a_chunk_addr(a_base, row0 + 4 + rsub, j);
    const int2* prA = s_pr + row0 + rsub;
    AttRow A, B;
    att_load<T>(A, 0, prA, QKj, XSj, BVj, tokbase);
#pragma unroll 1
    for (int it = 0; it < NIT; it += 2) {
        att_load<T>(B, 4, prA, QKj, XSj, BVj, tokbase);
        att_compute<T>(A, vcA, vcB, stA);
        if (it + 2 < NIT) att_load<T>(A, 8, prA, QKj, XSj, BVj, tokbase);
        att_compute<T>(B, vcA, vcB, stB);
        prA += 8; stA += 1024u; stB += 1024u;
    }
}

// NK k-steps of one [128 x N] += A[128 x 16NK] . B[N x 16NK]^T ; A / B tiles in K-major SWIZZLE_128B
template <int NK, int N>
__device__ __forceinline__ void mma_group(uint32_t a_base, uint32_t b_base, int b_rows, uint32_t d_tmem, uint32_t accum_first) {
    constexpr uint32_t idesc = umma_idesc_f16(128, N);
#pragma unroll
    for (int k = 0; k < NK; ++k) {
        uint32_t aoff = (uint32_t)(k >> 2) * 16384u + (uint32_t)(k & 3) * 32u;
        uint32_t boff = (uint32_t)(k >> 2) * (uint32_t)(b_rows * 128) + (uint32_t)(k & 3) * 32u;
        umma_f16(d_tmem, umma_desc(a_base + aoff), umma_desc(b_base + boff), idesc, (k > 0) ? 1u : accum_first);
    }
}
// same with the A tile in TMEM: k-step k reads the 8 columns [8k, 8k + 8)
template <int NK, int N>
__device__ __forceinline__ void mma_group_ts(uint32_t a_tmem, uint32_t b_base, int b_rows, uint32_t d_tmem, uint32_t accum_first) {
    constexpr uint32_t idesc = umma_idesc_f16(128, N);
#pragma unroll
    for (int k = 0; k < NK; ++k) {
        uint32_t boff = (uint32_t)(k >> 2) * (uint32_t)(b_rows * 128) + (uint32_t)(k & 3) * 32u;
        umma_f16_ts(d_tmem, a_tmem + (uint32_t)k * 8u, umma_desc(b_base + boff), idesc, (k > 0) ? 1u : accum_first);
    }
}
// one MMA round trip for a warpgroup: the A tile is complete in shared memory.  Generic-proxy writes of the A tile are
// made visible to the tensor core (async proxy), the warpgroup syncs, one thread issues and commits to the mbarrier.
#define ISSUE_BEGIN() fence_async_smem(); tc_fence_before(); wg_bar(wg); if (wtid == 0) { tc_fence_after();
#define ISSUE_END() umma_commit(bar_mma); }
// A operand written to TMEM by the warpgroup's own tcgen05.st: no shared-memory proxy fence needed
#define ISSUE_BEGIN_TS() tmem_st_wait(); tc_fence_before(); wg_bar(wg); if (wtid == 0) { tc_fence_after();
template <int NK, int N>
__device__ __forceinline__ void issue_gemm(int wg, int wtid, uint32_t a_base, uint32_t b_base, int b_rows, uint32_t d_tmem,
                                           uint32_t accum_first, uint32_t bar_mma) {
    ISSUE_BEGIN()
    mma_group<NK, N>(a_base, b_base, b_rows, d_tmem, accum_first);
    ISSUE_END()
}

// W [N][K] row-major fp32 -> K-major SWIZZLE_128B fp16 image (64 halves = 128 B per swizzle row)
static inline void swizzle_image(const float* W, int N, int K, __half* img) {
    for (int n = 0; n < N; ++n)
        for (int k = 0; k < K; ++k) {
            int atom = k / 64, kk = k % 64, chunk = kk / 8;
            size_t byte = (size_t)atom * N * 128 + (size_t)(n / 8) * 1024 + (size_t)(n % 8) * 128 + (size_t)((chunk ^ (n % 8)) * 16) +
                          (size_t)(kk % 8) * 2;
            img[byte / 2] = __float2half_rn(W[(size_t)n * K + k]);
        }
}

