// score_umma128.cu -- K2 on the tensor cores for d_model = 128 (the reference's 4DN configuration: `dimensions` 128, Modules.py:1029-1095
// is d-generic).  Same math, folding and rounding points as the d_model = 64 fast variant (score_umma2.cu), widened:
//   [y_c | g] = ctx . [fc1c ; W0' fc1c]^T   (M128 N256 K128, A = ctx tile in shared memory)   h = swish(rstd(y_c) g + b0')   -> TMEM A operand
//   z_c       = y_c + h . W1c^T + b1c       (M128 N128 K128, A in TMEM; b1c added by the epilogue)   u' = (z_c rstd + S')^2    -> TMEM A operand
//   c         = u' . C0'^T                  (M128 N64  K128, A in TMEM)                        o = cw1 . swish(c + cb0 / 2) + cb1
// n_head x d_k = n_head x d_v = 128 whatever d_model is (Higashi_wrapper.py:834-845), so the attention in front of MMA 1 -- the
// chunk-major Q / K / V tables, attn_weights, attn_ctx of score_rows.cuh -- is the d_model = 64 code unchanged.  What d_model = 128
// costs is tensor memory: the fused first accumulator is 256 fp32 columns, so a CTA holds TWO row tiles (2 x 256 = 512 columns; the
// d = 64 kernels hold three) and the ctx operand stays in shared memory (no columns left for it).  Shared memory: weight images
// 112 KB (F 64 KB, W1c 32 KB, C0' 16 KB) + 2 x 32 KB ctx tiles = 176 KB.  384 threads: 2 x 128 epilogue threads at 232 registers
// (all 32 S' quads of a row are requested before the wait on MMA 2) + a warpgroup of issue warps (two of its four warps idle:
// setmaxnreg is a warpgroup-wide instruction, a two-warp tail hangs on it).
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "score_rows.cuh"

#define W8_WG 2
#define W8_D 128
#define W8_B_F 0                 // [256 x 128]: rows 0..127 fc1c, rows 128..255 W0' fc1c  (2 K atoms x 32 KB)
#define W8_B_W1 65536            // [128 x 128] W1c                                          (2 K atoms x 16 KB)
#define W8_B_C0 98304            // [64 x 128]  C0' = gamma^2 C0 / 2                         (2 K atoms x 8 KB)
#define W8_WIMG_BYTES 114688
#define W8_SM_A 114688           // W8_WG x 32 KB ctx tiles
#define W8_SM_BAR (W8_SM_A + W8_WG * SM_A_BYTES)
#define W8_SM_TMEM (W8_SM_BAR + 64)
#define W8_SM_TOTAL (W8_SM_TMEM + 16)
#define C8_B0F 0
#define C8_CW1 128
#define C8_CB1 192
#define C8_B1C 200
#define C8_CB0H 328
#define C8_COUNT 392

struct Umma128Params {
    const uint4* QKc; const float4* XSc; const uint4* binVc; const uint4* cellV;
    const float4* binSb;     // [n_grp][32 quads][G]
    const float4* cellSb;    // [n_cells][32 quads]
    const unsigned char* wimg;
    float cst[C8_COUNT];
    const int2* pairs; const float* pair_bias; long long P;
    int n_grp, cell_start, n_batch, tiles_per_cell, act;
    float* out; __half* out16;
};

__device__ __forceinline__ float sumsq32_128(const uint32_t (&r)[32], float2& vA, float2& vB, float2& vC, float2& vD) {
#pragma unroll
    for (int i = 0; i < 16; i += 4) {
        float2 a = f2(r[2 * i], r[2 * i + 1]), b = f2(r[2 * i + 2], r[2 * i + 3]);
        float2 c2 = f2(r[2 * i + 4], r[2 * i + 5]), d2 = f2(r[2 * i + 6], r[2 * i + 7]);
        vA = __ffma2_rn(a, a, vA); vB = __ffma2_rn(b, b, vB); vC = __ffma2_rn(c2, c2, vC); vD = __ffma2_rn(d2, d2, vD);
    }
    return 0.f;
}

__global__ void __launch_bounds__(W8_WG * 128 + 128, 1) score_umma128_kernel(Umma128Params p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, wg = tid >> 7, wtid = tid & 127, warp_in_wg = (tid >> 5) & 3;
    const uint32_t sraw = smem_u32(smem_raw);
    const uint32_t sbase = (sraw + 1023u) & ~1023u;
    unsigned char* smem = smem_raw + (sbase - sraw);
    const float* cst = p.cst;
    const uint32_t bar_w = sbase + W8_SM_BAR;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + W8_SM_TMEM);

    if (tid == 0) {
        mbar_init(bar_w, 1);
        for (int i = 0; i < W8_WG; ++i) mbar_init(sbase + W8_SM_BAR + 8 + i * 8, 1);
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
        mbar_expect_tx(bar_w, W8_WIMG_BYTES);
        tma_bulk_g2s(sbase, p.wimg, 57344, bar_w);                     // two bulk copies (a single one is limited to < 1 MB but keep pieces modest)
        tma_bulk_g2s(sbase + 57344, p.wimg + 57344, W8_WIMG_BYTES - 57344, bar_w);
    }
    const uint32_t tmem_base = *tmem_slot;
    mbar_wait(bar_w, 0);
    const long long n_tiles = (long long)p.n_batch * p.tiles_per_cell;
    const long long n_quads = (n_tiles + W8_WG - 1) / W8_WG;

    if (wg == W8_WG) {
        // ---------------------------------------------------------------- issue warps: warp w serves warpgroup w
        asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");
        const int sw = __shfl_sync(0xffffffffu, warp_in_wg, 0);
        if (sw < W8_WG) {
        const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
        const uint32_t s_ty = tb + (uint32_t)(sw * 256), s_th = s_ty + 128, s_bar = sbase + W8_SM_BAR + 8 + sw * 8;
        const uint32_t s_a = sbase + W8_SM_A + (uint32_t)sw * SM_A_BYTES;
        uint32_t lead_p;
        asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(lead_p));
        const bool lead = lead_p != 0;
        long long n_tok = 0;
        for (long long q = blockIdx.x; q < n_quads; q += gridDim.x) n_tok += 3;
        for (long long i = 0; i < n_tok; ++i) {
            named_bar_sync(5 + sw, 160); tc_fence_after();                            // ctx tile complete in shared memory
            if (lead) { mma_group<8, 256>(s_a, sbase + W8_B_F, 256, s_ty, 0u); umma_commit(s_bar); }
            __syncwarp();
            named_bar_sync(5 + sw, 160); tc_fence_after();                            // h complete in TMEM
            if (lead) { mma_group_ts<8, 128>(s_th, sbase + W8_B_W1, 128, s_ty, 1u); umma_commit(s_bar); }
            __syncwarp();
            named_bar_sync(5 + sw, 160); tc_fence_after();                            // u' complete in TMEM
            if (lead) { mma_group_ts<8, 64>(s_th, sbase + W8_B_C0, 64, s_ty, 0u); umma_commit(s_bar); }
            __syncwarp();
        }
        }
        tc_fence_before();
        __syncthreads();
        return;
    }
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
    const uint32_t bar_mma = sbase + W8_SM_BAR + 8 + wg * 8;
    const uint32_t t_lane = ((uint32_t)(warp_in_wg * 32)) << 16;
    const uint32_t t_y = tmem_base + (uint32_t)(wg * 256), t_h = t_y + 128;
    const uint32_t a_base = sbase + W8_SM_A + wg * SM_A_BYTES;
    const uint32_t a_row_x = a_base + (uint32_t)(wtid >> 3) * 1024u + (uint32_t)(wtid & 7) * (128u + 16u);
    uint32_t phase = 0;
    auto decode = [&](long long q) {
        TileRow tr;
        const bool tile_ok = q * W8_WG + wg < n_tiles;
        const long long tile = tile_ok ? q * W8_WG + wg : n_tiles - 1;
        tr.cb = (int)(tile / p.tiles_per_cell);
        const int pt = (int)(tile - (long long)tr.cb * p.tiles_per_cell);
        tr.pi = (long long)pt * 128 + wtid;
        tr.valid = tile_ok && tr.pi < p.P;
        tr.pr = __ldg(p.pairs + (tr.pi < p.P ? tr.pi : (p.P - 1)));
        tr.bias = tr.valid ? __ldg(p.pair_bias + tr.pi) : 0.f;
        return tr;
    };
    auto weights = [&](const TileRow& tr, int tk, uint32_t (&w)[8]) {
        const uint4* QKs = p.QKc + (size_t)tr.cb * (size_t)(p.n_grp * 32 * HSG_GRP);
        const float4* XSs = p.XSc + (size_t)tr.cb * (size_t)(p.n_grp * 4 * HSG_GRP);
        if (tk == 0) attn_weights<0>(QKs, XSs, (uint32_t)tr.pr.x, (uint32_t)tr.pr.y, w);
        else if (tk == 1) attn_weights<1>(QKs, XSs, (uint32_t)tr.pr.x, (uint32_t)tr.pr.y, w);
        else attn_weights<2>(QKs, XSs, (uint32_t)tr.pr.x, (uint32_t)tr.pr.y, w);
    };
    auto ctx_build = [&](const TileRow& tr, int tk, const uint32_t (&w)[8]) {
        const uint4* cellVrow = p.cellV + (size_t)(p.cell_start + tr.cb) * 16u;
        if (tk == 0) attn_ctx<0>(p.binVc, cellVrow, (uint32_t)tr.pr.x, (uint32_t)tr.pr.y, a_row_x, w);
        else if (tk == 1) attn_ctx<1>(p.binVc, cellVrow, (uint32_t)tr.pr.x, (uint32_t)tr.pr.y, a_row_x, w);
        else attn_ctx<2>(p.binVc, cellVrow, (uint32_t)tr.pr.x, (uint32_t)tr.pr.y, a_row_x, w);
        fence_async_smem();
    };
    auto request = [&]() { tmem_st_wait(); tc_fence_before(); named_bar_arrive(5 + wg, 160); };
    auto wait_mma = [&]() { mbar_wait(bar_mma, phase); phase ^= 1; tc_fence_after(); };

    // after MMA 1: rstd of y_c (columns [0, 128) of t_y), h = swish(rstd g + b0') -> fp16 A operand over the consumed g columns
    auto epi_h = [&]() {
        float2 rs;
        {
            float2 vA = make_float2(0.f, 0.f), vB = vA, vC = vA, vD = vA;
#pragma unroll
            for (int part = 0; part < 4; ++part) {
                uint32_t r[32];
                tmem_ld32(t_y + t_lane + part * 32, r);
                tmem_ld_wait();
                sumsq32_128(r, vA, vB, vC, vD);
            }
            vA = __fadd2_rn(__fadd2_rn(vA, vB), __fadd2_rn(vC, vD));
            const float rstd = rsqrtf((vA.x + vA.y) * (1.0f / 128.0f) + HSG_LN_EPS);
            rs = make_float2(rstd, rstd);
        }
#pragma unroll
        for (int part = 0; part < 4; ++part) {
            uint32_t r[32];
            tmem_ld32(t_h + t_lane + part * 32, r);
            tmem_ld_wait();
            uint32_t hp[16];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const float* bq = cst + C8_B0F + part * 32 + c * 8;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    float2 w = __ffma2_rn(f2(r[c * 8 + 2 * q], r[c * 8 + 2 * q + 1]), rs, make_float2(bq[2 * q], bq[2 * q + 1]));
                    float2 th = make_float2(fast_tanh(w.x), fast_tanh(w.y));
                    float2 hh = __ffma2_rn(w, th, w);
                    hp[c * 4 + q] = pack_h2(hh.x, hh.y);
                }
            }
            tmem_st16(t_h + t_lane + part * 16, hp);          // h columns [16 part, 16 part + 16) overwrite g columns already consumed
        }
    };
    auto load_s = [&](int t, int cell, int2 pr, float4 (&sreg)[32]) {
        if (t == 0) {
            const float4* Srow = p.cellSb + (size_t)cell * 32u;
#pragma unroll
            for (int c = 0; c < 32; ++c) sreg[c] = ldg_nc_v4_volatile(Srow + c);
        } else {
            const float4* Scol = p.binSb + grp_off<32>((uint32_t)(t == 1 ? pr.x : pr.y));
#pragma unroll
            for (int c = 0; c < 32; ++c) sreg[c] = ldg_nc_v4_volatile(Scol + c * HSG_GRP);
        }
    };
    // after MMA 2: z_c = acc + b1c (centred by construction) -> rstd, u' = (z_c rstd + S')^2 -> fp16 A operand over the consumed h columns
    auto epi_u = [&](const float4 (&sreg)[32]) {
        float2 rs;
        {
            float2 vA = make_float2(0.f, 0.f), vB = vA, vC = vA, vD = vA;
#pragma unroll
            for (int part = 0; part < 4; ++part) {
                uint32_t r[32];
                tmem_ld32(t_y + t_lane + part * 32, r);
                tmem_ld_wait();
                const float* bb = cst + C8_B1C + part * 32;
#pragma unroll
                for (int i = 0; i < 16; i += 4) {
                    float2 a = __fadd2_rn(f2(r[2 * i], r[2 * i + 1]), make_float2(bb[2 * i], bb[2 * i + 1]));
                    float2 b = __fadd2_rn(f2(r[2 * i + 2], r[2 * i + 3]), make_float2(bb[2 * i + 2], bb[2 * i + 3]));
                    float2 c2 = __fadd2_rn(f2(r[2 * i + 4], r[2 * i + 5]), make_float2(bb[2 * i + 4], bb[2 * i + 5]));
                    float2 d2 = __fadd2_rn(f2(r[2 * i + 6], r[2 * i + 7]), make_float2(bb[2 * i + 6], bb[2 * i + 7]));
                    vA = __ffma2_rn(a, a, vA); vB = __ffma2_rn(b, b, vB); vC = __ffma2_rn(c2, c2, vC); vD = __ffma2_rn(d2, d2, vD);
                }
            }
            vA = __fadd2_rn(__fadd2_rn(vA, vB), __fadd2_rn(vC, vD));
            const float rstd = rsqrtf((vA.x + vA.y) * (1.0f / 128.0f) + HSG_LN_EPS);
            rs = make_float2(rstd, rstd);
        }
#pragma unroll
        for (int qt = 0; qt < 8; ++qt) {
            uint32_t r[16], up[8];
            tmem_ld16(t_y + t_lane + qt * 16, r);
            tmem_ld_wait();
            const float* bb = cst + C8_B1C + qt * 16;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const float4 s0 = sreg[qt * 4 + 2 * c], s1 = sreg[qt * 4 + 2 * c + 1];
                const float svv[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    float2 z = __fadd2_rn(f2(r[c * 8 + 2 * q], r[c * 8 + 2 * q + 1]), make_float2(bb[c * 8 + 2 * q], bb[c * 8 + 2 * q + 1]));
                    float2 u = __ffma2_rn(z, rs, make_float2(svv[2 * q], svv[2 * q + 1]));
                    u = __fmul2_rn(u, u);
                    up[c * 4 + q] = pack_h2(u.x, u.y);
                }
            }
            tmem_st8(t_h + t_lane + qt * 8, up);
        }
    };
    // after MMA 3: o_t = cw1 . swish(c + cb0 / 2) + cb1
    auto epi_o = [&]() {
        float2 oA = make_float2(0.f, 0.f), oB = oA;
#pragma unroll
        for (int part = 0; part < 2; ++part) {
            uint32_t r[32];
            tmem_ld32(t_y + t_lane + part * 32, r);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
                const float* wq = cst + C8_CW1 + part * 32 + 2 * i;
                const float* cb = cst + C8_CB0H + part * 32 + 2 * i;
                float2 w0 = __fadd2_rn(f2(r[2 * i], r[2 * i + 1]), make_float2(cb[0], cb[1]));
                float2 w1 = __fadd2_rn(f2(r[2 * i + 2], r[2 * i + 3]), make_float2(cb[2], cb[3]));
                float2 h0 = __ffma2_rn(w0, make_float2(fast_tanh(w0.x), fast_tanh(w0.y)), w0);
                float2 h1 = __ffma2_rn(w1, make_float2(fast_tanh(w1.x), fast_tanh(w1.y)), w1);
                oA = __ffma2_rn(h0, make_float2(wq[0], wq[1]), oA);
                oB = __ffma2_rn(h1, make_float2(wq[2], wq[3]), oB);
            }
        }
        return ((oA.x + oA.y) + (oB.x + oB.y)) + cst[C8_CB1];
    };

    // software pipeline over the warpgroup's own chain: the softmax weights of the NEXT token are computed while MMA 1 of the current
    // one is in flight, its ctx tile is built while MMA 3 is in flight (the A tile is free once MMA 1 has completed); across tiles too
    if ((long long)blockIdx.x < n_quads) {
        uint32_t w[8];
        TileRow cur = decode(blockIdx.x);
        weights(cur, 0, w);
        ctx_build(cur, 0, w);
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
                epi_h();
                request();                                                     // MMA 2
                float4 sreg[32];
                load_s(t, cell, cur.pr, sreg);
                wait_mma();
                epi_u(sreg);
                request();                                                     // MMA 3
                if (more) ctx_build(t < 2 ? cur : nxt, t < 2 ? t + 1 : 0, w);  // MMA 3 is in flight, MMA 1 of this token has read the tile
                wait_mma();
                osum += epi_o();
                if (more) request();                                           // MMA 1 of the next token (orders the TMEM reads above)
                else tc_fence_before();
            }
            if (cur.valid) {
                const float m = osum * (1.0f / 3.0f) + cur.bias;
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

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
int umma128_wimg_bytes() { return W8_WIMG_BYTES; }
int umma128_cst_count() { return C8_COUNT; }

// weight images and constants of the d_model = 128 fast variant (the foldings of umma_build_weights, score_umma.cu, at d = 128)
int umma128_build_weights(const float* fc1, const float* w0, const float* b0, const float* pff_g, const float* pff_b,
                          const float* w1, const float* b1, const float* ln1_g, const float* c0, const float* cb0,
                          const float* cw1, const float* cb1, unsigned char* wimg_host, float* cst_host) {
    const int d = W8_D;
    std::vector<float> w0f((size_t)d * d);
    for (int n = 0; n < d; ++n) {
        double acc = b0[n];
        for (int k = 0; k < d; ++k) {
            w0f[(size_t)n * d + k] = 0.5f * w0[(size_t)n * d + k] * pff_g[k];   // LN gain and the swish 1/2 folded into W0
            acc += (double)w0[(size_t)n * d + k] * pff_b[k];                // and the LN bias into b0
        }
        cst_host[C8_B0F + n] = (float)(0.5 * acc);
    }
    memset(wimg_host, 0, W8_WIMG_BYTES);
    for (int i = 0; i < d / 2; ++i) { cst_host[C8_CB0H + i] = 0.5f * cb0[i]; cst_host[C8_CW1 + i] = cw1[i]; }
    cst_host[C8_CB1] = cb1[0];
    // every LayerNorm input is produced centred, and LN_pff . W0' is folded into fc1
    std::vector<double> fc1c((size_t)d * HSG_HD);
    for (int k = 0; k < HSG_HD; ++k) {
        double m = 0;
        for (int n = 0; n < d; ++n) m += fc1[(size_t)n * HSG_HD + k];
        m /= d;
        for (int n = 0; n < d; ++n) fc1c[(size_t)n * HSG_HD + k] = fc1[(size_t)n * HSG_HD + k] - m;
    }
    std::vector<float> comb((size_t)2 * d * HSG_HD);
    for (int n = 0; n < d; ++n)
        for (int k = 0; k < HSG_HD; ++k) {
            comb[(size_t)n * HSG_HD + k] = (float)fc1c[(size_t)n * HSG_HD + k];
            double acc = 0;
            for (int j = 0; j < d; ++j) acc += (double)w0f[(size_t)n * d + j] * fc1c[(size_t)j * HSG_HD + k];
            comb[(size_t)(d + n) * HSG_HD + k] = (float)acc;
        }
    swizzle_image(comb.data(), 2 * d, HSG_HD, reinterpret_cast<__half*>(wimg_host + W8_B_F));
    std::vector<float> w1c((size_t)d * d);
    for (int k = 0; k < d; ++k) {
        double m = 0;
        for (int n = 0; n < d; ++n) m += w1[(size_t)n * d + k];
        m /= d;
        for (int n = 0; n < d; ++n) w1c[(size_t)n * d + k] = (float)(w1[(size_t)n * d + k] - m);
    }
    swizzle_image(w1c.data(), d, d, reinterpret_cast<__half*>(wimg_host + W8_B_W1));
    double mb = 0;
    for (int i = 0; i < d; ++i) mb += b1[i];
    mb /= d;
    for (int i = 0; i < d; ++i) cst_host[C8_B1C + i] = (float)(b1[i] - mb);
    // u' = u / gamma^2 (the S table is divided by gamma), so gamma^2 goes into C0's columns together with the swish 1/2
    std::vector<float> c0h((size_t)(d / 2) * d);
    for (int n = 0; n < d / 2; ++n)
        for (int k = 0; k < d; ++k) c0h[(size_t)n * d + k] = 0.5f * c0[(size_t)n * d + k] * ln1_g[k] * ln1_g[k];
    swizzle_image(c0h.data(), d / 2, d, reinterpret_cast<__half*>(wimg_host + W8_B_C0));
    return 0;
}

int umma128_gain_foldable(const float* ln1_g) {
    for (int i = 0; i < W8_D; ++i) if (!(fabsf(ln1_g[i]) >= 0.05f)) return 0;
    return 1;
}

// fp32 static tables -> 16-bit V rows (chunk-major groups for bins) and S' = (beta - S) / gamma in d / 4 quads per row
__global__ void convert_static128_kernel(const float* __restrict__ V, const float* __restrict__ S, const float* __restrict__ ln1_b,
                                         const float* __restrict__ ln1_g, long long rows, int d, int grouped,
                                         __half* __restrict__ V16, float* __restrict__ Sb) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int qpb = d / 4;
    if (i < rows * HSG_HD) {
        const long long r = i / HSG_HD; const int col = (int)(i - r * HSG_HD);
        const long long dst = grouped ? (((r >> 5) * 16 + (col >> 3)) * HSG_GRP + (r & 31)) * 8 + (col & 7) : i;
        V16[dst] = __float2half_rn(V[i]);
    }
    if (i < rows * d) {
        const long long r = i / d; const int f = (int)(i - r * d);
        const long long dst = grouped ? (((r >> 5) * qpb + (f >> 2)) * HSG_GRP + (r & 31)) * 4 + (f & 3) : i;
        Sb[dst] = (ln1_b[f] - S[i]) / ln1_g[f];
    }
}

int launch_convert_static128(hsg_ctx* c, const float* V, const float* S, long long rows, int grouped, void* V16, float* Sb, cudaStream_t s) {
    const long long n = rows * std::max(HSG_HD, c->d);
    convert_static128_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(V, S, c->w[HSG_W_LN1_B], c->w[HSG_W_LN1_G], rows, c->d, grouped,
                                                                          (__half*)V16, Sb);
    c->launches++;
    HSG_CUDA(cudaGetLastError());
    return 0;
}

int launch_score_pairs_umma128(hsg_ctx* c, int cell_start, int n_batch, float* out, void* out16, cudaStream_t s) {
    if (c->P == 0 || n_batch == 0) return 0;
    Umma128Params p;
    p.QKc = (const uint4*)c->QK16; p.XSc = (const float4*)c->XS; p.binVc = (const uint4*)c->binV16; p.cellV = (const uint4*)c->cellV16;
    p.binSb = (const float4*)c->binSb; p.cellSb = (const float4*)c->cellSb; p.wimg = c->wimg;
    memcpy(p.cst, c->h_cst, sizeof(p.cst));
    p.pairs = c->pairs; p.pair_bias = c->pair_bias; p.P = c->P;
    p.n_grp = (c->n_bins + HSG_GRP - 1) / HSG_GRP; p.cell_start = cell_start; p.n_batch = n_batch;
    p.tiles_per_cell = (int)((c->P + 127) / 128); p.act = c->act; p.out = out; p.out16 = (__half*)out16;
    const long long q8 = ((long long)n_batch * p.tiles_per_cell + W8_WG - 1) / W8_WG;
    const int g8 = (int)(q8 < c->n_sm ? q8 : c->n_sm);
    static bool attr8[64] = {};
    if (!attr8[c->device & 63]) {
        HSG_CUDA(cudaFuncSetAttribute(score_umma128_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, W8_SM_TOTAL + 1024));
        attr8[c->device & 63] = true;
    }
    score_umma128_kernel<<<g8, W8_WG * 128 + 128, W8_SM_TOTAL + 1024, s>>>(p);
    c->launches++;
    HSG_CUDA(cudaGetLastError());
    return 0;
}
