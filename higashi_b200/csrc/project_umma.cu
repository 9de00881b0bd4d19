// project_umma.cu -- K1b (tensor-core mode): per-(cell, bin) token projection on tcgen05 / TMEM (sm_100a).
//
//   dyn = swish(nn([neigh | E_self]))                         (GraphSageEncoder_with_weights, Modules.py:1480)
//   Q   = w_qs . layer_norm1(dyn),  K = w_ks . layer_norm2(dyn)   (MultiHeadAttention, Modules.py:1035-1044)
//   qcK = Q_cell,h . K_b,h / 4,  kcQ = Q_b,h . K_cell,h / 4       (the cell-token logits, hoisted out of the scorer)
//
// One CTA (128 threads, thread r <-> token r <-> TMEM lane r) per tile of 128 consecutive tokens, two CTAs per SM:
//   * A1 = [neigh | E_self] is converted to fp16 and written by the CTA itself into the canonical K-major SWIZZLE_128B
//     layout (coalesced 8-lanes-per-row reads); the weights arrive once per CTA with one TMA bulk copy;
//   * GEMM1  D1[128 x 64]  = A1 . G^T            (K = 128)  -> epilogue: + bias, swish, ONE set of LayerNorm statistics
//   * GEMM2  D2[128 x 256] = xhat . [Wq' ; Wk']^T (K = 64)  with the two LayerNorm gains folded into the weights and the
//     LayerNorm biases folded into the output bias: Q|K come out of a single N = 256 MMA;
//   * epilogue 2 adds the bias, accumulates the cell cross terms against the cell's K / Q vectors (staged in shared memory),
//     rounds to fp16 (Q pre-scaled by log2(e)/4, chunks in HSG_CHUNK_POS order, see score_umma.cu) into a swizzled staging tile and the CTA copies the tile
//     out with fully coalesced 128-bit stores (a tile's Q|K rows are one contiguous 64 KB block of the table).
// fp16 operands / fp32 accumulation: measured effect on the final score 1.9e-4 -> 2.0e-4 (scripts/bf16_emulation.py).
// Templated on d_model: 64 (K = 128 / 64, two CTAs per SM) and 128 (the reference's 4DN configuration: GEMM1 [128 x 128] over
// K = 256, GEMM2 [128 x 256] over K = 128; 128 KB of weight images + a 64 KB A tile -> one CTA per SM; chunk-major tables only).
#include "umma_common.cuh"

// shared-memory map as a function of d_model D (64: two CTAs per SM; 128: one):
//   [0, PJ_B_QK)        G image   [D x 2D]  fp16, 2D/64 K-atoms of D x 128 bytes
//   [PJ_B_QK, PJ_WIMG)  Q|K image [256 x D] fp16, D/64 K-atoms of 32 KB
//   [PJ_A, PJ_CELL)     A1 (2D/64 atoms of 16 KB); A2 = its first D/64 atoms; (row-major tables only) output staging reuses it
#define PJ_B_G 0
__host__ __device__ constexpr int pj_b_qk(int D) { return D * 2 * D * 2; }
__host__ __device__ constexpr int pj_wimg(int D) { return pj_b_qk(D) + 256 * D * 2; }
__host__ __device__ constexpr int pj_a(int D) { return pj_wimg(D); }
__host__ __device__ constexpr int pj_cell(int D) { return pj_a(D) + (2 * D / 64) * 16384; }
#define PJ_SLOTS 4               // PJ_SLOTS x 256 floats: [cell K (128) | cell Q (128)] of the cells this tile touches
__host__ __device__ constexpr int pj_bar(int D) { return pj_cell(D) + PJ_SLOTS * 1024; }
__host__ __device__ constexpr int pj_tmem(int D) { return pj_bar(D) + 16; }
__host__ __device__ constexpr int pj_total(int D) { return pj_tmem(D) + 16; }
#define PJC_GB 0                 // constants: sage bias [D], Q bias [128], K bias [128]
__host__ __device__ constexpr int pjc_bq(int D) { return D; }
__host__ __device__ constexpr int pjc_bk(int D) { return D + 128; }
#define PJC_MAX 384

struct ProjParams {
    const float* neigh; const float* E_bin; const float* cellQ; const float* cellK;
    const unsigned char* wimg;
    float cst[PJC_MAX];
    __half* QK16; float* XS;
    long long n_tok;
    int n_bins, cell_start, n_tiles;
    int chunk_major;         // 1: the grouped chunk-major tables of score_umma2.cu: QK16 [slot][n_grp][32 chunks][32 bins], XS [slot][n_grp][4 quads][32 bins]
};

template <int D>
__global__ void __launch_bounds__(128, D == 64 ? 2 : 1) project_umma_kernel(ProjParams p) {
    constexpr int PJ_B_QK = pj_b_qk(D), PJ_WIMG_BYTES = pj_wimg(D), PJ_A = pj_a(D), PJ_CELL = pj_cell(D), PJ_BAR = pj_bar(D), PJ_TMEM = pj_tmem(D);
    constexpr int PJC_BQ = pjc_bq(D), PJC_BK = pjc_bk(D);
    constexpr int DQ = D / 64;                   // 8-chunk groups per operand half / K-atoms of the second GEMM
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t sraw = smem_u32(smem_raw);
    const uint32_t sbase = (sraw + 1023u) & ~1023u;
    unsigned char* smem = smem_raw + (sbase - sraw);
    const float* cst = p.cst;
    const uint32_t bar_w = sbase + PJ_BAR, bar_mma = sbase + PJ_BAR + 8;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + PJ_TMEM);
    float* s_cell = reinterpret_cast<float*>(smem + PJ_CELL);

    if (tid == 0) {
        mbar_init(bar_w, 1);
        mbar_init(bar_mma, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 32) {     // 256 TMEM columns per CTA (two CTAs per SM)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (tid == 0) {
        mbar_expect_tx(bar_w, PJ_WIMG_BYTES);
        tma_bulk_g2s(sbase, p.wimg, PJ_WIMG_BYTES, bar_w);
    }
    const uint32_t t0 = *tmem_slot;
    const uint32_t t_lane = ((uint32_t)(warp * 32)) << 16;
    const uint32_t a_base = sbase + PJ_A;
    mbar_wait(bar_w, 0);

    uint32_t phase = 0;
    const int j = lane & 7, rsub = lane >> 3;
    const size_t n_grp = (size_t)((p.n_bins + HSG_GRP - 1) / HSG_GRP);
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        const long long tok0 = (long long)tile * 128;
        const int cb0 = (int)(tok0 / p.n_bins);
        const int g0 = (int)(tok0 - (long long)cb0 * p.n_bins);
        {   // the K / Q vectors of the cells this tile touches
            const long long last = (tok0 + 127 < p.n_tok ? tok0 + 127 : p.n_tok - 1);
            const int nslots = (int)(last / p.n_bins) - cb0 + 1;
            for (int i = tid; i < nslots * 256; i += 128) {
                const int slot = i >> 8, e = i & 255;
                const long long cell = (long long)p.cell_start + cb0 + slot;
                s_cell[i] = (e < 128) ? __ldg(p.cellK + cell * 128 + e) : __ldg(p.cellQ + cell * 128 + (e - 128));
            }
        }
        // ---------------------------------------------------------------- A1 = [neigh | E_self] -> fp16, swizzled
#pragma unroll 2
        for (int it = 0; it < 8; ++it) {
            const int rw = warp * 32 + it * 4 + rsub;
            long long tok = tok0 + rw;
            if (tok >= p.n_tok) tok = p.n_tok - 1;
            int g = g0 + (int)(tok - tok0);
            while (g >= p.n_bins) g -= p.n_bins;
#pragma unroll
            for (int q = 0; q < DQ; ++q) {
                const float4* sn = reinterpret_cast<const float4*>(p.neigh + tok * D + (q * 8 + j) * 8);
                const float4* se = reinterpret_cast<const float4*>(p.E_bin + (long long)g * D + (q * 8 + j) * 8);
                const float4 n0 = __ldg(sn), n1 = __ldg(sn + 1), e0 = __ldg(se), e1 = __ldg(se + 1);
                sts128(a_chunk_addr(a_base, rw, q * 8 + j), pack_h2(n0.x, n0.y), pack_h2(n0.z, n0.w), pack_h2(n1.x, n1.y), pack_h2(n1.z, n1.w));
                sts128(a_chunk_addr(a_base, rw, D / 8 + q * 8 + j), pack_h2(e0.x, e0.y), pack_h2(e0.z, e0.w), pack_h2(e1.x, e1.y), pack_h2(e1.z, e1.w));
            }
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            mma_group<2 * D / 16, D>(a_base, sbase + PJ_B_G, D, t0, 0u);
            umma_commit(bar_mma);
        }
        mbar_wait(bar_mma, phase); phase ^= 1;
        tc_fence_after();

        // ---------------------------------------------------------------- dyn = swish(. + b) ; xhat = LayerNorm statistics only
        {
            float x[D];
            float s = 0.f;
#pragma unroll
            for (int cc = 0; cc < D / 32; ++cc) {
                uint32_t r[32];
                tmem_ld32(t0 + t_lane + cc * 32, r);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const float v = __uint_as_float(r[i]) + cst[PJC_GB + cc * 32 + i];
                    const float d = v * rcp_approx(1.0f + ex2_approx(v * -1.4426950408889634f));      // v * sigmoid(v)
                    x[cc * 32 + i] = d;
                    s += d;
                }
            }
            const float mean = s * (1.0f / (float)D);
            float vs = 0.f;
#pragma unroll
            for (int i = 0; i < D; ++i) { x[i] -= mean; vs = fmaf(x[i], x[i], vs); }
            const float rstd = rsqrtf(vs * (1.0f / (float)D) + HSG_LN_EPS);
#pragma unroll
            for (int c = 0; c < D / 8; ++c)
                sts128(a_chunk_addr(a_base, tid, c), pack_h2(x[c * 8] * rstd, x[c * 8 + 1] * rstd), pack_h2(x[c * 8 + 2] * rstd, x[c * 8 + 3] * rstd),
                       pack_h2(x[c * 8 + 4] * rstd, x[c * 8 + 5] * rstd), pack_h2(x[c * 8 + 6] * rstd, x[c * 8 + 7] * rstd));
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            mma_group<D / 16, 256>(a_base, sbase + PJ_B_QK, 256, t0, 0u);
            umma_commit(bar_mma);
        }
        mbar_wait(bar_mma, phase); phase ^= 1;
        tc_fence_after();

        // ---------------------------------------------------------------- Q | K: bias, cross terms, fp16, coalesced copy-out
        const long long tokr = tok0 + tid;
        int slot_r = 0, g_r = 0;
        {
            int g = g0 + tid;
            while (g >= p.n_bins) { g -= p.n_bins; ++slot_r; }
            g_r = g;
            if (slot_r >= PJ_SLOTS) slot_r = PJ_SLOTS - 1;         // rows past n_tok only (never stored)
        }
        float xs[16];
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const float* cv = s_cell + slot_r * 256 + half * 128;          // Q half pairs with the cell's K, K half with its Q
            const float* bias = cst + (half == 0 ? PJC_BQ : PJC_BK);
            const float qs = (half == 0) ? HSG_L2E_QUARTER : 1.0f;
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
                uint32_t r[32];
                tmem_ld32(t0 + t_lane + half * 128 + cc * 32, r);
                tmem_ld_wait();
                float crA = 0.f, crB = 0.f;
#pragma unroll
                for (int c8 = 0; c8 < 4; ++c8) {
                    const float4 k0 = *reinterpret_cast<const float4*>(cv + cc * 32 + c8 * 8);
                    const float4 k1 = *reinterpret_cast<const float4*>(cv + cc * 32 + c8 * 8 + 4);
                    const float kk[8] = {k0.x, k0.y, k0.z, k0.w, k1.x, k1.y, k1.z, k1.w};
                    float v[8];
                    float cr = 0.f;
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        v[q] = __uint_as_float(r[c8 * 8 + q]) + bias[cc * 32 + c8 * 8 + q];
                        cr = fmaf(v[q], kk[q], cr);
                    }
                    if (c8 < 2) crA += cr; else crB += cr;
                    if (p.chunk_major) {
                        // thread = token: consecutive threads write consecutive 16-byte chunks of table[chunk][bin] -- coalesced, no staging
                        if (tokr < p.n_tok)
                            reinterpret_cast<uint4*>(p.QK16)[(((size_t)(cb0 + slot_r) * n_grp + (size_t)(g_r >> 5)) * 32u + (size_t)(half * 16 + cc * 4 + c8)) * HSG_GRP + (size_t)(g_r & 31)] =
                                make_uint4(pack_h2(v[0] * qs, v[1] * qs), pack_h2(v[2] * qs, v[3] * qs), pack_h2(v[4] * qs, v[5] * qs), pack_h2(v[6] * qs, v[7] * qs));
                    } else
                    sts128(a_chunk_addr(a_base, tid, HSG_CHUNK_POS(cc * 4 + c8)), pack_h2(v[0] * qs, v[1] * qs), pack_h2(v[2] * qs, v[3] * qs),
                           pack_h2(v[4] * qs, v[5] * qs), pack_h2(v[6] * qs, v[7] * qs));
                }
                // half 0: kcQ = Q_b . K_cell -> XS[8 + h] ; half 1: qcK = Q_cell . K_b -> XS[h]
                xs[(half == 0 ? 8 : 0) + HSG_XS_POS(cc * 2)] = crA * HSG_L2E_QUARTER;          // heads 2 cc and 2 cc + 1 of this 32-column chunk
                xs[(half == 0 ? 8 : 0) + HSG_XS_POS(cc * 2 + 1)] = crB * HSG_L2E_QUARTER;
            }
            if (p.chunk_major) continue;                      // rows already stored (block-uniform)
            __syncthreads();
            // copy-out: lane j of a row group moves chunks j and 8 + j (8 lanes = one contiguous 128-byte line)
#pragma unroll 2
            for (int it = 0; it < 8; ++it) {
                const int rw = warp * 32 + it * 4 + rsub;
                const long long tok = tok0 + rw;
                if (tok < p.n_tok) {
                    uint4 a, b;
                    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w) : "r"(a_chunk_addr(a_base, rw, j)));
                    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "r"(a_chunk_addr(a_base, rw, 8 + j)));
                    uint4* dst = reinterpret_cast<uint4*>(p.QK16 + tok * 256 + half * 128);
                    dst[j] = a;
                    dst[8 + j] = b;
                }
            }
            __syncthreads();
        }
        if (p.chunk_major) {
            if (tokr < p.n_tok) {
                float4* xd = reinterpret_cast<float4*>(p.XS) + ((size_t)(cb0 + slot_r) * n_grp + (size_t)(g_r >> 5)) * (4u * HSG_GRP) + (size_t)(g_r & 31);
                xd[0] = make_float4(xs[0], xs[1], xs[2], xs[3]); xd[HSG_GRP] = make_float4(xs[4], xs[5], xs[6], xs[7]);
                xd[2 * HSG_GRP] = make_float4(xs[8], xs[9], xs[10], xs[11]); xd[3 * HSG_GRP] = make_float4(xs[12], xs[13], xs[14], xs[15]);
            }
        } else if (tokr < p.n_tok) {
            float4* xd = reinterpret_cast<float4*>(p.XS + tokr * 16);
            xd[0] = make_float4(xs[0], xs[1], xs[2], xs[3]); xd[1] = make_float4(xs[4], xs[5], xs[6], xs[7]);
            xd[2] = make_float4(xs[8], xs[9], xs[10], xs[11]); xd[3] = make_float4(xs[12], xs[13], xs[14], xs[15]);
        }
        if (p.chunk_major) __syncthreads();      // the next tile overwrites s_cell: every thread must be done with its cross terms
        tc_fence_before();      // TMEM reads done before the next tile's MMA overwrites the columns (ordered by its __syncthreads)
    }

    tc_fence_before();
    __syncthreads();
    if (tid < 32) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(t0), "r"(256) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
int proj_wimg_bytes(int d) { return d == 128 ? pj_wimg(128) : pj_wimg(64); }
int proj_cst_count() { return PJC_MAX; }
int proj_min_bins() { return (127 + PJ_SLOTS - 2) / (PJ_SLOTS - 1); }        // a 128-token tile must touch at most PJ_SLOTS cells

// sage [d][2d], wq / wk [128][d], LayerNorm gains / biases [d]
int proj_build_weights(int d, const float* sage_w, const float* sage_b, const float* wq, const float* wk, const float* g1, const float* b1,
                       const float* g2, const float* b2, unsigned char* wimg_host, float* cst_host) {
    const int b_qk = d == 128 ? pj_b_qk(128) : pj_b_qk(64), bq0 = d, bk0 = d + 128;
    memset(wimg_host, 0, (size_t)proj_wimg_bytes(d));
    swizzle_image(sage_w, d, 2 * d, reinterpret_cast<__half*>(wimg_host + PJ_B_G));
    std::vector<float> qk((size_t)2 * HSG_HD * d);
    for (int n = 0; n < HSG_HD; ++n) {
        double bq = 0, bk = 0;
        for (int k = 0; k < d; ++k) {
            qk[(size_t)n * d + k] = wq[(size_t)n * d + k] * g1[k];
            qk[(size_t)(HSG_HD + n) * d + k] = wk[(size_t)n * d + k] * g2[k];
            bq += (double)wq[(size_t)n * d + k] * b1[k];
            bk += (double)wk[(size_t)n * d + k] * b2[k];
        }
        cst_host[bq0 + n] = (float)bq;
        cst_host[bk0 + n] = (float)bk;
    }
    swizzle_image(qk.data(), 2 * HSG_HD, d, reinterpret_cast<__half*>(wimg_host + b_qk));
    for (int i = 0; i < d; ++i) cst_host[PJC_GB + i] = sage_b[i];
    return 0;
}

template <int D>
static int launch_tpu(hsg_ctx* c, const ProjParams& p, cudaStream_t s) {
    static bool attr_set[64] = {};               // per device: the attribute belongs to the device's function instance
    if (!attr_set[c->device & 63]) {
        HSG_CUDA(cudaFuncSetAttribute(project_umma_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, pj_total(D) + 1024));
        attr_set[c->device & 63] = true;
    }
    const int per_sm = D == 64 ? 2 : 1;
    const int grid = p.n_tiles < per_sm * c->n_sm ? p.n_tiles : per_sm * c->n_sm;
    project_umma_kernel<D><<<grid, 128, pj_total(D) + 1024, s>>>(p);
    c->launches++;
    HSG_CUDA(cudaGetLastError());
    return 0;
}

int launch_token_project_umma(hsg_ctx* c, int cell_start, int n_batch, cudaStream_t s) {
    const long long n_tok = (long long)n_batch * c->n_bins;
    if (n_tok == 0) return 0;
    ProjParams p;
    p.neigh = c->neigh; p.E_bin = c->E_bin; p.cellQ = c->cellQ; p.cellK = c->cellK; p.wimg = c->pj_wimg;
    memcpy(p.cst, c->h_pj_cst, sizeof(p.cst));
    p.QK16 = (__half*)c->QK16; p.XS = c->XS; p.n_tok = n_tok; p.n_bins = c->n_bins; p.cell_start = cell_start;
    p.n_tiles = (int)((n_tok + 127) / 128);
    p.chunk_major = c->k2v2 ? 1 : 0;
    HSG_CHECK(c->d == 64 || (c->d == 128 && p.chunk_major), "tensor-core projection: d_model 64, or 128 with chunk-major tables");
    return c->d == 128 ? launch_tpu<128>(c, p, s) : launch_tpu<64>(c, p, s);
}
