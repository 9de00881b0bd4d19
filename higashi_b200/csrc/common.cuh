// common.cuh -- internal declarations shared by the libhsg translation units.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <string>
#include <vector>

#include "../../include/hsg.h"

#define HSG_NHEAD 8
#define HSG_DK 16
#define HSG_HD 128                 // n_head * d_k = n_head * d_v   (Higashi_wrapper.py:834-845)
#define HSG_LN_EPS 1e-5f
// cell cross-term table XS [token][16]: floats 0..7 = qcK[head], 8..15 = kcQ[head]
#define HSG_XS_POS(h) (h)
// 16-bit Q / K / V rows (128 halves = 16 chunks of 16 bytes; head h = logical chunks 2h, 2h + 1) are stored with the chunks
// permuted: position h holds the first and position 8 + h the second half of head h.  Scorer lane j of a row reads positions j
// and 8 + j: the 8 lanes of a row still cover two fully used 128-byte lines per load, and every lane owns ONE complete head
// (no cross-lane reduction of the logits, one softmax per lane).  The ctx A tile keeps the permuted order; the K columns of
// the fc1 weight image are permuted to match (umma_build_weights).
#define HSG_CHUNK_POS(c) ((((c) >> 1)) + (((c) & 1) << 3))
#define HSG_COL_POS(col) (HSG_CHUNK_POS((col) >> 3) * 8 + ((col) & 7))
// high-accuracy scorer tables: a 128-float Q / K / V row is stored as [4 quads][8 heads][4 floats], so that the 8 lanes of a row
// (lane = head) read one contiguous 128-byte line per 128-bit load.  Column c = 16 h + 4 e + r  ->  32 e + 4 h + r
#define HSG_HI_POS(c) ((((c) >> 2) & 3) * 32 + ((c) >> 4) * 4 + ((c) & 3))
// score_umma2.cu tables: bins in groups of 32, chunk-major inside a group (table[bin >> 5][chunk][bin & 31])
#define HSG_GRP 32
#define HSG_L2E_QUARTER 0.36067376022224085f      // log2(e) / 4: attention logits of the tensor-core scorer are kept in log2 units
#define HSG_PDF0 0.3989422804014327f   // scipy.stats.norm.pdf(0)   (Impute.py:19)
// the streaming gather copies 16-byte aligned spans of row_ptr / col / val: the device arrays carry this many spare elements
#define HSG_CSR_PAD 4

void hsg_set_error(const std::string& msg);

// In-kernel bounds checks for a debug build (`make DEBUG_ASSERTS=1`): compute-sanitizer is not available on the GPU pool, so the
// indices a wrong table / list / tile entry would send out of range are checked in the kernels themselves; a failed check prints
// its location and traps (the launch fails, nothing hangs).  Compiled out of the product build.
#ifndef HSG_DEBUG_ASSERTS
#define HSG_DEBUG_ASSERTS 0
#endif
#if HSG_DEBUG_ASSERTS
#include <cstdio>
#define HSG_DASSERT(cond)                                                                                         \
    do {                                                                                                          \
        if (!(cond)) { printf("HSG_DASSERT failed: %s at %s:%d (block %d thread %d)\n", #cond, __FILE__, __LINE__, \
                              (int)blockIdx.x, (int)threadIdx.x); __trap(); }                                     \
    } while (0)
#else
#define HSG_DASSERT(cond) do { } while (0)
#endif

#define HSG_CUDA(call)                                                                         \
    do {                                                                                       \
        cudaError_t e__ = (call);                                                              \
        if (e__ != cudaSuccess) {                                                              \
            hsg_set_error(std::string(#call) + " failed: " + cudaGetErrorString(e__) + " at " + \
                          __FILE__ + ":" + std::to_string(__LINE__));                          \
            return 1;                                                                          \
        }                                                                                      \
    } while (0)

#define HSG_CHECK(cond, msg)                                                                   \
    do {                                                                                       \
        if (!(cond)) {                                                                         \
            hsg_set_error(std::string(msg) + " (" + #cond + ")");                              \
            return 2;                                                                          \
        }                                                                                      \
    } while (0)

// Weights derived once per hsg_prepare (transposed so that the output index is contiguous).
struct DerivedWeights {
    float* sageT = nullptr;    // [2d][d]      encode1.dynamic_nn.nn.weight^T
    float* wqT = nullptr;      // [d][128]
    float* wkT = nullptr;      // [d][128]
    float* wvT = nullptr;      // [d][128]
    float* fc1T = nullptr;     // [128][d]
    float* fc2T = nullptr;     // [128][d]
    float* w0T = nullptr;      // [d][d]       pff_n1 layer 0
    float* w1T = nullptr;      // [d][d]       pff_n1 layer 1
    float* c0T = nullptr;      // [d][d/2]     pff_classifier layer 0
};

struct hsg_ctx {
    int device = 0;
    int d = 0, n_cells = 0, n_chrom = 0, cf = 0, n_bins = 0;
    std::vector<int> bins, bin_off;          // host copies
    int* d_bin_off = nullptr;                // [C+1]
    int* d_bin_chrom = nullptr;              // [n_bins]

    float* w[HSG_W_COUNT] = {nullptr};
    int64_t wcount[HSG_W_COUNT] = {0};

    float* E_cell = nullptr;                 // [n_cells, d]   off-hooked cell table
    float* E_bin = nullptr;                  // [n_bins, d]    off-hooked bin tables, chromosomes concatenated
    std::vector<char> table_set;             // per branch

    // contacts: packed CSR
    int64_t* row_ptr = nullptr; int32_t* col = nullptr; float* val = nullptr;
    int64_t n_rows = 0, nnz = 0;
    int64_t cap_rows = 0, cap_nnz = 0;          // allocated capacity of the contact arrays (re-uploads reuse the buffers)
    // second set of contact buffers: hsg_stage_contacts fills it on its own stream while the current set is in use,
    // hsg_commit_contacts swaps the two (streaming datasets: the upload of the next block overlaps the imputation of this one)
    int64_t* row_ptr2 = nullptr; int32_t* col2 = nullptr; float* val2 = nullptr;
    int64_t cap_rows2 = 0, cap_nnz2 = 0, staged_rows = 0, staged_nnz = 0;
    cudaStream_t h2d_stream = nullptr; cudaEvent_t staged_ev = nullptr; bool staged = false;
    // hsg_refresh_contacts: per-cell extents of the installed matrix (host), the event its copies signal and the event the last
    // imputation call recorded (copies wait for it: they may overwrite rows that call still reads)
    std::vector<int64_t> h_cell_ptr, h_cell_ptr2;
    std::vector<double> chrom_row_avg, chrom_row_avg2;   // mean entries per row on each chromosome (sizes the streaming gather's stages)
    cudaEvent_t refresh_ev = nullptr, work_ev = nullptr; bool refresh_pending = false, work_pending = false;
    // neighbours
    int32_t* nbr = nullptr; float* nbr_w = nullptr; int k1 = 1; bool weighted = false;

    // pairs
    int64_t P = 0;
    int2* pairs = nullptr;                   // [P] genome-wide 0-based (bin_i, bin_j)
    float* pair_bias = nullptr;              // [P] extra_proba2 on the pair (cell-invariant, only_model=True)
    std::vector<int> sel_chroms;
    std::vector<int64_t> sel_P, sel_off;
    std::vector<int2> h_pairs;

    // prepared state
    bool prepared = false;
    int precision = HSG_PREC_FP32, ltr = 0, act = HSG_ACT_SIGMOID;
    DerivedWeights dw;
    float* cellQ = nullptr; float* cellK = nullptr; float* cellV = nullptr; float* cellS = nullptr;  // [n_cells, 128|d]
    float* binV = nullptr;  float* binS = nullptr;                                                   // [n_bins, 128|d]

    // per-batch workspace
    int batch_cells = 0;
    float* neigh = nullptr;                  // [batch, n_bins, d]
    int* ovf_list = nullptr; int64_t ovf_cap = 0;   // K1a overflow tokens (+ counter in the last element)
    bool gather_sparse = true;               // sparse-row K1a fast path enabled (HSG_GATHER_SPARSE=0 disables it)
    bool gather_hash = true;                 // hash-merge K1a for long chromosomes enabled (HSG_GATHER_HASH=0 disables it)
    bool gather_chrom = true;                // chromosome-block K1a for dense rows enabled (HSG_GATHER_CHROM=0 disables it)
    bool gather_stream = false;              // streaming K1a (gather_stream.cu): opt-in with HSG_GATHER_STREAM=1 (measured slower than the register kernel, see its header)
    int gather_stream_cap = 0, gather_stream_lcap = 0;   // test hooks: force a small stage / candidate-list capacity (HSG_GATHER_STREAM_CAP / _LCAP)
    float* QK = nullptr;                     // fp32 [qk_cells, n_bins, 2, 128] (fp32 mode: batch; tensor-core mode: 1 cell for hsg_fix_cell)
    // tensor-core mode (HSG_PREC_TC16): 16-bit token tables + pre-swizzled weight images
    void* QK16 = nullptr;                    // half [batch, n_bins, 256]
    float* XS = nullptr;                     // [batch, n_bins, 16] cross terms with the cell token
    void* binV16 = nullptr; float* binSb = nullptr; void* cellV16 = nullptr; float* cellSb = nullptr;
    float* binV32p = nullptr; float* cellV32p = nullptr;   // high-accuracy variant: fp32 V rows in (quad e, head h) order (HSG_HI_POS)
    unsigned char* wimg = nullptr; float* cst = nullptr; float h_cst[512] = {0};
    unsigned char* pj_wimg = nullptr; float h_pj_cst[384] = {0}; bool use_proj_umma = false;   // tensor-core K1b (project_umma.cu)
    bool use_umma = false;
    // generation 4 of the fast scorer (score_umma4.cu): tile table of the pair list, P tables of the bins and cells
    int* tiles4 = nullptr; int n_tiles4 = 0; void* binP = nullptr; void* cellP = nullptr; bool k2v4 = false; int k2gen = 4; bool k2_dense_tiles = false;
    bool k2d128 = false;                     // d_model = 128 runs score_umma128.cu (chunk-major tables of score_umma2.cu, 32 S' quads per row)
    bool k2v2 = false;                       // fast tensor-core variant runs score_umma2.cu: chunk-major QK16 / XS / binV16 tables
    bool umma_hi = false;                    // high-accuracy tensor-core variant (fp32 tables, hi/lo operand splits)
    int n_sm = 148;
    int fixed_cell = -1;                     // 0-based cell whose tokens sit in QK slot 0 (hsg_fix_cell)

    cudaStream_t stream = nullptr;           // internal work stream
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev[8] = {nullptr};
    float* stage_out[2] = {nullptr, nullptr}; size_t stage_bytes = 0;
    cudaEvent_t stage_ev[2] = {nullptr, nullptr};
    bool profiling = false;
    float last_ms[4] = {0, 0, 0, 0};
    int64_t launches = 0;
    void* ins_ws = nullptr; size_t ins_ws_bytes = 0;      // hsg_insulation workspace (grow-only)
};

// ---- kernel launchers (definitions in the .cu files) ---------------------------------------
int launch_embed_table(hsg_ctx* c, const float* x, int64_t rows, int in_dim, const float* w, const float* b,
                       int apply_swish, float* out, cudaStream_t s);
int launch_static_tokens(hsg_ctx* c, const float* E, int64_t rows, float* Q, float* K, float* V, float* S, cudaStream_t s);
int launch_pair_bias(hsg_ctx* c, cudaStream_t s);
int launch_gather(hsg_ctx* c, const int* cells_dev_unused, int cell_start, int n_cells_batch, cudaStream_t s);
int try_launch_gather_stream(hsg_ctx* c, int cell_start, int n_cells_batch, int max_nc, int* ovf_count, cudaStream_t s, bool* done);
int launch_token_project(hsg_ctx* c, int cell_start, int n_cells_batch, int mode, cudaStream_t s);   // mode 0: fp32 QK, 1: fp16 QK + XS, 2: fp32 QK + XS
int launch_score_pairs_umma(hsg_ctx* c, int cell_start, int n_cells_batch, float* out, cudaStream_t s);
int launch_convert_static(hsg_ctx* c, const float* V, const float* S, long long rows, void* V16, float* Sb, float* V32p, cudaStream_t s);
int launch_convert_static2(hsg_ctx* c, const float* V, const float* S, long long rows, int chunk_major, void* V16, float* Sb, cudaStream_t s);
int launch_score_pairs_umma2(hsg_ctx* c, int cell_start, int n_cells_batch, float* out, void* out16, cudaStream_t s);
int launch_score_pairs_umma4(hsg_ctx* c, int cell_start, int n_cells_batch, float* out, void* out16, cudaStream_t s);
int launch_score_pairs_umma128(hsg_ctx* c, int cell_start, int n_cells_batch, float* out, void* out16, cudaStream_t s);
int launch_convert_static128(hsg_ctx* c, const float* V, const float* S, long long rows, int grouped, void* V16, float* Sb, cudaStream_t s);
int umma128_build_weights(const float* fc1, const float* w0, const float* b0, const float* pff_g, const float* pff_b,
                          const float* w1, const float* b1, const float* ln1_g, const float* c0, const float* cb0,
                          const float* cw1, const float* cb1, unsigned char* wimg_host, float* cst_host);
int umma128_wimg_bytes();
int umma128_cst_count();
int umma128_gain_foldable(const float* ln1_g);
int launch_score_pairs_umma5(hsg_ctx* c, int cell_start, int n_cells_batch, float* out, void* out16, cudaStream_t s);
int launch_build_ptable(hsg_ctx* c, const float* V, const float* F, long long rows, void* P, cudaStream_t s);
void build_tiles4(const int2* pairs, long long P, std::vector<int>& out, bool allow_dense);
int umma_build_weights(const float* fc1, const float* w0, const float* b0, const float* pff_g, const float* pff_b,
                       const float* w1, const float* b1, const float* ln1_g, const float* c0, const float* cb0,
                       const float* cw1, const float* cb1, int hi, int natural_k, unsigned char* wimg_host, float* cst_host, float* comb_out = nullptr);
int umma_gain_foldable(const float* ln1_g);
int umma_range_level(const float* wq, const float* wk, const float* wv, const float* g1, const float* b1, const float* g2, const float* b2,
                     const float* g3, const float* b3, const float* w0, const float* b0, const float* pff_g, const float* pff_b,
                     const float* ln1_g, const float* ln1_b, const float* S, long long s_rows, int d);
void train_release(hsg_ctx* c);     // train.cu: frees every per-context training / sampler buffer (called by hsg_destroy)
int umma_wimg_bytes();
int proj_wimg_bytes(int d);
int proj_cst_count();
int proj_min_bins();
int proj_build_weights(int d, const float* sage_w, const float* sage_b, const float* wq, const float* wk, const float* g1, const float* b1,
                       const float* g2, const float* b2, unsigned char* wimg_host, float* cst_host);
int launch_token_project_umma(hsg_ctx* c, int cell_start, int n_cells_batch, cudaStream_t s);
int umma_cst_count();
int launch_score_pairs_f32(hsg_ctx* c, int cell_start, int n_cells_batch, float* out, cudaStream_t s);
int launch_score_triplets_f32(hsg_ctx* c, const int64_t* x, int64_t B, float* out, cudaStream_t s);

// ---- small device helpers -------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// fp32-mode activations: accurate libm versions (parity bar 1e-5 absolute on the score)
__device__ __forceinline__ float swish_acc(float x) { return x / (1.0f + expf(-x)); }
__device__ __forceinline__ float sigmoid_acc(float x) { return 1.0f / (1.0f + expf(-x)); }
// torch.nn.functional.softplus(beta=1, threshold=20)
__device__ __forceinline__ float softplus_acc(float x) { return x > 20.0f ? x : log1pf(expf(x)); }
__device__ __forceinline__ float apply_act(float x, int act) {
    if (act == HSG_ACT_SIGMOID) return sigmoid_acc(x);
    if (act == HSG_ACT_SOFTPLUS) return softplus_acc(x);
    return x;
}
