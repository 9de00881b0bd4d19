// api.cu -- the C ABI of libhsg.so (declared in include/hsg.h): context, uploads, preparation and the
// imputation driver that strings K1a -> K1b -> K2 over batches of cells.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <cub/cub.cuh>

#include <cuda_fp16.h>

#include "common.cuh"

static thread_local std::string g_err;
void hsg_set_error(const std::string& m) { g_err = m; }

extern "C" const char* hsg_last_error(void) { return g_err.c_str(); }
extern "C" int hsg_abi_version(void) { return HSG_ABI_VERSION; }

#define CTX_GUARD(c)                                                         \
    HSG_CHECK((c) != nullptr, "null context");                               \
    HSG_CUDA(cudaSetDevice((c)->device))

template <typename T>
static int dev_alloc(T** p, size_t n) {
    if (*p) { cudaFree(*p); *p = nullptr; }
    if (n == 0) n = 1;
    HSG_CUDA(cudaMalloc((void**)p, n * sizeof(T)));
    return 0;
}
template <typename T>
static void dev_free(T*& p) { if (p) { cudaFree(p); p = nullptr; } }

// host copies of the weights (needed to derive the transposed layouts)
struct HostWeights { std::vector<float> v[HSG_W_COUNT]; };
static HostWeights* hw(hsg_ctx* c);
struct CtxExtra { HostWeights hwts; };
#include <map>
static std::map<hsg_ctx*, CtxExtra*> g_extra;
static HostWeights* hw(hsg_ctx* c) { return &g_extra[c]->hwts; }

extern "C" int hsg_create(hsg_ctx** out, int device_ordinal) {
    HSG_CHECK(out != nullptr, "null out pointer");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        hsg_set_error(std::string("no CUDA device available (libhsg has no CPU fallback): ") + cudaGetErrorString(e));
        return 1;
    }
    HSG_CHECK(device_ordinal >= 0 && device_ordinal < n, "device ordinal out of range");
    HSG_CUDA(cudaSetDevice(device_ordinal));
    cudaDeviceProp prop;
    HSG_CUDA(cudaGetDeviceProperties(&prop, device_ordinal));
    if (prop.major < 10) {
        hsg_set_error("libhsg is built for sm_100a (B200); found compute capability " + std::to_string(prop.major) + "." +
                      std::to_string(prop.minor));
        return 1;
    }
    hsg_ctx* c = new hsg_ctx();
    c->device = device_ordinal;
    c->n_sm = prop.multiProcessorCount;
    if (const char* e = getenv("HSG_GATHER_SPARSE")) c->gather_sparse = (e[0] != '0');
    if (const char* e = getenv("HSG_GATHER_CHROM")) c->gather_chrom = (e[0] != '0');
    if (const char* e = getenv("HSG_GATHER_HASH")) c->gather_hash = (e[0] != '0');
    if (const char* e = getenv("HSG_GATHER_STREAM")) c->gather_stream = (e[0] != '0');
    if (const char* e = getenv("HSG_GATHER_STREAM_CAP")) c->gather_stream_cap = atoi(e);
    if (const char* e = getenv("HSG_GATHER_STREAM_LCAP")) c->gather_stream_lcap = atoi(e);
    HSG_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    HSG_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 8; ++i) HSG_CUDA(cudaEventCreate(&c->ev[i]));
    for (int i = 0; i < 2; ++i) HSG_CUDA(cudaEventCreateWithFlags(&c->stage_ev[i], cudaEventDisableTiming));
    HSG_CUDA(cudaEventCreateWithFlags(&c->work_ev, cudaEventDisableTiming));
    HSG_CUDA(cudaEventCreateWithFlags(&c->refresh_ev, cudaEventDisableTiming));
    g_extra[c] = new CtxExtra();
    *out = c;
    return 0;
}

static void free_derived(hsg_ctx* c) {
    dev_free(c->dw.sageT); dev_free(c->dw.wqT); dev_free(c->dw.wkT); dev_free(c->dw.wvT); dev_free(c->dw.fc1T);
    dev_free(c->dw.fc2T); dev_free(c->dw.w0T); dev_free(c->dw.w1T); dev_free(c->dw.c0T);
    dev_free(c->cellQ); dev_free(c->cellK); dev_free(c->cellV); dev_free(c->cellS); dev_free(c->binV); dev_free(c->binS);
    dev_free(c->neigh); dev_free(c->QK); dev_free(c->pair_bias);
    { __half* p = (__half*)c->QK16; dev_free(p); c->QK16 = nullptr; }
    { __half* p = (__half*)c->binV16; dev_free(p); c->binV16 = nullptr; }
    { __half* p = (__half*)c->cellV16; dev_free(p); c->cellV16 = nullptr; }
    dev_free(c->XS); dev_free(c->binSb); dev_free(c->cellSb); dev_free(c->binV32p); dev_free(c->cellV32p); dev_free(c->wimg); dev_free(c->cst); dev_free(c->pj_wimg); dev_free(c->ovf_list);
    dev_free(c->tiles4); c->n_tiles4 = 0; c->k2v4 = false;
    { __half* p = (__half*)c->binP; dev_free(p); c->binP = nullptr; }
    { __half* p = (__half*)c->cellP; dev_free(p); c->cellP = nullptr; }
    c->ovf_cap = 0;                          // the overflow list is gone: the next gather must allocate it again
    c->use_umma = false;
    dev_free(c->stage_out[0]); dev_free(c->stage_out[1]); c->stage_bytes = 0;
    c->prepared = false;
}

extern "C" int hsg_destroy(hsg_ctx* c) {
    if (!c) return 0;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    train_release(c);
    free_derived(c);
    for (int i = 0; i < HSG_W_COUNT; ++i) dev_free(c->w[i]);
    dev_free(c->E_cell); dev_free(c->E_bin); dev_free(c->d_bin_off); dev_free(c->d_bin_chrom);
    dev_free(c->row_ptr); dev_free(c->col); dev_free(c->val); dev_free(c->nbr); dev_free(c->nbr_w); dev_free(c->pairs);
    dev_free(c->row_ptr2); dev_free(c->col2); dev_free(c->val2);
    if (c->ins_ws) cudaFree(c->ins_ws);
    if (c->h2d_stream) cudaStreamDestroy(c->h2d_stream);
    if (c->staged_ev) cudaEventDestroy(c->staged_ev);
    if (c->work_ev) cudaEventDestroy(c->work_ev);
    if (c->refresh_ev) cudaEventDestroy(c->refresh_ev);
    for (int i = 0; i < 8; ++i) if (c->ev[i]) cudaEventDestroy(c->ev[i]);
    for (int i = 0; i < 2; ++i) if (c->stage_ev[i]) cudaEventDestroy(c->stage_ev[i]);
    if (c->stream) cudaStreamDestroy(c->stream);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    auto it = g_extra.find(c);
    if (it != g_extra.end()) { delete it->second; g_extra.erase(it); }
    delete c;
    return 0;
}

extern "C" int hsg_set_dims(hsg_ctx* c, int d_model, int n_cells, int n_chrom, const int32_t* bins, int cf) {
    CTX_GUARD(c);
    HSG_CHECK(d_model == 64 || d_model == 128, "d_model must be 64 or 128");
    HSG_CHECK(n_cells > 0 && n_chrom > 0 && cf >= 0 && bins != nullptr, "bad dimensions");
    c->d = d_model; c->n_cells = n_cells; c->n_chrom = n_chrom; c->cf = cf;
    c->bins.assign(bins, bins + n_chrom);
    c->bin_off.assign(n_chrom + 1, 0);
    for (int j = 0; j < n_chrom; ++j) {
        HSG_CHECK(bins[j] > 0, "empty chromosome");
        c->bin_off[j + 1] = c->bin_off[j] + bins[j];
    }
    c->n_bins = c->bin_off[n_chrom];
    std::vector<int> chrom_of(c->n_bins);
    for (int j = 0; j < n_chrom; ++j)
        for (int b = c->bin_off[j]; b < c->bin_off[j + 1]; ++b) chrom_of[b] = j;
    if (dev_alloc(&c->d_bin_off, n_chrom + 1) || dev_alloc(&c->d_bin_chrom, c->n_bins)) return 1;
    HSG_CUDA(cudaMemcpy(c->d_bin_off, c->bin_off.data(), (n_chrom + 1) * sizeof(int), cudaMemcpyHostToDevice));
    HSG_CUDA(cudaMemcpy(c->d_bin_chrom, chrom_of.data(), c->n_bins * sizeof(int), cudaMemcpyHostToDevice));
    if (dev_alloc(&c->E_cell, (size_t)n_cells * d_model) || dev_alloc(&c->E_bin, (size_t)c->n_bins * d_model)) return 1;
    c->table_set.assign(n_chrom + 1, 0);
    c->prepared = false;
    return 0;
}

static int64_t expected_count(hsg_ctx* c, int slot) {
    const int d = c->d, C = c->n_chrom, cf = c->cf;
    const int xin = 2 * (C + 1) + cf;
    switch (slot) {
        case HSG_W_QS: case HSG_W_KS: case HSG_W_VS: return (int64_t)HSG_HD * d;
        case HSG_W_FC1: case HSG_W_FC2: return (int64_t)d * HSG_HD;
        case HSG_W_ALN1_G: case HSG_W_ALN1_B: case HSG_W_ALN2_G: case HSG_W_ALN2_B: case HSG_W_ALN3_G: case HSG_W_ALN3_B:
        case HSG_W_PFF_LN_G: case HSG_W_PFF_LN_B: case HSG_W_PFF_B0: case HSG_W_PFF_B1:
        case HSG_W_LN1_G: case HSG_W_LN1_B: case HSG_W_LN2_G: case HSG_W_LN2_B: case HSG_W_SAGE_B: return d;
        case HSG_W_PFF_W0: case HSG_W_PFF_W1: return (int64_t)d * d;
        case HSG_W_CLS_W0: case HSG_W_VAR_W0: case HSG_W_PRB_W0: return (int64_t)(d / 2) * d;
        case HSG_W_CLS_B0: case HSG_W_VAR_B0: case HSG_W_PRB_B0: case HSG_W_CLS_W1: case HSG_W_VAR_W1: case HSG_W_PRB_W1: return d / 2;
        case HSG_W_CLS_B1: case HSG_W_VAR_B1: case HSG_W_PRB_B1: return 1;
        case HSG_W_SAGE_W: return (int64_t)d * 2 * d;
        case HSG_W_XP_W0: case HSG_W_XP2_W0: case HSG_W_XP3_W0: return 4LL * xin;
        case HSG_W_XP_B0: case HSG_W_XP2_B0: case HSG_W_XP3_B0: case HSG_W_XP_W1: case HSG_W_XP2_W1: case HSG_W_XP3_W1: return 4;
        case HSG_W_XP_B1: case HSG_W_XP2_B1: case HSG_W_XP3_B1: return 1;
        case HSG_W_ATTR: return (int64_t)(c->n_cells + 1 + c->n_bins) * (C + 1);
        case HSG_W_CELL_FEATS: return (int64_t)(c->n_cells + 1) * cf;
        default: return -1;
    }
}

extern "C" int hsg_set_weight(hsg_ctx* c, int slot, const float* host, int64_t count) {
    CTX_GUARD(c);
    HSG_CHECK(c->d > 0, "call hsg_set_dims first");
    HSG_CHECK(slot >= 0 && slot < HSG_W_COUNT && host != nullptr, "bad weight slot");
    int64_t want = expected_count(c, slot);
    if (want != count) {
        hsg_set_error("weight slot " + std::to_string(slot) + ": expected " + std::to_string(want) + " values, got " +
                      std::to_string(count));
        return 2;
    }
    if (dev_alloc(&c->w[slot], (size_t)count)) return 1;
    HSG_CUDA(cudaMemcpy(c->w[slot], host, count * sizeof(float), cudaMemcpyHostToDevice));
    c->wcount[slot] = count;
    hw(c)->v[slot].assign(host, host + count);
    c->prepared = false;
    return 0;
}

static float* table_ptr(hsg_ctx* c, int branch, int64_t* rows) {
    if (branch == 0) { *rows = c->n_cells; return c->E_cell; }
    *rows = c->bins[branch - 1];
    return c->E_bin + (size_t)c->bin_off[branch - 1] * c->d;
}

extern "C" int hsg_embed_table(hsg_ctx* c, int branch, const float* x, int64_t rows, int in_dim, const float* w,
                               const float* b, float* table_out) {
    CTX_GUARD(c);
    HSG_CHECK(c->d > 0, "call hsg_set_dims first");
    HSG_CHECK(branch >= 0 && branch <= c->n_chrom && x && w && b && in_dim > 0, "bad embed_table arguments");
    int64_t want = 0;
    float* dst = table_ptr(c, branch, &want);
    HSG_CHECK(rows == want, "embed_table: row count does not match the node type");
    float *dx = nullptr, *dwt = nullptr, *db = nullptr;
    if (dev_alloc(&dx, (size_t)rows * in_dim) || dev_alloc(&dwt, (size_t)c->d * in_dim) || dev_alloc(&db, (size_t)c->d)) return 1;
    HSG_CUDA(cudaMemcpy(dx, x, (size_t)rows * in_dim * sizeof(float), cudaMemcpyHostToDevice));
    HSG_CUDA(cudaMemcpy(dwt, w, (size_t)c->d * in_dim * sizeof(float), cudaMemcpyHostToDevice));
    HSG_CUDA(cudaMemcpy(db, b, (size_t)c->d * sizeof(float), cudaMemcpyHostToDevice));
    int rc = launch_embed_table(c, dx, rows, in_dim, dwt, db, branch > 0 ? 1 : 0, dst, c->stream);
    if (rc == 0) {
        cudaError_t e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) { hsg_set_error(std::string("embed_table kernel: ") + cudaGetErrorString(e)); rc = 1; }
    }
    if (rc == 0 && table_out) {
        cudaError_t e = cudaMemcpy(table_out, dst, (size_t)rows * c->d * sizeof(float), cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { hsg_set_error(std::string("embed_table copy: ") + cudaGetErrorString(e)); rc = 1; }
    }
    cudaFree(dx); cudaFree(dwt); cudaFree(db);
    if (rc == 0) { c->table_set[branch] = 1; c->prepared = false; }
    return rc;
}

extern "C" int hsg_set_table(hsg_ctx* c, int branch, const float* table, int64_t rows) {
    CTX_GUARD(c);
    HSG_CHECK(c->d > 0, "call hsg_set_dims first");
    HSG_CHECK(branch >= 0 && branch <= c->n_chrom && table, "bad set_table arguments");
    int64_t want = 0;
    float* dst = table_ptr(c, branch, &want);
    HSG_CHECK(rows == want, "set_table: row count does not match the node type");
    HSG_CUDA(cudaMemcpy(dst, table, (size_t)rows * c->d * sizeof(float), cudaMemcpyHostToDevice));
    c->table_set[branch] = 1; c->prepared = false;
    return 0;
}

// mean number of entries per row on every chromosome (rows of long chromosomes are denser: the streaming gather sizes its
// shared-memory stages for the densest one)
static void chrom_row_stats(hsg_ctx* c, const int64_t* row_ptr, std::vector<double>& out) {
    out.assign((size_t)c->n_chrom, 0.0);
    for (int j = 0; j < c->n_chrom; ++j) {
        int64_t tot = 0;
        for (int i = 0; i < c->n_cells; ++i) {
            const int64_t r0 = (int64_t)i * c->n_bins + c->bin_off[j];
            tot += row_ptr[r0 + c->bins[j]] - row_ptr[r0];
        }
        out[(size_t)j] = (double)tot / std::max<double>(1.0, (double)c->n_cells * c->bins[j]);
    }
}

extern "C" int hsg_debug_tile_table(const int32_t* pairs_xy, int64_t P, int allow_dense, int32_t* tiles_out, int64_t tiles_cap) {
    if (!pairs_xy || !tiles_out || P < 0) return -1;
    std::vector<int> t;
    build_tiles4(reinterpret_cast<const int2*>(pairs_xy), P, t, allow_dense != 0);
    const int64_t n = (int64_t)t.size() / 8;
    if (n > tiles_cap) return -1;
    memcpy(tiles_out, t.data(), t.size() * sizeof(int));
    return (int)n;
}

extern "C" int hsg_set_contacts(hsg_ctx* c, const int64_t* row_ptr, const int32_t* col, const float* val, int64_t n_rows,
                                int64_t nnz) {
    CTX_GUARD(c);
    HSG_CHECK(c->d > 0, "call hsg_set_dims first");
    HSG_CHECK(row_ptr && (nnz == 0 || (col && val)), "null contact arrays");
    HSG_CHECK(n_rows == (int64_t)c->n_cells * c->n_bins, "contacts: n_rows must be n_cells * n_bins");
    HSG_CHECK(row_ptr[0] == 0 && row_ptr[n_rows] == nnz, "contacts: row_ptr does not span nnz");
    // a re-upload of a same-sized (or smaller) matrix reuses the device buffers: no cudaFree / cudaMalloc round trip
    if (c->col != nullptr) HSG_CUDA(cudaDeviceSynchronize());      // work queued on any stream may still read the old contents
    if (n_rows + 1 > c->cap_rows) {
        if (dev_alloc(&c->row_ptr, (size_t)n_rows + 1 + HSG_CSR_PAD)) return 1;
        c->cap_rows = n_rows + 1;
    }
    if (nnz > c->cap_nnz || c->col == nullptr) {
        if (dev_alloc(&c->col, (size_t)nnz + HSG_CSR_PAD) || dev_alloc(&c->val, (size_t)nnz + HSG_CSR_PAD)) return 1;
        c->cap_nnz = nnz;
    }
    // three back-to-back DMA transfers on the context stream (truly asynchronous from pinned memory), one wait at the end:
    // the caller may reuse its arrays as soon as this returns
    HSG_CUDA(cudaMemcpyAsync(c->row_ptr, row_ptr, (size_t)(n_rows + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, c->stream));
    if (nnz > 0) {
        HSG_CUDA(cudaMemcpyAsync(c->col, col, (size_t)nnz * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
        HSG_CUDA(cudaMemcpyAsync(c->val, val, (size_t)nnz * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    }
    HSG_CUDA(cudaStreamSynchronize(c->stream));
    c->n_rows = n_rows; c->nnz = nnz;
    c->h_cell_ptr.resize((size_t)c->n_cells + 1);
    for (int i = 0; i <= c->n_cells; ++i) c->h_cell_ptr[i] = row_ptr[(int64_t)i * c->n_bins];
    chrom_row_stats(c, row_ptr, c->chrom_row_avg);
    return 0;
}

extern "C" int hsg_stage_contacts(hsg_ctx* c, const int64_t* row_ptr, const int32_t* col, const float* val, int64_t n_rows,
                                  int64_t nnz) {
    CTX_GUARD(c);
    HSG_CHECK(c->d > 0, "call hsg_set_dims first");
    HSG_CHECK(row_ptr && (nnz == 0 || (col && val)), "null contact arrays");
    HSG_CHECK(n_rows == (int64_t)c->n_cells * c->n_bins, "contacts: n_rows must be n_cells * n_bins");
    HSG_CHECK(row_ptr[0] == 0 && row_ptr[n_rows] == nnz, "contacts: row_ptr does not span nnz");
    HSG_CHECK(!c->staged, "hsg_stage_contacts: the previous staged matrix has not been committed");
    if (!c->h2d_stream) {
        HSG_CUDA(cudaStreamCreateWithFlags(&c->h2d_stream, cudaStreamNonBlocking));
        HSG_CUDA(cudaEventCreateWithFlags(&c->staged_ev, cudaEventDisableTiming));
    }
    if (n_rows + 1 > c->cap_rows2) {
        if (dev_alloc(&c->row_ptr2, (size_t)n_rows + 1 + HSG_CSR_PAD)) return 1;
        c->cap_rows2 = n_rows + 1;
    }
    if (nnz > c->cap_nnz2 || c->col2 == nullptr) {
        if (dev_alloc(&c->col2, (size_t)nnz + HSG_CSR_PAD) || dev_alloc(&c->val2, (size_t)nnz + HSG_CSR_PAD)) return 1;
        c->cap_nnz2 = nnz;
    }
    HSG_CUDA(cudaMemcpyAsync(c->row_ptr2, row_ptr, (size_t)(n_rows + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, c->h2d_stream));
    if (nnz > 0) {
        HSG_CUDA(cudaMemcpyAsync(c->col2, col, (size_t)nnz * sizeof(int32_t), cudaMemcpyHostToDevice, c->h2d_stream));
        HSG_CUDA(cudaMemcpyAsync(c->val2, val, (size_t)nnz * sizeof(float), cudaMemcpyHostToDevice, c->h2d_stream));
    }
    HSG_CUDA(cudaEventRecord(c->staged_ev, c->h2d_stream));
    c->h_cell_ptr2.resize((size_t)c->n_cells + 1);
    for (int i = 0; i <= c->n_cells; ++i) c->h_cell_ptr2[i] = row_ptr[(int64_t)i * c->n_bins];
    chrom_row_stats(c, row_ptr, c->chrom_row_avg2);
    c->staged_rows = n_rows; c->staged_nnz = nnz; c->staged = true;
    return 0;
}

extern "C" int hsg_commit_contacts(hsg_ctx* c) {
    CTX_GUARD(c);
    HSG_CHECK(c->staged, "hsg_commit_contacts without hsg_stage_contacts");
    HSG_CUDA(cudaEventSynchronize(c->staged_ev));       // the staged copy has landed
    // nothing queued may still read the current matrix: wait for the last imputation call's event (not for the whole device)
    if (c->work_pending) { HSG_CUDA(cudaEventSynchronize(c->work_ev)); c->work_pending = false; }
    c->h_cell_ptr.swap(c->h_cell_ptr2); c->chrom_row_avg.swap(c->chrom_row_avg2);
    std::swap(c->row_ptr, c->row_ptr2); std::swap(c->col, c->col2); std::swap(c->val, c->val2);
    std::swap(c->cap_rows, c->cap_rows2); std::swap(c->cap_nnz, c->cap_nnz2);
    c->n_rows = c->staged_rows; c->nnz = c->staged_nnz; c->staged = false;
    return 0;
}

extern "C" int hsg_set_neighbors(hsg_ctx* c, const int32_t* nbr, const float* w, int k1) {
    CTX_GUARD(c);
    HSG_CHECK(c->d > 0, "call hsg_set_dims first");
    if (nbr == nullptr) { dev_free(c->nbr); dev_free(c->nbr_w); c->k1 = 1; c->weighted = false; return 0; }
    HSG_CHECK(w != nullptr && k1 >= 1, "bad neighbour arguments");
    for (int64_t i = 0; i < (int64_t)c->n_cells * k1; ++i)
        HSG_CHECK(nbr[i] >= 0 && nbr[i] < c->n_cells, "neighbour id out of range");
    if (c->nbr == nullptr || k1 != c->k1 || !c->weighted) {
        if (dev_alloc(&c->nbr, (size_t)c->n_cells * k1) || dev_alloc(&c->nbr_w, (size_t)c->n_cells * k1)) return 1;
    }
    HSG_CUDA(cudaMemcpyAsync(c->nbr, nbr, (size_t)c->n_cells * k1 * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
    HSG_CUDA(cudaMemcpyAsync(c->nbr_w, w, (size_t)c->n_cells * k1 * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    HSG_CUDA(cudaStreamSynchronize(c->stream));
    c->k1 = k1; c->weighted = true;
    return 0;
}

int knn_build(hsg_ctx* c, const float* embed_host, int dim, int k1, int32_t* nbr_out, float* w_out);

extern "C" int hsg_build_neighbors(hsg_ctx* c, const float* embed, int dim, int k1, int32_t* nbr_out, float* w_out) {
    CTX_GUARD(c);
    HSG_CHECK(c->d > 0, "call hsg_set_dims first");
    HSG_CHECK(embed && nbr_out && w_out && dim > 0, "bad kNN arguments");
    if (knn_build(c, embed, dim, k1, nbr_out, w_out)) return 1;
    return hsg_set_neighbors(c, nbr_out, w_out, k1);
}

// ---------------------------------------------------------------------------------------------
// K3: triplet enumeration on the device (generate_binpair, Impute.py:49-61), bit-exact.
//   usable-prefix per chromosome -> per-row pair counts -> exclusive scan -> fill.
// ---------------------------------------------------------------------------------------------
__global__ void pair_count_kernel(const int* __restrict__ row_chrom_off, const int* __restrict__ row_nc,
                                  const int* __restrict__ usable_prefix, const int* __restrict__ row_up_off,
                                  int n_rows, int min_bin, int max_bin, int64_t* __restrict__ counts) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    int off = row_chrom_off[r], nc = row_nc[r];
    int b1 = r - off;                                // rows are ordered chromosome by chromosome
    const int* up = usable_prefix + row_up_off[r];   // up[b] = #usable bins < b, length nc+1
    long long cnt = 0;
    if (up[b1 + 1] - up[b1] == 1) {
        long long lo = (long long)b1 + min_bin, hi = (long long)b1 + max_bin;
        if (hi > nc) hi = nc;
        if (lo < hi) cnt = up[hi] - up[lo];
    }
    counts[r] = cnt;
}

__global__ void pair_fill_kernel(const int* __restrict__ row_chrom_off, const int* __restrict__ row_nc,
                                 const int* __restrict__ usable_prefix, const int* __restrict__ row_up_off,
                                 const int* __restrict__ row_goff, int n_rows, int min_bin, int max_bin,
                                 const int64_t* __restrict__ starts, int2* __restrict__ pairs) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    int off = row_chrom_off[r], nc = row_nc[r];
    int b1 = r - off;
    const int* up = usable_prefix + row_up_off[r];
    if (up[b1 + 1] - up[b1] != 1) return;
    long long lo = (long long)b1 + min_bin, hi = (long long)b1 + max_bin;
    if (hi > nc) hi = nc;
    int64_t w = starts[r];
    int g = row_goff[r];
    for (long long b2 = lo; b2 < hi; ++b2)
        if (up[b2 + 1] - up[b2] == 1) pairs[w++] = make_int2(g + b1, g + (int)b2);
}

extern "C" int hsg_set_pairs(hsg_ctx* c, int n_sel, const int32_t* chroms, int min_bin, int max_bin, const int32_t* skip,
                             int n_skip, int64_t* P_total, int64_t* per_chrom_P) {
    CTX_GUARD(c);
    HSG_CHECK(c->d > 0, "call hsg_set_dims first");
    HSG_CHECK(n_sel > 0 && chroms != nullptr && min_bin >= 0, "bad pair arguments");
    int max_nc = *std::max_element(c->bins.begin(), c->bins.end());
    if (max_bin < 0 || max_bin > max_nc + 1) max_bin = max_nc + 1;
    // host-side row descriptors (O(n_bins) integers)
    std::vector<int> row_chrom_off, row_nc, row_up_off, row_goff, up;
    std::vector<int> sel_row0(n_sel + 1, 0);
    for (int s = 0; s < n_sel; ++s) {
        int j = chroms[s];
        HSG_CHECK(j >= 0 && j < c->n_chrom, "selected chromosome out of range");
        int nc = c->bins[j];
        std::vector<char> usable(nc, 1);
        for (int q = 0; q < n_skip; ++q) {
            if (skip[3 * q] != j) continue;
            for (int b = std::max(0, skip[3 * q + 1]); b < std::min(nc, skip[3 * q + 2]); ++b) usable[b] = 0;
        }
        int up_off = (int)up.size();
        up.push_back(0);
        for (int b = 0; b < nc; ++b) up.push_back(up.back() + (usable[b] ? 1 : 0));
        int row0 = (int)row_nc.size();
        sel_row0[s] = row0;
        for (int b = 0; b < nc; ++b) {
            row_chrom_off.push_back(row0); row_nc.push_back(nc); row_up_off.push_back(up_off);
            row_goff.push_back(c->bin_off[j]);
        }
    }
    int n_rows = (int)row_nc.size();
    sel_row0[n_sel] = n_rows;
    int *d_off = nullptr, *d_nc = nullptr, *d_upoff = nullptr, *d_goff = nullptr, *d_up = nullptr;
    int64_t *d_cnt = nullptr, *d_start = nullptr;
    void* d_tmp = nullptr;
    int rc = 0;
    auto cleanup = [&]() {
        cudaFree(d_off); cudaFree(d_nc); cudaFree(d_upoff); cudaFree(d_goff); cudaFree(d_up); cudaFree(d_cnt);
        cudaFree(d_start); cudaFree(d_tmp);
    };
#define TRY(x) do { if ((x) != cudaSuccess) { hsg_set_error(std::string("set_pairs: ") + cudaGetErrorString(cudaGetLastError())); cleanup(); return 1; } } while (0)
    TRY(cudaMalloc(&d_off, n_rows * sizeof(int))); TRY(cudaMalloc(&d_nc, n_rows * sizeof(int)));
    TRY(cudaMalloc(&d_upoff, n_rows * sizeof(int))); TRY(cudaMalloc(&d_goff, n_rows * sizeof(int)));
    TRY(cudaMalloc(&d_up, up.size() * sizeof(int)));
    TRY(cudaMalloc(&d_cnt, (n_rows + 1) * sizeof(int64_t))); TRY(cudaMalloc(&d_start, (n_rows + 1) * sizeof(int64_t)));
    TRY(cudaMemcpy(d_off, row_chrom_off.data(), n_rows * sizeof(int), cudaMemcpyHostToDevice));
    TRY(cudaMemcpy(d_nc, row_nc.data(), n_rows * sizeof(int), cudaMemcpyHostToDevice));
    TRY(cudaMemcpy(d_upoff, row_up_off.data(), n_rows * sizeof(int), cudaMemcpyHostToDevice));
    TRY(cudaMemcpy(d_goff, row_goff.data(), n_rows * sizeof(int), cudaMemcpyHostToDevice));
    TRY(cudaMemcpy(d_up, up.data(), up.size() * sizeof(int), cudaMemcpyHostToDevice));
    TRY(cudaMemsetAsync(d_cnt, 0, (n_rows + 1) * sizeof(int64_t), c->stream));
    unsigned grid = (unsigned)((n_rows + 127) / 128);
    pair_count_kernel<<<grid, 128, 0, c->stream>>>(d_off, d_nc, d_up, d_upoff, n_rows, min_bin, max_bin, d_cnt);
    c->launches++;
    size_t tmp_bytes = 0;
    TRY(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_cnt, d_start, n_rows + 1, c->stream));
    TRY(cudaMalloc(&d_tmp, tmp_bytes));
    TRY(cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, d_cnt, d_start, n_rows + 1, c->stream));
    c->launches++;
    std::vector<int64_t> starts(n_rows + 1);
    TRY(cudaMemcpyAsync(starts.data(), d_start, (n_rows + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost, c->stream));
    TRY(cudaStreamSynchronize(c->stream));
    int64_t P = starts[n_rows];
    dev_free(c->pairs);
    if (dev_alloc(&c->pairs, (size_t)std::max<int64_t>(P, 1))) { cleanup(); return 1; }
    if (P > 0) {
        pair_fill_kernel<<<grid, 128, 0, c->stream>>>(d_off, d_nc, d_up, d_upoff, d_goff, n_rows, min_bin, max_bin, d_start,
                                                      c->pairs);
        c->launches++;
        TRY(cudaGetLastError());
    }
    c->h_pairs.resize((size_t)P);
    if (P > 0) TRY(cudaMemcpyAsync(c->h_pairs.data(), c->pairs, (size_t)P * sizeof(int2), cudaMemcpyDeviceToHost, c->stream));
    TRY(cudaStreamSynchronize(c->stream));
#undef TRY
    cleanup();
    c->P = P;
    c->sel_chroms.assign(chroms, chroms + n_sel);
    c->sel_P.assign(n_sel, 0); c->sel_off.assign(n_sel + 1, 0);
    for (int s = 0; s < n_sel; ++s) {
        c->sel_off[s] = starts[sel_row0[s]];
        c->sel_P[s] = starts[sel_row0[s + 1]] - starts[sel_row0[s]];
        if (per_chrom_P) per_chrom_P[s] = c->sel_P[s];
    }
    c->sel_off[n_sel] = P;
    if (P_total) *P_total = P;
    c->prepared = false;
    return rc;
}

extern "C" int hsg_get_coordinates(hsg_ctx* c, int sel_slot, int64_t* host_out) {
    CTX_GUARD(c);
    HSG_CHECK(sel_slot >= 0 && sel_slot < (int)c->sel_chroms.size() && host_out, "bad coordinate request");
    int goff = c->bin_off[c->sel_chroms[sel_slot]];
    int64_t lo = c->sel_off[sel_slot], n = c->sel_P[sel_slot];
    for (int64_t i = 0; i < n; ++i) {
        host_out[2 * i] = c->h_pairs[lo + i].x - goff;
        host_out[2 * i + 1] = c->h_pairs[lo + i].y - goff;
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------
static int upload_transposed(float** dst, const std::vector<float>& src, int rows, int cols) {
    // src is [rows][cols] row-major; dst becomes [cols][rows]
    std::vector<float> t((size_t)rows * cols);
    for (int r = 0; r < rows; ++r)
        for (int q = 0; q < cols; ++q) t[(size_t)q * rows + r] = src[(size_t)r * cols + q];
    if (dev_alloc(dst, t.size())) return 1;
    HSG_CUDA(cudaMemcpy(*dst, t.data(), t.size() * sizeof(float), cudaMemcpyHostToDevice));
    return 0;
}

extern "C" int hsg_prepare(hsg_ctx* c, int precision, int local_transfer_range, int activation) {
    CTX_GUARD(c);
    HSG_CHECK(c->d > 0, "call hsg_set_dims first");
    HSG_CHECK(precision == HSG_PREC_FP32 || precision == HSG_PREC_TC16 || precision == HSG_PREC_TC16_HI, "unknown precision");
    HSG_CHECK(local_transfer_range >= 0 && local_transfer_range <= 8, "local_transfer_range out of range");
    HSG_CHECK(activation >= HSG_ACT_NONE && activation <= HSG_ACT_SOFTPLUS, "unknown activation");
    static const int needed[] = {HSG_W_QS, HSG_W_KS, HSG_W_VS, HSG_W_FC1, HSG_W_FC2, HSG_W_ALN1_G, HSG_W_ALN1_B, HSG_W_ALN2_G,
                                 HSG_W_ALN2_B, HSG_W_ALN3_G, HSG_W_ALN3_B, HSG_W_PFF_LN_G, HSG_W_PFF_LN_B, HSG_W_PFF_W0,
                                 HSG_W_PFF_B0, HSG_W_PFF_W1, HSG_W_PFF_B1, HSG_W_LN1_G, HSG_W_LN1_B, HSG_W_LN2_G, HSG_W_LN2_B,
                                 HSG_W_CLS_W0, HSG_W_CLS_B0, HSG_W_CLS_W1, HSG_W_CLS_B1, HSG_W_SAGE_W, HSG_W_SAGE_B,
                                 HSG_W_XP2_W0, HSG_W_XP2_B0, HSG_W_XP2_W1, HSG_W_XP2_B1, HSG_W_ATTR};
    for (int s : needed)
        if (c->w[s] == nullptr) { hsg_set_error("hsg_prepare: weight slot " + std::to_string(s) + " was never set"); return 2; }
    for (int b = 0; b <= c->n_chrom; ++b)
        if (!c->table_set[b]) { hsg_set_error("hsg_prepare: embedding table of branch " + std::to_string(b) + " missing"); return 2; }
    HSG_CHECK(c->row_ptr != nullptr, "hsg_prepare: contacts not set");
    free_derived(c);
    c->precision = precision; c->ltr = local_transfer_range; c->act = activation;
    const int d = c->d;
    HostWeights* h = hw(c);
    if (upload_transposed(&c->dw.sageT, h->v[HSG_W_SAGE_W], d, 2 * d)) return 1;
    if (upload_transposed(&c->dw.wqT, h->v[HSG_W_QS], HSG_HD, d)) return 1;
    if (upload_transposed(&c->dw.wkT, h->v[HSG_W_KS], HSG_HD, d)) return 1;
    if (upload_transposed(&c->dw.wvT, h->v[HSG_W_VS], HSG_HD, d)) return 1;
    if (upload_transposed(&c->dw.fc1T, h->v[HSG_W_FC1], d, HSG_HD)) return 1;
    if (upload_transposed(&c->dw.fc2T, h->v[HSG_W_FC2], d, HSG_HD)) return 1;
    if (upload_transposed(&c->dw.w0T, h->v[HSG_W_PFF_W0], d, d)) return 1;
    if (upload_transposed(&c->dw.w1T, h->v[HSG_W_PFF_W1], d, d)) return 1;
    if (upload_transposed(&c->dw.c0T, h->v[HSG_W_CLS_W0], d / 2, d)) return 1;
    // static tokens
    if (dev_alloc(&c->cellQ, (size_t)c->n_cells * HSG_HD) || dev_alloc(&c->cellK, (size_t)c->n_cells * HSG_HD) ||
        dev_alloc(&c->cellV, (size_t)c->n_cells * HSG_HD) || dev_alloc(&c->cellS, (size_t)c->n_cells * d) ||
        dev_alloc(&c->binV, (size_t)c->n_bins * HSG_HD) || dev_alloc(&c->binS, (size_t)c->n_bins * d)) return 1;
    if (launch_static_tokens(c, c->E_cell, c->n_cells, c->cellQ, c->cellK, c->cellV, c->cellS, c->stream)) return 1;
    if (launch_static_tokens(c, c->E_bin, c->n_bins, nullptr, nullptr, c->binV, c->binS, c->stream)) return 1;
    // pair bias
    if (dev_alloc(&c->pair_bias, (size_t)std::max<int64_t>(c->P, 1))) return 1;
    if (launch_pair_bias(c, c->stream)) return 1;
    // Tensor-core mode: d_model = 64 (fast and high-accuracy variants) and d_model = 128 (fast variant only: a model that needs the
    // high-accuracy variant -- softplus / raw head, layer_norm1 gain too close to zero, fp16 range -- runs the fp32 kernels there).
    c->use_umma = (precision != HSG_PREC_FP32 && (d == 64 || d == 128));
    c->use_proj_umma = false;
    c->k2d128 = false;
    // softplus / raw heads pass the logit error through with slope ~1: they get the high-accuracy variant
    // (so does a layer_norm1 gain too close to zero for the fast variant's gain folding)
    c->umma_hi = c->use_umma && (precision == HSG_PREC_TC16_HI || activation != HSG_ACT_SIGMOID ||
                                 !(d == 64 ? umma_gain_foldable(h->v[HSG_W_LN1_G].data()) : umma128_gain_foldable(h->v[HSG_W_LN1_G].data())));
    if (d == 128 && c->umma_hi) { c->use_umma = false; c->umma_hi = false; }
    if (c->use_umma) {
        // fp16 RANGE guard: the 16-bit operands (Q / K / V rows, ctx, h, the squared difference u) top out at 65504.  Bound every one
        // of them from the weights and from the S tables of THIS model; a model outside the fast variant's range takes the
        // high-accuracy variant (fp32 tables and attention, unfolded layer_norm1), one outside that range the fp32 kernels.
        HSG_CUDA(cudaStreamSynchronize(c->stream));
        std::vector<float> hS((size_t)(c->n_bins + c->n_cells) * d);
        HSG_CUDA(cudaMemcpy(hS.data(), c->binS, (size_t)c->n_bins * d * sizeof(float), cudaMemcpyDeviceToHost));
        HSG_CUDA(cudaMemcpy(hS.data() + (size_t)c->n_bins * d, c->cellS, (size_t)c->n_cells * d * sizeof(float), cudaMemcpyDeviceToHost));
        const int lvl = umma_range_level(h->v[HSG_W_QS].data(), h->v[HSG_W_KS].data(), h->v[HSG_W_VS].data(), h->v[HSG_W_ALN1_G].data(),
                                         h->v[HSG_W_ALN1_B].data(), h->v[HSG_W_ALN2_G].data(), h->v[HSG_W_ALN2_B].data(),
                                         h->v[HSG_W_ALN3_G].data(), h->v[HSG_W_ALN3_B].data(), h->v[HSG_W_PFF_W0].data(),
                                         h->v[HSG_W_PFF_B0].data(), h->v[HSG_W_PFF_LN_G].data(), h->v[HSG_W_PFF_LN_B].data(),
                                         h->v[HSG_W_LN1_G].data(), h->v[HSG_W_LN1_B].data(), hS.data(), (long long)(c->n_bins + c->n_cells), d);
        if (lvl >= 1) c->umma_hi = true;
        if (lvl >= 2 || (lvl >= 1 && d == 128)) { c->use_umma = false; c->umma_hi = false; }
    }
    {
        const char* v1 = getenv("HSG_K2V1");
        c->k2v2 = c->use_umma && !c->umma_hi && (d == 128 || !(v1 && v1[0] == '1'));
        c->k2d128 = c->use_umma && d == 128;
    }
    size_t per_cell;
    if (c->k2d128) {
        std::vector<unsigned char> wimg(umma128_wimg_bytes());
        std::vector<float> cst(umma128_cst_count(), 0.f);
        umma128_build_weights(h->v[HSG_W_FC1].data(), h->v[HSG_W_PFF_W0].data(), h->v[HSG_W_PFF_B0].data(),
                              h->v[HSG_W_PFF_LN_G].data(), h->v[HSG_W_PFF_LN_B].data(), h->v[HSG_W_PFF_W1].data(),
                              h->v[HSG_W_PFF_B1].data(), h->v[HSG_W_LN1_G].data(), h->v[HSG_W_CLS_W0].data(),
                              h->v[HSG_W_CLS_B0].data(), h->v[HSG_W_CLS_W1].data(), h->v[HSG_W_CLS_B1].data(), wimg.data(), cst.data());
        if (dev_alloc(&c->wimg, wimg.size())) return 1;
        HSG_CUDA(cudaMemcpy(c->wimg, wimg.data(), wimg.size(), cudaMemcpyHostToDevice));
        HSG_CHECK(cst.size() <= 512, "constant block too large");
        memcpy(c->h_cst, cst.data(), cst.size() * sizeof(float));
        __half *bv = nullptr, *cv = nullptr;
        const size_t bins_pad = (size_t)((c->n_bins + HSG_GRP - 1) / HSG_GRP) * HSG_GRP;
        if (dev_alloc(&bv, bins_pad * HSG_HD) || dev_alloc(&cv, (size_t)c->n_cells * HSG_HD) ||
            dev_alloc(&c->binSb, bins_pad * d) || dev_alloc(&c->cellSb, (size_t)c->n_cells * d)) return 1;
        HSG_CUDA(cudaMemsetAsync(bv, 0, bins_pad * HSG_HD * sizeof(__half), c->stream));
        HSG_CUDA(cudaMemsetAsync(c->binSb, 0, bins_pad * d * sizeof(float), c->stream));
        c->binV16 = bv; c->cellV16 = cv;
        if (launch_convert_static128(c, c->binV, c->binS, c->n_bins, 1, c->binV16, c->binSb, c->stream)) return 1;
        if (launch_convert_static128(c, c->cellV, c->cellS, c->n_cells, 0, c->cellV16, c->cellSb, c->stream)) return 1;
        // K1b on the tensor cores for d_model = 128 too (project_umma.cu, chunk-major tables)
        c->use_proj_umma = c->n_bins >= proj_min_bins() && !(getenv("HSG_PROJ128_F32") && getenv("HSG_PROJ128_F32")[0] == '1');
        if (c->use_proj_umma) {
            std::vector<unsigned char> pimg(proj_wimg_bytes(d));
            std::vector<float> pcst(proj_cst_count(), 0.f);
            proj_build_weights(d, h->v[HSG_W_SAGE_W].data(), h->v[HSG_W_SAGE_B].data(), h->v[HSG_W_QS].data(), h->v[HSG_W_KS].data(),
                               h->v[HSG_W_ALN1_G].data(), h->v[HSG_W_ALN1_B].data(), h->v[HSG_W_ALN2_G].data(),
                               h->v[HSG_W_ALN2_B].data(), pimg.data(), pcst.data());
            if (dev_alloc(&c->pj_wimg, pimg.size())) return 1;
            HSG_CUDA(cudaMemcpy(c->pj_wimg, pimg.data(), pimg.size(), cudaMemcpyHostToDevice));
            memcpy(c->h_pj_cst, pcst.data(), pcst.size() * sizeof(float));
        }
        per_cell = (size_t)c->n_bins * (2 * HSG_HD * sizeof(__half) + 16 * sizeof(float) + d * sizeof(float));
    } else if (c->use_umma) {
        std::vector<unsigned char> wimg(umma_wimg_bytes());
        std::vector<float> cst(umma_cst_count(), 0.f);
        std::vector<float> fcomb((size_t)2 * 64 * HSG_HD);
        umma_build_weights(h->v[HSG_W_FC1].data(), h->v[HSG_W_PFF_W0].data(), h->v[HSG_W_PFF_B0].data(),
                           h->v[HSG_W_PFF_LN_G].data(), h->v[HSG_W_PFF_LN_B].data(), h->v[HSG_W_PFF_W1].data(),
                           h->v[HSG_W_PFF_B1].data(), h->v[HSG_W_LN1_G].data(), h->v[HSG_W_CLS_W0].data(),
                           h->v[HSG_W_CLS_B0].data(), h->v[HSG_W_CLS_W1].data(), h->v[HSG_W_CLS_B1].data(), c->umma_hi ? 1 : 0, c->k2v2 ? 1 : 0,
                           wimg.data(), cst.data(), fcomb.data());
        if (dev_alloc(&c->wimg, wimg.size()) || dev_alloc(&c->cst, cst.size())) return 1;
        HSG_CUDA(cudaMemcpy(c->wimg, wimg.data(), wimg.size(), cudaMemcpyHostToDevice));
        HSG_CUDA(cudaMemcpy(c->cst, cst.data(), cst.size() * sizeof(float), cudaMemcpyHostToDevice));
        HSG_CHECK(cst.size() <= 512, "constant block too large");
        memcpy(c->h_cst, cst.data(), cst.size() * sizeof(float));
        // K1b on the tensor cores (fast variant only; a 128-token tile may touch at most 4 cells)
        c->use_proj_umma = !c->umma_hi && c->n_bins >= proj_min_bins();
        if (c->use_proj_umma) {
            std::vector<unsigned char> pimg(proj_wimg_bytes(d));
            std::vector<float> pcst(proj_cst_count(), 0.f);
            proj_build_weights(d, h->v[HSG_W_SAGE_W].data(), h->v[HSG_W_SAGE_B].data(), h->v[HSG_W_QS].data(), h->v[HSG_W_KS].data(),
                               h->v[HSG_W_ALN1_G].data(), h->v[HSG_W_ALN1_B].data(), h->v[HSG_W_ALN2_G].data(),
                               h->v[HSG_W_ALN2_B].data(), pimg.data(), pcst.data());
            if (dev_alloc(&c->pj_wimg, pimg.size())) return 1;
            HSG_CUDA(cudaMemcpy(c->pj_wimg, pimg.data(), pimg.size(), cudaMemcpyHostToDevice));
            memcpy(c->h_pj_cst, pcst.data(), pcst.size() * sizeof(float));
        }
        __half *bv = nullptr, *cv = nullptr;
        const size_t bins_pad = (size_t)((c->n_bins + HSG_GRP - 1) / HSG_GRP) * HSG_GRP;     // score_umma2.cu tables hold whole groups of 32 bins
        if (dev_alloc(&bv, bins_pad * HSG_HD) || dev_alloc(&cv, (size_t)c->n_cells * HSG_HD) ||
            dev_alloc(&c->binSb, bins_pad * d) || dev_alloc(&c->cellSb, (size_t)c->n_cells * d)) return 1;
        HSG_CUDA(cudaMemsetAsync(bv, 0, bins_pad * HSG_HD * sizeof(__half), c->stream));
        HSG_CUDA(cudaMemsetAsync(c->binSb, 0, bins_pad * d * sizeof(float), c->stream));
        c->binV16 = bv; c->cellV16 = cv;
        if (c->umma_hi && (dev_alloc(&c->binV32p, (size_t)c->n_bins * HSG_HD) || dev_alloc(&c->cellV32p, (size_t)c->n_cells * HSG_HD))) return 1;
        if (c->k2v2) {
            if (launch_convert_static2(c, c->binV, c->binS, c->n_bins, 1, c->binV16, c->binSb, c->stream)) return 1;
            if (launch_convert_static2(c, c->cellV, c->cellS, c->n_cells, 0, c->cellV16, c->cellSb, c->stream)) return 1;
            // generation 4 (default; HSG_K2_GEN=2 / 3 select the earlier kernels): P tables + tile table of the pair list
            const char* gen = getenv("HSG_K2_GEN");
            c->k2v4 = !(gen && (gen[0] == '2' || gen[0] == '3')) && c->P > 0;
            c->k2gen = (gen && gen[0] == '5') ? 5 : 4;            // HSG_K2_GEN=5: generation 4's MMA chain with two threads per row (measured 28 % slower)
            if (c->k2v4) {
                float* dF = nullptr; __half *bp = nullptr, *cp = nullptr;
                if (dev_alloc(&dF, fcomb.size()) || dev_alloc(&bp, (size_t)c->n_bins * 128 * 8) || dev_alloc(&cp, (size_t)c->n_cells * 128 * 8)) return 1;
                c->binP = bp; c->cellP = cp;
                HSG_CUDA(cudaMemcpyAsync(dF, fcomb.data(), fcomb.size() * sizeof(float), cudaMemcpyHostToDevice, c->stream));
                if (launch_build_ptable(c, c->binV, dF, c->n_bins, c->binP, c->stream)) return 1;
                if (launch_build_ptable(c, c->cellV, dF, c->n_cells, c->cellP, c->stream)) return 1;
                HSG_CUDA(cudaStreamSynchronize(c->stream));
                dev_free(dF);
                std::vector<int> tiles;
                // host copy of the pair list kept by hsg_set_pairs; dense tiles for the short rows at chromosome ends: generation 4 only
                // (they pay when they remove more than ~4 % of the tiles -- C2: 1904 -> 1762 per cell; on C3 / C5 the tails are a
                //  small share and the kernel instantiation without the dense-tile code is the faster one)
                const char* dt = getenv("HSG_K2_DENSE_TILES");
                build_tiles4(c->h_pairs.data(), c->P, tiles, false);
                c->k2_dense_tiles = false;
                if (c->k2gen == 4 && !(dt && dt[0] == '0')) {
                    std::vector<int> tiles_d;
                    build_tiles4(c->h_pairs.data(), c->P, tiles_d, true);
                    if ((dt && dt[0] == '1') || (double)tiles_d.size() <= 0.96 * (double)tiles.size()) { tiles.swap(tiles_d); c->k2_dense_tiles = true; }
                }
                c->n_tiles4 = (int)(tiles.size() / 8);
                if (dev_alloc(&c->tiles4, tiles.size())) return 1;
                HSG_CUDA(cudaMemcpy(c->tiles4, tiles.data(), tiles.size() * sizeof(int), cudaMemcpyHostToDevice));
            }
        } else {
        if (launch_convert_static(c, c->binV, c->binS, c->n_bins, c->binV16, c->binSb, c->umma_hi ? c->binV32p : nullptr, c->stream)) return 1;
        if (launch_convert_static(c, c->cellV, c->cellS, c->n_cells, c->cellV16, c->cellSb, c->umma_hi ? c->cellV32p : nullptr, c->stream)) return 1;
        }
        per_cell = (size_t)c->n_bins * ((c->umma_hi ? 2 * HSG_HD * sizeof(float) : 2 * HSG_HD * sizeof(__half)) + 16 * sizeof(float) + d * sizeof(float));
    } else {
        per_cell = (size_t)c->n_bins * (2 * HSG_HD + d) * sizeof(float);
    }
    // batch workspace: keep the token tables of a batch within ~96 MB so K2 finds them in L2
    int batch = (int)std::max<size_t>(1, std::min<size_t>(64, (96u << 20) / per_cell));
    c->batch_cells = batch;
    if (dev_alloc(&c->neigh, (size_t)batch * c->n_bins * d)) return 1;
    if (c->use_umma) {
        __half* qk16 = nullptr;
        const size_t bins_pad = (size_t)((c->n_bins + HSG_GRP - 1) / HSG_GRP) * HSG_GRP;
        if (dev_alloc(&qk16, c->umma_hi ? (size_t)1 : (size_t)batch * bins_pad * 2 * HSG_HD) ||
            dev_alloc(&c->XS, (size_t)batch * bins_pad * 16) ||
            dev_alloc(&c->QK, (size_t)(c->umma_hi ? batch : 1) * c->n_bins * 2 * HSG_HD)) return 1;
        c->QK16 = qk16;
    } else {
        if (dev_alloc(&c->QK, (size_t)batch * c->n_bins * 2 * HSG_HD)) return 1;
    }
    if (c->k2v2) {      // the padding bins of the last group are never written by K1b: keep them finite
        HSG_CUDA(cudaMemsetAsync(c->QK16, 0, (size_t)batch * (size_t)((c->n_bins + HSG_GRP - 1) / HSG_GRP) * HSG_GRP * 2 * HSG_HD * sizeof(__half), c->stream));
        HSG_CUDA(cudaMemsetAsync(c->XS, 0, (size_t)batch * (size_t)((c->n_bins + HSG_GRP - 1) / HSG_GRP) * HSG_GRP * 16 * sizeof(float), c->stream));
    }
    HSG_CUDA(cudaStreamSynchronize(c->stream));
    c->fixed_cell = -1;
    c->prepared = true;
    return 0;
}

static int run_batch(hsg_ctx* c, int cell0, int nb, float* out, cudaStream_t s, bool prof, void* out16 = nullptr) {
    if (prof) HSG_CUDA(cudaEventRecord(c->ev[0], s));
    if (launch_gather(c, nullptr, cell0, nb, s)) return 1;
    if (prof) HSG_CUDA(cudaEventRecord(c->ev[1], s));
    if (c->use_umma && c->use_proj_umma) { if (launch_token_project_umma(c, cell0, nb, s)) return 1; }
    else if (launch_token_project(c, cell0, nb, c->use_umma ? (c->umma_hi ? 2 : 1) : 0, s)) return 1;
    if (prof) HSG_CUDA(cudaEventRecord(c->ev[2], s));
    if (c->k2d128) { if (launch_score_pairs_umma128(c, cell0, nb, out, out16, s)) return 1; }
    else if (c->k2v4) { if ((c->k2gen == 4 ? launch_score_pairs_umma4 : launch_score_pairs_umma5)(c, cell0, nb, out, out16, s)) return 1; }
    else if (c->k2v2) { if (launch_score_pairs_umma2(c, cell0, nb, out, out16, s)) return 1; }
    else if (c->use_umma) { if (launch_score_pairs_umma(c, cell0, nb, out, s)) return 1; }
    else if (launch_score_pairs_f32(c, cell0, nb, out, s)) return 1;
    if (prof) {
        HSG_CUDA(cudaEventRecord(c->ev[3], s));
        HSG_CUDA(cudaEventSynchronize(c->ev[3]));
        float a = 0, b = 0, g = 0;
        cudaEventElapsedTime(&a, c->ev[0], c->ev[1]);
        cudaEventElapsedTime(&b, c->ev[1], c->ev[2]);
        cudaEventElapsedTime(&g, c->ev[2], c->ev[3]);
        c->last_ms[0] += a; c->last_ms[1] += b; c->last_ms[2] += g; c->last_ms[3] += a + b + g;
    }
    return 0;
}

// work queued on `s` must see the rows hsg_refresh_contacts uploaded on the upload stream: an event wait, not a device-wide sync
static int wait_for_uploads(hsg_ctx* c, cudaStream_t s) {
    if (c->refresh_pending) {
        HSG_CUDA(cudaStreamWaitEvent(s, c->refresh_ev, 0));
        c->refresh_pending = false;
    }
    return 0;
}

static int impute_dev(hsg_ctx* c, int cell_start, int cell_end, void* out_dev, bool half, cudaStream_t s) {
    HSG_CHECK(c->prepared, "call hsg_prepare first");
    HSG_CHECK(cell_start >= 0 && cell_end <= c->n_cells && cell_start <= cell_end, "cell range out of bounds");
    HSG_CHECK(out_dev != nullptr || cell_start == cell_end || c->P == 0, "null output");
    HSG_CHECK(!half || c->k2v2, "16-bit score blocks need the tensor-core scorer with a sigmoid head (HSG_PREC_TC16, HSG_ACT_SIGMOID, d_model 64)");
    if (wait_for_uploads(c, s)) return 1;
    for (int i = 0; i < 4; ++i) c->last_ms[i] = 0.f;
    for (int c0 = cell_start; c0 < cell_end; c0 += c->batch_cells) {
        int nb = std::min(c->batch_cells, cell_end - c0);
        const size_t off = (size_t)(c0 - cell_start) * c->P;
        if (run_batch(c, c0, nb, half ? nullptr : (float*)out_dev + off, s, c->profiling, half ? (void*)((__half*)out_dev + off) : nullptr)) return 1;
    }
    HSG_CUDA(cudaEventRecord(c->work_ev, s));        // hsg_refresh_contacts orders its copies behind the reads queued here
    c->work_pending = true;
    c->fixed_cell = -1;
    return 0;
}

extern "C" int hsg_impute_cells(hsg_ctx* c, int cell_start, int cell_end, float* out_dev, void* stream) {
    CTX_GUARD(c);
    return impute_dev(c, cell_start, cell_end, out_dev, false, (cudaStream_t)stream);
}

extern "C" int hsg_impute_cells_f16(hsg_ctx* c, int cell_start, int cell_end, void* out_dev_half, void* stream) {
    CTX_GUARD(c);
    return impute_dev(c, cell_start, cell_end, out_dev_half, true, (cudaStream_t)stream);
}

static int impute_host(hsg_ctx* c, int cell_start, int cell_end, void* out_host, bool half) {
    HSG_CHECK(c->prepared, "call hsg_prepare first");
    HSG_CHECK(cell_start >= 0 && cell_end <= c->n_cells && cell_start <= cell_end, "cell range out of bounds");
    HSG_CHECK(out_host != nullptr || cell_start == cell_end || c->P == 0, "null output");
    HSG_CHECK(!half || c->k2v2, "16-bit score blocks need the tensor-core scorer with a sigmoid head (HSG_PREC_TC16, HSG_ACT_SIGMOID, d_model 64)");
    const size_t esz = half ? sizeof(__half) : sizeof(float);
    size_t need = (size_t)c->batch_cells * std::max<int64_t>(c->P, 1) * sizeof(float);
    if (c->stage_bytes < need) {
        if (dev_alloc(&c->stage_out[0], need / sizeof(float)) || dev_alloc(&c->stage_out[1], need / sizeof(float))) return 1;
        c->stage_bytes = need;
    }
    if (wait_for_uploads(c, c->stream)) return 1;
    for (int i = 0; i < 4; ++i) c->last_ms[i] = 0.f;
    cudaEvent_t copied[2] = {c->ev[4], c->ev[5]};
    bool used[2] = {false, false};
    int it = 0;
    for (int c0 = cell_start; c0 < cell_end; c0 += c->batch_cells, ++it) {
        int nb = std::min(c->batch_cells, cell_end - c0), slot = it & 1;
        if (used[slot]) HSG_CUDA(cudaStreamWaitEvent(c->stream, copied[slot], 0));   // staging buffer free again
        if (run_batch(c, c0, nb, half ? nullptr : c->stage_out[slot], c->stream, c->profiling, half ? (void*)c->stage_out[slot] : nullptr)) return 1;
        HSG_CUDA(cudaEventRecord(c->stage_ev[slot], c->stream));
        HSG_CUDA(cudaStreamWaitEvent(c->copy_stream, c->stage_ev[slot], 0));
        HSG_CUDA(cudaMemcpyAsync((char*)out_host + (size_t)(c0 - cell_start) * c->P * esz, c->stage_out[slot],
                                 (size_t)nb * c->P * esz, cudaMemcpyDeviceToHost, c->copy_stream));
        HSG_CUDA(cudaEventRecord(copied[slot], c->copy_stream));
        used[slot] = true;
    }
    HSG_CUDA(cudaEventRecord(c->work_ev, c->stream));
    c->work_pending = true;
    HSG_CUDA(cudaStreamSynchronize(c->stream));
    HSG_CUDA(cudaStreamSynchronize(c->copy_stream));
    c->fixed_cell = -1;
    return 0;
}

extern "C" int hsg_impute_cells_host(hsg_ctx* c, int cell_start, int cell_end, float* out_host) {
    CTX_GUARD(c);
    return impute_host(c, cell_start, cell_end, out_host, false);
}

extern "C" int hsg_impute_cells_host_f16(hsg_ctx* c, int cell_start, int cell_end, void* out_host_half) {
    CTX_GUARD(c);
    return impute_host(c, cell_start, cell_end, out_host_half, true);
}

extern "C" int hsg_refresh_contacts(hsg_ctx* c, int cell_start, int cell_end, const int32_t* col, const float* val, int64_t nnz_range) {
    CTX_GUARD(c);
    HSG_CHECK(c->row_ptr != nullptr && !c->h_cell_ptr.empty(), "hsg_refresh_contacts: call hsg_set_contacts first");
    HSG_CHECK(cell_start >= 0 && cell_end <= c->n_cells && cell_start <= cell_end, "cell range out of bounds");
    const int64_t lo = c->h_cell_ptr[cell_start], hi = c->h_cell_ptr[cell_end];
    HSG_CHECK(nnz_range == hi - lo, "hsg_refresh_contacts: the block must hold exactly the stored entries of these cells' rows");
    if (nnz_range == 0) return 0;
    HSG_CHECK(col != nullptr && val != nullptr, "null contact arrays");
    if (!c->h2d_stream) HSG_CUDA(cudaStreamCreateWithFlags(&c->h2d_stream, cudaStreamNonBlocking));
    // imputation work already queued may still read these rows: the copies go behind it (event, no device-wide sync)
    if (c->work_pending) { HSG_CUDA(cudaStreamWaitEvent(c->h2d_stream, c->work_ev, 0)); c->work_pending = false; }
    HSG_CUDA(cudaMemcpyAsync(c->col + lo, col, (size_t)nnz_range * sizeof(int32_t), cudaMemcpyHostToDevice, c->h2d_stream));
    HSG_CUDA(cudaMemcpyAsync(c->val + lo, val, (size_t)nnz_range * sizeof(float), cudaMemcpyHostToDevice, c->h2d_stream));
    HSG_CUDA(cudaEventRecord(c->refresh_ev, c->h2d_stream));
    c->refresh_pending = true;
    return 0;
}

extern "C" int hsg_scorer_variant(hsg_ctx* c, int* variant_out) {
    HSG_CHECK(c != nullptr && variant_out != nullptr, "bad arguments");
    HSG_CHECK(c->prepared, "call hsg_prepare first");
    *variant_out = !c->use_umma ? HSG_PREC_FP32 : (c->umma_hi ? HSG_PREC_TC16_HI : HSG_PREC_TC16);
    return 0;
}

extern "C" int hsg_fix_cell(hsg_ctx* c, int cell_1based, void* stream) {
    CTX_GUARD(c);
    HSG_CHECK(c->prepared, "call hsg_prepare first");
    HSG_CHECK(cell_1based >= 1 && cell_1based <= c->n_cells, "cell id out of range");
    cudaStream_t s = (cudaStream_t)stream;
    if (launch_gather(c, nullptr, cell_1based - 1, 1, s)) return 1;
    if (launch_token_project(c, cell_1based - 1, 1, 0, s)) return 1;
    c->fixed_cell = cell_1based - 1;
    return 0;
}

extern "C" int hsg_score_triplets(hsg_ctx* c, const int64_t* x_dev, int64_t B, float* out_dev, void* stream) {
    CTX_GUARD(c);
    HSG_CHECK(c->prepared, "call hsg_prepare first");
    HSG_CHECK(c->fixed_cell >= 0, "call hsg_fix_cell first (GraphSageEncoder_with_weights.fix_cell2)");
    HSG_CHECK(B >= 0 && (B == 0 || (x_dev && out_dev)), "bad triplet buffers");
    return launch_score_triplets_f32(c, x_dev, B, out_dev, (cudaStream_t)stream);
}

extern "C" int64_t hsg_launch_count(hsg_ctx* c) { return c ? c->launches : 0; }
extern "C" int hsg_last_timing(hsg_ctx* c, float* ms4) {
    HSG_CHECK(c && ms4, "bad arguments");
    for (int i = 0; i < 4; ++i) ms4[i] = c->last_ms[i];
    return 0;
}
extern "C" int hsg_set_profiling(hsg_ctx* c, int enabled) {
    HSG_CHECK(c != nullptr, "null context");
    c->profiling = enabled != 0;
    return 0;
}
