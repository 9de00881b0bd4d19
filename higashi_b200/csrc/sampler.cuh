// sampler.cuh -- training batch producer on the device, included by train.cu.
//
//   hsg_sample_batch  : negatives by rejection sampling against the packed contact CSR (generate_negative_cpu + check_nonzero,
//                       Higashi_wrapper.py:71-168), labels / weights (one_thread_generate_neg :236-259), and the neighbour lists
//                       of all 2B bin tokens as the CSR that hsg_score_fwd consumes (one_thread_generate_neg :261-333 +
//                       to_neighs_to_mask :171-221): k+1 neighbour-cell rows weighted and merged (sum_duplicates), 40 % partner
//                       removal in training, log1p, L1 normalisation.
// The batch stays in the training workspace; hsg_score_fwd(x_host = NULL) runs on it.  The reference sampler is unseeded, so
// parity is statistical for the draws and exact for everything deterministic (neighbour lists with removal off, labels,
// weights, the acceptance rule).  Random numbers are counter-based (splitmix64 of (seed, stream, index, counter)).
#pragma once

struct SamplerWs {
    int64_t* pos = nullptr; float* pos_w = nullptr; int cap_pos = 0;
    int64_t* neg = nullptr; int* flag = nullptr; int* scan = nullptr; int cap_neg = 0;
    int64_t* x_all = nullptr; float* y_all = nullptr; float* w_all = nullptr; int cap_all = 0;
    int* cand_cnt = nullptr; int* cand_off = nullptr; int cap_rows = 0;
    int* s_col = nullptr; float* s_val = nullptr; float* s_sum = nullptr; int* s_first = nullptr; int64_t cap_cand = 0;
    int* counts = nullptr;          // device: [0] = hyperedges in the batch, [1] = neighbour candidates
    int B = 0; int64_t nnz = 0; bool ready = false;
    // asynchronous production: begin() runs on `side`, sizes land in pinned memory, end() finishes on the caller's stream
    cudaStream_t side = nullptr; cudaEvent_t sized = nullptr, consumed = nullptr; int* h_counts = nullptr;
    bool pending = false, consumed_valid = false;
    int p_Bp = 0, p_chrom = 0, p_training = 0; unsigned long long p_seed = 0;
    float* draws = nullptr; int n_draws = 0;      // hsg_sample_set_removal_draws: caller-supplied uniform draws of the partner removal
};
static std::map<hsg_ctx*, SamplerWs*> g_sampler;

__device__ __forceinline__ uint32_t rnd_u32(unsigned long long seed, unsigned stream, unsigned long long idx, unsigned ctr) {
    unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (idx + 1) + ((unsigned long long)stream << 48) + 0xD1B54A32D192ED03ull * (ctr + 1);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; z ^= z >> 31;
    return (uint32_t)(z >> 32);
}
__device__ __forceinline__ float rnd_uniform(unsigned long long seed, unsigned stream, unsigned long long idx, unsigned ctr) {
    return (float)(rnd_u32(seed, stream, idx, ctr) >> 8) * (1.0f / 16777216.0f);
}
__device__ __forceinline__ int rnd_below(int n, unsigned long long seed, unsigned stream, unsigned long long idx, unsigned ctr) {
    return (int)(((unsigned long long)rnd_u32(seed, stream, idx, ctr) * (unsigned long long)(n > 0 ? n : 1)) >> 32);
}

// any stored contact in the 3x3 neighbourhood of (r, c) of cell `cell0` (check_nonzero, mem_efficient branch: rows / columns
// clamped to [0, n-2]) -- `wide` = 0 tests the single entry (the other branch of check_nonzero)
__device__ bool contact_near(const int64_t* __restrict__ row_ptr, const int32_t* __restrict__ col, int64_t row0, int n, int r, int c,
                             int wide) {
    const int hi_clip = n > 1 ? n - 2 : 0;
    for (int dr = -wide; dr <= wide; ++dr) {
        int rr = min(max(r + dr, 0), wide ? hi_clip : n - 1);
        int c_lo = wide ? min(max(c - 1, 0), hi_clip) : c, c_hi = wide ? min(max(c + 1, 0), hi_clip) : c;
        for (int64_t e = row_ptr[row0 + rr]; e < row_ptr[row0 + rr + 1]; ++e) {
            int cc = col[e];
            if (cc >= c_lo && cc <= c_hi) return true;
        }
    }
    return false;
}

// one thread per (positive j, negative i)
__global__ void neg_sample_kernel(const int64_t* __restrict__ pos, int Bp, int neg_num, float pair_ratio, int max_bin, int wide,
                                  const int64_t* __restrict__ row_ptr, const int32_t* __restrict__ col, int n_bins, int n_cells,
                                  int bin_off_c, int n, int bin_base, unsigned long long seed, int64_t* __restrict__ neg,
                                  int* __restrict__ flag) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= Bp * neg_num) return;
    const int j = q / neg_num;
    const int64_t p0 = pos[j * 3], p1 = pos[j * 3 + 1], p2 = pos[j * 3 + 2];
    const bool hard = rnd_uniform(seed, 1, q, 0) <= pair_ratio;
    const int change = rnd_below(3, seed, 1, q, 1);
    int ok = 0;
    int64_t c0 = p0, c1 = p1, c2 = p2;
    for (int trial = 0; trial < 49 && !ok; ++trial) {
        const unsigned ctr = 2 + trial * 4;
        int64_t n0 = p0;
        int b1 = (int)(p1 - bin_base), b2 = (int)(p2 - bin_base);
        if (hard) {
            if (change == 0) n0 = rnd_below(n_cells, seed, 1, q, ctr) + 1;
            else {
                const int other = (int)(p2 - bin_base);           // the reference always windows around element [2] (wrapper:131-133)
                const int lo = max(0, other - max_bin), hi = min(n, other + max_bin);
                const int nb = lo + rnd_below(hi - lo, seed, 1, q, ctr);
                if (change == 1) b1 = nb; else b2 = nb;
            }
        } else {
            n0 = rnd_below(n_cells, seed, 1, q, ctr) + 1;
            b1 = rnd_below(n, seed, 1, q, ctr + 1);
            const int lo = max(0, b1 - max_bin), hi = min(n, b1 + max_bin);
            b2 = lo + rnd_below(hi - lo, seed, 1, q, ctr + 2);
        }
        if (b1 > b2) { int t = b1; b1 = b2; b2 = t; }
        const int dist = b2 - b1;
        if (dist >= max_bin || dist <= 1) continue;                // not a suitable sample: back to the positive, which is rejected
        const int64_t row0 = (int64_t)(n0 - 1) * n_bins + bin_off_c;
        if (!contact_near(row_ptr, col, row0, n, b1, b2, wide)) { ok = 1; c0 = n0; c1 = b1 + bin_base; c2 = b2 + bin_base; }
    }
    neg[q * 3] = c0; neg[q * 3 + 1] = c1; neg[q * 3 + 2] = c2;
    flag[q] = ok;
}

// exclusive scan of n ints by ONE block (n is a few thousand); total -> *total_out
__global__ void block_scan_kernel(const int* __restrict__ in, int n, int* __restrict__ out, int* __restrict__ total_out, int add) {
    __shared__ int warp_tot[32];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += blockDim.x) {
        const int i = base + threadIdx.x;
        const int v = i < n ? in[i] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, incl, o); if ((threadIdx.x & 31) >= o) incl += t; }
        if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = incl;
        __syncthreads();
        if (threadIdx.x < 32) {
            int w = threadIdx.x < (blockDim.x >> 5) ? warp_tot[threadIdx.x] : 0, wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, wi, o); if (threadIdx.x >= o) wi += t; }
            warp_tot[threadIdx.x] = wi - w;
        }
        __syncthreads();
        const int excl = carry + warp_tot[threadIdx.x >> 5] + incl - v;
        if (i < n) out[i] = excl;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total_out = carry + add;
}

// x = [positives ; accepted negatives], y = [1 ; 0], w = [edge weight ; correction]   (wrapper:246-259)
__global__ void build_batch_kernel(const int64_t* __restrict__ pos, const float* __restrict__ pos_w, int Bp, const int64_t* __restrict__ neg,
                                   const int* __restrict__ flag, const int* __restrict__ scan, int n_neg, float correction,
                                   int64_t* __restrict__ x, float* __restrict__ y, float* __restrict__ w) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < Bp) {
        x[i * 3] = pos[i * 3]; x[i * 3 + 1] = pos[i * 3 + 1]; x[i * 3 + 2] = pos[i * 3 + 2];
        y[i] = 1.0f; w[i] = pos_w ? pos_w[i] : 1.0f;
    } else if (i < Bp + n_neg) {
        const int q = i - Bp;
        if (flag[q]) {
            const int o = Bp + scan[q];
            x[o * 3] = neg[q * 3]; x[o * 3 + 1] = neg[q * 3 + 1]; x[o * 3 + 2] = neg[q * 3 + 2];
            y[o] = 0.0f; w[o] = correction;
        }
    }
}

// neighbour candidates of token row r (hyperedge r/2, bin r&1): sum of the k1 neighbour cells' row lengths
__global__ void cand_count_kernel(const int64_t* __restrict__ x, const int* __restrict__ counts, int rows_max,
                                  const int64_t* __restrict__ row_ptr, const int32_t* __restrict__ nbr, int k1, int weighted, int n_bins,
                                  int bin_off_c, int bin_base, int* __restrict__ cand_cnt) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows_max) return;
    int cnt = 0;
    if (r < 2 * counts[0]) {
        const int64_t cell0 = x[(r >> 1) * 3] - 1;
        const int b = (int)(x[(r >> 1) * 3 + 1 + (r & 1)] - bin_base);
        for (int q = 0; q < k1; ++q) {
            const int64_t src = weighted ? nbr[cell0 * k1 + q] : cell0;
            const int64_t row = src * n_bins + bin_off_c + b;
            cnt += (int)(row_ptr[row + 1] - row_ptr[row]);
        }
    }
    cand_cnt[r] = cnt;
}

// one warp per token row: gather (weighted) candidates in neighbour order, merge duplicates (sums in neighbour order), sort by
// column, optional partner removal, log1p, L1 normalisation -> CSR segment [indptr[r], indptr[r] + rowcnt[r])
__global__ void neigh_rows_kernel(const int64_t* __restrict__ x, int B, const int64_t* __restrict__ row_ptr, const int32_t* __restrict__ col,
                                  const float* __restrict__ val, const int32_t* __restrict__ nbr, const float* __restrict__ nbr_w, int k1,
                                  int weighted, int n_bins, int bin_off_c, int bin_base, int training, unsigned long long seed,
                                  const float* __restrict__ draws, int n_draws, const int* __restrict__ cand_off, int* __restrict__ s_col, float* __restrict__ s_val,
                                  float* __restrict__ s_sum, int* __restrict__ s_first, int* __restrict__ indptr, int* __restrict__ rowcnt,
                                  int* __restrict__ out_col, float* __restrict__ out_val) {
    const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (r >= 2 * B) return;
    const int64_t cell0 = x[(r >> 1) * 3] - 1;
    const int b = (int)(x[(r >> 1) * 3 + 1 + (r & 1)] - bin_base);
    const int partner = (int)(x[(r >> 1) * 3 + 2 - (r & 1)] - bin_base);
    const int off = cand_off[r];
    int n = 0;
    for (int q = 0; q < k1; ++q) {
        const int64_t src = weighted ? nbr[cell0 * k1 + q] : cell0;
        const float w = weighted ? nbr_w[cell0 * k1 + q] : 1.0f;
        const int64_t row = src * n_bins + bin_off_c + b;
        const int64_t lo = row_ptr[row];
        const int len = (int)(row_ptr[row + 1] - lo);
        for (int i = lane; i < len; i += 32) {
            s_col[off + n + i] = col[lo + i];
            s_val[off + n + i] = weighted ? __fmul_rn(val[lo + i], w) : val[lo + i];
        }
        n += len;
    }
    __syncwarp();
    // phase 1: owners (first occurrence of a column) and their sums in candidate order
    int n_own = 0;
    for (int e0 = 0; e0 < n; e0 += 32) {
        const int e = e0 + lane;
        bool first = false;
        if (e < n) {
            const int c = s_col[off + e];
            first = true;
            float sum = 0.f;
            for (int e2 = 0; e2 < n; ++e2)
                if (s_col[off + e2] == c) { if (e2 < e) { first = false; break; } sum = __fadd_rn(sum, s_val[off + e2]); }
            s_first[off + e] = first ? 1 : 0;
            s_sum[off + e] = sum;
        }
        n_own += __popc(__ballot_sync(0xffffffffu, first));
    }
    __syncwarp();
    // partner removal (training, 40 %, only if another neighbour remains)                 (wrapper:296-300)
    const float u_rm = (draws != nullptr && r < n_draws) ? draws[r] : rnd_uniform(seed, 2, (unsigned long long)r, 0);
    bool drop = training && n_own > 1 && u_rm >= 0.6f;
    // phase 2: rank among the kept owners = sorted position; log1p; L1 normalisation
    float part = 0.f;
    int partner_present = 0;
    for (int e0 = 0; e0 < n; e0 += 32) {
        const int e = e0 + lane;
        if (e < n && s_first[off + e]) {
            if (drop && s_col[off + e] == partner) partner_present = 1;
            else part += log1pf(s_sum[off + e]);
        }
    }
    part = warp_sum(part);
    partner_present = __any_sync(0xffffffffu, partner_present);
    for (int e0 = 0; e0 < n; e0 += 32) {
        const int e = e0 + lane;
        if (e < n && s_first[off + e]) {
            const int c = s_col[off + e];
            if (drop && c == partner) continue;
            int rank = 0;
            for (int e2 = 0; e2 < n; ++e2) {
                const int c2 = s_col[off + e2];
                if (s_first[off + e2] && c2 < c && !(drop && c2 == partner)) ++rank;
            }
            const float v = log1pf(s_sum[off + e]);
            out_col[off + rank] = c;
            out_val[off + rank] = part > 0.f ? __fdiv_rn(v, part) : v;
        }
    }
    if (lane == 0) {
        indptr[r] = off;
        rowcnt[r] = n_own - ((drop && partner_present) ? 1 : 0);
        if (r == 2 * B - 1) indptr[2 * B] = off + n;
    }
}

template <typename T>
static int grow(T** p, size_t n) {
    if (*p) cudaFree(*p);
    *p = nullptr;
    HSG_CUDA(cudaMalloc((void**)p, std::max<size_t>(n, 1) * sizeof(T)));
    return 0;
}

static int sampler_ws_create(hsg_ctx* c) {
    SamplerWs*& sw = g_sampler[c];
    if (sw) return 0;
    sw = new SamplerWs();
    HSG_CUDA(cudaMalloc((void**)&sw->counts, 2 * sizeof(int)));
    HSG_CUDA(cudaStreamCreateWithFlags(&sw->side, cudaStreamNonBlocking));
    HSG_CUDA(cudaEventCreateWithFlags(&sw->sized, cudaEventDisableTiming));
    HSG_CUDA(cudaEventCreateWithFlags(&sw->consumed, cudaEventDisableTiming));
    HSG_CUDA(cudaHostAlloc((void**)&sw->h_counts, 2 * sizeof(int), cudaHostAllocDefault));
    return 0;
}

// Phase 1 (asynchronous, on the producer's own stream): negatives, labels / weights, candidate counts; the two sizes the host
// needs (hyperedges in the batch, neighbour candidates) are copied to pinned memory and an event is recorded.
extern "C" int hsg_sample_batch_begin(hsg_ctx* c, const int64_t* pos_host, const float* pos_w_host, int Bp, int chrom, int neg_num,
                                      float pair_ratio, int max_bin, int wide_check, int training, float neg_weight, uint64_t seed) {
    HSG_CHECK(c && c->d > 0, "context not initialised");
    HSG_CUDA(cudaSetDevice(c->device));
    HSG_CHECK(c->row_ptr != nullptr, "hsg_sample_batch: contacts not set (hsg_set_contacts)");
    HSG_CHECK(pos_host && Bp > 0 && chrom >= 0 && chrom < c->n_chrom && neg_num >= 0, "bad arguments");
    HSG_CHECK(!c->weighted || c->nbr != nullptr, "neighbour graph flagged but missing");
    if (sampler_ws_create(c)) return 1;
    SamplerWs* sw = g_sampler[c];
    HSG_CHECK(!sw->pending, "hsg_sample_batch_begin: the previous batch has not been finished (hsg_sample_batch_end)");
    cudaStream_t s = sw->side;
    // the scratch of the previous batch may still be read by work queued on the consumer's stream
    if (sw->consumed_valid) HSG_CUDA(cudaStreamWaitEvent(s, sw->consumed, 0));
    const int n = c->bins[chrom], bin_base = c->n_cells + 1 + c->bin_off[chrom], n_neg = Bp * neg_num, Bmax = Bp + n_neg;
    for (int i = 0; i < Bp; ++i) {
        HSG_CHECK(pos_host[i * 3] >= 1 && pos_host[i * 3] <= c->n_cells, "positive: cell id out of range");
        HSG_CHECK(pos_host[i * 3 + 1] >= bin_base && pos_host[i * 3 + 2] < bin_base + n && pos_host[i * 3 + 1] <= pos_host[i * 3 + 2],
                  "positive: bins must be sorted ids of the batch chromosome");
    }
    const bool regrow = sw->cap_pos < Bp || sw->cap_neg < n_neg || sw->cap_all < Bmax || sw->cap_rows < 2 * Bmax;
    if (regrow) HSG_CUDA(cudaDeviceSynchronize());          // buffers are about to be replaced
    if (sw->cap_pos < Bp) { if (grow(&sw->pos, (size_t)Bp * 3) || grow(&sw->pos_w, (size_t)Bp)) return 1; sw->cap_pos = Bp; }
    if (sw->cap_neg < n_neg) {
        if (grow(&sw->neg, (size_t)n_neg * 3) || grow(&sw->flag, (size_t)n_neg) || grow(&sw->scan, (size_t)n_neg)) return 1;
        sw->cap_neg = n_neg;
    }
    if (sw->cap_all < Bmax) {
        if (grow(&sw->x_all, (size_t)Bmax * 3) || grow(&sw->y_all, (size_t)Bmax) || grow(&sw->w_all, (size_t)Bmax)) return 1;
        sw->cap_all = Bmax;
    }
    if (sw->cap_rows < 2 * Bmax) {
        if (grow(&sw->cand_cnt, (size_t)2 * Bmax) || grow(&sw->cand_off, (size_t)2 * Bmax)) return 1;
        sw->cap_rows = 2 * Bmax;
    }
    // pageable sources: the copies return once the data is staged, the caller may reuse its arrays
    HSG_CUDA(cudaMemcpyAsync(sw->pos, pos_host, (size_t)Bp * 3 * sizeof(int64_t), cudaMemcpyHostToDevice, s));
    if (pos_w_host) HSG_CUDA(cudaMemcpyAsync(sw->pos_w, pos_w_host, (size_t)Bp * sizeof(float), cudaMemcpyHostToDevice, s));
    if (n_neg > 0) {
        neg_sample_kernel<<<nblk(n_neg), 256, 0, s>>>(sw->pos, Bp, neg_num, pair_ratio, max_bin, wide_check ? 1 : 0, c->row_ptr, c->col,
                                                      c->n_bins, c->n_cells, c->bin_off[chrom], n, bin_base, seed, sw->neg, sw->flag);
        TR_CHECK_LAUNCH(c);
        block_scan_kernel<<<1, 1024, 0, s>>>(sw->flag, n_neg, sw->scan, sw->counts, Bp); TR_CHECK_LAUNCH(c);
    } else {
        sw->h_counts[0] = Bp;
        HSG_CUDA(cudaMemcpyAsync(sw->counts, sw->h_counts, sizeof(int), cudaMemcpyHostToDevice, s));
    }
    build_batch_kernel<<<nblk(Bmax), 256, 0, s>>>(sw->pos, pos_w_host ? sw->pos_w : nullptr, Bp, sw->neg, sw->flag, sw->scan, n_neg,
                                                  neg_weight, sw->x_all, sw->y_all, sw->w_all);
    TR_CHECK_LAUNCH(c);
    cand_count_kernel<<<nblk(2 * Bmax), 256, 0, s>>>(sw->x_all, sw->counts, 2 * Bmax, c->row_ptr, c->nbr, c->k1, c->weighted ? 1 : 0,
                                                     c->n_bins, c->bin_off[chrom], bin_base, sw->cand_cnt);
    TR_CHECK_LAUNCH(c);
    block_scan_kernel<<<1, 1024, 0, s>>>(sw->cand_cnt, 2 * Bmax, sw->cand_off, sw->counts + 1, 0); TR_CHECK_LAUNCH(c);
    HSG_CUDA(cudaMemcpyAsync(sw->h_counts, sw->counts, 2 * sizeof(int), cudaMemcpyDeviceToHost, s));
    HSG_CUDA(cudaEventRecord(sw->sized, s));
    sw->pending = true; sw->ready = false;
    sw->p_Bp = Bp; sw->p_chrom = chrom; sw->p_training = training; sw->p_seed = seed;
    return 0;
}

// Phase 2 (on the consumer's stream): waits for the sizes (normally long since there), sizes the training workspace, moves the
// batch in and builds the neighbour CSR.  Nothing here waits for work queued on `stream` itself.
extern "C" int hsg_sample_batch_end(hsg_ctx* c, int* B_out, int64_t* nnz_out, void* stream) {
    HSG_CHECK(c && B_out && nnz_out, "bad arguments");
    HSG_CUDA(cudaSetDevice(c->device));
    auto it = g_sampler.find(c);
    HSG_CHECK(it != g_sampler.end() && it->second->pending, "hsg_sample_batch_end without hsg_sample_batch_begin");
    SamplerWs* sw = it->second;
    cudaStream_t s = (cudaStream_t)stream;
    HSG_CUDA(cudaEventSynchronize(sw->sized));
    sw->pending = false;
    const int B = sw->h_counts[0];
    const int64_t n_cand = sw->h_counts[1];
    const int chrom = sw->p_chrom, bin_base = c->n_cells + 1 + c->bin_off[chrom];
    if (sw->cap_cand < n_cand) {
        HSG_CUDA(cudaDeviceSynchronize());
        if (grow(&sw->s_col, (size_t)n_cand) || grow(&sw->s_val, (size_t)n_cand) || grow(&sw->s_sum, (size_t)n_cand) ||
            grow(&sw->s_first, (size_t)n_cand)) return 1;
        sw->cap_cand = n_cand;
    }
    {
        auto wit = g_ws.find(c);
        const bool realloc_ws = wit == g_ws.end() || !wit->second || wit->second->maxB < B || wit->second->max_nnz < (int)n_cand;
        if (realloc_ws) HSG_CUDA(cudaDeviceSynchronize());   // the workspace is about to be replaced: let queued steps finish
        // grow with head-room so that the next slightly larger batch does not reallocate again
        if (realloc_ws && ws_alloc(c, B + B / 8, (int)(n_cand + n_cand / 4))) return 1;
    }
    TrainWs* w = g_ws[c];
    HSG_CUDA(cudaStreamWaitEvent(s, sw->sized, 0));
    HSG_CUDA(cudaMemcpyAsync(w->x, sw->x_all, (size_t)B * 3 * sizeof(int64_t), cudaMemcpyDeviceToDevice, s));
    HSG_CUDA(cudaMemcpyAsync(w->y, sw->y_all, (size_t)B * sizeof(float), cudaMemcpyDeviceToDevice, s));
    HSG_CUDA(cudaMemcpyAsync(w->w, sw->w_all, (size_t)B * sizeof(float), cudaMemcpyDeviceToDevice, s));
    neigh_rows_kernel<<<(2 * B + 7) / 8, 256, 0, s>>>(w->x, B, c->row_ptr, c->col, c->val, c->nbr, c->nbr_w, c->k1, c->weighted ? 1 : 0,
                                                      c->n_bins, c->bin_off[chrom], bin_base, sw->p_training ? 1 : 0, sw->p_seed, sw->draws,
                                                      sw->n_draws, sw->cand_off,
                                                      sw->s_col, sw->s_val, sw->s_sum, sw->s_first, w->indptr, w->rowcnt, w->cols, w->vals);
    TR_CHECK_LAUNCH(c);
    HSG_CUDA(cudaEventRecord(sw->consumed, s));
    sw->consumed_valid = true;
    sw->B = B; sw->nnz = n_cand; sw->ready = true;
    w->sampled = true; w->chrom = chrom;
    *B_out = B; *nnz_out = n_cand;
    return 0;
}

extern "C" int hsg_sample_batch(hsg_ctx* c, const int64_t* pos_host, const float* pos_w_host, int Bp, int chrom, int neg_num,
                                float pair_ratio, int max_bin, int wide_check, int training, float neg_weight, uint64_t seed,
                                int* B_out, int64_t* nnz_out, void* stream) {
    if (hsg_sample_batch_begin(c, pos_host, pos_w_host, Bp, chrom, neg_num, pair_ratio, max_bin, wide_check, training, neg_weight, seed))
        return 1;
    return hsg_sample_batch_end(c, B_out, nnz_out, stream);
}

// reproducibility hook: the uniform draws of the 40 % partner removal (one per bin-token row, `np.random.rand` of wrapper:270)
// come from this array for all following batches (rows past n use the generator); NULL / n = 0 restores the generator
extern "C" int hsg_sample_set_removal_draws(hsg_ctx* c, const float* draws_host, int n) {
    HSG_CHECK(c && c->d > 0, "context not initialised");
    HSG_CUDA(cudaSetDevice(c->device));
    HSG_CHECK(n >= 0 && (n == 0 || draws_host != nullptr), "bad arguments");
    if (sampler_ws_create(c)) return 1;
    SamplerWs* sw = g_sampler[c];
    HSG_CUDA(cudaDeviceSynchronize());
    if (sw->draws) { cudaFree(sw->draws); sw->draws = nullptr; }
    sw->n_draws = 0;
    if (n > 0 && draws_host) {
        HSG_CUDA(cudaMalloc((void**)&sw->draws, (size_t)n * sizeof(float)));
        HSG_CUDA(cudaMemcpy(sw->draws, draws_host, (size_t)n * sizeof(float), cudaMemcpyHostToDevice));
        sw->n_draws = n;
    }
    return 0;
}

// debug / test access to the batch the producer left on the device (any pointer may be NULL)
extern "C" int hsg_sampled_fetch(hsg_ctx* c, int64_t* x_host, float* y_host, float* w_host, int32_t* indptr_host, int32_t* rowcnt_host,
                                 int32_t* cols_host, float* vals_host, void* stream) {
    HSG_CHECK(c != nullptr, "null context");
    HSG_CUDA(cudaSetDevice(c->device));
    auto it = g_sampler.find(c);
    HSG_CHECK(it != g_sampler.end() && it->second->ready, "no sampled batch");
    SamplerWs* sw = it->second;
    TrainWs* w = g_ws[c];
    cudaStream_t s = (cudaStream_t)stream;
    const int B = sw->B;
    if (x_host) HSG_CUDA(cudaMemcpyAsync(x_host, w->x, (size_t)B * 3 * sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    if (y_host) HSG_CUDA(cudaMemcpyAsync(y_host, w->y, (size_t)B * sizeof(float), cudaMemcpyDeviceToHost, s));
    if (w_host) HSG_CUDA(cudaMemcpyAsync(w_host, w->w, (size_t)B * sizeof(float), cudaMemcpyDeviceToHost, s));
    if (indptr_host) HSG_CUDA(cudaMemcpyAsync(indptr_host, w->indptr, (size_t)(2 * B + 1) * sizeof(int), cudaMemcpyDeviceToHost, s));
    if (rowcnt_host) HSG_CUDA(cudaMemcpyAsync(rowcnt_host, w->rowcnt, (size_t)(2 * B) * sizeof(int), cudaMemcpyDeviceToHost, s));
    if (cols_host && sw->nnz) HSG_CUDA(cudaMemcpyAsync(cols_host, w->cols, (size_t)sw->nnz * sizeof(int), cudaMemcpyDeviceToHost, s));
    if (vals_host && sw->nnz) HSG_CUDA(cudaMemcpyAsync(vals_host, w->vals, (size_t)sw->nnz * sizeof(float), cudaMemcpyDeviceToHost, s));
    HSG_CUDA(cudaStreamSynchronize(s));
    return 0;
}
