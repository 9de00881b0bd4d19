/* hsg.h -- C ABI of libhsg.so: the B200-native Hyper-SAGNN hyperedge scorer
 * (Higashi's one data-parallel hot path: scoring (cell, bin_i, bin_j) triplets).
 *
 * The reference (ma-compbio/Higashi) has no FFI layer: its boundary for this path is the Python
 * module API.  Each entry point below names the reference interface it replaces (file:line
 * relative to the reference repo); INTEGRATION.md shows the ctypes binding a maintainer adds.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; hsg_last_error() then holds a
 *     message (thread-local).  The host mirror raises a Python exception from it.
 *   - plain pointers and sizes only.  "host" pointers are ordinary (ideally pinned) host memory,
 *     "dev" pointers are device memory on the context's GPU.  `stream` is a cudaStream_t passed
 *     as void* (NULL = legacy default stream).
 *   - one context per device; calls on one context are serialised by the caller.
 *   - ids follow the reference's node-id space where stated: 0 = pad, 1..n_cells = cells,
 *     then bins in chromosome order (Process.py:795-797); "local" ids are 0-based.
 *   - there is NO CPU fallback: every compute call fails if the device is unusable.
 */
#ifndef HSG_H_
#define HSG_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct hsg_ctx hsg_ctx;

#define HSG_ABI_VERSION 1

/* precision of the fused scorer (BASELINE north_star: 1e-5 abs in fp32 mode, 1e-3 in bf16 mode) */
#define HSG_PREC_FP32 0   /* CUDA-core fp32 everywhere                                            */
#define HSG_PREC_TC16 1   /* tcgen05 tensor cores: 16-bit (fp16) operands, fp32 accumulation in TMEM, for the
                             d_model projections; LayerNorm / softmax / residual / square stay fp32.
                             fp16 rather than bf16 operands: same tensor throughput, and the only one of the
                             two that meets the 1e-3 score bar (DESIGN.md "precision")                     */

#define HSG_PREC_TC16_HI 2 /* tensor cores, high-accuracy variant: fp32 Q/K/V tables and attention, hi/lo fp16 splits of the
                              fc1 / classifier weights and of the squared-difference operand.  HSG_PREC_TC16 selects it
                              automatically for softplus / raw heads, where the logit error is not damped by a sigmoid    */

/* activation applied to the `mean` head (Impute.py:218-221, Modules.py:814-817) */
#define HSG_ACT_NONE     0
#define HSG_ACT_SIGMOID  1   /* loss_mode == "classification" */
#define HSG_ACT_SOFTPLUS 2   /* zinb / rank / regression      */

/* Weight slots.  Shapes use d = d_model, H*dk = 128, C = n_chrom, cf = cell_feats width.
 * Names on the right are the reference's state_dict keys (Modules.py:657-710, 975-1027, 1098-1134,
 * 1387-1424).  All tensors are fp32, row-major, exactly as stored in the state_dict.            */
enum hsg_weight {
    HSG_W_QS = 0,        /* [128,d]  encode1.mul_head_attn.w_qs.weight                     */
    HSG_W_KS,            /* [128,d]  encode1.mul_head_attn.w_ks.weight                     */
    HSG_W_VS,            /* [128,d]  encode1.mul_head_attn.w_vs.weight                     */
    HSG_W_FC1,           /* [d,128]  encode1.mul_head_attn.fc1.w_stack.0.weight            */
    HSG_W_FC2,           /* [d,128]  encode1.mul_head_attn.fc2.w_stack.0.weight            */
    HSG_W_ALN1_G, HSG_W_ALN1_B,   /* [d] encode1.mul_head_attn.layer_norm1.{weight,bias}  */
    HSG_W_ALN2_G, HSG_W_ALN2_B,   /* [d] encode1.mul_head_attn.layer_norm2.{weight,bias}  */
    HSG_W_ALN3_G, HSG_W_ALN3_B,   /* [d] encode1.mul_head_attn.layer_norm3.{weight,bias}  */
    HSG_W_PFF_LN_G, HSG_W_PFF_LN_B, /* [d] encode1.pff_n1.layer_norm.{weight,bias}        */
    HSG_W_PFF_W0, HSG_W_PFF_B0,   /* [d,d],[d] encode1.pff_n1.w_stack.0.{weight[:,:,0],bias} */
    HSG_W_PFF_W1, HSG_W_PFF_B1,   /* [d,d],[d] encode1.pff_n1.w_stack.1.{weight[:,:,0],bias} */
    HSG_W_LN1_G, HSG_W_LN1_B,     /* [d] layer_norm1.{weight,bias}                         */
    HSG_W_LN2_G, HSG_W_LN2_B,     /* [d] layer_norm2.{weight,bias}                         */
    HSG_W_CLS_W0, HSG_W_CLS_B0,   /* [d/2,d],[d/2] pff_classifier.w_stack.0                */
    HSG_W_CLS_W1, HSG_W_CLS_B1,   /* [1,d/2],[1]   pff_classifier.w_stack.1                */
    HSG_W_VAR_W0, HSG_W_VAR_B0, HSG_W_VAR_W1, HSG_W_VAR_B1,     /* pff_classifier_var.*   */
    HSG_W_PRB_W0, HSG_W_PRB_B0, HSG_W_PRB_W1, HSG_W_PRB_B1,     /* pff_classifier_proba.* */
    HSG_W_SAGE_W, HSG_W_SAGE_B,   /* [d,2d],[d] encode1.dynamic_nn.nn.{weight,bias}        */
    HSG_W_XP_W0, HSG_W_XP_B0, HSG_W_XP_W1, HSG_W_XP_B1,         /* extra_proba.w_stack.*  : [4,2(C+1)+cf],[4],[1,4],[1] */
    HSG_W_XP2_W0, HSG_W_XP2_B0, HSG_W_XP2_W1, HSG_W_XP2_B1,     /* extra_proba2.w_stack.* */
    HSG_W_XP3_W0, HSG_W_XP3_B0, HSG_W_XP3_W1, HSG_W_XP3_B1,     /* extra_proba3.w_stack.* */
    HSG_W_ATTR,          /* [n_cells+1+n_bins, C+1] attribute_dict_embedding.weight        */
    HSG_W_CELL_FEATS,    /* [n_cells+1, cf]         Hyper_SAGNN.cell_feats (buffer)        */
    HSG_W_COUNT
};

/* ---- lifetime -------------------------------------------------------------------------- */
int  hsg_abi_version(void);
const char* hsg_last_error(void);
/* replaces: module-level `device` selection (Modules.py:25) / torch.cuda.set_device (Impute.py:333) */
int  hsg_create(hsg_ctx** out, int device_ordinal);
int  hsg_destroy(hsg_ctx* ctx);

/* ---- model ----------------------------------------------------------------------------- */
/* replaces: Hyper_SAGNN.__init__ shape bookkeeping (Modules.py:658-710); bins[C] = bins per chromosome */
int  hsg_set_dims(hsg_ctx* ctx, int d_model, int n_cells, int n_chrom, const int32_t* bins, int cf);
/* replaces: nn.Module.load_state_dict / .to(device) for one tensor */
int  hsg_set_weight(hsg_ctx* ctx, int slot, const float* host, int64_t count);
/* replaces: MultipleEmbedding.off_hook (Modules.py:605-631) = TiedAutoEncoder.encoder (:235-253) over
 * all nodes of one branch.  branch 0 = cells (linear), branch j>0 = chromosome j-1 (swish).
 * x host [rows,in_dim], w host [d,in_dim] (wstack.i.weight_list.0), b host [d] (wstack.i.bias_list.0).
 * The [rows,d] table stays on the device; table_out (host, may be NULL) receives a copy.           */
int  hsg_embed_table(hsg_ctx* ctx, int branch, const float* x, int64_t rows, int in_dim,
                     const float* w, const float* b, float* table_out);
/* alternative to hsg_embed_table: hand over an already computed table (SparseEmbedding(embed), Modules.py:627) */
int  hsg_set_table(hsg_ctx* ctx, int branch, const float* table, int64_t rows);

/* ---- data ------------------------------------------------------------------------------ */
/* replaces: np.load(sparse_nondiag_adj_nbr_1.npy) (Impute.py:154) -- all (cell, chromosome) contact
 * matrices as one CSR: row (cell, chrom j, local bin b) = cell*n_bins + bin_off[j] + b, columns local. */
int  hsg_set_contacts(hsg_ctx* ctx, const int64_t* row_ptr, const int32_t* col, const float* val,
                      int64_t n_rows, int64_t nnz);
/* streaming variant: hsg_stage_contacts copies the NEXT matrix into a second set of device buffers on its own stream and
 * returns at once (the host arrays must stay untouched -- pinned memory makes the copy truly asynchronous -- until the
 * commit); hsg_commit_contacts waits for that copy and swaps the two sets.  Imputation calls issued in between keep using
 * the current matrix, so the upload of block i+1 overlaps the imputation of block i.                                    */
int  hsg_stage_contacts(hsg_ctx* ctx, const int64_t* row_ptr, const int32_t* col, const float* val, int64_t n_rows, int64_t nnz);
int  hsg_commit_contacts(hsg_ctx* ctx);
/* replaces: weighted_info.npy = (cell_neighbor_list, weight_dict) (Higashi_wrapper.py:1520-1545, Impute.py:245)
 * nbr [n_cells,k1] 0-based cell ids (self first), w [n_cells,k1].  nbr == NULL => no neighbours (k1 = 1).   */
int  hsg_set_neighbors(hsg_ctx* ctx, const int32_t* nbr, const float* w, int k1);
/* replaces: Higashi.get_cell_neighbor (Higashi_wrapper.py:1303-1324): kNN graph from cell embeddings
 * embed host [n_cells,dim]; fills nbr_out [n_cells,k1] (0-based) and w_out [n_cells,k1] and installs them. */
int  hsg_build_neighbors(hsg_ctx* ctx, const float* embed, int dim, int k1, int32_t* nbr_out, float* w_out);

/* ---- triplet enumeration (K3) ------------------------------------------------------------ */
/* replaces: generate_binpair + skip set + big_samples assembly (Impute.py:49-61, 186-237).
 * chroms[n_sel]: chromosome indices of impute_list in order; skip[n_skip][3] = (chrom, first_bin, end_bin)
 * local; max_bin < 0 => unlimited.  P_total and per_chrom_P[n_sel] are outputs.                           */
int  hsg_set_pairs(hsg_ctx* ctx, int n_sel, const int32_t* chroms, int min_bin, int max_bin,
                   const int32_t* skip, int n_skip, int64_t* P_total, int64_t* per_chrom_P);
/* the `coordinates` dataset (Impute.py:228-230): int64 [P_c,2] local 0-based, for selected chromosome slot */
int  hsg_get_coordinates(hsg_ctx* ctx, int sel_slot, int64_t* host_out);

/* ---- imputation (K1 + K2 + K3 epilogue) --------------------------------------------------- */
/* replaces: model.eval(); only_model=True; off_hook(); start_fix() (Impute.py:163-177) and fixes the
 * run options.  Must be called after weights/tables/contacts/pairs are set.                       */
int  hsg_prepare(hsg_ctx* ctx, int precision, int local_transfer_range, int activation);
/* replaces: the hot loop of impute_process (Impute.py:253-277): prep_one + fix_cell2 + predict for cells
 * [cell_start, cell_end).  out_dev: device float [n_cells_in_range, P_total], row-major.               */
int  hsg_impute_cells(hsg_ctx* ctx, int cell_start, int cell_end, float* out_dev, void* stream);
/* same, result delivered to HOST memory (D2H overlapped with compute); out_host [n, P_total]. */
int  hsg_impute_cells_host(hsg_ctx* ctx, int cell_start, int cell_end, float* out_host);
/* reduced-precision score blocks (the sink's option): the same scores rounded once to IEEE fp16 by the
 * scorer's epilogue -- half the bytes written to HBM and moved to the host; the sink widens them back to the fp32 `cell_%d`
 * datasets of Impute.py:274-277.  Sigmoid outputs lie in [0, 1]: the rounding adds at most 2.44e-4.  Tensor-core scorer with a
 * sigmoid head only (HSG_PREC_TC16, HSG_ACT_SIGMOID, d_model 64); any other configuration fails.  out: __half [n, P_total]. */
int  hsg_impute_cells_f16(hsg_ctx* ctx, int cell_start, int cell_end, void* out_dev_half, void* stream);
int  hsg_impute_cells_host_f16(hsg_ctx* ctx, int cell_start, int cell_end, void* out_host_half);
/* replaces: the per-cell host-to-device copy of fix_cell2 (Modules.py:1432-1469, the `.to(device)` of the cell's sparse
 * matrices) for a block of cells: re-uploads the stored entries (col / val, nnz_range of them = the extent of the rows of cells
 * [cell_start, cell_end) in the matrix installed by hsg_set_contacts; the row extents do not change) on the context's upload
 * stream and returns at once.  The copies are ordered behind the imputation work already queued (it may read these rows) and
 * the next hsg_impute_cells* call waits for them -- events, no device-wide synchronisation.  col / val must stay untouched
 * (pinned memory makes the copy asynchronous) until that call has been issued.                                             */
int  hsg_refresh_contacts(hsg_ctx* ctx, int cell_start, int cell_end, const int32_t* col, const float* val, int64_t nnz_range);
/* which scorer hsg_prepare selected: HSG_PREC_FP32, HSG_PREC_TC16 or HSG_PREC_TC16_HI.  A model whose fp16 operands could leave
 * the fp16 range (bounds from the weights and the S tables, DESIGN.md "range guard") is moved to the next safer variant.      */
int  hsg_scorer_variant(hsg_ctx* ctx, int* variant_out);

/* ---- explicit triplets (Hyper_SAGNN.predict / forward) ------------------------------------- */
/* replaces: GraphSageEncoder_with_weights.fix_cell2 (Modules.py:1432-1469) for one 1-based cell id */
int  hsg_fix_cell(hsg_ctx* ctx, int cell_1based, void* stream);
/* replaces: Hyper_SAGNN.predict (Modules.py:786-827) after fix_cell2: x dev int64 [B,3] node ids
 * (cell+1, bin_i, bin_j); out dev float [B] = act(mean head).                                       */
int  hsg_score_triplets(hsg_ctx* ctx, const int64_t* x_dev, int64_t B, float* out_dev, void* stream);

/* ---- training (hyperedge minibatches) --------------------------------------------------------
 * Parameters and gradients live in two flat fp32 device buffers owned by the caller (the host mirror makes its
 * nn.Parameters views of `params`, runs torch.optim.Adam on them and all-reduces `grads` with NCCL in one call).
 * off_slot[HSG_W_COUNT]: element offset of each weight slot inside the flat buffers (-1 = not trainable / not bound);
 * off_proj_w / off_proj_b [n_chrom+1]: offsets of MultipleEmbedding.wstack.{i}.weight_list.0 [d,in_i] / bias_list.0 [d].
 * replaces: optimizer parameter registration + .to(device) (Higashi_wrapper.py:1426-1427)                              */
int  hsg_train_bind(hsg_ctx* ctx, float* params_dev, float* grads_dev, int64_t total, const int64_t* off_slot,
                    const int64_t* off_proj_w, const int64_t* off_proj_b);
/* raw node features of one branch (SparseEmbedding.embedding, Modules.py:101-142): x host [rows,in_dim] */
int  hsg_train_set_features(hsg_ctx* ctx, int branch, const float* x_host, int64_t rows, int in_dim);
/* replaces: Hyper_SAGNN.forward (Modules.py:726-783) for one minibatch of one chromosome + the loss head of
 * Higashi.forward_batch_hyperedge (Higashi_wrapper.py:861-913).  stage 1 = MultipleEmbedding for all tokens,
 * stage 2 = GraphSageEncoder_with_weights.forward_on_hook (Modules.py:1485-1553) with the neighbour lists of
 * to_neighs_to_mask (Higashi_wrapper.py:171-221) given as CSR over the 2B bin-token rows (indptr [2B+1], cols = LOCAL bin
 * of the chromosome, vals = normalised weights).  x host int64 [B,3] node ids; y, w host float [B] (NULL = no loss).
 * loss_mode 0 = classification (weighted BCE with logits), 1 = regression (MSE against w), 2 = zinb (-log_zinb_positive,
 * Modules.py:30-80, needs hsg_train_options), 3 = rank (pairwise BCE, Higashi_wrapper.py:880-888).  Eval-mode semantics
 * (dropout off).  mean_out / loss_out: host outputs (may be NULL).  x_host = NULL: run on the batch hsg_sample_batch
 * left on the device (B and nnz as it returned them; the neighbour / label pointers are ignored).                         */
int  hsg_score_fwd(hsg_ctx* ctx, int stage, int chrom, const int64_t* x_host, int B, const int32_t* indptr_host,
                   const int32_t* cols_local_host, const float* vals_host, int nnz, const float* y_host, const float* w_host,
                   int loss_mode, int only_model, float* mean_out_host, float* loss_out_host, void* stream);
/* per-run constants of the loss heads: cell2weight [n_cells] (Higashi_wrapper.py:722,898 `cell_feats1`, zinb mode) and the
 * rank threshold (config `rank_thres`, Higashi_wrapper.py:884).  cell2weight may be NULL.                              */
int  hsg_train_options(hsg_ctx* ctx, const float* cell2weight_host, float rank_thres);
/* training-mode dropout (0.3 after fc1/fc2, 0.4 inside pff_n1: Modules.py:685-686); masks are counter-based on `seed`,
 * identical in hsg_score_fwd and hsg_score_bwd; change the seed every step.  enabled = 0: eval-mode semantics.          */
int  hsg_train_set_dropout(hsg_ctx* ctx, int enabled, uint64_t seed);
/* the three heads of the last forward batch (var / proba are produced by loss_mode 2 = zinb only) and, in zinb mode, the
 * clamped mean `pred` that forward_batch_hyperedge returns; float [B] each, HOST OR DEVICE memory (copied with
 * cudaMemcpyDefault on `stream`, which is synchronised before the call returns), any may be NULL                     */
int  hsg_score_heads(hsg_ctx* ctx, float* mean_host, float* var_host, float* proba_host, float* pred_host, void* stream);
/* ---- training batch producer on the device ------------------------------
 * replaces: generate_negative_cpu + check_nonzero (Higashi_wrapper.py:71-168), the label / weight assembly and the neighbour-
 * list builder of one_thread_generate_neg (:236-333) and to_neighs_to_mask (:171-221) for ONE minibatch of ONE chromosome.
 * pos host int64 [Bp,3] (cell id, bin_i < bin_j global ids), pos_w host float [Bp] edge weights (NULL = 1); neg_num
 * negatives per positive by rejection sampling (<= 50 trials, hard negatives with probability pair_ratio, 1 < |bj-bi| <
 * max_bin, candidate rejected while it touches a stored contact: 3x3 neighbourhood if wide_check else the entry itself);
 * neg_weight = weight of a negative (1 classification, 0 otherwise); training != 0 enables the 40 % partner removal.
 * Uses the contacts (hsg_set_contacts) and, if set, the kNN graph (hsg_set_neighbors).  The batch [positives ; accepted
 * negatives] with labels, weights and the neighbour CSR stays on the device: call hsg_score_fwd with x_host = NULL,
 * B = *B_out, nnz = *nnz_out.  Counter-based random numbers: same (seed, inputs) -> same batch.                          */
int  hsg_sample_batch(hsg_ctx* ctx, const int64_t* pos_host, const float* pos_w_host, int Bp, int chrom, int neg_num,
                      float pair_ratio, int max_bin, int wide_check, int training, float neg_weight, uint64_t seed,
                      int* B_out, int64_t* nnz_out, void* stream);
/* the same in two phases, so that producing batch i+1 overlaps the training step on batch i: _begin runs asynchronously on
 * the producer's own stream (it may be called while a step is in flight), _end finishes on `stream` without waiting for the
 * work already queued there.  hsg_sample_batch = _begin + _end.                                                         */
int  hsg_sample_batch_begin(hsg_ctx* ctx, const int64_t* pos_host, const float* pos_w_host, int Bp, int chrom, int neg_num,
                            float pair_ratio, int max_bin, int wide_check, int training, float neg_weight, uint64_t seed);
int  hsg_sample_batch_end(hsg_ctx* ctx, int* B_out, int64_t* nnz_out, void* stream);
/* reproducibility hook: the uniform draws of the 40 % partner removal (`np.random.rand(len(x), 2)` of Higashi_wrapper.py:270, one
 * per bin-token row 2r + t) are taken from draws_host [n] for all following batches instead of the counter-based generator
 * (rows past n use the generator); NULL / n = 0 restores the generator.                                                   */
int  hsg_sample_set_removal_draws(hsg_ctx* ctx, const float* draws_host, int n);
/* copies the produced batch to the host (tests / debugging); row r of the neighbour CSR = [indptr[r], indptr[r] + rowcnt[r]);
 * any pointer may be NULL                                                                                               */
int  hsg_sampled_fetch(hsg_ctx* ctx, int64_t* x_host, float* y_host, float* w_host, int32_t* indptr_host, int32_t* rowcnt_host,
                       int32_t* cols_host, float* vals_host, void* stream);
/* upstream gradients of the heads from an EXTERNAL loss (Hyper_SAGNN.forward under torch autograd, Modules.py:726-783):
 * device float [B] each, dvar / dproba may be NULL; call between hsg_score_fwd (run without y / w) and hsg_score_bwd        */
int  hsg_score_set_head_grads(hsg_ctx* ctx, const float* dmean_dev, const float* dvar_dev, const float* dproba_dev, void* stream);
/* replaces: loss.backward() (Higashi_wrapper.py:1006) for the batch of the last hsg_score_fwd: ACCUMULATES into grads */
int  hsg_score_bwd(hsg_ctx* ctx, void* stream);

/* stage 1 (train_for_embeddings): the cell branch is on-hook (hsg_train_set_features(branch 0) + bound projection) and the
 * cell auto-encoder reconstruction loss over ALL cells is added (Higashi_wrapper.py:869-873; TiedAutoEncoder.forward with
 * return_recon, Modules.py:255-288).  off_reverse_w / off_recon_b: offsets of wstack.0.reverse_weight_list.0 [d,F] and
 * wstack.0.recon_bias_list.0 [F] in the flat buffers.  hsg_recon_loss computes the loss and its gradients into a scratch
 * block and returns ||d loss / d weight_list.0||; hsg_recon_apply adds scale * scratch to the bound gradients (train_epoch
 * picks scale = ratio1 from the two gradient norms, Higashi_wrapper.py:970-1002).                                         */
int  hsg_train_bind_recon(hsg_ctx* ctx, int64_t off_reverse_w, int64_t off_recon_b);
int  hsg_recon_loss(hsg_ctx* ctx, float* loss_host, float* gnorm_w0_host, void* stream);
int  hsg_recon_apply(hsg_ctx* ctx, float scale, void* stream);

/* ---- first consumer on the device ----------------------------------------
 * replaces: the per-cell body of scTAD.gen_tad (scTAD.py:91-104) for the cells of one score block:
 *   m[x, y] += proba ; m = m + m^T ; m *= mask ; sqrt_norm (Higashi_analysis.py:581-588) ;
 *   insulation_score (Higashi_TAD.py:7-19) ; score[discard_rows] = 1 ; call_tads (Higashi_TAD.py:22-27).
 * scores_dev: DEVICE float [n_batch, P_total], the block hsg_impute_cells wrote (row = cell, leading dimension = the total
 * pair count of hsg_set_pairs); sel_slot picks the chromosome.  window_bins = int(window_ins / res), tad_order =
 * int(window_tad / res / 2).  keep_host uint8 [n] (1 = bin kept, i.e. the diagonal of scTAD's `mask`; NULL = all kept),
 * discard_host uint8 [n] (1 = row of discard_rows; NULL = none).  Outputs (host): score float [n_batch, n],
 * border uint8 [n_batch, n] (1 where call_tads reports a boundary).  The n x n maps are never materialised.  Sums run in
 * 64-bit fixed point (order-independent, exact ties): scores must be finite with |score| < 1e5, window_bins <= 1000.       */
int  hsg_insulation(hsg_ctx* ctx, int sel_slot, const float* scores_dev, int n_batch, int window_bins, int tad_order,
                    const unsigned char* keep_host, const unsigned char* discard_host, float* score_out_host,
                    unsigned char* border_out_host, void* stream);

/* ---- introspection ------------------------------------------------------------------------ */
/* HOST-ONLY test hook (no device, no context): the generation-4 tile table of a pair list in generate_binpair order
 * (utils.py:generate_binpair via Impute.py:49; score_umma4.cu build_tiles4).  `pairs_xy` = P (bin_i, bin_j) int32 pairs;
 * `tiles_out` receives 8 ints per tile: (p0, cnt, s1, s2, bin_i of segment 0, 1, 2, dense flag).  Returns the number of
 * tiles, or -1 when `tiles_cap` (in tiles) is too small.                                                                */
int  hsg_debug_tile_table(const int32_t* pairs_xy, int64_t P, int allow_dense, int32_t* tiles_out, int64_t tiles_cap);
/* number of kernels this library launched since the context was created (bench.py gpu_launches) */
int64_t hsg_launch_count(hsg_ctx* ctx);
/* device time (ms, CUDA events on the work stream) of the last hsg_impute_cells* call, split by stage:
 * ms[0] = gather (K1a), ms[1] = token projection (K1b), ms[2] = fused scorer (K2), ms[3] = total */
int  hsg_last_timing(hsg_ctx* ctx, float* ms4);
int  hsg_set_profiling(hsg_ctx* ctx, int enabled);

#ifdef __cplusplus
}
#endif
#endif /* HSG_H_ */
