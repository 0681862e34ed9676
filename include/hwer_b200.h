/* hwer_b200.h -- C ABI of the B200-native serving hot path of
 * faizanahemad/Hybrid-Weighted-Embedding-Recommender ("hwer").
 *
 * The reference is pure Python and has no FFI; the seams this library sits
 * behind are the methods of hwer/recommendation_base.py and the helpers of
 * hwer/utils.py / hwer/validation.py cited on each entry point below (paths are
 * relative to the reference checkout).  INTEGRATION.md shows the ctypes stub a
 * reference maintainer would add.
 *
 * Conventions
 *  - every pointer named *_dev is a CUDA device pointer on the index's device;
 *    every other pointer is host memory;
 *  - tables are row-major, C-contiguous; `stream` is a cudaStream_t passed as
 *    void* (NULL = the legacy default stream); calls only ENQUEUE work unless
 *    stated otherwise;
 *  - return value: 0 on success, a negative hwer_status otherwise;
 *    hwer_last_error() returns a thread-local human-readable message;
 *  - there is no CPU fallback: without a CUDA device of compute capability 10.x
 *    every compute entry point fails with HWER_E_CUDA / HWER_E_ARCH.
 */
#ifndef HWER_B200_H
#define HWER_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum hwer_status {
    HWER_OK = 0,
    HWER_E_INVALID = -1,     /* bad argument (NULL, negative size, d_pad not a multiple of 64, ...)          */
    HWER_E_CUDA = -2,        /* a CUDA runtime / driver call failed                                         */
    HWER_E_ARCH = -3,        /* device is not sm_100 (B200)                                                 */
    HWER_E_K_TOO_LARGE = -4, /* k > rows of the index: sklearn's KDTree.query raises ValueError here        */
    HWER_E_OVERFLOW = -5,    /* candidate lists overflowed: re-run hwer_topk with cap >= *needed_cap        */
    HWER_E_NOMEM = -6,
    HWER_E_PEER = -7         /* multi-GPU exchange: a peer GPU did not report in time                        */
} hwer_status;

#define HWER_MODE_EXACT 0 /* bf16 tensor-core filter with a proven margin + fp64 re-score from the fp32 table */
#define HWER_MODE_BF16 1  /* bf16 tensor-core scores returned as they are                                    */

typedef struct hwer_index hwer_index_t;
typedef struct hwer_exchange hwer_exchange_t;
#define HWER_IPC_HANDLE_BYTES 64

const char* hwer_last_error(void);
int hwer_version(void);

/* Row width of the bf16 shadow table for logical width d (next multiple of 64). */
int32_t hwer_shadow_width(int32_t d);

/* V = unit(alpha * unit(C) + (1 - alpha) * unit(G)) row-wise; unit(a) = a / ||a||_2 without epsilon.
 * Replaces: GcnNCF.prepare_for_knn, hwer/gcn_ncf.py:447-456 (the blend slot; the reference computes the
 *           alpha = 0 case) and utils.unit_length, hwer/utils.py:43-44.
 * content_dev may be NULL (plain unit_length of collab_dev); alpha_rows_dev (per-row alpha, [n]) may be NULL;
 * out_bf16_dev (NULL to skip) receives the [n, d_pad] bf16 shadow with zero padding. */
int hwer_blend_normalize(const float* content_dev, const float* collab_dev, float alpha,
                         const float* alpha_rows_dev, int64_t n, int32_t d, float* out_f32_dev,
                         void* out_bf16_dev, int32_t d_pad, void* stream);

/* bf16 shadow (round-to-nearest-even, zero-padded to d_pad columns) of an already normalised fp32 table: the
 * operand the tensor-core scorer streams when __build_knn__ (hwer/recommendation_base.py:105-110) is handed a
 * finished table.  No reference counterpart (the reference's KDTree keeps a float64 copy instead, :74). */
int hwer_make_shadow(const float* table_dev, int64_t n, int32_t d, void* out_bf16_dev, int32_t d_pad, void* stream);

/* Row-norm statistics of an [n, d] fp32 table.
 * Replaces: utils.unit_length_violations(a, axis=1, epsilon), hwer/utils.py:51-57, asserted by
 *           RecommendationBase.__build_knn__, hwer/recommendation_base.py:105-107.
 * out5_dev = {violations, mean |norm - 1|, positive violations, negative violations, max norm} (doubles). */
int hwer_norm_stats(const float* table_dev, int64_t n, int32_t d, float epsilon, double* out5_dev, void* stream);

/* Index over one node type's rows (the per-type KDTree of MultiKNN.__init__, hwer/recommendation_base.py:65-76).
 * The fp32 table and its bf16 shadow are BORROWED: they must outlive the index.  max_norm is the largest row
 * norm (from hwer_norm_stats); it scales the exact-mode admission margin. */
int hwer_index_create(hwer_index_t** out, const float* table_f32_dev, const void* shadow_bf16_dev, int64_t n,
                      int32_t d, int32_t d_pad, float max_norm, int32_t device);
int hwer_index_destroy(hwer_index_t* index);

/* Exact top-k by dot product of B queries against every row of the index.
 * Replaces: MultiKNN.query, hwer/recommendation_base.py:78-83 (sklearn KDTree.query, exact Euclidean k-NN:
 *           on unit rows the order equals descending dot product) for a whole batch of anchors, i.e. the loop
 *           of validation.model_get_topk_knn, hwer/validation.py:30-35.
 * queries_dev [B, d] fp32 (any norm).  Results are ordered (score descending, row ascending); rows are local to
 * the index plus idx_offset; missing entries (fewer than k finite scores) are row -1 / score -inf.
 * cap = per-query candidate-list capacity (0 = automatic).  out_score64_dev may be NULL.
 * Asynchronous; call hwer_topk_finish before trusting the outputs.
 * The index owns its scratch memory and grows it on demand (the first call of a given batch size / k synchronises
 * the device): candidate lists [min(B, chunk), cap] u64 (<= 1 GiB) and the filter rounds' spill buffers
 * (58 - 465 MB: one 48-byte-entry append buffer per epilogue thread of the grid). */
int hwer_topk(hwer_index_t* index, const float* queries_dev, int32_t B, int32_t k, int32_t mode, uint32_t cap,
              int64_t idx_offset, int64_t* out_idx_dev, float* out_score_dev, double* out_score64_dev,
              void* stream);
/* Synchronises `stream` and reports candidate-list overflow of the topk calls enqueued since the last finish.
 * On HWER_E_OVERFLOW the outputs of the queries that did NOT overflow are valid; a query that did carries row -2
 * in column 0 of out_idx_dev.  Re-run those (or everything) with cap >= *needed_cap; a list is at most 16384
 * entries long, so when *needed_cap exceeds that (tens of thousands of rows inside the bf16 margin of the k-th
 * score: duplicated or default cold-start embeddings) answer the marked queries with hwer_topk_exhaustive. */
int hwer_topk_finish(hwer_index_t* index, void* stream, uint32_t* needed_cap);

/* The same exact answer (same fp64 scores, same (score desc, row asc) order) by exhaustive search: every row of
 * the index is scored in fp64 and a stable radix sort keeps the first k -- about a millisecond per million rows and
 * query, 24 bytes of scratch per row.  The path of last resort for queries hwer_topk cannot hold in its candidate
 * lists; like sklearn's KDTree (hwer/recommendation_base.py:79) it always answers. */
int hwer_topk_exhaustive(hwer_index_t* index, const float* queries_dev, int32_t B, int32_t k, int64_t idx_offset,
                         int64_t* out_idx_dev, float* out_score_dev, double* out_score64_dev, void* stream);

/* Measurement aid: when enabled, every score-filter launch of hwer_topk (with the spill-extract kernel that rides
 * behind it) is bracketed by CUDA events on the
 * launching stream.  hwer_profile_read synchronises `stream`, returns the summed filter-kernel time and the
 * number of kernels launched (filter / everything else) since the last read, and resets the counters. */
int hwer_profile(hwer_index_t* index, int enable);
int hwer_profile_read(hwer_index_t* index, void* stream, double* filter_ms, int64_t* filter_launches,
                      int64_t* other_launches);
/* Where a step's time goes: summed durations (ms) since the last hwer_profile_read of the brackets around
 * [0] the score-filter launches (+ spill extract), [1] the select launches (incl. waits on peer thresholds),
 * [2] the final re-score + sort, [3] the peer exchange (signal, owner merge, collect).  Call BEFORE
 * hwer_profile_read, which resets the events. */
int hwer_profile_stages(hwer_index_t* index, void* stream, double* out4_ms);
/* Per-launch view of the same events (call BEFORE hwer_profile_read, which resets them): the durations, in launch
 * order, of the score-filter launches since the last read.  *n_out = number of launches recorded; at most `cap`
 * are written to out_ms (milliseconds). */
int hwer_profile_launches(hwer_index_t* index, void* stream, double* out_ms, int32_t cap, int32_t* n_out);

/* Debug/validation aid: the full bf16 tensor-core score matrix out[n, ld] (ld >= B) for small problems. */
int hwer_debug_scores(hwer_index_t* index, const float* queries_dev, int32_t B, float* out_dev, int64_t ld,
                      void* stream);

/* G shards x [B, k] -> [B, k] under the same ordering rule (after the NCCL all-gather of per-GPU results).
 * No reference counterpart (the reference is single-process); SURVEY.md section 8(e). */
int hwer_merge_topk(const double* scores_dev, const int64_t* idx_dev, int32_t G, int32_t B, int32_t k,
                    int64_t* out_idx_dev, float* out_score_dev, double* out_score64_dev, void* stream);

/* ---- item-sharded search over the GPUs of one node, exchange over NVLink peer memory (no reference counterpart:
 * the reference is single-process).  One process per GPU; GPU g indexes catalogue rows
 * [g*N/G, (g+1)*N/G).  Every rank allocates one exchange buffer (hwer_exchange_bytes, hwer_peer_alloc), passes its
 * 64-byte IPC handle to the others through any host channel (torch.distributed in the Python mirror), opens theirs
 * (hwer_peer_open) and builds an exchange object over the `world` base pointers, its own included, in rank order.
 *
 * hwer_topk_sharded = hwer_topk on the local shard whose last kernel stores each query's result straight into
 * the buffer of the GPU that owns the query (queries are split into `world` contiguous owner ranges), an owner-side
 * merge that delivers the final rows to every rank, and a copy-out: out_* hold the same [B, k] result on every
 * rank, bit-identical to a single-GPU hwer_topk over the whole table.  Collective: every rank must call it with
 * the same B, k and mode, in the same order.  Each local shard needs at least k rows. */
int64_t hwer_exchange_bytes(int32_t world, int32_t b_cap, int32_t k_cap);
int hwer_peer_alloc(int64_t bytes, void** dev_ptr_out, unsigned char* handle_out /* [HWER_IPC_HANDLE_BYTES] */);
int hwer_peer_open(const unsigned char* handle, void** dev_ptr_out);
int hwer_peer_close(void* dev_ptr);
int hwer_peer_free(void* dev_ptr);
int hwer_exchange_create(hwer_exchange_t** out, int32_t world, int32_t rank, int32_t b_cap, int32_t k_cap,
                         void* const* bases /* [world] device pointers */, int32_t device);
/* largest_shard_rows: rows of the largest shard over all ranks (every rank must pass the same value: the ranks walk
 * one common round schedule because, with share_thresholds != 0 (default), every round exchanges per-query
 * thresholds -- each shard publishes its ceil(k/G)-th best score so far, the min over shards bounds the global
 * k-th best from below, so a shard admits ~1/G as many candidates per round.  0 = the shards are known to be of
 * equal size (same number of 128-row tiles). */
int hwer_exchange_configure(hwer_exchange_t* exchange, int64_t largest_shard_rows, int32_t share_thresholds);
int hwer_exchange_destroy(hwer_exchange_t* exchange);
#define HWER_PHASE_SEARCH 1  /* local search; its last kernel stores into the owners' buffers; publish "scattered" */
#define HWER_PHASE_MERGE 2   /* wait for every source, merge the owned queries, deliver to every rank, publish      */
#define HWER_PHASE_COLLECT 4 /* wait for every owner, copy the [B, k] result into out_*                             */
#define HWER_PHASE_ALL 7     /* the normal call; a host may also enqueue the phases one by one, in this order       */
#define HWER_PHASE_OWNED 8   /* modifier (pass it with MERGE and COLLECT): every rank keeps only the queries it merged:
                              * out_* receive rows [rank * ceil(B / world), ...) of the result, packed from row 0 (the
                              * "reduce-scatter" form of the exchange: a serving tier that answers from all ranks needs
                              * each result once, so winners are not delivered to the other ranks and a rank copies 1/world
                              * of the result to its host)                                                           */
int hwer_topk_sharded(hwer_index_t* index, hwer_exchange_t* exchange, const float* queries_dev, int32_t B, int32_t k,
                      int32_t mode, uint32_t cap, int64_t idx_offset, int64_t* out_idx_dev, float* out_score_dev,
                      double* out_score64_dev, int32_t phases, void* stream);
/* Synchronises `stream`; HWER_E_PEER if any wait on a peer timed out since the exchange was created. */
int hwer_exchange_error(hwer_exchange_t* exchange, void* stream);
/* hwer_topk_finish after hwer_topk_sharded reports the outcome of the COLLECTIVE: the capacity demand is the largest
 * over all ranks (it travels with the exchange flags, so every rank takes the same retry decision without a
 * host-side collective) and a peer timeout on this rank surfaces as HWER_E_PEER. */

/* out[p] = (dot(row src[p], row dst[p]) + 1) / 2; a row id outside [0, n) means "node not seen in training"
 * and scores with clip(row 0, 1e-6, 1e-5).
 * Replaces: RecommendationBase.predict / get_embeddings, hwer/recommendation_base.py:135-151
 *           (and GcnNCF.predict's cosine branch, hwer/gcn_ncf.py:330-334). */
int hwer_pair_score(const float* table_dev, int64_t n, int32_t d, const int64_t* src_dev, const int64_t* dst_dev,
                    int64_t P, float* out_dev, void* stream);

/* Query vectors of a batch of find_closest_neighbours calls:
 *   out[q] = average of { unit(row anchor[q]), unit(mean(rows pos[q])), -unit(mean(rows neg[q])) } over the parts
 *   present, not re-normalised.
 * Replaces: RecommendationBase.find_closest_neighbours' embedding composition, hwer/recommendation_base.py:164-170
 *           (get_average_embeddings :153-155, get_embeddings :146-151), identical in hwer/gcn_ncf.py:369-376.
 * pos/neg are CSR lists over the queries (ptr [B+1], rows); either ptr may be NULL.  Row ids outside [0, n) are
 * nodes never trained on: clip(row 0, 1e-6, 1e-5).  d <= 1024. */
int hwer_compose_queries(const float* table_dev, int64_t n, int32_t d, const int64_t* anchor_rows_dev,
                         const int64_t* pos_ptr_dev, const int64_t* pos_rows_dev, const int64_t* neg_ptr_dev,
                         const int64_t* neg_rows_dev, int32_t B, float* out_dev, void* stream);

/* unit(mean(rows of list l)) for L CSR lists of row ids (ptr [L+1], rows); ids outside [0, n) are nodes never
 * trained on: clip(row 0, 1e-6, 1e-5).  An empty list yields NaN like numpy.  d <= 1024.
 * Replaces: RecommendationBase.get_average_embeddings, hwer/recommendation_base.py:153-155. */
int hwer_average_embeddings(const float* table_dev, int64_t n, int32_t d, const int64_t* ptr_dev,
                            const int64_t* rows_dev, int32_t L, float* out_dev, void* stream);

/* out[p] = table[rows[p]], or clip(table[0], 1e-6, 1e-5) for a row id outside [0, n).
 * Replaces: RecommendationBase.get_embeddings, hwer/recommendation_base.py:146-151. */
int hwer_gather_rows(const float* table_dev, int64_t n, int32_t d, const int64_t* rows_dev, int64_t P, float* out_dev,
                     void* stream);

/* Rows local to a per-type index -> global rows: out = row_map ? row_map[row] : row + offset; negative ids
 * (missing results) pass through.  Replaces: the bidict lookup of MultiKNN.query, hwer/recommendation_base.py:80. */
int hwer_map_rows(const int64_t* rows_dev, int64_t count, const int64_t* row_map_dev, int64_t offset,
                  int64_t* out_dev, void* stream);

/* The step every find_closest_neighbours variant ends with: score the k retrieved rows of each of B anchors in
 * the caller's convention and order each anchor's list by that score with a STABLE sort (equal scores keep the
 * order of `rows_dev`, as Python's sorted() does), all in one launch.
 *   HWER_SCORE_PAIR    (table[anchor] . table[row] + 1) / 2 in fp32, descending  -- RecommendationBase.
 *                      find_closest_neighbours' rescoring through predict, hwer/recommendation_base.py:172-174
 *   HWER_SCORE_DIST    (2 - ||table[row] - query||) / 2 in fp64, descending      -- GcnNCF.find_closest_neighbours,
 *                      cosine branch, hwer/gcn_ncf.py:378-383 (query = the composed embedding of :369-376)
 *   HWER_SCORE_GIVEN   given_dev[B, k] (e.g. hwer_ncf_score output), descending  -- NCF branch, hwer/gcn_ncf.py:384-386
 *   HWER_SCORE_EUCLID  ||table[row] - query|| in fp64, ASCENDING                 -- MultiKNN.query,
 *                      hwer/recommendation_base.py:79-82
 * rows_dev [B, k]: rows as returned by hwer_topk (-1 = none; these sort last and come out as row -1, score -inf,
 * +inf for EUCLID).  row_map_dev (may be NULL): rows are local to a gathered per-type index and row_map[row] is
 * the row of table_dev; out_rows_dev then holds the mapped rows.  anchor_rows_dev [B] (PAIR), queries_dev [B, d]
 * (DIST, EUCLID), given_dev [B, k] (GIVEN): the others may be NULL.  out_score_dev [B, k] doubles.  k <= 8192. */
#define HWER_SCORE_PAIR 0
#define HWER_SCORE_DIST 1
#define HWER_SCORE_GIVEN 2
#define HWER_SCORE_EUCLID 3
int hwer_rerank(const float* table_dev, int64_t n, int32_t d, const int64_t* rows_dev, const int64_t* row_map_dev,
                int32_t B, int32_t k, int32_t convention, const int64_t* anchor_rows_dev, const float* queries_dev,
                const float* given_dev, int64_t* out_rows_dev, double* out_score_dev, void* stream);

/* HR@topn and binary NDCG@topn of one positive among M sampled negatives per user.
 * Replaces: the per-user sort / top-10 / binary_ndcg_v2 loop of validation.ncf_eval, hwer/validation.py:82-96.
 * scores_dev [U, 1 + M] fp32, column 0 = the positive (equal-scored negatives rank behind it, as the reference's
 * stable sort leaves them).  out2_dev = {mean(rank < topn), mean(rank < topn ? 1 / log2(rank + 2) / (1 + 1e-8) : 0)};
 * rank_dev (NULL to skip) receives the [U] ranks (number of negatives scored strictly higher). */
int hwer_hit_rank_metrics(const float* scores_dev, int32_t U, int32_t M, int32_t topn, double* out2_dev,
                          int32_t* rank_dev, void* stream);

/* NCF re-rank: out[p] = sigmoid(w_out . MLP([h[src[p]] || h[dst[p]]]) + b_out), fp32.
 * Replaces: NCF.forward, hwer/ncf.py:7-27, as driven by GcnNCF.predict, hwer/gcn_ncf.py:336-361, and by the NCF
 *           branch of GcnNCF.find_closest_neighbours, hwer/gcn_ncf.py:384-386.
 * h_dev [n_rows, F] is the reference's prediction_artifacts["h"] (row 0 = padding node; callers pass node row + 1,
 * rows outside [0, n_rows) read row 0 like the reference's "unknown node -> 0").  The MLP has `depth` Linear +
 * LeakyReLU(0.01) layers of widths F*iw -> F*ow with iw = 4 if layer == 2 else 2, ow = 1 if layer == depth else
 * (4 if layer == 1 else 2) (ncf.py:12-16).  params_dev = [W1 (out x in row-major, torch layout), b1, ...,
 * W_depth, b_depth, w_out (F), b_out (1)], hwer_ncf_param_count(F, depth) floats.  F must be a multiple of 4. */
int64_t hwer_ncf_param_count(int32_t F, int32_t depth);
int hwer_ncf_score(const float* h_dev, int64_t n_rows, int32_t F, int32_t depth, const float* params_dev,
                   const int64_t* src_dev, const int64_t* dst_dev, int64_t P, float* out_dev, void* stream);

/* GCN inference: the collaborative table the serving path consumes, computed from a trained graph-convolution model.
 * Replaces: get_gcn_vectors, hwer/gcn_ncf.py:260-279 = GraphConvModule.forward in eval mode, hwer/gcn.py:162-193
 * (content projection gcn.py:40-44,59-63; GraphConv layers gcn.py:104-128; EMA with `previous` gcn.py:186-191), over
 * the whole graph.  The neighbour sample of every block -- what DGL's NeighborSampler draws (two random in-neighbours
 * and a self loop per node, gcn_ncf.py:262-272) -- is an input: nbr_ptr_dev[i] / nbr_idx_dev[i] are HOST arrays of
 * `layers` device pointers, CSR over all n nodes (ptr [n + 1], idx node ids).
 *   node_emb [(n + 1), F] (row v + 1 = node v, row 0 = padding node), content [n, C], proj_w [F, C], proj_b / ln_g /
 *   ln_b [F], fc0_w [4F, F (layers + 1)], fc0_b [4F], fc1_w [F, 4F], fc1_b [F] (torch nn.Linear layouts),
 *   previous [>= n, F] updated in place (NULL: no EMA), out [n, F].  C and F multiples of 4 (pad with zero columns). */
int hwer_gcn_infer(const float* node_emb_dev, const float* content_dev, int64_t n, int32_t C, int32_t F, int32_t layers,
                   const float* proj_w_dev, const float* proj_b_dev, const float* ln_g_dev, const float* ln_b_dev,
                   const int64_t* const* nbr_ptr_dev, const int64_t* const* nbr_idx_dev, const float* fc0_w_dev,
                   const float* fc0_b_dev, const float* fc1_w_dev, const float* fc1_b_dev, float* previous_dev, float ema,
                   float* out_dev, void* stream);

/* Ranking metrics for U users in one pass.
 * Replaces: the per-user loops of validation.extraction_efficiency, hwer/validation.py:133-174, with
 *           utils.reciprocal_rank / ndcg / binary_ndcg / recall, hwer/utils.py:71-121.
 * topk_dev [U, kret] item ids in rank order (-1 = none); train CSR rows sorted ascending by item id; validation
 * CSR rows sorted by relevance descending (train items may be present, they are skipped).
 * cutoffs_dev ascending, last <= 256.  out_dev has 3*n_cut + 3 doubles:
 *   [3c] recall@c, [3c+1] graded ndcg@c, [3c+2] binary ndcg@c (means over users with a validation row),
 *   [3*n_cut] MRR, [3*n_cut+1] distinct items among all users' filtered top-max_cut, [3*n_cut+2] #validation users.
 * per_user_dev (NULL to skip) receives the [U, 3*n_cut+1] per-user values. */
int hwer_eval_metrics(const int64_t* topk_dev, int32_t U, int32_t kret, const int64_t* train_ptr_dev,
                      const int64_t* train_idx_dev, const int64_t* val_ptr_dev, const int64_t* val_idx_dev,
                      const float* val_rel_dev, const int32_t* cutoffs_dev, int32_t n_cut, int64_t n_items,
                      double* out_dev, double* per_user_dev, void* stream);

/* Link-prediction metrics of P scored pairs with 0/1 labels.
 * Replaces: the sklearn calls of validation.link_prediction_accuracy, hwer/validation.py:52-59 --
 *           average_precision_score(labels, scores), precision_recall_fscore_support(labels, scores >= 0.5,
 *           average='binary') and accuracy_score(labels, scores >= 0.5).
 * scores_dev [P] fp32 (hwer_pair_score / hwer_ncf_score output), labels_dev [P] u8.  out8_dev receives 8 doubles:
 *   average precision (step-wise over distinct scores, as sklearn), precision, recall, accuracy at
 *   `threshold` (score >= threshold is a predicted link; undefined ratios are 0 like sklearn's zero_division),
 *   then the confusion counts tp, fp, fn, tn.  0 < P < 2^31 - 1. */
int hwer_link_metrics(const float* scores_dev, const uint8_t* labels_dev, int64_t P, float threshold, double* out8_dev,
                      void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HWER_B200_H */
