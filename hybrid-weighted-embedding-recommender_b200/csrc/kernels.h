// Internal launcher declarations (C++ side of the C-ABI in include/hwer_b200.h).
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace hwer {

constexpr int kTileItems = 128;      // catalogue rows per tcgen05 tile (UMMA M)
constexpr int kKBlock = 64;          // bf16 elements per 128-byte swizzled smem row
constexpr int kMaxNQ = 256;          // queries per CTA-resident query block (UMMA N)
constexpr int kSmemBudget = 227 * 1024;
constexpr int kMaxPeers = 8;         // GPUs of one node

// Proven bound on |bf16-tensor-core score - exact score| / (|q| * |x|):
// two round-to-nearest bf16 roundings (2^-9 each) plus fp32 accumulation of
// d_pad products.  See DESIGN.md "Exactness".
// (+ 2^-22: a query block may be staged as q * (1 / thr), two more fp32 roundings of the query operand.)
inline float bf16_score_eps_rel(int d_pad) {
    return 0.00390625f + 3.8147e-6f + 2.3841858e-7f + float(d_pad) * 2.3841858e-7f;
}
inline float f32_score_eps_rel(int d) { return float(d) * 2.3841858e-7f; }

struct FilterParams {
    const float* queries;     // [B, d] fp32, device
    int B;
    int d;                    // logical embedding width
    int kb;                   // d_pad / 64
    int nq;                   // queries per block (multiple of 32, <= 256)
    int nqb;                  // number of query blocks = ceil(B / nq)
    int slots;                // CTAs cooperating on one query block (split of the item tiles)
    int qb_step;              // query blocks advanced per outer iteration (= gridDim.x / slots)
    int extra;                // CTAs beyond slots * nqb that share the tail tiles [tile_split, tile_end) (score_filter.cu)
    int tile_split;
    int no_extra_ctas;        // A/B knob
    const float* thr;         // [B] current per-query admission threshold
    const float* floor;       // [B] finite lower bound of any score of the query (stands in for thr = -inf)
    const float* qmax;        // [B] largest |component| of the query (overflow guard of the scaled filter)
    int fused_extract;        // 1: the filter CTA turns its own spill entries into candidates at the end of each query
                              // block; 0 (A/B knob): spill_extract_kernel is launched behind the filter kernel
    int force_bias;           // A/B knob: filter every query block with the bias MMA (never scale the queries)
    unsigned long long* cand; // [B, cap] packed candidate keys
    unsigned int* cnt;        // [B] candidate counters (may exceed cap => overflow)
    unsigned int cap;
    long long n_items;
    int tile_begin, tile_end; // item-tile range of this round (positions in the visiting order)
    long long tile_mul;       // visiting order: physical tile = (position * tile_mul) % tile_mod
    long long tile_mod;
    long long tile_step;      // (slots * tile_mul) % tile_mod: visiting-order stride of one CTA
    int stages;               // smem pipeline depth
    int acc_stages;           // TMEM accumulator stages
    int tmem_cols;            // power of two >= acc_stages * nq
    int stream_once;          // 1: item tiles are read by one CTA only -> L2 evict-first
    int dense;                // 1: round 0, every score is written at its position in the round (no atomics)
    uint4* spill;             // [spill_ctas * 512 threads][spill_cap][3] hit entries of the filter rounds (48 B each)
    unsigned int* spill_cnt;  // [spill_ctas * 512] entries written by each epilogue thread of the last launch
    int spill_cap;            // entries per thread
    int spill_ctas;           // CTAs the spill buffers were sized for (>= grid)
    float* dump;              // debug: full score matrix [n_items, dump_ld] (nullptr in production)
    long long dump_ld;
};

cudaError_t launch_filter_tc(const CUtensorMap& tmap, FilterParams p, int num_sms, cudaStream_t stream);
size_t filter_tc_smem_bytes(int nq, int kb, int stages);
size_t filter_tc_spill_entries(int num_sms, int spill_cap);   // 48-byte entries to allocate for FilterParams::spill
size_t filter_tc_spill_buffers(int num_sms);                  // counters to allocate for FilterParams::spill_cnt
int filter_tc_pick_stages(int nq, int kb);

// Generic CUDA-core variant for widths the tensor-core tile does not cover (d_pad > 256).
cudaError_t launch_filter_simt(const float* table, long long n_items, int d, const float* queries, int B,
                               const float* thr, unsigned long long* cand, unsigned int* cnt, unsigned int cap,
                               long long row_begin, long long row_end, int num_sms, cudaStream_t stream);

cudaError_t launch_query_margin(const float* queries, int B, int d, float factor, float max_norm, float* margin,
                                float* floor, float* qmax, unsigned int* cnt, unsigned int* ovf, float* thr,
                                cudaStream_t stream);
cudaError_t launch_fill_f32(float* p, long long n, float v, cudaStream_t stream);

// Cross-GPU threshold sharing for a row-sharded catalogue (DESIGN.md "Multi-GPU"): with G shards, the min over
// shards of each shard's ceil(K/G)-th best score so far is a lower bound of the GLOBAL K-th best, and a much
// tighter one than a shard's own K-th best -- so every shard admits ~1/G as many candidates per round.
//   kSelShared : select the k_share-th best of the local list, store {epoch, score} into every peer's slot of the
//                query, wait for the G words of the query, thr = min - margin, compact the local list (one kernel)
enum : int { kSelLocal = 0, kSelShared = 1 };
struct SelExchange {
    int mode;                        // kSelLocal: no exchange (single GPU)
    int world, rank;
    int b_cap;                       // row length of thr_x
    int k_share;                     // ceil(K / world)
    long long q0;                    // global query index of this launch's first query
    unsigned int epoch;              // this round's epoch; a slot is ready when its epoch is >= this
    unsigned long long* thr_x[kMaxPeers];   // per rank: [world][b_cap] words {epoch << 32 | fp32 score bits}
    unsigned int* flags;             // this rank's flag block ([17] = error)
};

cudaError_t launch_select_compact(unsigned long long* cand, unsigned int* cnt, unsigned int cap, int B, int K,
                                  int fixed_count, const float* margin, float* thr, unsigned int* needed_cap,
                                  unsigned int* ovf, const SelExchange* sx, int dense_warp, cudaStream_t stream);

// Where final_kernel stores a query's result when the catalogue is sharded over several GPUs: straight into the
// exchange buffer of the GPU that owns (merges) that query, over NVLink peer stores (exchange.cu).
struct PeerDst {
    int world;                       // 0 = disabled (results go to the local out_* arrays)
    int rank;
    int q_per_owner;                 // Bq = ceil(B / world): owner(q) = q / Bq
    int q_cap;                       // rows per source rank in every exchange buffer (>= Bq)
    int k_cap;                       // columns per row (>= K)
    long long q0;                    // global query index of this launch's first query (batch chunking)
    double* xs[kMaxPeers];           // [world][q_cap][k_cap] fp64 scores, on each owner
    long long* xi[kMaxPeers];        // [world][q_cap][k_cap] global rows
};

cudaError_t launch_final(const unsigned long long* cand, const unsigned int* cnt, unsigned int cap, int B, int K,
                         int exact, const float* table, int d, const float* queries, long long idx_offset,
                         long long* out_idx, float* out_score, double* out_score64, unsigned int* needed_cap,
                         unsigned int* ovf, const PeerDst* peer, int allow_small, cudaStream_t stream);

// Exhaustive exact search of ONE query at a time (the answer of last resort for a query whose candidate list cannot
// hold everything inside the bf16 margin, e.g. tens of thousands of duplicate rows): fp64 scores of every row with
// final_kernel's arithmetic, a stable descending radix sort, the first k.  scratch: see bruteforce_scratch_bytes.
size_t bruteforce_scratch_bytes(long long n);
cudaError_t launch_bruteforce_topk(const float* table, long long n, int d, const float* query, int k,
                                   long long idx_offset, long long* out_idx, float* out_score, double* out_score64,
                                   void* scratch, size_t scratch_bytes, cudaStream_t stream);

// Peer exchange (exchange.cu): publish "my scatter is complete" on every peer; owner-side wait + merge + delivery
// of the merged rows to every rank; final wait + copy-out.
struct ExchangeView {
    int world, rank, q_per_owner, q_cap, k_cap, b_cap;
    unsigned int* flags[kMaxPeers];      // per rank: [2][kMaxPeers] epochs, [16] counter, [17] error, [40..47] needed cap per source
    unsigned long long* thr_x[kMaxPeers];   // per rank: [world][b_cap] shared-threshold words {epoch, score}
    long long sched_rows;                // rows of the largest shard: every rank walks the same round schedule
    double* xs[kMaxPeers];
    long long* xi[kMaxPeers];
    long long* out_idx[kMaxPeers];       // per rank: [b_cap, k_cap] merged rows (row stride k of the call)
    float* out_score[kMaxPeers];
    double* out_score64[kMaxPeers];
};
cudaError_t exchange_preload();      // loads the exchange kernels now (see exchange.cu)
// `needed` (phase 0, may be NULL): this rank's candidate-capacity demand, delivered to every peer with the flag
cudaError_t launch_exchange_signal(const ExchangeView& v, int phase, unsigned int epoch, const unsigned int* needed,
                                   cudaStream_t stream);
cudaError_t launch_exchange_merge(const ExchangeView& v, int B, int K, unsigned int epoch, int owned,
                                  cudaStream_t stream);
cudaError_t launch_exchange_collect(const ExchangeView& v, int B, int K, unsigned int epoch, long long* out_idx,
                                    float* out_score, double* out_score64, unsigned int* status2, int owned,
                                    cudaStream_t stream);

cudaError_t launch_merge(const double* scores, const long long* idx, int G, int B, int K, long long* out_idx,
                         float* out_score, double* out_score64, cudaStream_t stream);

cudaError_t launch_blend_normalize(const float* content, const float* collab, float alpha, const float* alpha_rows,
                                   long long n, int d, float* out_f32, void* out_bf16, int d_pad,
                                   cudaStream_t stream);
cudaError_t launch_make_shadow(const float* table, long long n, int d, void* out_bf16, int d_pad, cudaStream_t stream);
cudaError_t launch_norm_stats(const float* v, long long n, int d, float eps, double* out5, cudaStream_t stream);

cudaError_t launch_pair_score(const float* table, long long n, int d, const long long* src, const long long* dst,
                              long long P, float* out, cudaStream_t stream);

cudaError_t launch_compose_queries(const float* table, long long n, int d, const long long* anchor,
                                   const long long* pos_ptr, const long long* pos_rows, const long long* neg_ptr,
                                   const long long* neg_rows, int B, float* out, cudaStream_t stream);

// unit(mean(rows of list l)) per CSR list (get_average_embeddings, hwer/recommendation_base.py:153-155)
cudaError_t launch_average_embeddings(const float* table, long long n, int d, const long long* ptr,
                                      const long long* rows, int L, float* out, cudaStream_t stream);

// rerank.cu: score conventions + stable per-anchor ordering of the k retrieved rows, and the row gathers around them
enum : int { kScorePair = 0, kScoreDist = 1, kScoreGiven = 2, kScoreEuclid = 3 };
cudaError_t launch_rerank(const float* table, long long n, int d, const long long* rows, const long long* row_map,
                          int B, int k, int conv, const long long* anchor_rows, const float* queries,
                          const float* given, long long* out_rows, double* out_score, cudaStream_t stream);
cudaError_t launch_map_rows(const long long* rows, long long count, const long long* row_map, long long offset,
                            long long* out, cudaStream_t stream);
cudaError_t launch_gather_rows(const float* table, long long n, int d, const long long* rows, long long P, float* out,
                               cudaStream_t stream);
cudaError_t launch_hit_rank(const float* scores, int U, int M1, int topn, int* rank, double* out2, cudaStream_t stream);

cudaError_t launch_eval(const long long* topk, int U, int Kret, const long long* train_ptr,
                        const long long* train_idx, const long long* val_ptr, const long long* val_idx,
                        const float* val_rel, const int* cutoffs, int n_cut, long long n_items,
                        double* per_user, unsigned int* seen_bitmap, cudaStream_t stream);
cudaError_t launch_eval_reduce(const double* per_user, int U, int M, const long long* val_ptr,
                               const unsigned int* seen_bitmap, long long n_items, int n_cut, double* out,
                               cudaStream_t stream);

// Link-prediction metrics (link_metrics.cu): out8 = ap, precision, recall, accuracy, tp, fp, fn, tn.
cudaError_t launch_link_metrics(const float* score, const unsigned char* label, long long P, float thr, double* out8,
                                cudaStream_t stream);

// NCF re-rank (ncf.cu): params = [W1 (out x in, row-major), b1, ..., W_depth, b_depth, w_out (F), b_out (1)].
int ncf_layer_in(int F, int depth, int layer);
int ncf_layer_out(int F, int depth, int layer);
long long ncf_param_count(int F, int depth);
cudaError_t launch_ncf_score(const float* h, long long n_rows, int F, int depth, const float* params,
                             const long long* src, const long long* dst, long long P, float* out, float* ws0,
                             float* ws1, long long chunk, cudaStream_t stream);

// tcgen05 path of the NCF re-rank (ncf_tc.cu): split-bf16 operands, F % 64 == 0, F <= 256.
bool ncf_tc_supported(int F, int depth);
long long ncf_tc_chunk(int num_sms);
size_t ncf_tc_scratch_elems(int F, int depth, long long chunk);      // bf16 elements
cudaError_t launch_ncf_score_tc(const float* h, long long n_rows, int F, int depth, const float* params,
                                const long long* src, const long long* dst, long long P, float* out, void* scratch,
                                long long chunk, int num_sms, cudaStream_t stream);

// fp32 GEMM of ncf.cu as a plain linear layer (gcn_infer.cu)
cudaError_t launch_linear_f32(const float* x, const float* w, const float* b, float* y, long long P, int in, int out,
                              float slope, cudaStream_t stream);

// GCN inference over the whole graph with explicit per-block neighbour lists (gcn_infer.cu)
size_t gcn_infer_scratch_floats(long long n, int F, int layers, long long chunk);
cudaError_t launch_gcn_infer(const float* node_emb, const float* content, int C, const float* proj_w, const float* proj_b,
                             const float* ln_g, const float* ln_b, long long n, int F, int layers,
                             const long long* const* nbr_ptr, const long long* const* nbr_idx, const float* fc0_w,
                             const float* fc0_b, const float* fc1_w, const float* fc1_b, float* previous, float ema,
                             float* out, float* scratch, long long chunk, cudaStream_t stream);

}  // namespace hwer
