// Pair scoring and ranking-metric evaluation kernels.
//
//   pair_score : hwer/recommendation_base.py:135-151  predict / get_embeddings
//                (row-wise dot of two gathered row sets, (s + 1) / 2; a node
//                that was never trained on gets clip(row 0, 1e-6, 1e-5)).
//   eval       : hwer/validation.py:133-174 extraction_efficiency's per-user
//                loop with hwer/utils.py:71-121 reciprocal_rank / ndcg /
//                binary_ndcg / recall, evaluated for every user in one launch,
//                plus the diversity bitmap of validation.py:144-145.
#include "common.cuh"
#include "kernels.h"

namespace hwer {

namespace {

__device__ __forceinline__ float unknown_clip(float v) { return fminf(fmaxf(v, 1e-6f), 1e-5f); }

__global__ void __launch_bounds__(256)
pair_score_kernel(const float* __restrict__ table, long long n, int d, const long long* __restrict__ src,
                  const long long* __restrict__ dst, long long P, float* __restrict__ out) {
    const int lane = lane_id();
    const long long warps_total = (long long)gridDim.x * 8;
    for (long long p = (long long)blockIdx.x * 8 + (threadIdx.x >> 5); p < P; p += warps_total) {
        const long long a = src[p], b = dst[p];
        const bool ua = a < 0 || a >= n, ub = b < 0 || b >= n;
        const float* xa = table + (size_t)(ua ? 0 : a) * d;
        const float* xb = table + (size_t)(ub ? 0 : b) * d;
        float s = 0.f;
        for (int j = lane; j < d; j += 32) {
            float va = xa[j], vb = xb[j];
            if (ua) va = unknown_clip(va);
            if (ub) vb = unknown_clip(vb);
            s = fmaf(va, vb, s);
        }
        s = warp_sum(s);
        if (lane == 0) out[p] = (s + 1.0f) / 2.0f;
    }
}

// ---------------------------------------------------------------------------
// Query composition: hwer/recommendation_base.py:164-170 (same in gcn_ncf.py:369-376)
//   embedding = average([unit(mean(anchor)), unit(mean(positive)), -unit(mean(negative))])
// over the parts that are present; NOT re-normalised.  One warp per query; a row id outside [0, n) is a node that
// was never trained on and contributes clip(row 0, 1e-6, 1e-5) (get_embeddings, :146-151).  d <= 1024.
// ---------------------------------------------------------------------------
constexpr int kComposeMaxPerLane = 32;

// PL = values per lane (d <= 32 * PL): 4 for the serving widths up to 128, 8 up to 256, 32 in general -- the general
// shape keeps 64 accumulators per lane and 32 predicated loads per row, which left compose_queries at 9 % of the
// copy bandwidth on 128-wide rows (bench.py aux_kernels, r2).
template <int PL>
__device__ __forceinline__ void mean_unit_accumulate(const float* __restrict__ table, long long n, int d,
                                                     const long long* __restrict__ rows, long long b, long long e,
                                                     float sign, float (&acc)[PL]) {
    const int lane = lane_id();
    float m[PL];
#pragma unroll
    for (int j = 0; j < PL; ++j) m[j] = 0.f;
    for (long long i = b; i < e; ++i) {
        const long long r = rows[i];
        const bool unk = r < 0 || r >= n;
        const float* x = table + (size_t)(unk ? 0 : r) * d;
#pragma unroll
        for (int j = 0; j < PL; ++j) {
            const int c = lane + 32 * j;
            if (c < d) {
                float v = x[c];
                if (unk) v = unknown_clip(v);
                m[j] += v;
            }
        }
    }
    const float inv_cnt = 1.0f / (float)(e - b);
    float ss = 0.f;
#pragma unroll
    for (int j = 0; j < PL; ++j) { m[j] *= inv_cnt; ss = fmaf(m[j], m[j], ss); }
    ss = warp_sum(ss);
    const float inv_norm = sign / sqrtf(ss);
#pragma unroll
    for (int j = 0; j < PL; ++j) acc[j] = fmaf(m[j], inv_norm, acc[j]);
}

template <int PL>
__global__ void __launch_bounds__(256)
compose_queries_kernel(const float* __restrict__ table, long long n, int d, const long long* __restrict__ anchor,
                       const long long* __restrict__ pos_ptr, const long long* __restrict__ pos_rows,
                       const long long* __restrict__ neg_ptr, const long long* __restrict__ neg_rows, int B,
                       float* __restrict__ out) {
    const int q = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (q >= B) return;
    float acc[PL];
#pragma unroll
    for (int j = 0; j < PL; ++j) acc[j] = 0.f;
    int parts = 1;
    mean_unit_accumulate<PL>(table, n, d, anchor, q, q + 1, 1.0f, acc);
    if (pos_ptr && pos_ptr[q + 1] > pos_ptr[q]) {
        mean_unit_accumulate<PL>(table, n, d, pos_rows, pos_ptr[q], pos_ptr[q + 1], 1.0f, acc);
        ++parts;
    }
    if (neg_ptr && neg_ptr[q + 1] > neg_ptr[q]) {
        mean_unit_accumulate<PL>(table, n, d, neg_rows, neg_ptr[q], neg_ptr[q + 1], -1.0f, acc);
        ++parts;
    }
    const float inv = 1.0f / (float)parts;
#pragma unroll
    for (int j = 0; j < PL; ++j) {
        const int c = lane_id() + 32 * j;
        if (c < d) out[(size_t)q * d + c] = acc[j] * inv;
    }
}

// get_average_embeddings (hwer/recommendation_base.py:153-155): unit(mean(rows)) of each CSR list, one warp each.
template <int PL>
__global__ void __launch_bounds__(256)
average_embeddings_kernel(const float* __restrict__ table, long long n, int d, const long long* __restrict__ ptr,
                          const long long* __restrict__ rows, int L, float* __restrict__ out) {
    const int l = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (l >= L) return;
    float acc[PL];
#pragma unroll
    for (int j = 0; j < PL; ++j) acc[j] = 0.f;
    mean_unit_accumulate<PL>(table, n, d, rows, ptr[l], ptr[l + 1], 1.0f, acc);     // empty list: 0/0 = NaN like numpy
#pragma unroll
    for (int j = 0; j < PL; ++j) {
        const int c = lane_id() + 32 * j;
        if (c < d) out[(size_t)l * d + c] = acc[j];
    }
}

// ---------------------------------------------------------------------------
// Evaluation: one warp per user.
// per_user row layout (M = 3 * n_cut + 1 doubles):
//   [3*c + 0] recall@cut[c]   [3*c + 1] ndcg@cut[c] (graded)   [3*c + 2] binary ndcg@cut[c]
//   [3*n_cut] reciprocal rank over the largest cutoff
// ---------------------------------------------------------------------------
constexpr int kEvalMaxPred = 256;   // largest supported cutoff
constexpr int kEvalWarps = 4;

__device__ __forceinline__ bool sorted_contains(const long long* __restrict__ a, long long lo, long long hi,
                                                long long key) {
    while (lo < hi) {
        const long long mid = (lo + hi) >> 1;
        const long long v = a[mid];
        if (v < key) lo = mid + 1; else if (v > key) hi = mid; else return true;
    }
    return false;
}

__global__ void __launch_bounds__(kEvalWarps * 32)
eval_kernel(const long long* __restrict__ topk, int U, int Kret, const long long* __restrict__ train_ptr,
            const long long* __restrict__ train_idx, const long long* __restrict__ val_ptr,
            const long long* __restrict__ val_idx, const float* __restrict__ val_rel,
            const int* __restrict__ cutoffs, int n_cut, long long n_items, double* __restrict__ per_user,
            unsigned int* __restrict__ seen_bitmap) {
    __shared__ long long preds_s[kEvalWarps][kEvalMaxPred];
    __shared__ float rel_s[kEvalWarps][kEvalMaxPred];     // relevance of each prediction (0 = miss)
    __shared__ unsigned char hit_s[kEvalWarps][kEvalMaxPred];
    const int lane = lane_id(), w = threadIdx.x >> 5;
    const int u = blockIdx.x * kEvalWarps + w;
    if (u >= U) return;
    const int M = 3 * n_cut + 1;
    const int max_cut = cutoffs[n_cut - 1];
    const long long tb = train_ptr[u], te = train_ptr[u + 1];
    const long long vb = val_ptr[u], ve = val_ptr[u + 1];

    // 1. predictions in rank order with train items removed (validation.py:136), first max_cut kept
    int n_pred = 0;
    for (int base = 0; base < Kret && n_pred < max_cut; base += 32) {
        const int i = base + lane;
        long long item = (i < Kret) ? topk[(size_t)u * Kret + i] : -1;
        const bool keep = item >= 0 && !sorted_contains(train_idx, tb, te, item);
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        const int pos = n_pred + __popc(m & ((1u << lane) - 1u));
        if (keep && pos < max_cut) preds_s[w][pos] = item;
        n_pred += __popc(m);
    }
    if (n_pred > max_cut) n_pred = max_cut;
    __syncwarp();
    // diversity: union of every user's filtered top-max_cut (validation.py:145)
    for (int i = lane; i < n_pred; i += 32) {
        const long long it = preds_s[w][i];
        if (it < n_items) atomicOr(&seen_bitmap[it >> 5], 1u << (it & 31));
    }
    // 2. relevance of each prediction: validation items not seen in training (validation.py:158-161)
    for (int i = lane; i < n_pred; i += 32) {
        const long long it = preds_s[w][i];
        float r = 0.f; unsigned char h = 0;
        for (long long j = vb; j < ve; ++j) {
            if (val_idx[j] == it) { r = val_rel[j]; h = 1; break; }
        }
        rel_s[w][i] = r; hit_s[w][i] = h;
    }
    __syncwarp();
    // 3. the user's true list, relevance-descending (host sorted), train items dropped
    //    -> ideal gains accumulate in rank order; lane 0 walks it (short lists)
    double* outp = per_user + (size_t)u * M;
    if (lane == 0) {
        double rr = 0.0;
        for (int i = 0; i < n_pred; ++i) if (hit_s[w][i]) { rr = 1.0 / (i + 1); break; }
        outp[3 * n_cut] = rr;
        int c = 0;
        // running sums over prediction ranks
        double dcg = 0, dcg_b = 0; int hits = 0;
        // running ideal sums over the filtered true list
        double idcg = 0, idcg_b = 0; int t_seen = 0; long long vj = vb;
        int i = 0;
        for (c = 0; c < n_cut; ++c) {
            const int L = cutoffs[c] < n_pred ? cutoffs[c] : n_pred;   // len(y_pred) at this cutoff
            for (; i < L; ++i) {
                const double disc = log2((double)i + 2.0);
                if (hit_s[w][i]) { dcg += (exp2((double)rel_s[w][i]) - 1.0) / disc; dcg_b += 1.0 / disc; ++hits; }
            }
            // ideal list truncated to len(y_pred) (utils.py:103)
            while (t_seen < L && vj < ve) {
                const long long it = val_idx[vj];
                if (!sorted_contains(train_idx, tb, te, it)) {
                    const double disc = log2((double)t_seen + 2.0);
                    idcg += (exp2((double)val_rel[vj]) - 1.0) / disc;
                    idcg_b += 1.0 / disc;
                    ++t_seen;
                }
                ++vj;
            }
            // |y_true| (full filtered length) for recall's denominator (utils.py:119)
            long long T = t_seen;
            for (long long j = vj; j < ve; ++j) T += sorted_contains(train_idx, tb, te, val_idx[j]) ? 0 : 1;
            const long long nrm = (long long)L < T ? (long long)L : T;
            outp[3 * c + 0] = (double)hits / (double)(nrm > 1 ? nrm : 1);
            outp[3 * c + 1] = dcg / (idcg + 1e-8);
            outp[3 * c + 2] = dcg_b / (idcg_b + 1e-8);
        }
    }
}

// Deterministic mean over validation users (users with a non-empty validation row).
// out[0..M) = means, out[M] = diversity numerator (popcount), out[M+1] = #validation users
__global__ void __launch_bounds__(256)
eval_reduce_kernel(const double* __restrict__ per_user, int U, int M, const long long* __restrict__ val_ptr,
                   const unsigned int* __restrict__ seen_bitmap, long long n_items, double* __restrict__ out) {
    __shared__ double sh[256];
    const int m = blockIdx.x;   // one CTA per output column; columns M and M+1 are the two counters
    double acc = 0.0;
    if (m < M) {
        for (int u = threadIdx.x; u < U; u += 256)
            if (val_ptr[u + 1] > val_ptr[u]) acc += per_user[(size_t)u * M + m];
    } else if (m == M) {
        const long long words = (n_items + 31) >> 5;
        for (long long i = threadIdx.x; i < words; i += 256) acc += (double)__popc(seen_bitmap[i]);
    } else {
        for (int u = threadIdx.x; u < U; u += 256) acc += (val_ptr[u + 1] > val_ptr[u]) ? 1.0 : 0.0;
    }
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[m] = sh[0];
}

__global__ void eval_finalize_kernel(double* out, int M) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        const double nu = out[M + 1];
        for (int m = 0; m < M; ++m) out[m] = nu > 0 ? out[m] / nu : 0.0;
    }
}

}  // namespace

cudaError_t launch_pair_score(const float* table, long long n, int d, const long long* src, const long long* dst,
                              long long P, float* out, cudaStream_t stream) {
    if (P <= 0) return cudaSuccess;
    long long blocks = (P + 7) / 8;
    if (blocks > 148LL * 16) blocks = 148LL * 16;
    pair_score_kernel<<<(int)blocks, 256, 0, stream>>>(table, n, d, src, dst, P, out);
    return cudaGetLastError();
}

cudaError_t launch_compose_queries(const float* table, long long n, int d, const long long* anchor,
                                   const long long* pos_ptr, const long long* pos_rows, const long long* neg_ptr,
                                   const long long* neg_rows, int B, float* out, cudaStream_t stream) {
    if (B <= 0) return cudaSuccess;
    if (d > 32 * kComposeMaxPerLane) return cudaErrorInvalidValue;
#define HWER_COMPOSE(PL)                                                                                              \
    compose_queries_kernel<PL><<<(B + 7) / 8, 256, 0, stream>>>(table, n, d, anchor, pos_ptr, pos_rows, neg_ptr, neg_rows, \
                                                                B, out)
    if (d <= 128) HWER_COMPOSE(4); else if (d <= 256) HWER_COMPOSE(8); else HWER_COMPOSE(kComposeMaxPerLane);
#undef HWER_COMPOSE
    return cudaGetLastError();
}

cudaError_t launch_average_embeddings(const float* table, long long n, int d, const long long* ptr,
                                      const long long* rows, int L, float* out, cudaStream_t stream) {
    if (L <= 0) return cudaSuccess;
    if (d > 32 * kComposeMaxPerLane) return cudaErrorInvalidValue;
    if (d <= 128) average_embeddings_kernel<4><<<(L + 7) / 8, 256, 0, stream>>>(table, n, d, ptr, rows, L, out);
    else if (d <= 256) average_embeddings_kernel<8><<<(L + 7) / 8, 256, 0, stream>>>(table, n, d, ptr, rows, L, out);
    else average_embeddings_kernel<kComposeMaxPerLane><<<(L + 7) / 8, 256, 0, stream>>>(table, n, d, ptr, rows, L, out);
    return cudaGetLastError();
}

cudaError_t launch_eval(const long long* topk, int U, int Kret, const long long* train_ptr,
                        const long long* train_idx, const long long* val_ptr, const long long* val_idx,
                        const float* val_rel, const int* cutoffs, int n_cut, long long n_items, double* per_user,
                        unsigned int* seen_bitmap, cudaStream_t stream) {
    if (U <= 0) return cudaSuccess;
    eval_kernel<<<(U + kEvalWarps - 1) / kEvalWarps, kEvalWarps * 32, 0, stream>>>(
        topk, U, Kret, train_ptr, train_idx, val_ptr, val_idx, val_rel, cutoffs, n_cut, n_items, per_user,
        seen_bitmap);
    return cudaGetLastError();
}

cudaError_t launch_eval_reduce(const double* per_user, int U, int M, const long long* val_ptr,
                               const unsigned int* seen_bitmap, long long n_items, int n_cut, double* out,
                               cudaStream_t stream) {
    (void)n_cut;
    eval_reduce_kernel<<<M + 2, 256, 0, stream>>>(per_user, U, M, val_ptr, seen_bitmap, n_items, out);
    eval_finalize_kernel<<<1, 32, 0, stream>>>(out, M);
    return cudaGetLastError();
}

}  // namespace hwer
