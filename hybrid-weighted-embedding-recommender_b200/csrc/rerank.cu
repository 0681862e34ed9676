// Score conventions + per-anchor ordering of a batch of find_closest_neighbours calls, and the small row
// gathers around them -- everything the reference does to the k retrieved rows after the k-NN search:
//
//   rerank (PAIR)    hwer/recommendation_base.py:172-174  scores = predict([(anchor, node)]) = (a . x + 1) / 2
//                                                         in the table's fp32, sorted descending
//   rerank (DIST)    hwer/gcn_ncf.py:378-383              (2 - dist) / 2, dist = KDTree64's Euclidean distance of the
//                                                         row to the COMPOSED query embedding, sorted descending
//   rerank (GIVEN)   hwer/gcn_ncf.py:384-386              NCF scores of the k rows, sorted descending
//   rerank (EUCLID)  hwer/recommendation_base.py:79-82    MultiKNN.query: (node, dist) sorted ascending
//   map_rows         hwer/recommendation_base.py:80       local row of a per-type index -> global row
//   gather_rows      hwer/recommendation_base.py:146-151  get_embeddings (unknown node -> clip(row 0, 1e-6, 1e-5))
//   hit_rank         hwer/validation.py:82-96             ncf_eval: rank of the positive among 1 + M scored items,
//                                                         HR@n and binary NDCG@n
//
// Python's sorted(..., reverse=True) is stable: equal scores keep the order the k-NN search returned them in
// (here: dot product descending, row ascending).  The sort key is (score image, position), so this is too.
#include "common.cuh"
#include "kernels.h"

namespace hwer {

namespace {

__device__ __forceinline__ float unknown_clip_r(float v) { return fminf(fmaxf(v, 1e-6f), 1e-5f); }

__device__ __forceinline__ int pow2_ge_dev(int v) {
    int p = 2;
    while (p < v) p <<= 1;
    return p;
}

// One CTA per anchor.  Phase 1: one warp per retrieved row computes its score (row gathers: latency-bound, all
// warps of the CTA in flight).  Phase 2: bitonic sort of (key, position) in shared memory.  Phase 3: write.
__global__ void __launch_bounds__(512)
rerank_kernel(const float* __restrict__ table, long long n, int d, const long long* __restrict__ rows,
              const long long* __restrict__ row_map, int k, int conv, const long long* __restrict__ anchor_rows,
              const float* __restrict__ queries, const float* __restrict__ given, long long* __restrict__ out_rows,
              double* __restrict__ out_score) {
    extern __shared__ unsigned long long rr_sm[];
    const int P = pow2_ge_dev(k);
    unsigned long long* key = rr_sm;                                  // [P] order-preserving image of the score
    unsigned int* pos = reinterpret_cast<unsigned int*>(rr_sm + P);   // [P] position in the input list
    const int q = blockIdx.x;
    const int lane = lane_id(), warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const long long* my_rows = rows + (size_t)q * k;

    if (conv == kScoreGiven) {
        for (int i = threadIdx.x; i < k; i += blockDim.x) {
            const long long r = my_rows[i];
            const double s = (double)given[(size_t)q * k + i];
            key[i] = r < 0 ? 0ull : (s == s ? f64_to_ordered(s) : 1ull);
            pos[i] = (unsigned int)i;
        }
    } else {
        const long long a = conv == kScorePair ? anchor_rows[q] : 0;
        const bool ua = a < 0 || a >= n;
        const float* xa = table + (size_t)(ua ? 0 : a) * d;
        const float* qv = queries ? queries + (size_t)q * d : nullptr;
        for (int i = warp; i < k; i += nwarps) {
            const long long r = my_rows[i];
            long long g = r;
            if (r >= 0 && row_map) g = row_map[r];
            unsigned long long kk = 0ull;
            if (r >= 0) {
                const bool ub = g < 0 || g >= n;
                const float* xb = table + (size_t)(ub ? 0 : g) * d;
                double s;
                if (conv == kScorePair) {
                    // the arithmetic of pair_score_kernel (pair_eval.cu), so predict() and this agree bit for bit
                    float acc = 0.f;
                    for (int j = lane; j < d; j += 32) {
                        float va = xa[j], vb = xb[j];
                        if (ua) va = unknown_clip_r(va);
                        if (ub) vb = unknown_clip_r(vb);
                        acc = fmaf(va, vb, acc);
                    }
                    acc = warp_sum(acc);
                    s = (double)((acc + 1.0f) / 2.0f);
                } else {
                    // KDTree64: float64 copy of the fp32 rows, sqrt of the summed squared differences
                    double acc = 0.0;
                    for (int j = lane; j < d; j += 32) {
                        const double df = (double)xb[j] - (double)qv[j];
                        acc = fma(df, df, acc);
                    }
                    const double dist = sqrt(warp_sum(acc));
                    s = conv == kScoreDist ? (2.0 - dist) / 2.0 : dist;
                }
                // EUCLID sorts ascending: complement the image so the one descending sort below serves both
                kk = s == s ? (conv == kScoreEuclid ? ~f64_to_ordered(s) : f64_to_ordered(s)) : 1ull;
                if (kk < 2ull) kk = s == s ? 2ull : 1ull;
            }
            if (lane == 0) { key[i] = kk; pos[i] = (unsigned int)i; }
        }
    }
    for (int i = k + threadIdx.x; i < P; i += blockDim.x) { key[i] = 0ull; pos[i] = 0xffffffffu; }
    __syncthreads();
    for (int kk = 2; kk <= P; kk <<= 1) {
        for (int j = kk >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < P; i += blockDim.x) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const bool desc = (i & kk) == 0;
                    const bool i_first = key[i] > key[ixj] || (key[i] == key[ixj] && pos[i] < pos[ixj]);
                    if (desc ? !i_first : i_first) {
                        const unsigned long long t = key[i]; key[i] = key[ixj]; key[ixj] = t;
                        const unsigned int p2 = pos[i]; pos[i] = pos[ixj]; pos[ixj] = p2;
                    }
                }
            }
            __syncthreads();
        }
    }
    for (int i = threadIdx.x; i < k; i += blockDim.x) {
        const size_t o = (size_t)q * k + i;
        const unsigned long long kk = key[i];
        const unsigned int p = pos[i];
        if (kk == 0ull || p >= (unsigned int)k) {
            out_rows[o] = -1;
            out_score[o] = conv == kScoreEuclid ? INFINITY : -INFINITY;
            continue;
        }
        const long long r = my_rows[p];
        out_rows[o] = row_map ? row_map[r] : r;
        out_score[o] = kk == 1ull ? __longlong_as_double(0x7ff8000000000000ll)
                                  : ordered_to_f64(conv == kScoreEuclid ? ~kk : kk);
    }
}

__global__ void __launch_bounds__(256)
map_rows_kernel(const long long* __restrict__ rows, long long count, const long long* __restrict__ row_map,
                long long offset, long long* __restrict__ out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count;
         i += (long long)gridDim.x * blockDim.x) {
        const long long r = rows[i];
        out[i] = r < 0 ? r : (row_map ? row_map[r] : r + offset);
    }
}

__global__ void __launch_bounds__(256)
gather_rows_kernel(const float* __restrict__ table, long long n, int d, const long long* __restrict__ rows,
                   long long P, float* __restrict__ out) {
    const long long total = P * (long long)d;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long p = i / d;
        const int c = (int)(i - p * d);
        const long long r = rows[p];
        const bool unk = r < 0 || r >= n;
        const float v = table[(size_t)(unk ? 0 : r) * d + c];
        out[i] = unk ? unknown_clip_r(v) : v;
    }
}

// d % 4 == 0: a thread moves 16 bytes (one row id load per 4 columns instead of one per column)
__global__ void __launch_bounds__(256)
gather_rows_v4_kernel(const float* __restrict__ table, long long n, int d4, const long long* __restrict__ rows,
                      long long P, float* __restrict__ out) {
    const long long total = P * (long long)d4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long p = i / d4;
        const int c = (int)(i - p * d4);
        const long long r = rows[p];
        const bool unk = r < 0 || r >= n;
        float4 v = __ldg(reinterpret_cast<const float4*>(table) + (size_t)(unk ? 0 : r) * d4 + c);
        if (unk) { v.x = unknown_clip_r(v.x); v.y = unknown_clip_r(v.y); v.z = unknown_clip_r(v.z); v.w = unknown_clip_r(v.w); }
        reinterpret_cast<float4*>(out)[i] = v;
    }
}

// ncf_eval: scores [U, 1 + M], column 0 = the positive.  A stable descending sort keeps the positive ahead of
// equal-scored negatives, so its rank is the number of strictly greater negatives.  One warp per user.
__global__ void __launch_bounds__(256)
hit_rank_kernel(const float* __restrict__ scores, int U, int M1, int* __restrict__ rank) {
    const int u = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (u >= U) return;
    const float* s = scores + (size_t)u * M1;
    const float s0 = s[0];
    int c = 0;
    for (int j = 1 + (int)lane_id(); j < M1; j += 32) c += (s[j] > s0) ? 1 : 0;
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if (lane_id() == 0) rank[u] = c;
}

// out2 = {mean(rank < topn), mean(rank < topn ? 1 / log2(rank + 2) / (1 + 1e-8) : 0)}: HR@n and
// utils.binary_ndcg_v2([positive], top-n list) (hwer/utils.py:101-115).  One CTA, fixed summation order.
__global__ void __launch_bounds__(256)
hit_rank_reduce_kernel(const int* __restrict__ rank, int U, int topn, double* __restrict__ out2) {
    __shared__ double sh[2][256];
    double hr = 0.0, nd = 0.0;
    for (int u = threadIdx.x; u < U; u += 256) {
        const int r = rank[u];
        if (r < topn) { hr += 1.0; nd += 1.0 / log2((double)r + 2.0) / (1.0 + 1e-8); }
    }
    sh[0][threadIdx.x] = hr;
    sh[1][threadIdx.x] = nd;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) { sh[0][threadIdx.x] += sh[0][threadIdx.x + s]; sh[1][threadIdx.x] += sh[1][threadIdx.x + s]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { out2[0] = sh[0][0] / (double)U; out2[1] = sh[1][0] / (double)U; }
}

}  // namespace

cudaError_t launch_rerank(const float* table, long long n, int d, const long long* rows, const long long* row_map,
                          int B, int k, int conv, const long long* anchor_rows, const float* queries,
                          const float* given, long long* out_rows, double* out_score, cudaStream_t stream) {
    if (B <= 0 || k <= 0) return cudaSuccess;
    int P = 2;
    while (P < k) P <<= 1;
    const size_t smem = (size_t)P * 12;
    if (smem > (size_t)kSmemBudget) return cudaErrorInvalidValue;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(rerank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    const int threads = P <= 128 ? 128 : (P <= 1024 ? 256 : 512);
    rerank_kernel<<<B, threads, smem, stream>>>(table, n, d, rows, row_map, k, conv, anchor_rows, queries, given,
                                                out_rows, out_score);
    return cudaGetLastError();
}

cudaError_t launch_map_rows(const long long* rows, long long count, const long long* row_map, long long offset,
                            long long* out, cudaStream_t stream) {
    if (count <= 0) return cudaSuccess;
    long long blocks = (count + 255) / 256;
    if (blocks > 148LL * 8) blocks = 148LL * 8;
    map_rows_kernel<<<(int)blocks, 256, 0, stream>>>(rows, count, row_map, offset, out);
    return cudaGetLastError();
}

cudaError_t launch_gather_rows(const float* table, long long n, int d, const long long* rows, long long P, float* out,
                               cudaStream_t stream) {
    if (P <= 0) return cudaSuccess;
    const bool v4 = (d & 3) == 0 && ((reinterpret_cast<uintptr_t>(table) | reinterpret_cast<uintptr_t>(out)) & 15u) == 0;
    long long blocks = (P * (v4 ? d / 4 : d) + 255) / 256;
    if (blocks > 148LL * 32) blocks = 148LL * 32;
    if (v4) gather_rows_v4_kernel<<<(int)blocks, 256, 0, stream>>>(table, n, d / 4, rows, P, out);
    else gather_rows_kernel<<<(int)blocks, 256, 0, stream>>>(table, n, d, rows, P, out);
    return cudaGetLastError();
}

cudaError_t launch_hit_rank(const float* scores, int U, int M1, int topn, int* rank, double* out2, cudaStream_t stream) {
    if (U <= 0) return cudaSuccess;
    hit_rank_kernel<<<(U + 7) / 8, 256, 0, stream>>>(scores, U, M1, rank);
    hit_rank_reduce_kernel<<<1, 256, 0, stream>>>(rank, U, topn, out2);
    return cudaGetLastError();
}

}  // namespace hwer
