// Fused score-and-select, stage 2: per-query selection over the candidate
// lists the filter kernels produce.
//
//   select_compact : K-th best approximate score -> next admission threshold
//                    (K-th score minus the proven bf16 error margin), list
//                    compacted to the entries still able to reach the top K.
//   final          : exact mode re-scores the surviving candidates from the
//                    fp32 table in fp64 (what sklearn's KDTree64 does with the
//                    reference's table, hwer/recommendation_base.py:74,79) and
//                    orders them (score desc, row asc); bf16 mode orders the
//                    tensor-core scores directly.
//   merge          : G shards x K -> K with the same ordering rule (multi-GPU).
//
// One CTA per query; lists are sorted in shared memory with a bitonic network.
#include "common.cuh"
#include "kernels.h"

namespace hwer {

namespace {

constexpr int kSelThreads = 256;
constexpr int kFinalThreads = 512;   // 16 warps re-score candidates concurrently (row gathers are latency bound)

__device__ __forceinline__ int next_pow2(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

// Descending bitonic sort of P (power of two) elements addressed through
// `before(i, j)` ("element i must precede element j") and `swap(i, j)`.
template <class Before, class Swap>
__device__ __forceinline__ void block_bitonic(int P, Before before, Swap swap) {
    for (int k = 2; k <= P; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < P; i += blockDim.x) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const bool desc = (i & k) == 0;
                    if (desc ? before(ixj, i) : before(i, ixj)) swap(i, ixj);
                }
            }
            __syncthreads();
        }
    }
}

// K-th largest key prefix by radix select (4 passes x 8 bits over the order-preserving score image), then an
// unordered compaction of everything at or above the new threshold.  O(c) work per query; a full sort of the
// list is only needed once, in final_kernel, over the ~1.4 K survivors.
__global__ void __launch_bounds__(kSelThreads)
select_compact_kernel(unsigned long long* __restrict__ cand, unsigned int* __restrict__ cnt, unsigned int cap, int K,
                      int fixed_count, const float* __restrict__ margin, float* __restrict__ thr,
                      unsigned int* needed_cap) {
    extern __shared__ unsigned long long keys[];
    __shared__ unsigned int hist[256];
    __shared__ unsigned int sel_prefix, sel_remaining, kept_s, valid_s;
    const int q = blockIdx.x;
    unsigned int c_raw = fixed_count >= 0 ? (unsigned int)fixed_count : cnt[q];
    if (c_raw > cap) {
        if (threadIdx.x == 0) atomicMax(needed_cap, c_raw);
        c_raw = cap;
    }
    const int c = (int)c_raw;
    unsigned long long* list = cand + (size_t)q * cap;
    if (threadIdx.x == 0) { sel_prefix = 0u; sel_remaining = (unsigned int)K; kept_s = 0u; valid_s = 0u; }
    __syncthreads();
    unsigned int nvalid = 0;
    for (int i = threadIdx.x; i < c; i += blockDim.x) {
        const unsigned long long k = list[i];
        keys[i] = k;
        nvalid += (k != 0ull) ? 1u : 0u;     // key 0 = empty slot of a dense round (tail rows, NaN scores)
    }
    for (int o = 16; o > 0; o >>= 1) nvalid += __shfl_xor_sync(0xffffffffu, nvalid, o);
    if (lane_id() == 0 && nvalid) atomicAdd(&valid_s, nvalid);
    __syncthreads();
    float t = __int_as_float(0xff800000);   // -inf: fewer than K candidates so far, admit everything
    if ((int)valid_s >= K) {
        unsigned int mask = 0u;
        for (int shift = 24; shift >= 0; shift -= 8) {
            for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0u;
            __syncthreads();
            const unsigned int prefix = sel_prefix;
            for (int i = threadIdx.x; i < c; i += blockDim.x) {
                const unsigned long long k = keys[i];
                const unsigned int hi = (unsigned int)(k >> 32);
                if (k != 0ull && (hi & mask) == prefix) atomicAdd(&hist[(hi >> shift) & 255u], 1u);
            }
            __syncthreads();
            if (threadIdx.x < 32) {
                // lane l owns bins [8l, 8l+8); walk from the top bin down to the one holding the K-th key
                const int l = threadIdx.x;
                unsigned int mine = 0u;
#pragma unroll
                for (int b = 0; b < 8; ++b) mine += hist[8 * l + b];
                unsigned int above = mine;      // inclusive suffix sum over lanes >= l
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const unsigned int x = __shfl_down_sync(0xffffffffu, above, o);
                    if (l + o < 32) above += x;
                }
                const unsigned int rem = sel_remaining;
                const unsigned int strictly_above = above - mine;      // keys in bins of higher lanes
                if (strictly_above < rem && rem <= above) {
                    unsigned int acc = strictly_above;
                    for (int b = 7; b >= 0; --b) {
                        const unsigned int h = hist[8 * l + b];
                        if (acc + h >= rem) {
                            sel_prefix = prefix | ((unsigned int)(8 * l + b) << shift);
                            sel_remaining = rem - acc;
                            break;
                        }
                        acc += h;
                    }
                }
            }
            mask |= 255u << shift;
            __syncthreads();
        }
        const float sk = ordered_to_f32(sel_prefix);
        const float m = margin ? margin[q] : 0.0f;
        t = sk - m;
        if (m > 0.0f) t -= 1e-6f * (fabsf(sk) + m);   // absorb the rounding of the subtraction itself
    }
    for (int i = threadIdx.x; i < c; i += blockDim.x) {
        const unsigned long long k = keys[i];
        if (k != 0ull && key_score(k) >= t) list[atomicAdd(&kept_s, 1u)] = k;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        cnt[q] = kept_s;
        thr[q] = t;
    }
}

// Exact fp64 dot of one fp32 table row with the fp32 query, one warp per
// candidate, fixed summation order (deterministic).
__device__ __forceinline__ double exact_dot(const float* __restrict__ x, const float* __restrict__ qv, int d) {
    double s = 0.0;
    for (int c = lane_id(); c < d; c += 32) s = fma((double)x[c], (double)qv[c], s);
    return warp_sum(s);
}

template <bool EXACT>
__global__ void __launch_bounds__(kFinalThreads)
final_kernel(const unsigned long long* __restrict__ cand, const unsigned int* __restrict__ cnt, unsigned int cap,
             int K, const float* __restrict__ table, int d, const float* __restrict__ queries, long long idx_offset,
             long long* __restrict__ out_idx, float* __restrict__ out_score, double* __restrict__ out_score64,
             unsigned int* needed_cap) {
    extern __shared__ unsigned long long sm[];
    const int q = blockIdx.x;
    unsigned int c_raw = cnt[q];
    if (c_raw > cap) {
        if (threadIdx.x == 0) atomicMax(needed_cap, c_raw);
        c_raw = cap;
    }
    const int c = (int)c_raw;
    const int P = next_pow2(c > 1 ? c : 2);
    unsigned long long* sk = sm;                                   // [P] ordered score (fp64 or fp32 image)
    uint32_t* rw = reinterpret_cast<uint32_t*>(sm + P);            // [P] row
    const unsigned long long* list = cand + (size_t)q * cap;
    if (EXACT) {
        const float* qv = queries + (size_t)q * d;
        const int warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
        for (int i = warp; i < c; i += nwarps) {
            const unsigned long long k = list[i];
            const uint32_t row = key_row(k);
            const double s = (k != 0ull) ? exact_dot(table + (size_t)row * d, qv, d) : 0.0;
            if (lane_id() == 0) {
                sk[i] = (k != 0ull && s == s) ? f64_to_ordered(s) : 0ull;
                rw[i] = (k != 0ull && s == s) ? row : 0xffffffffu;
            }
        }
        for (int i = c + threadIdx.x; i < P; i += blockDim.x) { sk[i] = 0ull; rw[i] = 0xffffffffu; }
    } else {
        for (int i = threadIdx.x; i < P; i += blockDim.x) {
            const unsigned long long k = (i < c) ? list[i] : 0ull;
            sk[i] = k >> 32;
            rw[i] = (k != 0ull) ? key_row(k) : 0xffffffffu;
        }
    }
    __syncthreads();
    block_bitonic(P,
                  [&](int a, int b) { return sk[a] > sk[b] || (sk[a] == sk[b] && rw[a] < rw[b]); },
                  [&](int a, int b) {
                      unsigned long long t = sk[a]; sk[a] = sk[b]; sk[b] = t;
                      uint32_t r = rw[a]; rw[a] = rw[b]; rw[b] = r;
                  });
    for (int i = threadIdx.x; i < K; i += blockDim.x) {
        const size_t o = (size_t)q * K + i;
        if (i < c && rw[i] != 0xffffffffu) {
            const double s = EXACT ? ordered_to_f64(sk[i]) : (double)ordered_to_f32((uint32_t)sk[i]);
            out_idx[o] = (long long)rw[i] + idx_offset;
            out_score[o] = (float)s;
            if (out_score64) out_score64[o] = s;
        } else {
            out_idx[o] = -1;
            out_score[o] = __int_as_float(0xff800000);
            if (out_score64) out_score64[o] = -INFINITY;
        }
    }
}

__global__ void __launch_bounds__(kSelThreads)
merge_kernel(const double* __restrict__ scores, const long long* __restrict__ idx, int G, int B, int K,
             long long* __restrict__ out_idx, float* __restrict__ out_score, double* __restrict__ out_score64) {
    extern __shared__ unsigned long long sm[];
    const int q = blockIdx.x;
    const int c = G * K;
    const int P = next_pow2(c > 1 ? c : 2);
    unsigned long long* sk = sm;
    long long* id = reinterpret_cast<long long*>(sm + P);
    for (int i = threadIdx.x; i < P; i += blockDim.x) {
        if (i < c) {
            const int g = i / K, j = i - g * K;
            const size_t o = ((size_t)g * B + q) * K + j;
            const long long ix = idx[o];
            sk[i] = ix < 0 ? 0ull : f64_to_ordered(scores[o]);
            id[i] = ix < 0 ? 0x7fffffffffffffffll : ix;
        } else {
            sk[i] = 0ull;
            id[i] = 0x7fffffffffffffffll;
        }
    }
    __syncthreads();
    block_bitonic(P,
                  [&](int a, int b) { return sk[a] > sk[b] || (sk[a] == sk[b] && id[a] < id[b]); },
                  [&](int a, int b) {
                      unsigned long long t = sk[a]; sk[a] = sk[b]; sk[b] = t;
                      long long r = id[a]; id[a] = id[b]; id[b] = r;
                  });
    for (int i = threadIdx.x; i < K; i += blockDim.x) {
        const size_t o = (size_t)q * K + i;
        const bool ok = id[i] != 0x7fffffffffffffffll;
        const double s = ok ? ordered_to_f64(sk[i]) : -INFINITY;
        out_idx[o] = ok ? id[i] : -1;
        out_score[o] = (float)s;
        if (out_score64) out_score64[o] = s;
    }
}

// Raises a kernel's opt-in dynamic shared-memory ceiling (static shared memory counts against the same 227 KB).
template <class Kern>
cudaError_t set_smem(Kern k, size_t bytes) {
    if (bytes <= 48 * 1024) return cudaSuccess;
    return cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

inline size_t pow2_ge(size_t v) {
    size_t p = 2;
    while (p < v) p <<= 1;
    return p;
}

}  // namespace

cudaError_t launch_select_compact(unsigned long long* cand, unsigned int* cnt, unsigned int cap, int B, int K,
                                  int fixed_count, const float* margin, float* thr, unsigned int* needed_cap,
                                  cudaStream_t stream) {
    if (B <= 0) return cudaSuccess;
    // stage only what can be there: a dense round holds fixed_count keys, a filter round at most cap
    const size_t n = fixed_count >= 0 ? (size_t)fixed_count : (size_t)cap;
    const size_t smem = (n < 1024 ? 1024 : n) * sizeof(unsigned long long);
    cudaError_t e = set_smem(select_compact_kernel, smem);
    if (e != cudaSuccess) return e;
    select_compact_kernel<<<B, kSelThreads, smem, stream>>>(cand, cnt, cap, K, fixed_count, margin, thr, needed_cap);
    return cudaGetLastError();
}

cudaError_t launch_final(const unsigned long long* cand, const unsigned int* cnt, unsigned int cap, int B, int K,
                         int exact, const float* table, int d, const float* queries, long long idx_offset,
                         long long* out_idx, float* out_score, double* out_score64, unsigned int* needed_cap,
                         cudaStream_t stream) {
    if (B <= 0) return cudaSuccess;
    const size_t smem = pow2_ge(cap) * (sizeof(unsigned long long) + sizeof(uint32_t));
    cudaError_t e;
    if (exact) {
        e = set_smem(final_kernel<true>, smem);
        if (e != cudaSuccess) return e;
        final_kernel<true><<<B, kFinalThreads, smem, stream>>>(cand, cnt, cap, K, table, d, queries, idx_offset, out_idx,
                                                            out_score, out_score64, needed_cap);
    } else {
        e = set_smem(final_kernel<false>, smem);
        if (e != cudaSuccess) return e;
        final_kernel<false><<<B, kFinalThreads, smem, stream>>>(cand, cnt, cap, K, table, d, queries, idx_offset,
                                                             out_idx, out_score, out_score64, needed_cap);
    }
    return cudaGetLastError();
}

cudaError_t launch_merge(const double* scores, const long long* idx, int G, int B, int K, long long* out_idx,
                         float* out_score, double* out_score64, cudaStream_t stream) {
    if (B <= 0 || K <= 0) return cudaSuccess;
    const size_t smem = pow2_ge((size_t)G * K) * (sizeof(unsigned long long) + sizeof(long long));
    if (smem > (size_t)kSmemBudget) return cudaErrorInvalidValue;
    cudaError_t e = set_smem(merge_kernel, smem);
    if (e != cudaSuccess) return e;
    merge_kernel<<<B, kSelThreads, smem, stream>>>(scores, idx, G, B, K, out_idx, out_score, out_score64);
    return cudaGetLastError();
}

}  // namespace hwer
