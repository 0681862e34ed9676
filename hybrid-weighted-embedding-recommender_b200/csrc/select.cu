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

__global__ void __launch_bounds__(kSelThreads)
select_compact_kernel(unsigned long long* __restrict__ cand, unsigned int* __restrict__ cnt, unsigned int cap, int K,
                      const float* __restrict__ margin, float* __restrict__ thr, unsigned int* needed_cap) {
    extern __shared__ unsigned long long keys[];
    __shared__ int kept_s;
    const int q = blockIdx.x;
    unsigned int c_raw = cnt[q];
    if (c_raw > cap) {
        if (threadIdx.x == 0) atomicMax(needed_cap, c_raw);
        c_raw = cap;
    }
    const int c = (int)c_raw;
    const int P = next_pow2(c > 1 ? c : 2);
    unsigned long long* list = cand + (size_t)q * cap;
    for (int i = threadIdx.x; i < P; i += blockDim.x) keys[i] = (i < c) ? list[i] : 0ull;
    if (threadIdx.x == 0) kept_s = 0;
    __syncthreads();
    block_bitonic(P, [&](int a, int b) { return keys[a] > keys[b]; },
                  [&](int a, int b) { unsigned long long t = keys[a]; keys[a] = keys[b]; keys[b] = t; });
    float t = __int_as_float(0xff800000);   // -inf: fewer than K candidates so far, admit everything
    if (c >= K) {
        const float sk = key_score(keys[K - 1]);
        const float m = margin ? margin[q] : 0.0f;
        t = sk - m;
        if (m > 0.0f) t -= 1e-6f * (fabsf(sk) + m);   // absorb the rounding of the subtraction itself
    }
    // keys are sorted descending: count the prefix still at or above the new threshold
    int local = 0;
    for (int i = threadIdx.x; i < c; i += blockDim.x) local += (key_score(keys[i]) >= t) ? 1 : 0;
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if (lane_id() == 0 && local) atomicAdd(&kept_s, local);
    __syncthreads();
    const int kept = kept_s;
    for (int i = threadIdx.x; i < kept; i += blockDim.x) list[i] = keys[i];
    if (threadIdx.x == 0) {
        cnt[q] = (unsigned int)kept;
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
__global__ void __launch_bounds__(kSelThreads)
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
            const uint32_t row = key_row(list[i]);
            const double s = exact_dot(table + (size_t)row * d, qv, d);
            if (lane_id() == 0) {
                sk[i] = f64_to_ordered(s);
                rw[i] = row;
            }
        }
        for (int i = c + threadIdx.x; i < P; i += blockDim.x) { sk[i] = 0ull; rw[i] = 0xffffffffu; }
    } else {
        for (int i = threadIdx.x; i < P; i += blockDim.x) {
            const unsigned long long k = (i < c) ? list[i] : 0ull;
            sk[i] = k >> 32;
            rw[i] = (i < c) ? key_row(k) : 0xffffffffu;
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
        if (i < c) {
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
                                  const float* margin, float* thr, unsigned int* needed_cap, cudaStream_t stream) {
    if (B <= 0) return cudaSuccess;
    const size_t smem = pow2_ge(cap) * sizeof(unsigned long long);
    cudaError_t e = set_smem(select_compact_kernel, smem);
    if (e != cudaSuccess) return e;
    select_compact_kernel<<<B, kSelThreads, smem, stream>>>(cand, cnt, cap, K, margin, thr, needed_cap);
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
        final_kernel<true><<<B, kSelThreads, smem, stream>>>(cand, cnt, cap, K, table, d, queries, idx_offset, out_idx,
                                                            out_score, out_score64, needed_cap);
    } else {
        e = set_smem(final_kernel<false>, smem);
        if (e != cudaSuccess) return e;
        final_kernel<false><<<B, kSelThreads, smem, stream>>>(cand, cnt, cap, K, table, d, queries, idx_offset,
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
