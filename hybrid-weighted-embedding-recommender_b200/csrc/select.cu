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
#include <cstdlib>
#include <cstring>

#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"
#include "kernels.h"

namespace hwer {

namespace {

constexpr int kSelThreads = 256;
constexpr int kFinalThreads = 512;   // 16 warps re-score candidates concurrently (row gathers are latency bound)
constexpr int kFinalSmallThreads = 128;

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

// Histogram increment, one key per lane.  Plain shared-memory atomics: the hardware merges same-address updates
// of a warp, and both ballot- and match-based aggregation measured slower (profiles/README.md, v7 notes).
__device__ __forceinline__ void hist_add_aggregated(unsigned int* hist, bool take, unsigned int bin) {
    if (take) atomicAdd(&hist[bin], 1u);
}

// ---- cross-GPU threshold sharing inside the select kernels (row-sharded catalogue, DESIGN.md "Multi-GPU") ----
// Every shard publishes, per query, the ceil(k / G)-th best score of its local list; the min over shards bounds the
// GLOBAL k-th best from below, so all shards filter the next round with ~1/G the hits.  One 64-bit word per
// (shard, query) carries {round epoch, score}: a single peer store publishes it and the reader needs no separate
// flag or fence -- the warp that selected query q publishes q, then spins on its OWN memory until the G words of q
// carry this round's epoch (or a later one: a shard may already be a round ahead, and its newer bound is just as
// valid and tighter).  No extra kernel launches, no grid-wide handshake: queries proceed independently.
__device__ __forceinline__ unsigned long long sel_ld_acquire_sys64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void sel_st_relaxed_sys64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void sel_publish_kth(const SelExchange& sx, long long q, float sk) {
    const unsigned long long w = ((unsigned long long)sx.epoch << 32) | (unsigned long long)__float_as_uint(sk);
    for (int g = 0; g < sx.world; ++g) sel_st_relaxed_sys64(sx.thr_x[g] + (size_t)sx.rank * sx.b_cap + q, w);
}
// Lane g < world waits for shard g's word of query q; returns the min over shards on every lane of the warp.
// Bounded spin: a dead peer must not hang the GPU (-inf = admit everything, and the error flag is raised).
__device__ __forceinline__ float sel_wait_global_kth(const SelExchange& sx, long long q, int lane) {
    float v = __int_as_float(0x7f800000);
    if (lane < sx.world) {
        const unsigned long long* slot = sx.thr_x[sx.rank] + (size_t)lane * sx.b_cap + q;
        const long long t0 = clock64();
        unsigned long long w = sel_ld_acquire_sys64(slot);
        while ((int)((unsigned int)(w >> 32) - sx.epoch) < 0) {
            if (clock64() - t0 > 6000000000ll) { atomicExch(sx.flags + 17, 1u); w = 0xff800000ull; break; }
            __nanosleep(64);
            w = sel_ld_acquire_sys64(slot);
        }
        v = __uint_as_float((unsigned int)w);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// K-th largest key prefix by radix select (4 passes x 8 bits over the order-preserving score image), then an
// unordered compaction of everything at or above the new threshold.  O(c) work per query; a full sort of the
// list is only needed once, in final_kernel, over the ~1.4 K survivors.
__global__ void __launch_bounds__(kSelThreads)
select_compact_kernel(unsigned long long* __restrict__ cand, unsigned int* __restrict__ cnt, unsigned int cap, int K,
                      int fixed_count, const float* __restrict__ margin, float* __restrict__ thr,
                      unsigned int* needed_cap, unsigned int* __restrict__ ovf,
                      const __grid_constant__ SelExchange sx) {
    extern __shared__ unsigned long long keys[];
    __shared__ unsigned int hist[256];
    __shared__ unsigned int sel_prefix, sel_remaining, kept_s, valid_s;
    __shared__ float global_kth_s;
    const int q = blockIdx.x;
    const bool shared = sx.mode == kSelShared;
    if (shared) K = sx.k_share;                       // what this shard publishes: its ceil(k / G)-th best
    unsigned int c_raw = fixed_count >= 0 ? (unsigned int)fixed_count : cnt[q];
    if (c_raw > cap) {
        if (threadIdx.x == 0) { atomicMax(needed_cap, c_raw); ovf[q] = 1u; }
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
    float sk = t;                           // the K-th best score of the list
    if ((int)valid_s >= K) {
        unsigned int mask = 0u;
        for (int shift = 24; shift >= 0; shift -= 8) {
            for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0u;
            __syncthreads();
            const unsigned int prefix = sel_prefix;
            for (int base = 0; base < c; base += blockDim.x) {       // whole warps iterate together (match_any)
                const int i = base + threadIdx.x;
                const unsigned long long k = i < c ? keys[i] : 0ull;
                const unsigned int hi = (unsigned int)(k >> 32);
                hist_add_aggregated(hist, k != 0ull && (hi & mask) == prefix, (hi >> shift) & 255u);
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
        sk = ordered_to_f32(sel_prefix);
    }
    if (shared) {                                     // publish first, then wait: no shard ever waits on a waiter
        if (threadIdx.x == 0) sel_publish_kth(sx, sx.q0 + q, sk);       // -inf = "cannot bound yet"
        if (threadIdx.x < 32) {
            const float g = sel_wait_global_kth(sx, sx.q0 + q, threadIdx.x);
            if (threadIdx.x == 0) global_kth_s = g;
        }
        __syncthreads();
        sk = global_kth_s;                            // lower bound of the global k-th best score
    }
    if (sk > __int_as_float(0xff800000)) {
        const float m = margin ? margin[q] : 0.0f;
        t = sk - m;
        if (m > 0.0f) t -= 1e-6f * (fabsf(sk) + m);   // absorb the rounding of the subtraction itself
    }
    for (int base = 0; base < c; base += blockDim.x) {
        const int i = base + threadIdx.x;
        const unsigned long long k = i < c ? keys[i] : 0ull;
        const bool keep = k != 0ull && key_score(k) >= t;
        const unsigned int m = __ballot_sync(0xffffffffu, keep);       // one slot-allocating atomic per warp
        unsigned int slot = 0u;
        if (lane_id() == 0 && m) slot = atomicAdd(&kept_s, (unsigned int)__popc(m));
        slot = __shfl_sync(0xffffffffu, slot, 0);
        if (keep) list[slot + __popc(m & ((1u << lane_id()) - 1u))] = k;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        cnt[q] = kept_s;
        thr[q] = t;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// The same selection with the keys held in REGISTERS and the K-th best found by BISECTION on the 32-bit score image
// (r2).  f(v) = #{keys with image >= v} is monotone, so the K-th largest image is the largest v with f(v) >= K:
// at most 32 steps of "compare my keys with the pivot, add up" -- per step a handful of integer compares per thread
// and ONE block-wide sum, against four radix passes of shared-memory histogram atomics over a staged copy of the
// list (heavily conflicting in the first pass: all scores of a query share their top byte).  The answer is the same
// number the radix select returns, so thresholds, candidate lists and results are bit-identical.
constexpr int kSelPreKeys = 512;    // score images the pre-selection may keep (2 KB of shared memory)

template <int KPT>      // keys per thread: the CTA handles lists of up to kSelThreads * KPT keys (the launcher checks cap)
__global__ void __launch_bounds__(kSelThreads)
select_compact_bisect_kernel(unsigned long long* __restrict__ cand, unsigned int* __restrict__ cnt, unsigned int cap,
                             int K, int fixed_count, const float* __restrict__ margin, float* __restrict__ thr,
                             unsigned int* needed_cap, unsigned int* __restrict__ ovf,
                             const __grid_constant__ SelExchange sx) {
    __shared__ unsigned int step_cnt[36];            // one counter per bisection step (never reset inside the loop)
    __shared__ unsigned int kept_s, valid_s, max_s, npre_s, lo0_s;
    __shared__ unsigned int tmax_s[kSelThreads];     // per-thread maximum score image (pre-selection)
    __shared__ unsigned int pre_s[kSelPreKeys];      // score images >= the pre-selection bound
    __shared__ float global_kth_s;
    const int q = blockIdx.x;
    const bool shared = sx.mode == kSelShared;
    if (shared) K = sx.k_share;                       // what this shard publishes: its ceil(k / G)-th best
    unsigned int c_raw = fixed_count >= 0 ? (unsigned int)fixed_count : cnt[q];
    if (c_raw > cap) {
        if (threadIdx.x == 0) { atomicMax(needed_cap, c_raw); ovf[q] = 1u; }
        c_raw = cap;
    }
    const int c = (int)c_raw;
    unsigned long long* list = cand + (size_t)q * cap;
    if (threadIdx.x < 36) step_cnt[threadIdx.x] = 0u;
    if (threadIdx.x == 0) { kept_s = 0u; valid_s = 0u; max_s = 0u; npre_s = 0u; lo0_s = 1u; }
    __syncthreads();
    unsigned long long key[KPT];
    unsigned int nvalid = 0u, vmax = 0u;
#pragma unroll
    for (int j = 0; j < KPT; ++j) {
        const int i = (int)threadIdx.x + j * kSelThreads;
        key[j] = i < c ? list[i] : 0ull;             // key 0 = empty slot of a dense round (tail rows, NaN scores)
        const unsigned int hi = (unsigned int)(key[j] >> 32);
        nvalid += key[j] != 0ull ? 1u : 0u;
        vmax = hi > vmax ? hi : vmax;
    }
    tmax_s[threadIdx.x] = vmax;
    nvalid = __reduce_add_sync(0xffffffffu, nvalid);
    vmax = __reduce_max_sync(0xffffffffu, vmax);
    if (lane_id() == 0) { if (nvalid) atomicAdd(&valid_s, nvalid); atomicMax(&max_s, vmax); }
    __syncthreads();
    float t = __int_as_float(0xff800000);   // -inf: fewer than K candidates so far, admit everything
    float sk = t;                           // the K-th best score of the list
    if ((int)valid_s >= K) {
        // invariant: f(lo) >= K.  Every valid key's image is >= 1 (0 is the empty key), f(1) = valid_s.
        unsigned int lo = 1u, hi = max_s;
        bool done = false;
        if (K <= kSelThreads) {
            // PRE-SELECTION (long lists, i.e. round 0): the K-th largest of the 256 per-thread maxima, lo0, is a lower
            // bound of the K-th largest key (those K maxima are K distinct keys >= lo0), and only ~K * 1.2 keys of
            // a 4096-key list reach it.  One warp finds lo0 (8 values per lane), the block compacts the keys >= lo0
            // into shared memory, and one warp finishes the bisection over them with a REDUX per step: 4 block
            // barriers instead of ~28, a third of the instructions.
            const int l = lane_id();
            if (threadIdx.x < 32) {
                unsigned int m[kSelThreads / 32], have = 0u;
#pragma unroll
                for (int j = 0; j < kSelThreads / 32; ++j) { m[j] = tmax_s[l + 32 * j]; have += m[j] != 0u ? 1u : 0u; }
                have = __reduce_add_sync(0xffffffffu, have);
                unsigned int a = 1u, b = hi;
                if (have >= (unsigned int)K) {
                    while (a < b) {
                        const unsigned int mid = a + ((b - a + 1u) >> 1);
                        unsigned int n = 0u;
#pragma unroll
                        for (int j = 0; j < kSelThreads / 32; ++j) n += m[j] >= mid ? 1u : 0u;
                        n = __reduce_add_sync(0xffffffffu, n);
                        if (n >= (unsigned int)K) a = mid; else b = mid - 1u;
                    }
                }
                if (l == 0) lo0_s = a;                // 1 = no usable bound (fewer than K threads hold a key)
            }
            __syncthreads();
            lo = lo0_s;
            // (a few per cent of the keys qualify: one shared-memory atomic per taker is far cheaper than a
            // ballot / popc / shuffle sequence per key)
#pragma unroll
            for (int j = 0; j < KPT; ++j) {
                const unsigned int h = (unsigned int)(key[j] >> 32);
                if (h >= lo) {                        // lo >= 1 excludes the empty key
                    const unsigned int slot = atomicAdd(&npre_s, 1u);
                    if (slot < (unsigned int)kSelPreKeys) pre_s[slot] = h;
                }
            }
            __syncthreads();
            const unsigned int npre = npre_s;         // = f(lo) >= K
            if (npre <= (unsigned int)kSelPreKeys) {
                if (threadIdx.x < 32) {
                    unsigned int m[kSelPreKeys / 32];
#pragma unroll
                    for (int j = 0; j < kSelPreKeys / 32; ++j) m[j] = (unsigned int)(l + 32 * j) < npre ? pre_s[l + 32 * j] : 0u;
                    unsigned int a = lo, b = hi;
                    while (a < b) {
                        const unsigned int mid = a + ((b - a + 1u) >> 1);
                        unsigned int n = 0u;
#pragma unroll
                        for (int j = 0; j < kSelPreKeys / 32; ++j) n += m[j] >= mid ? 1u : 0u;
                        n = __reduce_add_sync(0xffffffffu, n);
                        if (n >= (unsigned int)K) a = mid; else b = mid - 1u;
                    }
                    if (l == 0) lo0_s = a;
                }
                __syncthreads();
                lo = lo0_s;
                done = true;
            }
        }
        int step = 0;
        while (!done && lo < hi) {          // block-uniform: lo / hi derive from shared counters only
            const unsigned int mid = lo + ((hi - lo + 1u) >> 1);
            unsigned int n = 0u;
#pragma unroll
            for (int j = 0; j < KPT; ++j) n += (unsigned int)(key[j] >> 32) >= mid ? 1u : 0u;
            n = __reduce_add_sync(0xffffffffu, n);
            if (lane_id() == 0 && n) atomicAdd(&step_cnt[step], n);
            __syncthreads();
            if (step_cnt[step] >= (unsigned int)K) lo = mid; else hi = mid - 1u;
            ++step;
        }
        sk = ordered_to_f32(lo);
    }
    if (shared) {                                     // publish first, then wait: no shard ever waits on a waiter
        if (threadIdx.x == 0) sel_publish_kth(sx, sx.q0 + q, sk);       // -inf = "cannot bound yet"
        if (threadIdx.x < 32) {
            const float g = sel_wait_global_kth(sx, sx.q0 + q, threadIdx.x);
            if (threadIdx.x == 0) global_kth_s = g;
        }
        __syncthreads();
        sk = global_kth_s;                            // lower bound of the global k-th best score
    }
    if (sk > __int_as_float(0xff800000)) {
        const float m = margin ? margin[q] : 0.0f;
        t = sk - m;
        if (m > 0.0f) t -= 1e-6f * (fabsf(sk) + m);   // absorb the rounding of the subtraction itself
    }
    // every key is in a register by now, so compacting in place cannot overwrite an unread one.  Survivors are a
    // few per cent of a long list (round 0) but most of a short one: sparse lists take one shared-memory atomic per
    // survivor, dense ones one per warp (ballot)
    const unsigned int t_img = f32_to_ordered(t);     // key_score(k) >= t  <=>  image(k) >= image(t): no conversion per key
    if (c > 4 * K + 256) {
#pragma unroll
        for (int j = 0; j < KPT; ++j) {
            const unsigned long long k = key[j];
            if (k != 0ull && (unsigned int)(k >> 32) >= t_img) list[atomicAdd(&kept_s, 1u)] = k;
        }
    } else {
#pragma unroll
        for (int j = 0; j < KPT; ++j) {
            if (j * kSelThreads < c) {                // block-uniform: whole warps take part in the ballot
                const unsigned long long k = key[j];
                const bool keep = k != 0ull && (unsigned int)(k >> 32) >= t_img;
                const unsigned int m = __ballot_sync(0xffffffffu, keep);   // one slot-allocating atomic per warp
                unsigned int slot = 0u;
                if (lane_id() == 0 && m) slot = atomicAdd(&kept_s, (unsigned int)__popc(m));
                slot = __shfl_sync(0xffffffffu, slot, 0);
                if (keep) list[slot + __popc(m & ((1u << lane_id()) - 1u))] = k;
            }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        cnt[q] = kept_s;
        thr[q] = t;
    }
}

// ONE WARP per query, keys in registers (lists of up to 32 * kSelWarpRegKeys keys; longer ones take the radix path of
// select_compact_warp_kernel's body below): per bisection step a lane compares its keys with the pivot and one
// warp-wide integer add (REDUX) gives every lane the count -- no shared memory, no atomics, no barrier.
constexpr int kSelWarpRegKeys = 32;

template <int KPL>
__device__ __forceinline__ void select_warp_bisect(unsigned long long* __restrict__ glist, int c, int K, int q, int l,
                                                   const float* __restrict__ margin, float* __restrict__ thr,
                                                   unsigned int* __restrict__ cnt, const SelExchange& sx, bool shared) {
    unsigned long long key[KPL];
    unsigned int nvalid = 0u, vmax = 0u;
#pragma unroll
    for (int j = 0; j < KPL; ++j) {
        const int i = l + 32 * j;
        key[j] = i < c ? glist[i] : 0ull;
        const unsigned int hi = (unsigned int)(key[j] >> 32);
        nvalid += key[j] != 0ull ? 1u : 0u;
        vmax = hi > vmax ? hi : vmax;
    }
    nvalid = __reduce_add_sync(0xffffffffu, nvalid);
    vmax = __reduce_max_sync(0xffffffffu, vmax);
    float t = __int_as_float(0xff800000);
    float sk = t;
    if ((int)nvalid >= K) {
        unsigned int lo = 1u, hi = vmax;
        while (lo < hi) {
            const unsigned int mid = lo + ((hi - lo + 1u) >> 1);
            unsigned int n = 0u;
#pragma unroll
            for (int j = 0; j < KPL; ++j) n += (unsigned int)(key[j] >> 32) >= mid ? 1u : 0u;
            n = __reduce_add_sync(0xffffffffu, n);
            if (n >= (unsigned int)K) lo = mid; else hi = mid - 1u;
        }
        sk = ordered_to_f32(lo);
    }
    if (shared) {
        if (l == 0) sel_publish_kth(sx, sx.q0 + q, sk);
        __syncwarp();
        sk = sel_wait_global_kth(sx, sx.q0 + q, l);
    }
    if (sk > __int_as_float(0xff800000)) {
        const float m = margin ? margin[q] : 0.0f;
        t = sk - m;
        if (m > 0.0f) t -= 1e-6f * (fabsf(sk) + m);
    }
    unsigned int kept = 0u;
#pragma unroll
    for (int j = 0; j < KPL; ++j) {
        if (32 * j < c) {                                           // warp-uniform
            const unsigned long long k = key[j];
            const bool keep = k != 0ull && key_score(k) >= t;
            const unsigned int m = __ballot_sync(0xffffffffu, keep);
            if (keep) glist[kept + __popc(m & ((1u << l) - 1u))] = k;      // order-preserving, in place
            kept += __popc(m);
        }
    }
    if (l == 0) {
        cnt[q] = kept;
        thr[q] = t;
    }
}

// The same selection with ONE WARP per query, for the filter rounds of large batches: their lists are short
// (about k * (1 + 1.4 growth) keys), so a CTA per query spends its time in __syncthreads and leaves most of the
// machine idle, while thousands of independent warps finish in one wave.  The list is re-read from L1/L2 in every
// pass (a few KB per query) instead of being staged; the compaction is in place and order-preserving.
constexpr int kSelWarps = 4;
constexpr int kSelWarpKeys = 1024;          // keys staged per warp (8 KB); longer lists are re-read from L2

__global__ void __launch_bounds__(kSelWarps * 32)
select_compact_warp_kernel(unsigned long long* __restrict__ cand, unsigned int* __restrict__ cnt, unsigned int cap,
                           int B, int K, int fixed_count, const float* __restrict__ margin, float* __restrict__ thr,
                           unsigned int* needed_cap, unsigned int* __restrict__ ovf,
                           const __grid_constant__ SelExchange sx) {
    __shared__ unsigned long long keys_all[kSelWarps][kSelWarpKeys];
    __shared__ unsigned int hist_all[kSelWarps][256];
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    const int q = blockIdx.x * kSelWarps + w;
    if (q >= B) return;
    const bool shared = sx.mode == kSelShared;
    if (shared) K = sx.k_share;                       // what this shard publishes: its ceil(k / G)-th best
    unsigned int* hist = hist_all[w];
    unsigned int c_raw = fixed_count >= 0 ? (unsigned int)fixed_count : cnt[q];
    if (c_raw > cap) {
        if (l == 0) { atomicMax(needed_cap, c_raw); ovf[q] = 1u; }
        c_raw = cap;
    }
    const int c = (int)c_raw;
    unsigned long long* glist = cand + (size_t)q * cap;
    if (c <= 32 * kSelWarpRegKeys) {                  // the normal case: keys in registers, bisection (see below)
        if (c <= 32 * 8) select_warp_bisect<8>(glist, c, K, q, l, margin, thr, cnt, sx, shared);
        else if (c <= 32 * 16) select_warp_bisect<16>(glist, c, K, q, l, margin, thr, cnt, sx, shared);
        else select_warp_bisect<kSelWarpRegKeys>(glist, c, K, q, l, margin, thr, cnt, sx, shared);
        return;
    }
    // the list lives in shared memory when it fits (the normal case), else it is read through L2 every pass
    const bool staged = c <= kSelWarpKeys;
    volatile unsigned long long* list = staged ? keys_all[w] : glist;
    unsigned int nvalid = 0;
    for (int i = l; i < c; i += 32) {
        const unsigned long long k = glist[i];
        if (staged) keys_all[w][i] = k;
        nvalid += (k != 0ull) ? 1u : 0u;
    }
    __syncwarp();
    for (int o = 16; o > 0; o >>= 1) nvalid += __shfl_xor_sync(0xffffffffu, nvalid, o);
    float t = __int_as_float(0xff800000);   // -inf: fewer than K candidates so far, admit everything
    float sk = t;                           // the K-th best score of the list
    if ((int)nvalid >= K) {
        unsigned int mask = 0u, prefix = 0u, remaining = (unsigned int)K;
        for (int shift = 24; shift >= 0; shift -= 8) {
#pragma unroll
            for (int b = 0; b < 8; ++b) hist[8 * l + b] = 0u;
            __syncwarp();
            for (int base = 0; base < c; base += 32) {
                const int i = base + l;
                const unsigned long long k = i < c ? list[i] : 0ull;
                const unsigned int hi = (unsigned int)(k >> 32);
                hist_add_aggregated(hist, k != 0ull && (hi & mask) == prefix, (hi >> shift) & 255u);
            }
            __syncwarp();
            // lane l owns bins [8l, 8l+8); walk from the top bin down to the one holding the K-th key
            unsigned int mine = 0u, h8[8];
#pragma unroll
            for (int b = 0; b < 8; ++b) { h8[b] = hist[8 * l + b]; mine += h8[b]; }
            unsigned int above = mine;      // inclusive suffix sum over lanes >= l
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned int x = __shfl_down_sync(0xffffffffu, above, o);
                if (l + o < 32) above += x;
            }
            const unsigned int strictly_above = above - mine;
            unsigned int found = 0u, new_rem = 0u;
            if (strictly_above < remaining && remaining <= above) {
                unsigned int acc = strictly_above;
#pragma unroll
                for (int b = 7; b >= 0; --b) {
                    if (!found && acc + h8[b] >= remaining) {
                        found = 0x100u | (unsigned int)(8 * l + b);
                        new_rem = remaining - acc;
                    }
                    acc += h8[b];
                }
            }
            const unsigned int owner = __ballot_sync(0xffffffffu, found != 0u);
            const int src = __ffs(owner) - 1;                 // exactly one lane owns the K-th key's bin
            found = __shfl_sync(0xffffffffu, found, src);
            remaining = __shfl_sync(0xffffffffu, new_rem, src);
            prefix |= (found & 255u) << shift;
            mask |= 255u << shift;
            __syncwarp();
        }
        sk = ordered_to_f32(prefix);
    }
    if (shared) {                                     // publish first, then wait: no shard ever waits on a waiter
        if (l == 0) sel_publish_kth(sx, sx.q0 + q, sk);                 // -inf = "cannot bound yet"
        __syncwarp();
        sk = sel_wait_global_kth(sx, sx.q0 + q, l);   // lower bound of the global k-th best score
    }
    if (sk > __int_as_float(0xff800000)) {
        const float m = margin ? margin[q] : 0.0f;
        t = sk - m;
        if (m > 0.0f) t -= 1e-6f * (fabsf(sk) + m);   // absorb the rounding of the subtraction itself
    }
    unsigned int kept = 0u;
    for (int base = 0; base < c; base += 32) {
        const int i = base + l;
        const unsigned long long k = (i < c) ? list[i] : 0ull;
        const bool keep = k != 0ull && key_score(k) >= t;
        const unsigned int m = __ballot_sync(0xffffffffu, keep);
        // order-preserving, and in place when unstaged: the write position never passes the batch being read
        if (keep) glist[kept + __popc(m & ((1u << l) - 1u))] = k;
        kept += __popc(m);
        __syncwarp();
    }
    if (l == 0) {
        cnt[q] = kept;
        thr[q] = t;
    }
}

// Exact fp64 dot of one fp32 table row with the fp32 query, one warp per
// candidate, fixed summation order (deterministic): lane l accumulates columns
// l, l+32, ... in ascending order, then a butterfly over the lanes.
__device__ __forceinline__ double exact_dot(const float* __restrict__ x, const float* __restrict__ qv, int d) {
    double s = 0.0;
    for (int c = lane_id(); c < d; c += 32) s = fma((double)x[c], (double)qv[c], s);
    return warp_sum(s);
}

// Same sums for widths that are a multiple of 128 with 16-byte aligned rows: each lane fetches its columns with
// 128-bit loads, and two candidates are in flight per warp, because the row gathers (random 512-byte reads of a
// multi-GB table) are latency-bound.  Lane l owns columns 4l..4l+3 of every 128-column block; the per-lane partial
// sums differ from exact_dot's column assignment, so a table is always scored by ONE of the two variants
// (fp32 x fp32 products are exact in fp64; the final rounding of the sum depends on the order).
__device__ __forceinline__ void exact_dot2_v4(const float* __restrict__ xa, const float* __restrict__ xb,
                                              const float4 (&qreg)[4], int nblk, double& sa, double& sb) {
    sa = 0.0; sb = 0.0;
    const int l = lane_id();
#pragma unroll 1
    for (int b = 0; b < nblk; ++b) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(xa) + b * 32 + l);
        const float4 c = __ldg(reinterpret_cast<const float4*>(xb) + b * 32 + l);
        const float4 q = b == 0 ? qreg[0] : (b == 1 ? qreg[1] : (b == 2 ? qreg[2] : qreg[3]));
        sa = fma((double)a.x, (double)q.x, sa); sb = fma((double)c.x, (double)q.x, sb);
        sa = fma((double)a.y, (double)q.y, sa); sb = fma((double)c.y, (double)q.y, sb);
        sa = fma((double)a.z, (double)q.z, sa); sb = fma((double)c.z, (double)q.z, sb);
        sa = fma((double)a.w, (double)q.w, sa); sb = fma((double)c.w, (double)q.w, sb);
    }
    sa = warp_sum(sa);
    sb = warp_sum(sb);
}

// Four candidates in flight per warp (the small-list variant of final_kernel runs four warps per CTA): per
// candidate the same fma sequence as exact_dot2_v4, so the sums are bit-identical whichever variant scores a row.
__device__ __forceinline__ void exact_dot4_v4(const float* __restrict__ x0, const float* __restrict__ x1,
                                              const float* __restrict__ x2, const float* __restrict__ x3,
                                              const float4 (&qreg)[4], int nblk, double (&s)[4]) {
    s[0] = s[1] = s[2] = s[3] = 0.0;
    const int l = lane_id();
    if (nblk == 2) {
        // d = 256 (C3): both 128-column blocks of the four rows are requested before the first FMA -- eight loads in
        // flight instead of two dependent rounds of four; the FMA order per candidate is unchanged (block 0 then
        // block 1, x y z w), so the fp64 sums are the same bits
        float4 a[2][4];
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            a[b][0] = __ldg(reinterpret_cast<const float4*>(x0) + b * 32 + l);
            a[b][1] = __ldg(reinterpret_cast<const float4*>(x1) + b * 32 + l);
            a[b][2] = __ldg(reinterpret_cast<const float4*>(x2) + b * 32 + l);
            a[b][3] = __ldg(reinterpret_cast<const float4*>(x3) + b * 32 + l);
        }
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            const float4 q = qreg[b];
#pragma unroll
            for (int j = 0; j < 4; ++j) s[j] = fma((double)a[b][j].x, (double)q.x, s[j]);
#pragma unroll
            for (int j = 0; j < 4; ++j) s[j] = fma((double)a[b][j].y, (double)q.y, s[j]);
#pragma unroll
            for (int j = 0; j < 4; ++j) s[j] = fma((double)a[b][j].z, (double)q.z, s[j]);
#pragma unroll
            for (int j = 0; j < 4; ++j) s[j] = fma((double)a[b][j].w, (double)q.w, s[j]);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) s[j] = warp_sum(s[j]);
        return;
    }
#pragma unroll 1
    for (int b = 0; b < nblk; ++b) {
        const float4 a0 = __ldg(reinterpret_cast<const float4*>(x0) + b * 32 + l);
        const float4 a1 = __ldg(reinterpret_cast<const float4*>(x1) + b * 32 + l);
        const float4 a2 = __ldg(reinterpret_cast<const float4*>(x2) + b * 32 + l);
        const float4 a3 = __ldg(reinterpret_cast<const float4*>(x3) + b * 32 + l);
        const float4 q = b == 0 ? qreg[0] : (b == 1 ? qreg[1] : (b == 2 ? qreg[2] : qreg[3]));
        s[0] = fma((double)a0.x, (double)q.x, s[0]); s[1] = fma((double)a1.x, (double)q.x, s[1]);
        s[2] = fma((double)a2.x, (double)q.x, s[2]); s[3] = fma((double)a3.x, (double)q.x, s[3]);
        s[0] = fma((double)a0.y, (double)q.y, s[0]); s[1] = fma((double)a1.y, (double)q.y, s[1]);
        s[2] = fma((double)a2.y, (double)q.y, s[2]); s[3] = fma((double)a3.y, (double)q.y, s[3]);
        s[0] = fma((double)a0.z, (double)q.z, s[0]); s[1] = fma((double)a1.z, (double)q.z, s[1]);
        s[2] = fma((double)a2.z, (double)q.z, s[2]); s[3] = fma((double)a3.z, (double)q.z, s[3]);
        s[0] = fma((double)a0.w, (double)q.w, s[0]); s[1] = fma((double)a1.w, (double)q.w, s[1]);
        s[2] = fma((double)a2.w, (double)q.w, s[2]); s[3] = fma((double)a3.w, (double)q.w, s[3]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) s[j] = warp_sum(s[j]);
}

// Two launch shapes share this kernel.  SMALL: 128 threads, a list of at most `small_keys` survivors sorted in
// 12-24 KB of shared memory, six CTAs per SM -- the normal case (about 1.5 k survivors per query), where a
// 512-thread CTA with shared memory for a full-capacity list spent its time in barriers and left the row gathers
// (random 512-byte reads, latency-bound) with too little in flight.  BIG: the full-capacity shape, launched behind
// the small one for the queries it skipped (heavily tied lists); it exits at once for everyone else.
template <bool EXACT, bool SMALL>
__global__ void __launch_bounds__(SMALL ? kFinalSmallThreads : kFinalThreads, SMALL ? 6 : 2)
final_kernel(const unsigned long long* __restrict__ cand, const unsigned int* __restrict__ cnt, unsigned int cap,
             int K, const float* __restrict__ table, int d, const float* __restrict__ queries, long long idx_offset,
             long long* __restrict__ out_idx, float* __restrict__ out_score, double* __restrict__ out_score64,
             unsigned int* needed_cap, unsigned int* __restrict__ ovf, int small_keys,
             const __grid_constant__ PeerDst peer) {
    extern __shared__ unsigned long long sm[];
    const int q = blockIdx.x;
    unsigned int c_raw = cnt[q];
    // which of the two launches owns this query (small_keys == 0: there is only one launch)
    if (small_keys > 0 && ((c_raw > (unsigned int)small_keys) == SMALL)) return;
    if (c_raw > cap) {
        if (threadIdx.x == 0) { atomicMax(needed_cap, c_raw); ovf[q] = 1u; }
        c_raw = cap;
    }
    // a list that overflowed in ANY round of this call may have lost a true neighbour: the whole row is marked
    // (row -2 in column 0) so the caller can tell which queries to re-run with a larger cap / exhaustively
    const bool overflowed = ovf[q] != 0u || cnt[q] > cap;
    const int c = (int)c_raw;
    const int P = next_pow2(c > 1 ? c : 2);
    unsigned long long* sk = sm;                                   // [P] ordered score (fp64 or fp32 image)
    uint32_t* rw = reinterpret_cast<uint32_t*>(sm + P);            // [P] row
    const unsigned long long* list = cand + (size_t)q * cap;
    if (EXACT) {
        const float* qv = queries + (size_t)q * d;
        const int warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
        const bool vec = (d % 128 == 0) && d <= 512 && ((reinterpret_cast<uintptr_t>(table) | reinterpret_cast<uintptr_t>(qv)) & 15u) == 0;
        if (vec) {
            float4 qreg[4];
            const int nblk = d / 128;
#pragma unroll
            for (int b = 0; b < 4; ++b)
                qreg[b] = b < nblk ? __ldg(reinterpret_cast<const float4*>(qv) + b * 32 + lane_id()) : make_float4(0.f, 0.f, 0.f, 0.f);
            for (int i = 4 * warp; SMALL && i < c; i += 4 * nwarps) {
                unsigned long long kx[4];
                const float* xp[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    kx[j] = (i + j < c) ? list[i + j] : 0ull;
                    xp[j] = table + (size_t)(kx[j] != 0ull ? key_row(kx[j]) : 0u) * d;
                }
                double sx[4];
                exact_dot4_v4(xp[0], xp[1], xp[2], xp[3], qreg, nblk, sx);
                if (lane_id() == 0) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        if (i + j < c) {
                            const bool ok = kx[j] != 0ull && sx[j] == sx[j];
                            sk[i + j] = ok ? f64_to_ordered(sx[j]) : 0ull;
                            rw[i + j] = ok ? key_row(kx[j]) : 0xffffffffu;
                        }
                    }
                }
            }
            for (int i = 2 * warp; !SMALL && i < c; i += 2 * nwarps) {
                const unsigned long long ka = list[i];
                const unsigned long long kb = (i + 1 < c) ? list[i + 1] : 0ull;
                const uint32_t ra = key_row(ka), rb = key_row(kb);
                double sa, sb;
                exact_dot2_v4(table + (size_t)(ka != 0ull ? ra : 0u) * d, table + (size_t)(kb != 0ull ? rb : 0u) * d, qreg,
                              nblk, sa, sb);
                if (lane_id() == 0) {
                    sk[i] = (ka != 0ull && sa == sa) ? f64_to_ordered(sa) : 0ull;
                    rw[i] = (ka != 0ull && sa == sa) ? ra : 0xffffffffu;
                    if (i + 1 < c) {
                        sk[i + 1] = (kb != 0ull && sb == sb) ? f64_to_ordered(sb) : 0ull;
                        rw[i + 1] = (kb != 0ull && sb == sb) ? rb : 0xffffffffu;
                    }
                }
            }
        } else {
            for (int i = warp; i < c; i += nwarps) {
                const unsigned long long k = list[i];
                const uint32_t row = key_row(k);
                const double s = (k != 0ull) ? exact_dot(table + (size_t)row * d, qv, d) : 0.0;
                if (lane_id() == 0) {
                    sk[i] = (k != 0ull && s == s) ? f64_to_ordered(s) : 0ull;
                    rw[i] = (k != 0ull && s == s) ? row : 0xffffffffu;
                }
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
    if (peer.world > 0) {
        // sharded catalogue: this query's local result goes straight to the GPU that merges it (NVLink peer
        // stores, coalesced 8-byte words); that GPU learns about it from the flag published after this kernel
        const long long qg = peer.q0 + q;
        const int owner = (int)(qg / peer.q_per_owner);
        const size_t o = ((size_t)peer.rank * peer.q_cap + (size_t)(qg - (long long)owner * peer.q_per_owner)) * peer.k_cap;
        double* xs = peer.xs[owner] + o;
        long long* xi = peer.xi[owner] + o;
        for (int i = threadIdx.x; i < K; i += blockDim.x) {
            const bool ok = i < c && rw[i] != 0xffffffffu;
            const double s = EXACT ? ordered_to_f64(sk[i]) : (double)ordered_to_f32((uint32_t)sk[i]);
            xs[i] = ok ? s : -INFINITY;
            xi[i] = ok ? (long long)rw[i] + idx_offset : -1;
        }
        return;
    }
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
        if (overflowed && i == 0) out_idx[o] = -2;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// final, ONE WARP per query: the shape for the normal case (k <= 512, a few hundred survivors per query).  A CTA per
// query spends its life in barriers (list load -> gathers -> 36-phase bitonic sort -> store) with at most six
// queries in flight per SM; a warp per query needs no barrier at all, ~48 queries per SM are in flight at once, and
// the whole batch finishes in about the latency of one query.  Eight row gathers are in flight per warp; per
// candidate the fma sequence is that of exact_dot2_v4 / exact_dot, so every final shape returns the same bits.
constexpr int kFinalWarpQ = 4;          // warps (queries) per CTA
constexpr int kFinalWarpKeys = 1024;    // longest list a warp sorts (12 KB of shared memory per warp)

template <int NC>
__device__ __forceinline__ void exact_dotn_v4(const float* const (&x)[NC], const float4 (&qreg)[4], int nblk,
                                              double (&s)[NC]) {
#pragma unroll
    for (int j = 0; j < NC; ++j) s[j] = 0.0;
    const int l = lane_id();
#pragma unroll 1
    for (int b = 0; b < nblk; ++b) {
        float4 a[NC];
#pragma unroll
        for (int j = 0; j < NC; ++j) a[j] = __ldg(reinterpret_cast<const float4*>(x[j]) + b * 32 + l);
        const float4 q = b == 0 ? qreg[0] : (b == 1 ? qreg[1] : (b == 2 ? qreg[2] : qreg[3]));
#pragma unroll
        for (int j = 0; j < NC; ++j) {
            s[j] = fma((double)a[j].x, (double)q.x, s[j]);
            s[j] = fma((double)a[j].y, (double)q.y, s[j]);
            s[j] = fma((double)a[j].z, (double)q.z, s[j]);
            s[j] = fma((double)a[j].w, (double)q.w, s[j]);
        }
    }
#pragma unroll
    for (int j = 0; j < NC; ++j) s[j] = warp_sum(s[j]);
}

template <bool EXACT>
__global__ void __launch_bounds__(kFinalWarpQ * 32, 4)
final_warp_kernel(const unsigned long long* __restrict__ cand, const unsigned int* __restrict__ cnt, unsigned int cap,
                  int B, int K, const float* __restrict__ table, int d, const float* __restrict__ queries,
                  long long idx_offset, long long* __restrict__ out_idx, float* __restrict__ out_score,
                  double* __restrict__ out_score64, unsigned int* needed_cap, unsigned int* __restrict__ ovf,
                  const __grid_constant__ PeerDst peer) {
    __shared__ unsigned long long sk_all[kFinalWarpQ][kFinalWarpKeys];
    __shared__ uint32_t rw_all[kFinalWarpQ][kFinalWarpKeys];
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    const int q = blockIdx.x * kFinalWarpQ + w;
    if (q >= B) return;
    unsigned int c_raw = cnt[q];
    // longer lists belong to final_kernel<EXACT, false>, launched behind this kernel whenever the capacity allows them
    if (c_raw > (unsigned int)kFinalWarpKeys && cap > (unsigned int)kFinalWarpKeys) return;
    if (c_raw > cap) {
        if (l == 0) { atomicMax(needed_cap, c_raw); ovf[q] = 1u; }
        c_raw = cap;
        __syncwarp();
    }
    const int c = (int)c_raw;
    const int P = next_pow2(c > 1 ? c : 2);
    unsigned long long* sk = sk_all[w];
    uint32_t* rw = rw_all[w];
    const unsigned long long* list = cand + (size_t)q * cap;
    if (EXACT) {
        const float* qv = queries + (size_t)q * d;
        const bool vec = (d % 128 == 0) && d <= 512 && ((reinterpret_cast<uintptr_t>(table) | reinterpret_cast<uintptr_t>(qv)) & 15u) == 0;
        if (vec) {
            float4 qreg[4];
            const int nblk = d / 128;
#pragma unroll
            for (int b = 0; b < 4; ++b)
                qreg[b] = b < nblk ? __ldg(reinterpret_cast<const float4*>(qv) + b * 32 + l) : make_float4(0.f, 0.f, 0.f, 0.f);
            for (int i = 0; i < c; i += 8) {
                // lane j < 8 fetches key i + j; everyone learns the eight rows by shuffle
                const unsigned long long mine = (l < 8 && i + l < c) ? list[i + l] : 0ull;
                unsigned long long kx[8];
                const float* xp[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    kx[j] = __shfl_sync(0xffffffffu, mine, j);
                    xp[j] = table + (size_t)(kx[j] != 0ull ? key_row(kx[j]) : 0u) * d;
                }
                double sx[8];
                exact_dotn_v4<8>(xp, qreg, nblk, sx);
                if (l < 8 && i + l < c) {
                    double sv = sx[0];
#pragma unroll
                    for (int j = 1; j < 8; ++j) sv = l == j ? sx[j] : sv;
                    const bool ok = mine != 0ull && sv == sv;
                    sk[i + l] = ok ? f64_to_ordered(sv) : 0ull;
                    rw[i + l] = ok ? key_row(mine) : 0xffffffffu;
                }
            }
        } else {
            for (int i = 0; i < c; ++i) {
                const unsigned long long k = list[i];
                const uint32_t row = key_row(k);
                const double sv = (k != 0ull) ? exact_dot(table + (size_t)row * d, qv, d) : 0.0;
                if (l == 0) {
                    sk[i] = (k != 0ull && sv == sv) ? f64_to_ordered(sv) : 0ull;
                    rw[i] = (k != 0ull && sv == sv) ? row : 0xffffffffu;
                }
            }
        }
        for (int i = c + l; i < P; i += 32) { sk[i] = 0ull; rw[i] = 0xffffffffu; }
    } else {
        for (int i = l; i < P; i += 32) {
            const unsigned long long k = (i < c) ? list[i] : 0ull;
            sk[i] = k >> 32;
            rw[i] = (k != 0ull) ? key_row(k) : 0xffffffffu;
        }
    }
    __syncwarp();
    // descending bitonic sort by (score, then row ascending) -- one warp, no block barrier
    for (int kk = 2; kk <= P; kk <<= 1) {
        for (int j = kk >> 1; j > 0; j >>= 1) {
            for (int i = l; i < P; i += 32) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const bool i_first = sk[i] > sk[ixj] || (sk[i] == sk[ixj] && rw[i] < rw[ixj]);
                    if (((i & kk) == 0) ? !i_first : i_first) {
                        const unsigned long long t = sk[i]; sk[i] = sk[ixj]; sk[ixj] = t;
                        const uint32_t r = rw[i]; rw[i] = rw[ixj]; rw[ixj] = r;
                    }
                }
            }
            __syncwarp();
        }
    }
    const bool overflowed = ovf[q] != 0u || cnt[q] > cap;
    if (peer.world > 0) {
        // sharded catalogue: straight into the exchange buffer of the GPU that merges this query (see final_kernel)
        const long long qg = peer.q0 + q;
        const int owner = (int)(qg / peer.q_per_owner);
        const size_t o = ((size_t)peer.rank * peer.q_cap + (size_t)(qg - (long long)owner * peer.q_per_owner)) * peer.k_cap;
        double* xs = peer.xs[owner] + o;
        long long* xi = peer.xi[owner] + o;
        for (int i = l; i < K; i += 32) {
            const bool ok = i < c && rw[i] != 0xffffffffu;
            const double sv = EXACT ? ordered_to_f64(sk[i]) : (double)ordered_to_f32((uint32_t)sk[i]);
            xs[i] = ok ? sv : -INFINITY;
            xi[i] = ok ? (long long)rw[i] + idx_offset : -1;
        }
        return;
    }
    for (int i = l; i < K; i += 32) {
        const size_t o = (size_t)q * K + i;
        const bool ok = i < c && rw[i] != 0xffffffffu;
        const double sv = ok ? (EXACT ? ordered_to_f64(sk[i]) : (double)ordered_to_f32((uint32_t)sk[i])) : -INFINITY;
        out_idx[o] = ok ? (long long)rw[i] + idx_offset : -1;
        out_score[o] = (float)sv;
        if (out_score64) out_score64[o] = sv;
        if (overflowed && i == 0) out_idx[o] = -2;
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

// ---------------------------------------------------------------------------------------------------------------
// Exhaustive exact search of one query: the answer of last resort (see kernels.h).  Scores every row with the same
// fp64 arithmetic final_kernel uses for that table width (so results are bit-identical to the filtered path),
// then a stable descending radix sort of (score image, row): equal scores stay in ascending row order.
__global__ void __launch_bounds__(256)
bruteforce_score_kernel(const float* __restrict__ table, long long n, int d, const float* __restrict__ qv,
                        unsigned long long* __restrict__ keys, unsigned int* __restrict__ rows) {
    const bool vec = (d % 128 == 0) && d <= 512 && ((reinterpret_cast<uintptr_t>(table) | reinterpret_cast<uintptr_t>(qv)) & 15u) == 0;
    const long long warps = (long long)gridDim.x * 8;
    const long long w0 = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (vec) {
        float4 qreg[4];
        const int nblk = d / 128;
#pragma unroll
        for (int b = 0; b < 4; ++b)
            qreg[b] = b < nblk ? __ldg(reinterpret_cast<const float4*>(qv) + b * 32 + lane_id()) : make_float4(0.f, 0.f, 0.f, 0.f);
        for (long long r = 2 * w0; r < n; r += 2 * warps) {
            const long long r2 = r + 1 < n ? r + 1 : r;
            double sa, sb;
            exact_dot2_v4(table + (size_t)r * d, table + (size_t)r2 * d, qreg, nblk, sa, sb);
            if (lane_id() == 0) {
                keys[r] = sa == sa ? f64_to_ordered(sa) : 0ull;
                rows[r] = (unsigned int)r;
                if (r + 1 < n) { keys[r + 1] = sb == sb ? f64_to_ordered(sb) : 0ull; rows[r + 1] = (unsigned int)(r + 1); }
            }
        }
    } else {
        for (long long r = w0; r < n; r += warps) {
            const double s = exact_dot(table + (size_t)r * d, qv, d);
            if (lane_id() == 0) { keys[r] = s == s ? f64_to_ordered(s) : 0ull; rows[r] = (unsigned int)r; }
        }
    }
}

__global__ void bruteforce_emit_kernel(const unsigned long long* __restrict__ keys, const unsigned int* __restrict__ rows,
                                       int K, long long idx_offset, long long* __restrict__ out_idx,
                                       float* __restrict__ out_score, double* __restrict__ out_score64) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= K) return;
    const unsigned long long kk = keys[i];
    const double s = kk ? ordered_to_f64(kk) : -INFINITY;
    out_idx[i] = kk ? (long long)rows[i] + idx_offset : -1;
    out_score[i] = (float)s;
    if (out_score64) out_score64[i] = s;
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
                                  unsigned int* ovf, const SelExchange* sxp, int dense_warp, cudaStream_t stream) {
    if (B <= 0) return cudaSuccess;
    SelExchange sx;
    if (sxp) sx = *sxp; else memset(&sx, 0, sizeof sx);
    // large batches: one warp per query -- the short lists of the filter rounds, and (dense_warp) the dense round's
    // too: thousands of independent warps re-reading a 32 KB list through L2 beat a CTA per query in barriers
    // (a warp holds up to 1024 keys in registers: lists of k (1 + 1.4 growth) keys with k in the hundreds fit; top-1000
    // lists do not -- their warps fell back to re-reading the list from L2 in four radix passes, 9 ms of a 65 ms C5
    // step -- so large k takes the CTA-per-query kernel, whose registers hold up to 16 K keys)
    if (B >= 128 && (fixed_count < 0 || dense_warp) && K <= 256) {
        select_compact_warp_kernel<<<(B + kSelWarps - 1) / kSelWarps, kSelWarps * 32, 0, stream>>>(
            cand, cnt, cap, B, K, fixed_count, margin, thr, needed_cap, ovf, sx);
        return cudaGetLastError();
    }
    // stage only what can be there: a dense round holds fixed_count keys, a filter round at most cap
    const size_t n = fixed_count >= 0 ? (size_t)fixed_count : (size_t)cap;
    static const bool radix_only = getenv("HWER_SELECT_RADIX") != nullptr;      // A/B knob, read once
    if (!radix_only && n <= (size_t)kSelThreads * 64) {
        // keys in registers + bisection (bit-identical thresholds); KPT sized to what the list can hold
#define HWER_SEL_BISECT(KPT)                                                                                          \
        select_compact_bisect_kernel<KPT><<<B, kSelThreads, 0, stream>>>(cand, cnt, cap, K, fixed_count, margin, thr, \
                                                                         needed_cap, ovf, sx)
        if (n <= (size_t)kSelThreads * 8) HWER_SEL_BISECT(8);
        else if (n <= (size_t)kSelThreads * 16) HWER_SEL_BISECT(16);
        else if (n <= (size_t)kSelThreads * 32) HWER_SEL_BISECT(32);
        else HWER_SEL_BISECT(64);
#undef HWER_SEL_BISECT
        return cudaGetLastError();
    }
    const size_t smem = (n < 1024 ? 1024 : n) * sizeof(unsigned long long);
    cudaError_t e = set_smem(select_compact_kernel, smem);
    if (e != cudaSuccess) return e;
    select_compact_kernel<<<B, kSelThreads, smem, stream>>>(cand, cnt, cap, K, fixed_count, margin, thr, needed_cap, ovf,
                                                            sx);
    return cudaGetLastError();
}

cudaError_t launch_final(const unsigned long long* cand, const unsigned int* cnt, unsigned int cap, int B, int K,
                         int exact, const float* table, int d, const float* queries, long long idx_offset,
                         long long* out_idx, float* out_score, double* out_score64, unsigned int* needed_cap,
                         unsigned int* ovf, const PeerDst* peer, int allow_small, cudaStream_t stream) {
    if (B <= 0) return cudaSuccess;
    PeerDst pd;
    if (peer) pd = *peer; else memset(&pd, 0, sizeof pd);
    const size_t per_key = sizeof(unsigned long long) + sizeof(uint32_t);
    // lists of up to small_keys survivors (the normal case) go to the small shape; if that already covers the
    // capacity it is the only launch
    size_t small_keys = pow2_ge((size_t)(2 * K > 1024 ? 2 * K : 1024));
    const bool only_small = small_keys >= pow2_ge(cap);
    if (only_small) small_keys = pow2_ge(cap);
    const size_t smem_small = small_keys * per_key, smem_big = pow2_ge(cap) * per_key;
    const bool use_small = allow_small && smem_small <= 24 * 1024;
    cudaError_t e;
    if (allow_small == 1 && 2 * K <= kFinalWarpKeys) {
        // the normal case: a warp per query for lists of up to kFinalWarpKeys survivors, then the full-capacity CTA
        // shape for whatever is longer (it exits at once for every query the warp kernel answered)
        const int grid = (B + kFinalWarpQ - 1) / kFinalWarpQ;
        if (exact)
            final_warp_kernel<true><<<grid, kFinalWarpQ * 32, 0, stream>>>(cand, cnt, cap, B, K, table, d, queries, idx_offset,
                                                                        out_idx, out_score, out_score64, needed_cap, ovf, pd);
        else
            final_warp_kernel<false><<<grid, kFinalWarpQ * 32, 0, stream>>>(cand, cnt, cap, B, K, table, d, queries, idx_offset,
                                                                         out_idx, out_score, out_score64, needed_cap, ovf, pd);
        if (cap <= (unsigned int)kFinalWarpKeys) return cudaGetLastError();      // no list can be longer
        const int skip = kFinalWarpKeys < (int)cap ? kFinalWarpKeys : (int)cap;
        if (exact) {
            e = set_smem(final_kernel<true, false>, smem_big);
            if (e != cudaSuccess) return e;
            final_kernel<true, false><<<B, kFinalThreads, smem_big, stream>>>(cand, cnt, cap, K, table, d, queries, idx_offset,
                                                                           out_idx, out_score, out_score64, needed_cap, ovf, skip, pd);
        } else {
            e = set_smem(final_kernel<false, false>, smem_big);
            if (e != cudaSuccess) return e;
            final_kernel<false, false><<<B, kFinalThreads, smem_big, stream>>>(cand, cnt, cap, K, table, d, queries, idx_offset,
                                                                            out_idx, out_score, out_score64, needed_cap, ovf, skip, pd);
        }
        return cudaGetLastError();
    }
#define HWER_LAUNCH_FINAL(EX)                                                                                          \
    do {                                                                                                               \
        if (use_small) {                                                                                               \
            final_kernel<EX, true><<<B, kFinalSmallThreads, smem_small, stream>>>(                                     \
                cand, cnt, cap, K, table, d, queries, idx_offset, out_idx, out_score, out_score64, needed_cap, ovf,   \
                only_small ? 0 : (int)small_keys, pd);                                                                 \
            if (only_small) break;                                                                                     \
        }                                                                                                              \
        e = set_smem(final_kernel<EX, false>, smem_big);                                                               \
        if (e != cudaSuccess) return e;                                                                                \
        final_kernel<EX, false><<<B, kFinalThreads, smem_big, stream>>>(                                               \
            cand, cnt, cap, K, table, d, queries, idx_offset, out_idx, out_score, out_score64, needed_cap, ovf,       \
            use_small ? (int)small_keys : 0, pd);                                                                      \
    } while (0)
    if (exact) HWER_LAUNCH_FINAL(true); else HWER_LAUNCH_FINAL(false);
#undef HWER_LAUNCH_FINAL
    return cudaGetLastError();
}

namespace {
struct BruteLayout { size_t keys_in, keys_out, rows_in, rows_out, temp, temp_bytes, total; };
BruteLayout brute_layout(long long n) {
    BruteLayout L;
    size_t tb = 0;
    cub::DeviceRadixSort::SortPairsDescending(nullptr, tb, (const unsigned long long*)nullptr, (unsigned long long*)nullptr,
                                              (const unsigned int*)nullptr, (unsigned int*)nullptr, (int)n);
    auto al = [](size_t x) { return (x + 255) / 256 * 256; };
    size_t off = 0;
    L.keys_in = off; off += al((size_t)n * 8);
    L.keys_out = off; off += al((size_t)n * 8);
    L.rows_in = off; off += al((size_t)n * 4);
    L.rows_out = off; off += al((size_t)n * 4);
    L.temp = off; off += al(tb);
    L.temp_bytes = tb;
    L.total = off;
    return L;
}
}  // namespace

size_t bruteforce_scratch_bytes(long long n) { return brute_layout(n).total; }

cudaError_t launch_bruteforce_topk(const float* table, long long n, int d, const float* query, int k,
                                   long long idx_offset, long long* out_idx, float* out_score, double* out_score64,
                                   void* scratch, size_t scratch_bytes, cudaStream_t stream) {
    const BruteLayout L = brute_layout(n);
    if (scratch_bytes < L.total || n >= (1LL << 31)) return cudaErrorInvalidValue;
    unsigned char* b = static_cast<unsigned char*>(scratch);
    unsigned long long* keys_in = reinterpret_cast<unsigned long long*>(b + L.keys_in);
    unsigned long long* keys_out = reinterpret_cast<unsigned long long*>(b + L.keys_out);
    unsigned int* rows_in = reinterpret_cast<unsigned int*>(b + L.rows_in);
    unsigned int* rows_out = reinterpret_cast<unsigned int*>(b + L.rows_out);
    long long blocks = (n + 15) / 16;
    if (blocks > 148LL * 8) blocks = 148LL * 8;
    bruteforce_score_kernel<<<(int)blocks, 256, 0, stream>>>(table, n, d, query, keys_in, rows_in);
    size_t tb = L.temp_bytes;
    cudaError_t e = cub::DeviceRadixSort::SortPairsDescending(b + L.temp, tb, keys_in, keys_out, rows_in, rows_out, (int)n,
                                                              0, 64, stream);
    if (e != cudaSuccess) return e;
    bruteforce_emit_kernel<<<(k + 255) / 256, 256, 0, stream>>>(keys_out, rows_out, k, idx_offset, out_idx, out_score,
                                                                out_score64);
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
