// Multi-GPU result exchange over NVLink / NVSwitch peer memory.
//
// The catalogue is row-sharded: every GPU holds the exact top-k of ITS rows for all B queries.  Top-k of a union
// is the top-k of the per-shard top-ks, so one exchange step finishes the search.  Instead of an all-gather that
// lands all G x B x k candidates on every GPU, the queries are split into G owner ranges:
//
//   phase 0  final_kernel (select.cu) stores query q's k (score, row) pairs straight into the exchange buffer of
//            owner(q) -- plain coalesced peer stores, issued by the kernel that produces the values; a one-warp
//            kernel then publishes flag[0][rank] = epoch on every peer (release, system scope).
//   phase 1  exchange_merge_kernel on the owner spins (acquire, system scope) until all G sources have
//            published, merges its B/G queries (G*k keys each) under the (score desc, row asc) rule, stores the k
//            winners into the result area of EVERY rank (peer stores again) and the last CTA publishes
//            flag[1][rank] = epoch everywhere.
//   phase 2  exchange_collect_kernel waits for the G owner flags and copies the [B, k] result to the caller.
//
// Per GPU and step that is B*k*16 bytes out and in (6.5 MB at B = 4096, k = 100) instead of G times as much, no
// NCCL launch on the data path, and every wait is a few-microsecond spin on local memory.  Buffers are reused
// every step: a rank can only start step n+1 after it has seen every owner's phase-1 flag of step n, and an owner
// only publishes that flag after it has finished reading its exchange buffer, so no rank can overwrite data
// another one still needs (DESIGN.md "Multi-GPU").  Epochs increase monotonically; flags are never reset.
//
// No reference counterpart: the reference is single-process.
#include "common.cuh"
#include "kernels.h"

namespace hwer {

namespace {

constexpr int kMergeThreads = 256;
constexpr long long kSpinTimeoutCycles = 6000000000ll;   // ~3 s: a dead peer must not hang the GPU

__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// flags layout per rank (unsigned int[64]): [0..7] phase-0 epoch per source rank, [8..15] phase-1 epoch per owner,
// [16] merge CTA counter, [17] error (a wait timed out), [40..47] candidate-capacity demand of each source rank
__device__ bool wait_epoch(const unsigned int* flag, unsigned int epoch, unsigned int* err) {
    const long long t0 = clock64();
    while ((int)(ld_acquire_sys(flag) - epoch) < 0) {
        if (clock64() - t0 > kSpinTimeoutCycles) {
            atomicExch(err, 1u);
            return false;
        }
        __nanosleep(100);
    }
    return true;
}

__global__ void exchange_signal_kernel(const __grid_constant__ ExchangeView v, int phase, unsigned int epoch,
                                       const unsigned int* __restrict__ needed) {
    // this rank's candidate-capacity demand rides with the flag: every rank learns the largest one from its own
    // memory (exchange_collect_kernel), so an overflow retry is agreed without a host-side collective
    if (needed && (int)threadIdx.x < v.world) v.flags[threadIdx.x][40 + v.rank] = *needed;
    // everything this stream wrote before (the final kernels' peer stores) is ordered before the flags
    __threadfence_system();
    if ((int)threadIdx.x < v.world) st_release_sys(v.flags[threadIdx.x] + phase * kMaxPeers + v.rank, epoch);
}

__device__ __forceinline__ int next_pow2_i(int x) {
    int p = 2;
    while (p < x) p <<= 1;
    return p;
}

__global__ void __launch_bounds__(kMergeThreads)
exchange_merge_kernel(const __grid_constant__ ExchangeView v, int B, int K, unsigned int epoch, int owned) {
    extern __shared__ unsigned long long sm[];
    __shared__ int ok_s;
    unsigned int* myflags = v.flags[v.rank];
    if (threadIdx.x == 0) ok_s = 1;
    __syncthreads();
    if ((int)threadIdx.x < v.world && !wait_epoch(myflags + threadIdx.x, epoch, myflags + 17)) ok_s = 0;
    __syncthreads();
    const int ql = blockIdx.x;                                  // query local to this owner
    const long long qg = (long long)v.rank * v.q_per_owner + ql;
    const int c = v.world * K;
    const int P = next_pow2_i(c);
    unsigned long long* sk = sm;
    long long* id = reinterpret_cast<long long*>(sm + P);
    if (ok_s) {
        for (int i = threadIdx.x; i < P; i += blockDim.x) {
            unsigned long long s = 0ull;
            long long ix = 0x7fffffffffffffffll;
            if (i < c) {
                const int g = i / K, j = i - g * K;
                const size_t o = ((size_t)g * v.q_cap + ql) * v.k_cap + j;
                const long long r = v.xi[v.rank][o];
                if (r >= 0) { s = f64_to_ordered(v.xs[v.rank][o]); ix = r; }
            }
            sk[i] = s;
            id[i] = ix;
        }
        __syncthreads();
        for (int k = 2; k <= P; k <<= 1) {
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int i = threadIdx.x; i < P; i += blockDim.x) {
                    const int ixj = i ^ j;
                    if (ixj > i) {
                        const bool desc = (i & k) == 0;
                        const bool i_first = sk[i] > sk[ixj] || (sk[i] == sk[ixj] && id[i] < id[ixj]);
                        if (desc ? !i_first : i_first) {
                            unsigned long long t = sk[i]; sk[i] = sk[ixj]; sk[ixj] = t;
                            long long r = id[i]; id[i] = id[ixj]; id[ixj] = r;
                        }
                    }
                }
                __syncthreads();
            }
        }
        // deliver the k winners of this query to every rank's result area (owned: this rank keeps its queries'
        // results to itself -- HWER_PHASE_OWNED -- and nothing crosses NVLink a second time)
        const int r_first = owned ? v.rank : 0, r_count = owned ? 1 : v.world;
        for (int i = threadIdx.x; i < K * r_count; i += blockDim.x) {
            const int r = r_first + i / K, j = i - (i / K) * K;
            const bool ok = id[j] != 0x7fffffffffffffffll;
            const double s = ok ? ordered_to_f64(sk[j]) : -INFINITY;
            const size_t o = (size_t)qg * K + j;
            v.out_idx[r][o] = ok ? id[j] : -1;
            v.out_score64[r][o] = s;
            v.out_score[r][o] = (float)s;
        }
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int prev = atomicAdd(myflags + 16, 1u);
        if (prev == gridDim.x - 1) {          // last CTA of this owner: everything is delivered
            myflags[16] = 0u;
            __threadfence_system();
            for (int r = 0; r < v.world; ++r) st_release_sys(v.flags[r] + kMaxPeers + v.rank, epoch);
        }
    }
}

__global__ void __launch_bounds__(256)
exchange_collect_kernel(const __grid_constant__ ExchangeView v, int B, int K, unsigned int epoch,
                        long long* __restrict__ out_idx, float* __restrict__ out_score,
                        double* __restrict__ out_score64, unsigned int* __restrict__ status2, int owned) {
    unsigned int* myflags = v.flags[v.rank];
    if ((int)threadIdx.x < v.world) wait_epoch(myflags + kMaxPeers + threadIdx.x, epoch, myflags + 17);
    __syncthreads();
    if (status2 && blockIdx.x == 0 && (int)threadIdx.x < v.world) {
        // every source's phase-0 flag (and with it its capacity demand) arrived long before its owners finished
        if (wait_epoch(myflags + threadIdx.x, epoch, myflags + 17)) atomicMax(status2, __ldcv(myflags + 40 + threadIdx.x));
        if (__ldcv(myflags + 17)) atomicMax(status2 + 1, 1u);
    }
    // owned: only the rows this rank merged, [rank * q_per_owner, ...), packed at the front of out_*.  The wait for
    // every owner's flag above stays: it is what keeps a fast rank from overwriting exchange buffers an owner still reads
    long long own = (long long)B - (long long)v.rank * v.q_per_owner;
    own = own > v.q_per_owner ? v.q_per_owner : (own < 0 ? 0 : own);
    const size_t n = owned ? (size_t)own * K : (size_t)B * K;
    const size_t o0 = owned ? (size_t)v.rank * v.q_per_owner * K : 0;
    const long long* si = v.out_idx[v.rank] + o0;
    const float* ss = v.out_score[v.rank] + o0;
    const double* sd = v.out_score64[v.rank] + o0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        out_idx[i] = __ldcv(si + i);          // written by peers: never from a stale cache line
        out_score[i] = __ldcv(ss + i);
        if (out_score64) out_score64[i] = __ldcv(sd + i);
    }
}

}  // namespace

// CUDA loads a kernel's code on its first launch (lazy module loading), and that load can wait for the device to go
// idle.  A host thread that first launches an exchange kernel while a select kernel of the same process is already
// spinning on a peer (one process driving several ranks, or a peer that is late) would stall behind that spin, so
// every kernel of this file is loaded when the exchange object is created.
cudaError_t exchange_preload() {
    cudaFuncAttributes a;
    cudaError_t e = cudaFuncGetAttributes(&a, exchange_signal_kernel);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, exchange_merge_kernel);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, exchange_collect_kernel);
    return e;
}

cudaError_t launch_exchange_signal(const ExchangeView& v, int phase, unsigned int epoch, const unsigned int* needed,
                                   cudaStream_t stream) {
    exchange_signal_kernel<<<1, 32, 0, stream>>>(v, phase, epoch, needed);
    return cudaGetLastError();
}

cudaError_t launch_exchange_merge(const ExchangeView& v, int B, int K, unsigned int epoch, int owned,
                                  cudaStream_t stream) {
    long long own = (long long)B - (long long)v.rank * v.q_per_owner;
    if (own > v.q_per_owner) own = v.q_per_owner;
    if (own <= 0) return launch_exchange_signal(v, 1, epoch, nullptr, stream);      // nothing to merge: just report in
    int P = 2;
    while (P < v.world * K) P <<= 1;
    const size_t smem = (size_t)P * 16;
    if (smem > (size_t)kSmemBudget) return cudaErrorInvalidValue;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(exchange_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    exchange_merge_kernel<<<(unsigned)own, kMergeThreads, smem, stream>>>(v, B, K, epoch, owned);
    return cudaGetLastError();
}

cudaError_t launch_exchange_collect(const ExchangeView& v, int B, int K, unsigned int epoch, long long* out_idx,
                                    float* out_score, double* out_score64, unsigned int* status2, int owned,
                                    cudaStream_t stream) {
    const size_t n = owned ? (size_t)v.q_per_owner * K : (size_t)B * K;
    int grid = (int)((n + 256 * 8 - 1) / (256 * 8));
    if (grid < 1) grid = 1;
    if (grid > 296) grid = 296;
    exchange_collect_kernel<<<grid, 256, 0, stream>>>(v, B, K, epoch, out_idx, out_score, out_score64, status2, owned);
    return cudaGetLastError();
}

}  // namespace hwer
