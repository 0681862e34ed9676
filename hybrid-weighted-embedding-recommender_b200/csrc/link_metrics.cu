// Link-prediction metrics of a scored, labelled pair set, entirely on the device.
//
//   hwer/validation.py:52-59  link_prediction_accuracy: sklearn average_precision_score(labels, scores),
//                             precision_recall_fscore_support(labels, scores >= 0.5, average='binary')
//                             and accuracy_score(labels, scores >= 0.5).
//
// average_precision_score is the step-wise area under the precision/recall curve over the DISTINCT score values
// (sklearn.metrics._ranking: AP = sum_n (R_n - R_{n-1}) P_n): sort by score descending, and at the last element e
// of every group of equal scores take precision tp_e / (e + 1) weighted by the recall gained inside the group.
// The sort is cub's radix sort over the order-preserving u32 image of the fp32 score; everything after it is
// two scans, one reduction (fp64) and a handful of element-wise kernels.  Latency/sort-bound, ~32 B of scratch
// per pair; P is 11 |E| in the reference's harness.
#include <cub/cub.cuh>

#include "common.cuh"
#include "kernels.h"

namespace hwer {

namespace {

struct MaxInt {
    __device__ __forceinline__ int operator()(int a, int b) const { return a > b ? a : b; }
};

// keys for the sort + the confusion counts at `thr` (counts: tp, fp, fn, tn)
__global__ void __launch_bounds__(256)
lm_prepare_kernel(const float* __restrict__ score, const unsigned char* __restrict__ label, int P, float thr,
                  uint32_t* __restrict__ keys, int* __restrict__ lab, unsigned long long* __restrict__ counts) {
    unsigned int c[4] = {0u, 0u, 0u, 0u};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P; i += gridDim.x * blockDim.x) {
        float s = score[i];
        if (s == 0.0f) s = 0.0f;                       // -0.0 and +0.0 are one threshold
        const int y = label[i] != 0;
        keys[i] = f32_to_ordered(s);
        lab[i] = y;
        const int pred = s >= thr;
        c[(y ? 0 : 1) + (pred ? 0 : 2)] += 1u;         // y&pred -> tp(0), !y&pred -> fp(1), y&!pred -> fn(2), tn(3)
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        unsigned int v = c[j];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd(&counts[j], (unsigned long long)v);
    }
}

__global__ void __launch_bounds__(256)
lm_starts_kernel(const uint32_t* __restrict__ keys, int P, int* __restrict__ starts) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < P) starts[i] = (i > 0 && keys[i] != keys[i - 1]) ? i : 0;
}

// contribution of every group end to  sum_n (tp_n - tp_{n-1}) * tp_n / (n_n)   (divided by #positives at the end)
__global__ void __launch_bounds__(256)
lm_contrib_kernel(const uint32_t* __restrict__ keys, const int* __restrict__ tp, const int* __restrict__ gstart,
                  int P, double* __restrict__ contrib) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    double c = 0.0;
    if (i == P - 1 || keys[i + 1] != keys[i]) {
        const int s = gstart[i];
        const int before = s > 0 ? tp[s - 1] : 0;
        c = (double)(tp[i] - before) * ((double)tp[i] / (double)(i + 1));
    }
    contrib[i] = c;
}

__global__ void lm_finish_kernel(const double* __restrict__ ap_sum, const unsigned long long* __restrict__ counts,
                                 double* __restrict__ out) {
    if (threadIdx.x || blockIdx.x) return;
    const double tp = (double)counts[0], fp = (double)counts[1], fn = (double)counts[2], tn = (double)counts[3];
    const double pos = tp + fn, all = tp + fp + fn + tn;
    out[0] = pos > 0.0 ? *ap_sum / pos : 0.0;          // sklearn: no positives -> recall undefined, AP reported 0
    out[1] = (tp + fp) > 0.0 ? tp / (tp + fp) : 0.0;   // zero_division -> 0 (sklearn's default, with a warning)
    out[2] = pos > 0.0 ? tp / pos : 0.0;
    out[3] = all > 0.0 ? (tp + tn) / all : 0.0;
    out[4] = tp; out[5] = fp; out[6] = fn; out[7] = tn;
}

size_t up256(size_t x) { return (x + 255) / 256 * 256; }

}  // namespace

cudaError_t launch_link_metrics(const float* score, const unsigned char* label, long long P64, float thr, double* out8,
                                cudaStream_t stream) {
    const int P = (int)P64;
    // scratch layout
    size_t t_sort = 0, t_sum = 0, t_max = 0, t_red = 0;
    uint32_t* nk = nullptr; int* ni = nullptr; double* nd = nullptr;
    cudaError_t e;
    if ((e = cub::DeviceRadixSort::SortPairsDescending(nullptr, t_sort, nk, nk, ni, ni, P, 0, 32, stream))) return e;
    if ((e = cub::DeviceScan::InclusiveSum(nullptr, t_sum, ni, ni, P, stream))) return e;
    if ((e = cub::DeviceScan::InclusiveScan(nullptr, t_max, ni, ni, MaxInt(), P, stream))) return e;
    if ((e = cub::DeviceReduce::Sum(nullptr, t_red, nd, nd, P, stream))) return e;
    size_t t_cub = t_sort;
    if (t_sum > t_cub) t_cub = t_sum;
    if (t_max > t_cub) t_cub = t_max;
    if (t_red > t_cub) t_cub = t_red;
    const size_t a4 = up256((size_t)P * 4), a8 = up256((size_t)P * 8);
    const size_t total = 256 /*counts + ap*/ + 6 * a4 + a8 + up256(t_cub);
    unsigned char* base = nullptr;
    if ((e = cudaMallocAsync((void**)&base, total, stream))) return e;
    size_t off = 0;
    unsigned long long* counts = (unsigned long long*)(base + off);
    double* ap_sum = (double*)(base + off + 64); off += 256;
    uint32_t* keys_in = (uint32_t*)(base + off); off += a4;
    uint32_t* keys = (uint32_t*)(base + off); off += a4;
    int* lab_in = (int*)(base + off); off += a4;
    int* lab = (int*)(base + off); off += a4;
    int* tp = (int*)(base + off); off += a4;
    int* gstart = (int*)(base + off); off += a4;
    double* contrib = (double*)(base + off); off += a8;
    void* tmp = base + off;
    int* starts = lab_in;        // lab_in is dead after the sort
    const int threads = 256;
    const int blocks_all = (P + threads - 1) / threads;
    int blocks_grid = blocks_all < 148 * 8 ? blocks_all : 148 * 8;
    if (blocks_grid < 1) blocks_grid = 1;
    do {
        if ((e = cudaMemsetAsync(base, 0, 256, stream))) break;
        lm_prepare_kernel<<<blocks_grid, threads, 0, stream>>>(score, label, P, thr, keys_in, lab_in, counts);
        if ((e = cudaGetLastError())) break;
        size_t t = t_cub;
        if ((e = cub::DeviceRadixSort::SortPairsDescending(tmp, t, keys_in, keys, lab_in, lab, P, 0, 32, stream))) break;
        lm_starts_kernel<<<blocks_all, threads, 0, stream>>>(keys, P, starts);
        if ((e = cudaGetLastError())) break;
        t = t_cub;
        if ((e = cub::DeviceScan::InclusiveSum(tmp, t, lab, tp, P, stream))) break;
        t = t_cub;
        if ((e = cub::DeviceScan::InclusiveScan(tmp, t, starts, gstart, MaxInt(), P, stream))) break;
        lm_contrib_kernel<<<blocks_all, threads, 0, stream>>>(keys, tp, gstart, P, contrib);
        if ((e = cudaGetLastError())) break;
        t = t_cub;
        if ((e = cub::DeviceReduce::Sum(tmp, t, contrib, ap_sum, P, stream))) break;
        lm_finish_kernel<<<1, 32, 0, stream>>>(ap_sum, counts, out8);
        e = cudaGetLastError();
    } while (0);
    cudaError_t e2 = cudaFreeAsync(base, stream);
    return e ? e : e2;
}

}  // namespace hwer
