// NCF re-rank of retrieved candidates: the step right after top-k retrieval when
// the reference was trained with ncf_epochs > 0.
//
// Replaces: NCF.forward, hwer/ncf.py:7-27 (Linear/LeakyReLU stack over [h_src || h_dst], Linear(F, 1), sigmoid)
//           as called by GcnNCF.predict, hwer/gcn_ncf.py:336-361, and by the NCF branch of
//           GcnNCF.find_closest_neighbours, hwer/gcn_ncf.py:384-386.
//
// Layer widths follow ncf.py:12-16 exactly: layer i (1-based) of `depth` maps F*iw -> F*ow with
// iw = 4 if i == 2 else 2 and ow = 1 if i == depth else (4 if i == 1 else 2); GaussianNoise is the identity at
// inference (gcn.py:32-38, model.eval()).  Arithmetic is fp32 (the reference runs torch fp32 on the CPU): a
// register-tiled FFMA GEMM per layer with the gather of the two embedding rows fused into the first layer's
// operand loads and bias + LeakyReLU fused into every epilogue.  FFMA-bound: 2 * sum(in*out) FLOP per pair.
// (A tcgen05 version needs split-bf16 operands to hold the 1e-5 score tolerance; DESIGN.md lists it as next.)
#include "common.cuh"
#include "kernels.h"

namespace hwer {

namespace {

constexpr int BM = 128, BN = 128, BK = 8;      // CTA tile: 128 pairs x 128 outputs, K step 8
constexpr int kGemmThreads = 256;               // each thread owns an 8 x 8 micro-tile

struct LinearArgs {
    const float* x;            // [P, in] activations (nullptr: gather from `h`)
    const float* h;            // [n_rows, F] embedding table for the gathered first layer
    const long long* src;      // [P] rows of h (out of range -> row 0, the reference's padding row)
    const long long* dst;
    long long n_rows;
    int F;
    const float* w;            // [out, in] row-major (torch nn.Linear.weight)
    const float* b;            // [out]
    float* y;                  // [P, out]
    long long P;
    int in, out;
    float slope;               // LeakyReLU negative slope
};

__device__ __forceinline__ float4 load_a4(const LinearArgs& a, long long row, int k) {
    if (row >= a.P || k >= a.in) return make_float4(0.f, 0.f, 0.f, 0.f);
    if (a.x) return *reinterpret_cast<const float4*>(a.x + (size_t)row * a.in + k);
    long long r = k < a.F ? a.src[row] : a.dst[row];
    if (r < 0 || r >= a.n_rows) r = 0;
    return __ldg(reinterpret_cast<const float4*>(a.h + (size_t)r * a.F + (k < a.F ? k : k - a.F)));
}

__global__ void __launch_bounds__(kGemmThreads, 2)
ncf_linear_kernel(const __grid_constant__ LinearArgs a) {
    __shared__ float As[2][BK][BM + 4];
    __shared__ float Bs[2][BK][BN + 4];
    const int tid = threadIdx.x;
    const long long row0 = (long long)blockIdx.y * BM;
    const int col0 = blockIdx.x * BN;
    // loader mapping: 128 rows x 8 k per tile = 256 float4, one per thread
    const int lrow = tid >> 1, lk = (tid & 1) * 4;
    // compute mapping: thread (ty, tx) owns rows ty*8..+7, cols tx*8..+7 split in two 4-wide halves 64 apart
    const int ty = tid >> 4, tx = tid & 15;
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    auto load_b4 = [&](int k) -> float4 {
        const int o = col0 + lrow;
        if (o >= a.out || k + lk >= a.in) return make_float4(0.f, 0.f, 0.f, 0.f);
        return __ldg(reinterpret_cast<const float4*>(a.w + (size_t)o * a.in + k + lk));
    };
    auto store_tile = [&](int buf, const float4& av, const float4& bv) {
        As[buf][lk + 0][lrow] = av.x; As[buf][lk + 1][lrow] = av.y; As[buf][lk + 2][lrow] = av.z; As[buf][lk + 3][lrow] = av.w;
        Bs[buf][lk + 0][lrow] = bv.x; Bs[buf][lk + 1][lrow] = bv.y; Bs[buf][lk + 2][lrow] = bv.z; Bs[buf][lk + 3][lrow] = bv.w;
    };

    float4 av = load_a4(a, row0 + lrow, lk), bv = load_b4(0);
    store_tile(0, av, bv);
    __syncthreads();
    int buf = 0;
    for (int k0 = 0; k0 < a.in; k0 += BK) {
        const bool more = k0 + BK < a.in;
        if (more) { av = load_a4(a, row0 + lrow, k0 + BK + lk); bv = load_b4(k0 + BK); }
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float ra[8], rb[8];
            const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][64 + tx * 4]);
            ra[0] = a0.x; ra[1] = a0.y; ra[2] = a0.z; ra[3] = a0.w; ra[4] = a1.x; ra[5] = a1.y; ra[6] = a1.z; ra[7] = a1.w;
            rb[0] = b0.x; rb[1] = b0.y; rb[2] = b0.z; rb[3] = b0.w; rb[4] = b1.x; rb[5] = b1.y; rb[6] = b1.z; rb[7] = b1.w;
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(ra[i], rb[j], acc[i][j]);
        }
        if (more) {
            store_tile(buf ^ 1, av, bv);
            __syncthreads();
            buf ^= 1;
        }
    }
    // epilogue: bias + LeakyReLU, 128-bit stores (out % 4 == 0)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const long long row = row0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (row >= a.P) continue;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int col = col0 + half * 64 + tx * 4;
            if (col >= a.out) continue;
            float v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float t = acc[i][half * 4 + j] + __ldg(a.b + col + j);
                v[j] = t > 0.f ? t : a.slope * t;
            }
            *reinterpret_cast<float4*>(a.y + (size_t)row * a.out + col) = make_float4(v[0], v[1], v[2], v[3]);
        }
    }
}

// out[p] = sigmoid(x[p] . w + b): one warp per pair.
__global__ void ncf_out_kernel(const float* __restrict__ x, long long P, int F, const float* __restrict__ w,
                               const float* __restrict__ b, float* __restrict__ out) {
    const long long p = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (p >= P) return;
    float s = 0.f;
    for (int c = lane_id(); c < F; c += 32) s = fmaf(x[(size_t)p * F + c], __ldg(w + c), s);
    s = warp_sum(s);
    if (lane_id() == 0) out[p] = 1.0f / (1.0f + expf(-(s + __ldg(b))));
}

}  // namespace

int ncf_layer_in(int F, int depth, int layer) { (void)depth; return F * (layer == 2 ? 4 : 2); }
int ncf_layer_out(int F, int depth, int layer) { return F * (layer == depth ? 1 : (layer == 1 ? 4 : 2)); }

long long ncf_param_count(int F, int depth) {
    long long n = 0;
    for (int l = 1; l <= depth; ++l) n += (long long)ncf_layer_in(F, depth, l) * ncf_layer_out(F, depth, l) + ncf_layer_out(F, depth, l);
    return n + F + 1;
}

// y[P, out] = LeakyReLU_slope(x[P, in] w^T + b) with the same FFMA GEMM (slope 1 = no activation); in and out must be
// multiples of 4.  Used by the GCN inference path (gcn_infer.cu).
cudaError_t launch_linear_f32(const float* x, const float* w, const float* b, float* y, long long P, int in, int out,
                              float slope, cudaStream_t stream) {
    if (P <= 0) return cudaSuccess;
    if ((in & 3) || (out & 3)) return cudaErrorInvalidValue;
    LinearArgs a;
    a.x = x; a.h = nullptr; a.src = nullptr; a.dst = nullptr; a.n_rows = 0; a.F = 0;
    a.in = in; a.out = out; a.w = w; a.b = b; a.y = y; a.P = P; a.slope = slope;
    dim3 grid((unsigned)((out + BN - 1) / BN), (unsigned)((P + BM - 1) / BM));
    ncf_linear_kernel<<<grid, kGemmThreads, 0, stream>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_ncf_score(const float* h, long long n_rows, int F, int depth, const float* params,
                             const long long* src, const long long* dst, long long P, float* out, float* ws0,
                             float* ws1, long long chunk, cudaStream_t stream) {
    for (long long p0 = 0; p0 < P; p0 += chunk) {
        const long long pc = P - p0 < chunk ? P - p0 : chunk;
        const float* w = params;
        const float* x = nullptr;
        float* bufs[2] = {ws0, ws1};
        for (int l = 1; l <= depth; ++l) {
            LinearArgs a;
            a.x = x; a.h = h; a.src = src + p0; a.dst = dst + p0; a.n_rows = n_rows; a.F = F;
            a.in = ncf_layer_in(F, depth, l); a.out = ncf_layer_out(F, depth, l);
            a.w = w; a.b = w + (size_t)a.in * a.out; a.y = bufs[(l - 1) & 1]; a.P = pc; a.slope = 0.01f;
            dim3 grid((unsigned)((a.out + BN - 1) / BN), (unsigned)((pc + BM - 1) / BM));
            ncf_linear_kernel<<<grid, kGemmThreads, 0, stream>>>(a);
            cudaError_t e = cudaGetLastError();
            if (e != cudaSuccess) return e;
            x = a.y;
            w = a.b + a.out;
        }
        ncf_out_kernel<<<(unsigned)((pc + 7) / 8), 256, 0, stream>>>(x, pc, F, w, w + F, out + p0);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

}  // namespace hwer
