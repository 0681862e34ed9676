// Blend + L2-normalise, and the unit-norm check, as HBM-bound streaming kernels.
//
//   blend_normalize : V = unit(alpha * unit(C) + (1 - alpha) * unit(G))   per row
//       the alpha-weighted content/collaborative blend of BASELINE.json's
//       north_star, in the slot hwer/gcn_ncf.py:447-456 (prepare_for_knn) where
//       the reference today computes unit_length(G) only (== alpha = 0);
//       unit() is hwer/utils.py:43-44 (a / ||a||, no epsilon: a zero row -> NaN).
//       Writes the fp32 table and, optionally, the zero-padded bf16 shadow the
//       tensor-core scorer streams.
//   norm_stats      : hwer/utils.py:51-57 unit_length_violations (+ max norm).
//
// One warp per row, 128-bit coalesced loads, rows held in registers between
// the norm pass and the write pass (no second read from HBM).
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"

namespace hwer {

namespace {

constexpr int kRowThreads = 256;   // 8 warps -> 8 rows per CTA iteration
constexpr int kMaxVec = 4;         // register-resident path: d <= 32 lanes * 4 float4 * 4 = 512

__device__ __forceinline__ void store_bf16_row_chunk(__nv_bfloat16* dst, const float4& v) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y);
    __nv_bfloat162 b = __floats2bfloat162_rn(v.z, v.w);
    uint2 u;
    u.x = *reinterpret_cast<uint32_t*>(&a);
    u.y = *reinterpret_cast<uint32_t*>(&b);
    stg_stream_u2(reinterpret_cast<uint2*>(dst), u);
}

// VEC = float4 per lane (d4 = d/4 <= 32*VEC).  HAS_C: blend with a content row.  A warp works on kRowsPerWarp rows at
// a time: all their loads are issued before the first reduction, so a lane has 2 * VEC * kRowsPerWarp 128-bit loads
// in flight.  Measured at 10 M x 128 (scripts/tune_blend.py, profiles/r02_w_tune_blend.txt, best of 5 launches):
// one row per warp 5.10-5.98 TB/s, two 5.94-6.37, four rows with 32 CTAs per SM 6.50 TB/s = 0.99 of the copy bandwidth
// measured in the same process.
template <int VEC, bool HAS_C, int kRowsPerWarp = (VEC == 1 ? 4 : (VEC == 2 ? 2 : 1))>   // wider rows: fewer fit registers
__global__ void __launch_bounds__(kRowThreads)
blend_normalize_vec_kernel(const float* __restrict__ content, const float* __restrict__ collab, float alpha,
                           const float* __restrict__ alpha_rows, long long n, int d, float* __restrict__ out_f32,
                           __nv_bfloat16* __restrict__ out_bf16, int d_pad) {
    const int lane = lane_id();
    const int d4 = d >> 2;
    const long long warps_total = (long long)gridDim.x * (kRowThreads / 32);
    for (long long row0 = ((long long)blockIdx.x * (kRowThreads / 32) + (threadIdx.x >> 5)) * kRowsPerWarp; row0 < n;
         row0 += warps_total * kRowsPerWarp) {
        float4 g[kRowsPerWarp][VEC], c[kRowsPerWarp][VEC];
#pragma unroll
        for (int r = 0; r < kRowsPerWarp; ++r) {
            const long long row = row0 + r < n ? row0 + r : n - 1;      // a clamped duplicate row is computed, not stored
            const float4* gp = reinterpret_cast<const float4*>(collab + (size_t)row * d);
            const float4* cp = HAS_C ? reinterpret_cast<const float4*>(content + (size_t)row * d) : nullptr;
#pragma unroll
            for (int i = 0; i < VEC; ++i) {
                const int j = lane + 32 * i;
                g[r][i] = (j < d4) ? ldg_stream_f4(gp + j) : make_float4(0.f, 0.f, 0.f, 0.f);
                if (HAS_C) c[r][i] = (j < d4) ? ldg_stream_f4(cp + j) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
#pragma unroll
        for (int r = 0; r < kRowsPerWarp; ++r) {
            const long long row = row0 + r;
            if (row >= n) break;
            float sg = 0.f, sc = 0.f;
#pragma unroll
            for (int i = 0; i < VEC; ++i) {
                sg = fmaf(g[r][i].x, g[r][i].x, sg); sg = fmaf(g[r][i].y, g[r][i].y, sg);
                sg = fmaf(g[r][i].z, g[r][i].z, sg); sg = fmaf(g[r][i].w, g[r][i].w, sg);
                if (HAS_C) {
                    sc = fmaf(c[r][i].x, c[r][i].x, sc); sc = fmaf(c[r][i].y, c[r][i].y, sc);
                    sc = fmaf(c[r][i].z, c[r][i].z, sc); sc = fmaf(c[r][i].w, c[r][i].w, sc);
                }
            }
            const float ng = sqrtf(warp_sum(sg));
            float sv = 0.f;
            if (HAS_C) {
                const float nc = sqrtf(warp_sum(sc));
                const float a = alpha_rows ? alpha_rows[row] : alpha;
                // unit() of each source first, as the spec composes them; the two inner normalisations use one
                // reciprocal per row (an IEEE divide per element made this kernel instruction-bound), the final
                // one below stays a true division like numpy's
                const float wc = a * (1.0f / nc), wg = (1.0f - a) * (1.0f / ng);
#pragma unroll
                for (int i = 0; i < VEC; ++i) {
                    g[r][i].x = fmaf(wc, c[r][i].x, wg * g[r][i].x);
                    g[r][i].y = fmaf(wc, c[r][i].y, wg * g[r][i].y);
                    g[r][i].z = fmaf(wc, c[r][i].z, wg * g[r][i].z);
                    g[r][i].w = fmaf(wc, c[r][i].w, wg * g[r][i].w);
                    sv = fmaf(g[r][i].x, g[r][i].x, sv); sv = fmaf(g[r][i].y, g[r][i].y, sv);
                    sv = fmaf(g[r][i].z, g[r][i].z, sv); sv = fmaf(g[r][i].w, g[r][i].w, sv);
                }
            }
            const float nv = HAS_C ? sqrtf(warp_sum(sv)) : ng;
            float4* op = reinterpret_cast<float4*>(out_f32 + (size_t)row * d);
#pragma unroll
            for (int i = 0; i < VEC; ++i) {
                const int j = lane + 32 * i;
                float4 o;
                o.x = g[r][i].x / nv; o.y = g[r][i].y / nv; o.z = g[r][i].z / nv; o.w = g[r][i].w / nv;
                if (j < d4) {
                    stg_stream_f4(op + j, o);
                    if (out_bf16) store_bf16_row_chunk(out_bf16 + (size_t)row * d_pad + 4 * j, o);
                }
            }
            if (out_bf16) {   // zero the padding columns [d, d_pad)
                for (int col = d + 4 * lane; col < d_pad; col += 128)
                    stg_stream_u2(reinterpret_cast<uint2*>(out_bf16 + (size_t)row * d_pad + col), make_uint2(0u, 0u));
            }
        }
    }
}

// Any d (scalar accesses, rows re-read from cache for the second pass).
__global__ void __launch_bounds__(kRowThreads)
blend_normalize_generic_kernel(const float* __restrict__ content, const float* __restrict__ collab, float alpha,
                               const float* __restrict__ alpha_rows, long long n, int d, float* __restrict__ out_f32,
                               __nv_bfloat16* __restrict__ out_bf16, int d_pad) {
    const int lane = lane_id();
    const long long warps_total = (long long)gridDim.x * (kRowThreads / 32);
    for (long long row = (long long)blockIdx.x * (kRowThreads / 32) + (threadIdx.x >> 5); row < n;
         row += warps_total) {
        const float* g = collab + (size_t)row * d;
        const float* c = content ? content + (size_t)row * d : nullptr;
        float sg = 0.f, sc = 0.f;
        for (int j = lane; j < d; j += 32) {
            sg = fmaf(g[j], g[j], sg);
            if (c) sc = fmaf(c[j], c[j], sc);
        }
        const float ng = sqrtf(warp_sum(sg));
        const float nc = c ? sqrtf(warp_sum(sc)) : 1.0f;
        const float a = c ? (alpha_rows ? alpha_rows[row] : alpha) : 0.0f;
        const float b = 1.0f - a;
        float sv = 0.f;
        if (c) {
            for (int j = lane; j < d; j += 32) {
                const float v = a * (c[j] / nc) + b * (g[j] / ng);
                sv = fmaf(v, v, sv);
            }
        }
        const float nv = c ? sqrtf(warp_sum(sv)) : ng;
        for (int j = lane; j < d_pad || j < d; j += 32) {
            if (j < d) {
                const float v = c ? a * (c[j] / nc) + b * (g[j] / ng) : g[j];
                const float o = v / nv;
                out_f32[(size_t)row * d + j] = o;
                if (out_bf16) out_bf16[(size_t)row * d_pad + j] = __float2bfloat16_rn(o);
            } else if (out_bf16) {
                out_bf16[(size_t)row * d_pad + j] = __float2bfloat16_rn(0.0f);
            }
        }
    }
}

// Per-row norm statistics: partial[block] = {positive, negative, sum|norm-1|, max norm}
__global__ void __launch_bounds__(kRowThreads)
norm_stats_kernel(const float* __restrict__ v, long long n, int d, float eps, double* __restrict__ partial) {
    __shared__ double sh[4][kRowThreads / 32];
    const int lane = lane_id(), warp = threadIdx.x >> 5;
    const long long warps_total = (long long)gridDim.x * (kRowThreads / 32);
    double pos = 0, neg = 0, dev = 0, mx = 0;
    const bool vec = (d & 3) == 0;
    for (long long row = (long long)blockIdx.x * (kRowThreads / 32) + warp; row < n; row += warps_total) {
        float s = 0.f;
        if (vec) {
            const float4* p = reinterpret_cast<const float4*>(v + (size_t)row * d);
            for (int j = lane; j < (d >> 2); j += 32) {
                const float4 x = ldg_stream_f4(p + j);
                s = fmaf(x.x, x.x, s); s = fmaf(x.y, x.y, s); s = fmaf(x.z, x.z, s); s = fmaf(x.w, x.w, s);
            }
        } else {
            for (int j = lane; j < d; j += 32) { const float x = v[(size_t)row * d + j]; s = fmaf(x, x, s); }
        }
        const float nrm = sqrtf(warp_sum(s));
        if (lane == 0) {
            // comparisons happen in double like numpy's float32-vs-python-float compare
            if ((double)nrm > 1.0 + (double)eps) pos += 1;
            if ((double)nrm < 1.0 - (double)eps) neg += 1;
            dev += fabs((double)nrm - 1.0);   // NaN rows poison the mean exactly as np.mean would
            if (!(nrm <= mx)) mx = nrm;       // NaN propagates into max as well
        }
    }
    if (lane == 0) { sh[0][warp] = pos; sh[1][warp] = neg; sh[2][warp] = dev; sh[3][warp] = mx; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0, b = 0, c = 0, m = 0;
        for (int w = 0; w < kRowThreads / 32; ++w) {
            a += sh[0][w]; b += sh[1][w]; c += sh[2][w];
            if (!(sh[3][w] <= m)) m = sh[3][w];
        }
        partial[4 * blockIdx.x + 0] = a; partial[4 * blockIdx.x + 1] = b;
        partial[4 * blockIdx.x + 2] = c; partial[4 * blockIdx.x + 3] = m;
    }
}

// Fixed-order final reduction: out5 = {violations, mean|norm-1|, positive, negative, max norm}.  One CTA: thread t
// sums partials t, t + 256, ... in ascending order, then a fixed binary tree over the 256 threads -- the same
// association for a given grid size on every run (r1 walked the ~2400 partials with ONE thread: 53 us).
__global__ void __launch_bounds__(256)
norm_stats_reduce_kernel(const double* __restrict__ partial, int nblocks, long long n, double* __restrict__ out5) {
    __shared__ double sh[4][256];
    double a = 0, b = 0, c = 0, m = 0;
    for (int i = threadIdx.x; i < nblocks; i += 256) {
        a += partial[4 * i]; b += partial[4 * i + 1]; c += partial[4 * i + 2];
        if (!(partial[4 * i + 3] <= m)) m = partial[4 * i + 3];
    }
    sh[0][threadIdx.x] = a; sh[1][threadIdx.x] = b; sh[2][threadIdx.x] = c; sh[3][threadIdx.x] = m;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) {
            sh[0][threadIdx.x] += sh[0][threadIdx.x + s];
            sh[1][threadIdx.x] += sh[1][threadIdx.x + s];
            sh[2][threadIdx.x] += sh[2][threadIdx.x + s];
            if (!(sh[3][threadIdx.x + s] <= sh[3][threadIdx.x])) sh[3][threadIdx.x] = sh[3][threadIdx.x + s];   // NaN wins
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        out5[0] = sh[0][0] + sh[1][0];
        out5[1] = n > 0 ? sh[2][0] / (double)n : 0.0;
        out5[2] = sh[0][0];
        out5[3] = sh[1][0];
        out5[4] = sh[3][0];
    }
}

// fp32 table -> zero-padded bf16 shadow (round-to-nearest-even), no normalisation.
__global__ void __launch_bounds__(256)
make_shadow_kernel(const float* __restrict__ table, long long n, int d, __nv_bfloat16* __restrict__ out, int d_pad) {
    const long long total = n * (long long)(d_pad >> 1);   // one bf16x2 per thread step
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long row = i / (d_pad >> 1);
        const int col = (int)(i - row * (d_pad >> 1)) * 2;
        const float a = col < d ? table[(size_t)row * d + col] : 0.0f;
        const float b = col + 1 < d ? table[(size_t)row * d + col + 1] : 0.0f;
        reinterpret_cast<__nv_bfloat162*>(out)[i] = __floats2bfloat162_rn(a, b);
    }
}

// A/B knobs, read once (scripts/tune_blend.py): rows a warp keeps in flight, CTAs per SM of the persistent grid
int blend_knob(const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}

int row_grid(long long n) {
    long long blocks = (n + (kRowThreads / 32) - 1) / (kRowThreads / 32);
    static const int ctas_per_sm = blend_knob("HWER_BLEND_CTAS", 32);
    const long long cap = 148LL * ctas_per_sm;   // grid-stride over 32 CTAs per SM (8 resident at a time)
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

}  // namespace

cudaError_t launch_blend_normalize(const float* content, const float* collab, float alpha, const float* alpha_rows,
                                   long long n, int d, float* out_f32, void* out_bf16_v, int d_pad,
                                   cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    __nv_bfloat16* out_bf16 = reinterpret_cast<__nv_bfloat16*>(out_bf16_v);
    const int grid = row_grid(n);
    const bool vec_ok = (d % 4 == 0) && d <= 128 * kMaxVec && (!out_bf16 || d_pad % 4 == 0);
    if (!vec_ok) {
        blend_normalize_generic_kernel<<<grid, kRowThreads, 0, stream>>>(content, collab, alpha, alpha_rows, n, d,
                                                                        out_f32, out_bf16, d_pad);
        return cudaGetLastError();
    }
    const int vec = (d / 4 + 31) / 32;
    static const int rpw = blend_knob("HWER_BLEND_RPW", 0);      // 0 = the kernel's default (2 rows for d <= 256)
#define HWER_LAUNCH_BLEND_R(V, R)                                                                              \
    if (content)                                                                                               \
        blend_normalize_vec_kernel<V, true, R><<<grid, kRowThreads, 0, stream>>>(content, collab, alpha,        \
                                                                                alpha_rows, n, d, out_f32,     \
                                                                                out_bf16, d_pad);              \
    else                                                                                                       \
        blend_normalize_vec_kernel<V, false, R><<<grid, kRowThreads, 0, stream>>>(content, collab, alpha,       \
                                                                                 alpha_rows, n, d, out_f32,    \
                                                                                 out_bf16, d_pad);
#define HWER_LAUNCH_BLEND(V)                                                                                   \
    if (content)                                                                                               \
        blend_normalize_vec_kernel<V, true><<<grid, kRowThreads, 0, stream>>>(content, collab, alpha, alpha_rows, \
                                                                             n, d, out_f32, out_bf16, d_pad);  \
    else                                                                                                       \
        blend_normalize_vec_kernel<V, false><<<grid, kRowThreads, 0, stream>>>(content, collab, alpha,          \
                                                                              alpha_rows, n, d, out_f32,       \
                                                                              out_bf16, d_pad);
    if (vec == 1 && (rpw == 1 || rpw == 2)) {
        if (rpw == 1) { HWER_LAUNCH_BLEND_R(1, 1) } else { HWER_LAUNCH_BLEND_R(1, 2) }
        return cudaGetLastError();
    }
    switch (vec) {
        case 1: HWER_LAUNCH_BLEND(1) break;
        case 2: HWER_LAUNCH_BLEND(2) break;
        case 3: HWER_LAUNCH_BLEND(3) break;
        default: HWER_LAUNCH_BLEND(4) break;
    }
#undef HWER_LAUNCH_BLEND
#undef HWER_LAUNCH_BLEND_R
    return cudaGetLastError();
}

cudaError_t launch_make_shadow(const float* table, long long n, int d, void* out_bf16, int d_pad, cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    long long blocks = (n * (d_pad / 2) + 255) / 256;
    if (blocks > 148LL * 32) blocks = 148LL * 32;
    make_shadow_kernel<<<(int)blocks, 256, 0, stream>>>(table, n, d, reinterpret_cast<__nv_bfloat16*>(out_bf16), d_pad);
    return cudaGetLastError();
}

cudaError_t launch_norm_stats(const float* v, long long n, int d, float eps, double* out5, cudaStream_t stream) {
    const int grid = row_grid(n);
    double* partial = nullptr;
    cudaError_t e = cudaMallocAsync(&partial, sizeof(double) * 4 * grid, stream);
    if (e != cudaSuccess) return e;
    norm_stats_kernel<<<grid, kRowThreads, 0, stream>>>(v, n, d, eps, partial);
    norm_stats_reduce_kernel<<<1, 256, 0, stream>>>(partial, grid, n, out5);
    e = cudaGetLastError();
    cudaFreeAsync(partial, stream);
    return e;
}

}  // namespace hwer
