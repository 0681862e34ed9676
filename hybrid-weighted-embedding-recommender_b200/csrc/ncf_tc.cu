// NCF re-rank on the tcgen05 tensor cores.
//
// Replaces the same reference code as ncf.cu -- NCF.forward, hwer/ncf.py:7-27, driven by GcnNCF.predict,
// hwer/gcn_ncf.py:336-361, and the NCF branch of find_closest_neighbours, hwer/gcn_ncf.py:384-386 -- for feature
// widths that are a multiple of 64 (the serving widths: F = 64, 128, 256).
//
// The reference computes in fp32 and the score tolerance is 1e-5, so the operands are SPLIT: x = x_hi + x_lo with
// both halves bf16 (16 mantissa bits together), likewise the weights, and every K = 16 step issues three
// tcgen05.mma into the same fp32 TMEM accumulator:  x_hi w_hi + x_hi w_lo + x_lo w_hi  (the dropped x_lo w_lo term
// is 2^-16 of a product).  Measured against the fp32 reference MLP: <= 6e-7 on the sigmoid output for F = 64..256,
// depth 2..4 (tests/test_gpu_parity.py::test_ncf_*), at 3x the tensor work of a plain bf16 GEMM -- which is still
// an order of magnitude more than the fp32 FFMA pipe delivers.
//
// One persistent kernel launch per Linear + LeakyReLU layer (one CTA per SM):
//   warp 0      TMA producer: [128 pairs x 64] tiles of x_hi / x_lo and [N x 64] tiles of w_hi / w_lo, 128B swizzle
//   warp 1      MMA issuer + TMEM owner: 12 MMAs per 64-wide K block, two accumulator stages of N <= 256 columns
//   warps 2-17  epilogue: tcgen05.ld -> + bias -> LeakyReLU -> split to bf16 hi / lo -> next layer's operand arrays;
//               the last layer instead folds Linear(F, 1) + sigmoid into the epilogue (a fixed-order row reduction)
// The gather of [h[src] || h[dst]] and its split feed the first layer; activations never exist in fp32 in HBM.
#include <cstring>

#include "common.cuh"
#include "kernels.h"

namespace hwer {

namespace {

constexpr int kNcfM = 128;                    // pairs per tile (UMMA M)
constexpr int kNcfEpiWarps = 16;
constexpr int kNcfThreads = (2 + kNcfEpiWarps) * 32;
constexpr int kNcfXSlab = kNcfM * 128;        // one 64-wide K block of an activation tile: 16 KB

struct NcfLayerParams {
    long long p_rows;          // pairs in this chunk (rows >= p_rows of the padded tiles are zero)
    int m_tiles;               // padded rows / 128
    int n_tiles;               // out / n_tile
    int n_tile;                // columns per tile (64, 128 or 256)
    int in, out;
    int stages;
    const float* bias;         // [out]
    float slope;
    __nv_bfloat16* y_hi;       // [m_tiles * 128, out] next layer's operands (nullptr on the last layer)
    __nv_bfloat16* y_lo;
    const float* w_out;        // last layer: Linear(out, 1) weights, bias and the score output
    const float* b_out;
    float* score;              // [p_rows]
};

__device__ __forceinline__ void split_bf16(float v, __nv_bfloat16& hi, __nv_bfloat16& lo) {
    hi = __float2bfloat16_rn(v);
    lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}

// [h[src] || h[dst]] -> bf16 hi / lo rows of width 2F; rows >= P (padding up to a multiple of 128) are zero.
__global__ void __launch_bounds__(256)
ncf_gather_split_kernel(const float* __restrict__ h, long long n_rows, int F, const long long* __restrict__ src,
                        const long long* __restrict__ dst, long long P, long long p_pad, __nv_bfloat16* __restrict__ x_hi,
                        __nv_bfloat16* __restrict__ x_lo) {
    const int per_row = (2 * F) >> 2;                          // float4 chunks per output row
    const long long total = p_pad * per_row;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long p = i / per_row;
        const int c = (int)(i - p * per_row) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p < P) {
            long long r = c < F ? src[p] : dst[p];
            if (r < 0 || r >= n_rows) r = 0;                   // unknown node -> the padding row (gcn_ncf.py:341-342)
            v = __ldg(reinterpret_cast<const float4*>(h + (size_t)r * F + (c < F ? c : c - F)));
        }
        __nv_bfloat16 hi[4], lo[4];
        split_bf16(v.x, hi[0], lo[0]); split_bf16(v.y, hi[1], lo[1]);
        split_bf16(v.z, hi[2], lo[2]); split_bf16(v.w, hi[3], lo[3]);
        const size_t o = (size_t)p * 2 * F + c;
        *reinterpret_cast<uint2*>(x_hi + o) = *reinterpret_cast<const uint2*>(hi);
        *reinterpret_cast<uint2*>(x_lo + o) = *reinterpret_cast<const uint2*>(lo);
    }
}

__global__ void __launch_bounds__(256)
ncf_split_kernel(const float* __restrict__ w, long long n, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        split_bf16(w[i], hi[i], lo[i]);
}

template <bool LAST>
__global__ void __launch_bounds__(kNcfThreads, 1)
ncf_layer_tc_kernel(const __grid_constant__ CUtensorMap tm_xh, const __grid_constant__ CUtensorMap tm_xl,
                    const __grid_constant__ CUtensorMap tm_wh, const __grid_constant__ CUtensorMap tm_wl,
                    const __grid_constant__ NcfLayerParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int n_tile = p.n_tile;
    const int w_slab = n_tile * 128;                              // bytes of one 64-wide K block of a weight tile
    const int stage_bytes = 2 * kNcfXSlab + 2 * w_slab;           // x_hi, x_lo, w_hi, w_lo
    uint8_t* ring = smem;
    float* red_s = reinterpret_cast<float*>(ring + (size_t)p.stages * stage_bytes);     // [8 chunks][128 rows] (LAST)
    uint64_t* bars = reinterpret_cast<uint64_t*>(red_s + 8 * kNcfM);
    uint64_t* full_bar = bars;                      // [stages]  TMA -> MMA
    uint64_t* empty_bar = bars + p.stages;          // [stages]  MMA -> TMA
    uint64_t* tfull_bar = bars + 2 * p.stages;      // [2]       MMA -> epilogue
    uint64_t* tempty_bar = tfull_bar + 2;           // [2]       epilogue -> MMA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int chunks = n_tile / 32;                               // 32-column chunks per tile: 2, 4 or 8
    // epilogue warp (quad, part) reads TMEM lanes 32*quad.. and chunks part, part + 4 (as in score_filter.cu)
    const int parts_active = chunks < 4 ? chunks : 4;
    const uint32_t tmem_cols = n_tile * 2 <= 32 ? 32u : (n_tile * 2 <= 64 ? 64u : (n_tile * 2 <= 128 ? 128u : (n_tile * 2 <= 256 ? 256u : 512u)));

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_xh); tma_prefetch_desc(&tm_xl); tma_prefetch_desc(&tm_wh); tma_prefetch_desc(&tm_wl);
        for (int i = 0; i < p.stages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], 4 * parts_active); }
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, tmem_cols);
        tmem_relinquish();
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    const int k_blocks = p.in / 64;
    const int items = p.m_tiles * p.n_tiles;                      // work items, n fastest: CTAs of a wave share x tiles
    const uint32_t idesc = umma_idesc_bf16_f32(kNcfM, (uint32_t)n_tile);

    if (warp == 0) {
        if (lane == 0) {
            const uint64_t keep = l2_policy_evict_last();
            uint32_t stage = 0, phase = 0;
            for (int it = blockIdx.x; it < items; it += gridDim.x) {
                const int mt = it / p.n_tiles, nt = it - mt * p.n_tiles;
                for (int kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1u);
                    mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)stage_bytes);
                    uint8_t* dst = ring + (size_t)stage * stage_bytes;
                    tma_load_2d(dst, &tm_xh, kb * 64, mt * kNcfM, &full_bar[stage], keep);
                    tma_load_2d(dst + kNcfXSlab, &tm_xl, kb * 64, mt * kNcfM, &full_bar[stage], keep);
                    tma_load_2d(dst + 2 * kNcfXSlab, &tm_wh, kb * 64, nt * n_tile, &full_bar[stage], keep);
                    tma_load_2d(dst + 2 * kNcfXSlab + w_slab, &tm_wl, kb * 64, nt * n_tile, &full_bar[stage], keep);
                    if (++stage == (uint32_t)p.stages) { stage = 0; phase ^= 1u; }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        const bool leader = elect_one();
        const uint32_t sw_hi = (uint32_t)(umma_desc_k_sw128(0) >> 32);
        const uint32_t ring_lo = (uint32_t)umma_desc_k_sw128(smem_u32(ring));
        uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
        for (int it = blockIdx.x; it < items; it += gridDim.x) {
            mbar_wait(&tempty_bar[acc], acc_phase ^ 1u);
            const uint32_t d_tmem = tmem_base + acc * (uint32_t)n_tile;
            for (int kb = 0; kb < k_blocks; ++kb) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after_sync();
                if (leader) {
                    const uint32_t xh = ring_lo + ((stage * (uint32_t)stage_bytes) >> 4);
                    const uint32_t xl = xh + (kNcfXSlab >> 4);
                    const uint32_t wh = xh + ((2 * kNcfXSlab) >> 4);
                    const uint32_t wl = wh + ((uint32_t)w_slab >> 4);
#pragma unroll
                    for (int s = 0; s < 4; ++s) {                 // K = 16 steps of the 64-wide block (32 bytes each)
                        umma_bf16_lohi(d_tmem, xh + 2 * s, sw_hi, wh + 2 * s, sw_hi, idesc, (kb | s) ? 1u : 0u);
                        umma_bf16_lohi(d_tmem, xh + 2 * s, sw_hi, wl + 2 * s, sw_hi, idesc, 1u);
                        umma_bf16_lohi(d_tmem, xl + 2 * s, sw_hi, wh + 2 * s, sw_hi, idesc, 1u);
                    }
                    umma_commit(&empty_bar[stage]);
                    if (kb == k_blocks - 1) umma_commit(&tfull_bar[acc]);
                }
                __syncwarp();
                if (++stage == (uint32_t)p.stages) { stage = 0; phase ^= 1u; }
            }
            if (++acc == 2u) { acc = 0; acc_phase ^= 1u; }
        }
    } else if (((warp - 2) >> 2) < parts_active) {
        const uint32_t quad = (uint32_t)warp & 3u;
        const int part = (warp - 2) >> 2;
        const uint32_t lane_row = quad * 32u + (uint32_t)lane;
        uint32_t acc = 0, acc_phase = 0;
        for (int it = blockIdx.x; it < items; it += gridDim.x) {
            const int mt = it / p.n_tiles, nt = it - mt * p.n_tiles;
            const long long row = (long long)mt * kNcfM + lane_row;
            mbar_wait(&tfull_bar[acc], acc_phase);
            tc_fence_after_sync();
            float dot = 0.f;                                      // LAST: this thread's share of Linear(F, 1)
            for (int ch = part; ch < chunks; ch += 4) {
                uint32_t v[32];
                tmem_ld_x32(tmem_base + ((quad * 32u) << 16) + acc * (uint32_t)n_tile + 32u * (uint32_t)ch, v);
                tmem_ld_wait();
                const int col0 = nt * n_tile + 32 * ch;
                if (LAST) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        float t = __uint_as_float(v[j]) + __ldg(p.bias + col0 + j);
                        t = t > 0.f ? t : p.slope * t;
                        dot = fmaf(t, __ldg(p.w_out + col0 + j), dot);
                    }
                    red_s[ch * kNcfM + lane_row] = dot;           // one partial per (chunk, row); summed in order below
                    dot = 0.f;
                } else {
                    __align__(16) __nv_bfloat16 hi[32], lo[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        float t = __uint_as_float(v[j]) + __ldg(p.bias + col0 + j);
                        t = t > 0.f ? t : p.slope * t;
                        split_bf16(t, hi[j], lo[j]);
                    }
                    uint4* yh = reinterpret_cast<uint4*>(p.y_hi + (size_t)row * p.out + col0);
                    uint4* yl = reinterpret_cast<uint4*>(p.y_lo + (size_t)row * p.out + col0);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        yh[j] = reinterpret_cast<const uint4*>(hi)[j];
                        yl[j] = reinterpret_cast<const uint4*>(lo)[j];
                    }
                }
            }
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[acc]);        // accumulator stage drained
            if (++acc == 2u) { acc = 0; acc_phase ^= 1u; }
            if (LAST) {
                // the 4 * parts_active epilogue warps meet (named barrier 1), then the part-0 warps finish their rows
                asm volatile("bar.sync 1, %0;" ::"r"(128 * parts_active) : "memory");
                if (part == 0 && row < p.p_rows) {
                    float s = 0.f;
                    for (int ch = 0; ch < chunks; ++ch) s += red_s[ch * kNcfM + lane_row];
                    p.score[row] = 1.0f / (1.0f + expf(-(s + __ldg(p.b_out))));
                }
                asm volatile("bar.sync 1, %0;" ::"r"(128 * parts_active) : "memory");
            }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, tmem_cols);
}

size_t ncf_tc_smem_bytes(int n_tile, int stages) {
    return 1024 + (size_t)stages * (2 * kNcfXSlab + 2 * (size_t)n_tile * 128) + 8 * kNcfM * sizeof(float) +
           (2 * stages + 4) * sizeof(uint64_t) + 16;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn ncf_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* q = nullptr;
        cudaDriverEntryPointQueryResult r;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &q, cudaEnableDefault, &r) == cudaSuccess &&
            r == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(q);
    }
    return fn;
}

// bf16 [rows, cols] row-major, box [box_rows x 64 columns], 128B swizzle
bool ncf_tensor_map(CUtensorMap* m, const void* base, long long rows, int cols, int box_rows) {
    EncodeTiledFn enc = ncf_encode_fn();
    if (!enc) return false;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
    cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

bool ncf_tc_supported(int F, int depth) { return F % 64 == 0 && F >= 64 && F <= 256 && depth >= 1; }

long long ncf_tc_chunk(int num_sms) { return (long long)num_sms * 4 * kNcfM; }      // four 128-pair tiles per CTA

// bf16 elements of scratch: two ping-pong pairs (hi, lo) of [chunk, 4F] activations + split weights
static size_t ncf_params_padded(int F, int depth) { return ((size_t)ncf_param_count(F, depth) + 63) / 64 * 64; }
size_t ncf_tc_scratch_elems(int F, int depth, long long chunk) {
    return 4 * (size_t)chunk * 4 * F + 2 * ncf_params_padded(F, depth);
}

cudaError_t launch_ncf_score_tc(const float* h, long long n_rows, int F, int depth, const float* params,
                                const long long* src, const long long* dst, long long P, float* out, void* scratch,
                                long long chunk, int num_sms, cudaStream_t stream) {
    __nv_bfloat16* base = reinterpret_cast<__nv_bfloat16*>(scratch);
    const size_t act = (size_t)chunk * 4 * F;
    __nv_bfloat16* xh[2] = {base, base + 2 * act};
    __nv_bfloat16* xl[2] = {base + act, base + 3 * act};
    __nv_bfloat16* wh = base + 4 * act;
    __nv_bfloat16* wl = wh + ncf_params_padded(F, depth);            // keeps every weight tile 128-byte aligned
    const long long n_params = ncf_param_count(F, depth);
    ncf_split_kernel<<<148 * 4, 256, 0, stream>>>(params, n_params, wh, wl);       // biases are split too (unused)
    static bool attr_done[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !attr_done[dev]) {
        cudaError_t e = cudaFuncSetAttribute(ncf_layer_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(ncf_layer_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) attr_done[dev] = true;
    }
    // Tensor maps depend on the scratch layout only (ping-pong activation arrays of `chunk` rows, split weights), not
    // on the chunk being processed: encode them ONCE per call.  (r2: encoding 12 maps per chunk made the chunk loop
    // host-bound -- 5.0 ms per 819 K pairs of which 2.5 ms were kernels.)  A short last chunk reuses them: the layer
    // kernels only touch the first p_pad / 128 row tiles, and the gather zero-fills [pc, p_pad).
    struct LayerPlan { NcfLayerParams q; CUtensorMap mxh, mxl, mwh, mwl; size_t smem; bool last; };
    LayerPlan plan[16];
    if (depth > 16) return cudaErrorInvalidValue;
    {
        size_t w_off = 0;
        int cur = 0;
        for (int l = 1; l <= depth; ++l) {
            LayerPlan& L = plan[l - 1];
            const int in = ncf_layer_in(F, depth, l), outw = ncf_layer_out(F, depth, l);
            L.last = l == depth;
            NcfLayerParams& q = L.q;
            memset(&q, 0, sizeof q);
            q.n_tile = outw < 256 ? outw : 256;
            q.n_tiles = outw / q.n_tile;
            q.in = in; q.out = outw; q.slope = 0.01f;
            q.bias = params + w_off + (size_t)in * outw;
            q.stages = 2;
            while (q.stages < 6 && ncf_tc_smem_bytes(q.n_tile, q.stages + 1) <= (size_t)kSmemBudget) ++q.stages;
            if (!ncf_tensor_map(&L.mxh, xh[cur], chunk, in, kNcfM) || !ncf_tensor_map(&L.mxl, xl[cur], chunk, in, kNcfM) ||
                !ncf_tensor_map(&L.mwh, wh + w_off, outw, in, q.n_tile) || !ncf_tensor_map(&L.mwl, wl + w_off, outw, in, q.n_tile))
                return cudaErrorInvalidValue;
            L.smem = ncf_tc_smem_bytes(q.n_tile, q.stages);
            if (L.last) {
                q.w_out = params + w_off + (size_t)in * outw + outw;
                q.b_out = q.w_out + F;
            } else {
                q.y_hi = xh[cur ^ 1]; q.y_lo = xl[cur ^ 1];
            }
            w_off += (size_t)in * outw + outw;
            cur ^= 1;
        }
    }
    for (long long p0 = 0; p0 < P; p0 += chunk) {
        const long long pc = P - p0 < chunk ? P - p0 : chunk;
        const long long p_pad = (pc + kNcfM - 1) / kNcfM * kNcfM;
        {
            const long long total = p_pad * (2 * F / 4);
            long long blocks = (total + 255) / 256;
            if (blocks > 148LL * 16) blocks = 148LL * 16;
            ncf_gather_split_kernel<<<(int)blocks, 256, 0, stream>>>(h, n_rows, F, src + p0, dst + p0, pc, p_pad, xh[0], xl[0]);
        }
        for (int l = 1; l <= depth; ++l) {
            LayerPlan& L = plan[l - 1];
            NcfLayerParams q = L.q;
            q.p_rows = pc; q.m_tiles = (int)(p_pad / kNcfM);
            const int items = q.m_tiles * q.n_tiles;
            const int grid = items < num_sms ? items : num_sms;
            if (L.last) {
                q.score = out + p0;
                ncf_layer_tc_kernel<true><<<grid, kNcfThreads, L.smem, stream>>>(L.mxh, L.mxl, L.mwh, L.mwl, q);
            } else {
                ncf_layer_tc_kernel<false><<<grid, kNcfThreads, L.smem, stream>>>(L.mxh, L.mxl, L.mwh, L.mwl, q);
            }
            cudaError_t e = cudaGetLastError();
            if (e != cudaSuccess) return e;
        }
    }
    return cudaSuccess;
}

}  // namespace hwer
