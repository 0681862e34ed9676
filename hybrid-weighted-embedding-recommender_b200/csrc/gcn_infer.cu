// GCN inference: the producer of the collaborative table the serving path consumes.
//
// Replaces: get_gcn_vectors, hwer/gcn_ncf.py:260-279 -- GraphConvModule.forward in eval mode, hwer/gcn.py:162-193,
//           with its content projection (build_content_layer, gcn.py:59-63; mix_embeddings, gcn.py:40-44) and its
//           GraphConv layers (gcn.py:104-128) -- over the WHOLE graph at once: every NodeFlow layer holds every node.
// The neighbour sample of each block (what DGL's NeighborSampler draws: two random in-neighbours and a self loop per
// node, gcn_ncf.py:262-272) is an INPUT, one CSR list per block, so the result is a deterministic function of its
// arguments; tests/golden/reference_gcn.npz holds outputs of the reference module itself for such lists.
//
//   h0[v]  = unit(node_emb[v + 1] + LayerNorm(LeakyReLU_0.1(content[v] W^T + b)))
//   H_i[v] = [ mean_{u in nbr_i(v)} H_{i-1}[u]  ||  h0[v] ]          (width F (i + 1); DGL copy_src + sum, then / count)
//   out[v] = unit(fc1(LeakyReLU_0.01(fc0(H_L[v]))));  out = (1 - ema) out + ema previous[v];  previous[v] = out[v]
// unit() divides by max(norm, 1e-5).  The two dense layers are ncf.cu's fp32 GEMM (the reference computes in fp32);
// everything else is a warp per node, HBM-bound gathers.  The last stage runs in node chunks so the 4F-wide hidden
// layer never exists for the whole graph.
#include "common.cuh"
#include "kernels.h"

namespace hwer {

namespace {

constexpr int kGcnWarps = 8;

// pre[v] = LeakyReLU_0.1(content W^T + b) (from the GEMM)  ->  h0[v]
__global__ void __launch_bounds__(kGcnWarps * 32)
gcn_h0_kernel(const float* __restrict__ pre, const float* __restrict__ node_emb, const float* __restrict__ ln_g,
              const float* __restrict__ ln_b, long long n, int F, float* __restrict__ h0) {
    const long long v = (long long)blockIdx.x * kGcnWarps + (threadIdx.x >> 5);
    if (v >= n) return;
    const int l = lane_id();
    const float* p = pre + (size_t)v * F;
    float s = 0.f;
    for (int c = l; c < F; c += 32) s += p[c];
    const float mu = warp_sum(s) / (float)F;
    float q = 0.f;
    for (int c = l; c < F; c += 32) { const float t = p[c] - mu; q = fmaf(t, t, q); }
    const float rstd = 1.0f / sqrtf(warp_sum(q) / (float)F + 1e-5f);           // biased variance, eps 1e-5 (nn.LayerNorm)
    const float* e = node_emb + (size_t)(v + 1) * F;                            // row 0 is the padding node
    float nn = 0.f;
    for (int c = l; c < F; c += 32) {
        const float t = e[c] + ((p[c] - mu) * rstd * ln_g[c] + ln_b[c]);
        nn = fmaf(t, t, nn);
    }
    const float nrm = fmaxf(sqrtf(warp_sum(nn)), 1e-5f);
    for (int c = l; c < F; c += 32) {
        const float t = e[c] + ((p[c] - mu) * rstd * ln_g[c] + ln_b[c]);
        h0[(size_t)v * F + c] = t / nrm;
    }
}

// y[v - v0] = [ mean of x rows of nbr(v)  ||  h0[v] ]  for v in [v0, v1)
__global__ void __launch_bounds__(kGcnWarps * 32)
gcn_aggregate_kernel(const float* __restrict__ x, int din, const float* __restrict__ h0, int F,
                     const long long* __restrict__ ptr, const long long* __restrict__ idx, long long n, long long v0,
                     long long v1, float* __restrict__ y) {
    const long long v = v0 + (long long)blockIdx.x * kGcnWarps + (threadIdx.x >> 5);
    if (v >= v1) return;
    const int l = lane_id();
    const long long b = ptr[v], e = ptr[v + 1];
    const float cnt = (float)(e - b);
    float* out = y + (size_t)(v - v0) * (din + F);
    for (int c = l; c < din; c += 32) {
        float s = 0.f;
        for (long long j = b; j < e; ++j) {                     // list order, like DGL's sum reducer
            long long u = idx[j];
            if (u < 0 || u >= n) u = v;                         // defensive: a bad id degrades to the self loop
            s += x[(size_t)u * din + c];
        }
        out[c] = s / cnt;                                       // an empty list gives 0 / 0 like the reference
    }
    for (int c = l; c < F; c += 32) out[din + c] = h0[(size_t)v * F + c];
}

// z[v - v0] -> out[v] = unit(z), EMA with previous[v], previous[v] = out[v]
__global__ void __launch_bounds__(kGcnWarps * 32)
gcn_finish_kernel(const float* __restrict__ z, int F, long long v0, long long v1, float* __restrict__ previous,
                  float ema, float* __restrict__ out) {
    const long long v = v0 + (long long)blockIdx.x * kGcnWarps + (threadIdx.x >> 5);
    if (v >= v1) return;
    const int l = lane_id();
    const float* p = z + (size_t)(v - v0) * F;
    float nn = 0.f;
    for (int c = l; c < F; c += 32) nn = fmaf(p[c], p[c], nn);
    const float nrm = fmaxf(sqrtf(warp_sum(nn)), 1e-5f);
    for (int c = l; c < F; c += 32) {
        float t = p[c] / nrm;
        if (previous) {
            t = (1.0f - ema) * t + ema * previous[(size_t)v * F + c];          // indexed by node id (gcn.py:188)
            previous[(size_t)v * F + c] = t;
        }
        out[(size_t)v * F + c] = t;
    }
}

inline unsigned gcn_grid(long long rows) { return (unsigned)((rows + kGcnWarps - 1) / kGcnWarps); }

}  // namespace

// scratch (floats): h0 [n, F] + two ping-pong arrays [n, F * layers] (only for layers >= 2) + per-chunk
// [chunk, F (layers + 1)] + [chunk, 4F] + [chunk, F]; the projection's [chunk, F] reuses the hidden array.
size_t gcn_infer_scratch_floats(long long n, int F, int layers, long long chunk) {
    const size_t full = (size_t)n * F + (layers >= 2 ? 2 * (size_t)n * F * layers : 0);
    return full + (size_t)chunk * ((size_t)F * (layers + 1) + 4 * (size_t)F + F);
}

cudaError_t launch_gcn_infer(const float* node_emb, const float* content, int C, const float* proj_w, const float* proj_b,
                             const float* ln_g, const float* ln_b, long long n, int F, int layers,
                             const long long* const* nbr_ptr, const long long* const* nbr_idx, const float* fc0_w,
                             const float* fc0_b, const float* fc1_w, const float* fc1_b, float* previous, float ema,
                             float* out, float* scratch, long long chunk, cudaStream_t stream) {
    float* h0 = scratch;
    float* ping = h0 + (size_t)n * F;
    float* pong = ping + (layers >= 2 ? (size_t)n * F * layers : 0);
    float* last_in = pong + (layers >= 2 ? (size_t)n * F * layers : 0);          // [chunk, F (layers + 1)]
    float* hidden = last_in + (size_t)chunk * F * (layers + 1);                  // [chunk, 4F]
    float* z = hidden + (size_t)chunk * 4 * F;                                   // [chunk, F]
    cudaError_t e;
    // h0 for every node, in chunks (the GEMM output is the chunk's `hidden` array)
    for (long long v0 = 0; v0 < n; v0 += chunk) {
        const long long rows = n - v0 < chunk ? n - v0 : chunk;
        e = launch_linear_f32(content + (size_t)v0 * C, proj_w, proj_b, hidden, rows, C, F, 0.1f, stream);
        if (e != cudaSuccess) return e;
        gcn_h0_kernel<<<gcn_grid(rows), kGcnWarps * 32, 0, stream>>>(hidden, node_emb + (size_t)v0 * F, ln_g, ln_b, rows, F,
                                                                    h0 + (size_t)v0 * F);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
    }
    // blocks 1 .. layers - 1 over the whole graph: H_i = [mean H_{i-1} || h0]
    const float* x = h0;
    int din = F;
    for (int i = 0; i + 1 < layers; ++i) {
        float* y = (i & 1) ? pong : ping;
        gcn_aggregate_kernel<<<gcn_grid(n), kGcnWarps * 32, 0, stream>>>(x, din, h0, F, nbr_ptr[i], nbr_idx[i], n, 0, n, y);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
        x = y;
        din += F;
    }
    // the prediction block, chunk by chunk: aggregate -> fc0 + LeakyReLU -> fc1 -> unit + EMA
    for (long long v0 = 0; v0 < n; v0 += chunk) {
        const long long v1 = v0 + chunk < n ? v0 + chunk : n, rows = v1 - v0;
        gcn_aggregate_kernel<<<gcn_grid(rows), kGcnWarps * 32, 0, stream>>>(x, din, h0, F, nbr_ptr[layers - 1],
                                                                           nbr_idx[layers - 1], n, v0, v1, last_in);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
        e = launch_linear_f32(last_in, fc0_w, fc0_b, hidden, rows, din + F, 4 * F, 0.01f, stream);
        if (e != cudaSuccess) return e;
        e = launch_linear_f32(hidden, fc1_w, fc1_b, z, rows, 4 * F, F, 1.0f, stream);
        if (e != cudaSuccess) return e;
        gcn_finish_kernel<<<gcn_grid(rows), kGcnWarps * 32, 0, stream>>>(z, F, v0, v1, previous, ema, out);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
    }
    return cudaSuccess;
}

}  // namespace hwer
