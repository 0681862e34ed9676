// extern "C" entry points of include/hwer_b200.h: argument checking, workspace
// management, the TMA tensor map, and the round schedule of the fused
// score-and-select (filter -> select ... -> final).
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

#include "../../include/hwer_b200.h"
#include "kernels.h"

namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}
int fail_cuda(cudaError_t e, const char* what) {
    g_err = std::string(what) + ": " + cudaGetErrorString(e);
    return HWER_E_CUDA;
}
#define HWER_CUDA(call)                                    \
    do {                                                   \
        cudaError_t e_ = (call);                           \
        if (e_ != cudaSuccess) return fail_cuda(e_, #call); \
    } while (0)

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

long long gcd_ll(long long a, long long b) {
    while (b) { long long t = a % b; a = b; b = t; }
    return a;
}

}  // namespace

struct hwer_index {
    int device = 0;
    int num_sms = 0;
    const float* table = nullptr;
    const void* shadow = nullptr;
    long long n = 0;
    int d = 0, d_pad = 0;
    float max_norm = 1.0f;
    bool use_tc = false;
    CUtensorMap tmap;
    long long n_tiles = 0;
    long long tile_mul = 1;
    // workspace, grown on demand
    unsigned long long* cand = nullptr;
    unsigned int* cnt = nullptr;
    float* thr = nullptr;
    float* margin = nullptr;
    float* floor = nullptr;
    float* qmax = nullptr;
    size_t ws_queries = 0, ws_cap = 0;
    uint4* spill = nullptr;                // hit entries of the filter rounds (FilterParams::spill)
    unsigned int* spill_cnt = nullptr;
    int spill_cap = 0;
    unsigned int* needed_dev = nullptr;
    unsigned int* needed_host = nullptr;   // pinned
    unsigned int last_cap = 0;
    unsigned int* ovf = nullptr;           // [ws_queries] sticky per-query "candidate list overflowed" marks of a call
    // tuning knobs (HWER_FIRST_ROWS / HWER_GROWTH / HWER_LATE_ROWS), read once when the index is created
    long long env_first_rows = 0, env_late_rows = 0;
    int env_growth = 0;
    long long env_narrow_tiles = 0;
    // HWER_DISABLE bit mask for A/B runs (measured in profiles/r02_*_ab.txt): 1 = bias MMA instead of scaled queries,
    // 2 = only the full-capacity final shape; and three alternatives that measured no better and are off by default:
    // 4 = warp-per-query final, 8 = spill extraction fused into the filter kernel, 16 = warp-per-query dense select;
    // 32 = ENABLE extra CTAs on the SMs left over by slots * query blocks (148 - 9 * 16 = 4 at B = 4096).  Built and
    // measured in r2 (profiles/r02_s_ab_extra_ctas.txt): the four extra CTAs walk the tail tiles once per query block,
    // out of step with everyone else -- DRAM reads per step rose from 3.1 to 4.4 GB, the mid rounds got 10-20 % slower
    // and the long round did not change, so the grid stays at slots * query blocks (144 CTAs) by default
    int env_disable = 0;
    // optional live profiling of the dominant (filter) kernel with CUDA events on the launching stream
    bool prof = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev;
    std::vector<int> ev_stage;             // what each bracket timed: 0 filter (+ extract), 1 select, 2 final, 3 exchange
    size_t ev_used = 0;
    long long filter_launches = 0, other_launches = 0;
};

namespace {

int ensure_workspace(hwer_index* ix, size_t queries, size_t cap) {
    if (queries <= ix->ws_queries && cap <= ix->ws_cap) return HWER_OK;
    // grow-only in each dimension (alternating call shapes must not reallocate on every call), bounded by the
    // ~1 GiB the chunking of topk_impl aims at: when the product would exceed it, keep exactly what is asked for
    size_t q = queries > ix->ws_queries ? queries : ix->ws_queries;
    size_t c = cap > ix->ws_cap ? cap : ix->ws_cap;
    if (q * c * sizeof(unsigned long long) > ((size_t)5 << 28)) { q = queries; c = cap; }
    HWER_CUDA(cudaDeviceSynchronize());
    if (ix->cand) cudaFree(ix->cand);
    if (ix->cnt) cudaFree(ix->cnt);
    if (ix->thr) cudaFree(ix->thr);
    if (ix->margin) cudaFree(ix->margin);
    if (ix->floor) cudaFree(ix->floor);
    if (ix->ovf) cudaFree(ix->ovf);
    if (ix->qmax) cudaFree(ix->qmax);
    ix->floor = nullptr; ix->ovf = nullptr; ix->qmax = nullptr;
    ix->cand = nullptr; ix->cnt = nullptr; ix->thr = nullptr; ix->margin = nullptr;
    ix->ws_queries = ix->ws_cap = 0;
    if (cudaMalloc(&ix->cand, q * c * sizeof(unsigned long long)) != cudaSuccess ||
        cudaMalloc(&ix->cnt, q * sizeof(unsigned int)) != cudaSuccess ||
        cudaMalloc(&ix->ovf, q * sizeof(unsigned int)) != cudaSuccess ||
        cudaMalloc(&ix->qmax, q * sizeof(float)) != cudaSuccess ||
        cudaMalloc(&ix->thr, q * sizeof(float)) != cudaSuccess ||
        cudaMalloc(&ix->margin, q * sizeof(float)) != cudaSuccess ||
        cudaMalloc(&ix->floor, q * sizeof(float)) != cudaSuccess) {
        cudaGetLastError();
        return fail(HWER_E_NOMEM, "hwer_topk: cannot allocate the candidate workspace");
    }
    ix->ws_queries = q;
    ix->ws_cap = c;
    return HWER_OK;
}

// Per-thread spill buffers of the tensor-core filter: sized for twice the hits a thread expects in one round
// (~1.4 k g candidates per query and round, spread over every epilogue thread of the grid), 16..128 entries.
int ensure_spill(hwer_index* ix, int Bc, int k, int growth) {
    if (!ix->use_tc) return HWER_OK;
    const double per_thread = 1.4 * k * growth * (double)Bc / (double)hwer::filter_tc_spill_buffers(ix->num_sms);
    // ceiling: 128 entries per thread (465 MB), or 512 (1.9 GB) when HBM has room -- top-1000 at batch 4096
    // (config C5) expects ~150 entries per thread and round, and a full buffer means blocking appends
    int want = 16;
    while (want < 2.0 * per_thread + 8.0 && want < 512) want <<= 1;
    if (want <= ix->spill_cap) return HWER_OK;      // grow-only: a larger buffer serves every smaller request
    if (want > 128) {                               // (asked only when growing: cudaMemGetInfo costs ~1 ms)
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess || free_b < ((size_t)16 << 30)) want = 128;
        if (want <= ix->spill_cap) return HWER_OK;
    }
    HWER_CUDA(cudaDeviceSynchronize());
    if (ix->spill) cudaFree(ix->spill);
    if (ix->spill_cnt) cudaFree(ix->spill_cnt);
    ix->spill = nullptr; ix->spill_cnt = nullptr; ix->spill_cap = 0;
    if (cudaMalloc(&ix->spill, hwer::filter_tc_spill_entries(ix->num_sms, want) * 48) != cudaSuccess ||
        cudaMalloc(&ix->spill_cnt, hwer::filter_tc_spill_buffers(ix->num_sms) * sizeof(unsigned int)) != cudaSuccess) {
        cudaGetLastError();
        return fail(HWER_E_NOMEM, "hwer_topk: cannot allocate the spill buffers");
    }
    ix->spill_cap = want;
    return HWER_OK;
}

enum : int { kStageFilter = 0, kStageSelect = 1, kStageFinal = 2, kStageExchange = 3, kStages = 4 };

bool prof_begin(hwer_index* ix, cudaStream_t stream, int stage = kStageFilter) {
    if (!ix->prof) return false;
    if (ix->ev_used == ix->ev.size()) {
        cudaEvent_t a, b;
        if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) return false;
        ix->ev.emplace_back(a, b);
        ix->ev_stage.push_back(stage);
    }
    ix->ev_stage[ix->ev_used] = stage;
    cudaEventRecord(ix->ev[ix->ev_used].first, stream);
    return true;
}
void prof_end(hwer_index* ix, cudaStream_t stream, bool began) {
    if (began) cudaEventRecord(ix->ev[ix->ev_used++].second, stream);
}

struct Schedule {
    int growth;             // early rounds: score `growth` x the rows seen so far
    long long late_tiles;   // once this many tiles have been seen, rounds only double (growth 1)
    unsigned int cap;
    long long first_tiles;
};

// Round schedule: round 0 scores `first_tiles` tiles densely (open threshold, positional writes); each later round
// scores `growth` times what has been seen, so it admits about 1.4 * k * growth candidates per query however long it
// is (DESIGN.md "Rounds").  A hit costs the epilogue far more than a miss, a round costs a launch plus a
// select_compact pass, and both scale differently with the batch: the table below is measured on the C4 catalogue
// (scripts/tune_schedule.py, profiles/r01_v6_tune.txt).
int make_schedule(const hwer_index* ix, int B, int k, unsigned int cap_user, int world_share, Schedule* s) {
    const unsigned int max_cap = 16384;   // bounded by the shared-memory sort in select.cu
    // (re-tuned with the r2 kernels, profiles/r02_an_tune.txt: hits got cheaper, so batches up to 128 take growth 32
    // after an 8192-row round 0 -- B = 64: 0.565 -> 0.50 ms per step; B = 4096 keeps 4096 rows / growth 2)
    long long first_rows = B > 512 ? 4096 : 8192;
    int g = B <= 128 ? 32 : (B <= 512 ? 4 : 2);
    // large batches of large k are bound by hit handling, not by launches: plain doubling admits 1.4 k per round
    // instead of 2.8 k (C5 shard, 62.5 M rows, k = 1000, B = 4096: 59.1 -> 58.2 ms per step; growth 3: 64.8 ms)
    if (B > 512 && k >= 512) g = 1;
    // round 0 must leave every list with the k candidates its select needs (shards that share thresholds publish
    // their ceil(k / G)-th best, so they need that many)
    const long long k_need = world_share > 1 ? (k + world_share - 1) / world_share : k;
    // sharing shards see the union of G round-0 samples: a shorter dense round gives the same first bound
    if (world_share >= 8 && B > 512) first_rows = 2048;
    if (world_share > 1 && B > 16 && B <= 128) first_rows = 2048;
    if (first_rows < 2LL * k_need) first_rows = 2LL * k_need;
    if (ix->env_first_rows >= 2LL * k_need && ix->env_first_rows <= max_cap) first_rows = ix->env_first_rows;   // tuning knob
    s->first_tiles = (first_rows + hwer::kTileItems - 1) / hwer::kTileItems;
    first_rows = s->first_tiles * hwer::kTileItems;
    while (g > 1 && 3LL * k * g > max_cap) g >>= 1;
    if (ix->env_growth >= 1 && 3LL * k * ix->env_growth <= max_cap) g = ix->env_growth;   // tuning knob
    unsigned long long want = 3ULL * k * g;
    if (want < (unsigned long long)first_rows) want = first_rows;
    if (cap_user) {
        if (cap_user < first_rows) return fail(HWER_E_INVALID, "hwer_topk: cap smaller than the first round");
        want = cap_user;
    }
    unsigned long long cap = 1024;
    while (cap < want) cap <<= 1;
    if (cap > max_cap) return fail(HWER_E_INVALID, "hwer_topk: k (or cap) too large for the shared-memory selector");
    // Item shards that share thresholds every round (hwer_topk_sharded) each admit ~1/G of a round's candidates, so
    // their rounds can grow G times faster for the same lists: a 1.25 M-row shard of an 8-way split needs 3 filter
    // launches instead of 7, and every launch saved is ~100 us of extract + select + exchange on a ~2 ms step.
    // Measured on 2 / 4 / 8 B200 (scripts/tune_schedule_sharded.py, profiles/r02_k_tune_n*.txt): the growth that pays
    // stops at 8 for large batches (10 M rows over 8 GPUs, B = 4096: 1.40 ms per step at growth 8 after a 2048-row
    // round 0, 1.55 ms at 16 after 4096 rows) and at 32 for small ones.
    if (world_share > 1 && ix->env_growth < 1) {
        long long ge = (long long)g * world_share;
        const long long g_max = B > 512 ? 8 : (B > 128 ? 64 : 32);
        g = (int)(ge > g_max ? g_max : ge);
    }
    s->growth = g;
    s->late_tiles = 1LL << 40;            // optional switch to plain doubling once this many tiles have been seen
    if (ix->env_late_rows > 0) s->late_tiles = ix->env_late_rows / hwer::kTileItems;   // tuning knob
    s->cap = (unsigned int)cap;
    return HWER_OK;
}

}  // namespace

extern "C" {

const char* hwer_last_error(void) { return g_err.c_str(); }
int hwer_version(void) { return 100; }
int32_t hwer_shadow_width(int32_t d) { return (d + 63) / 64 * 64; }

int hwer_blend_normalize(const float* content_dev, const float* collab_dev, float alpha, const float* alpha_rows_dev,
                         int64_t n, int32_t d, float* out_f32_dev, void* out_bf16_dev, int32_t d_pad, void* stream) {
    if (!collab_dev || !out_f32_dev || n < 0 || d <= 0) return fail(HWER_E_INVALID, "hwer_blend_normalize: bad argument");
    if (out_bf16_dev && d_pad < d) return fail(HWER_E_INVALID, "hwer_blend_normalize: d_pad < d");
    HWER_CUDA(hwer::launch_blend_normalize(content_dev, collab_dev, alpha, alpha_rows_dev, n, d, out_f32_dev,
                                           out_bf16_dev, d_pad, (cudaStream_t)stream));
    return HWER_OK;
}

int hwer_make_shadow(const float* table_dev, int64_t n, int32_t d, void* out_bf16_dev, int32_t d_pad, void* stream) {
    if (!table_dev || !out_bf16_dev || n <= 0 || d <= 0 || d_pad < d || d_pad % 2)
        return fail(HWER_E_INVALID, "hwer_make_shadow: bad argument");
    HWER_CUDA(hwer::launch_make_shadow(table_dev, n, d, out_bf16_dev, d_pad, (cudaStream_t)stream));
    return HWER_OK;
}

int hwer_norm_stats(const float* table_dev, int64_t n, int32_t d, float epsilon, double* out5_dev, void* stream) {
    if (!table_dev || !out5_dev || n <= 0 || d <= 0) return fail(HWER_E_INVALID, "hwer_norm_stats: bad argument");
    HWER_CUDA(hwer::launch_norm_stats(table_dev, n, d, epsilon, out5_dev, (cudaStream_t)stream));
    return HWER_OK;
}

int hwer_index_create(hwer_index_t** out, const float* table_f32_dev, const void* shadow_bf16_dev, int64_t n,
                      int32_t d, int32_t d_pad, float max_norm, int32_t device) {
    if (!out || !table_f32_dev || n <= 0 || d <= 0) return fail(HWER_E_INVALID, "hwer_index_create: bad argument");
    if (n >= (1LL << 31) - 256) return fail(HWER_E_INVALID, "hwer_index_create: shard too large (rows must fit 31 bits)");
    HWER_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    HWER_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return fail(HWER_E_ARCH, "hwer_index_create: device is not sm_100 (B200)");
    hwer_index* ix = new hwer_index();
    ix->device = device;
    ix->num_sms = prop.multiProcessorCount;
    ix->table = table_f32_dev;
    ix->shadow = shadow_bf16_dev;
    ix->n = n;
    ix->d = d;
    ix->d_pad = d_pad;
    ix->max_norm = (max_norm > 0.0f && max_norm == max_norm) ? max_norm : 1.0f;
    ix->n_tiles = (n + hwer::kTileItems - 1) / hwer::kTileItems;
    ix->use_tc = shadow_bf16_dev != nullptr && d_pad >= d && d_pad % 64 == 0 && d_pad <= 256;
    if (shadow_bf16_dev && !ix->use_tc && d_pad <= 256) {
        delete ix;
        return fail(HWER_E_INVALID, "hwer_index_create: d_pad must be a multiple of 64 and >= d");
    }
    if (ix->use_tc) {
        EncodeTiledFn enc = encode_tiled_fn();
        if (!enc) { delete ix; return fail(HWER_E_CUDA, "cuTensorMapEncodeTiled entry point not found"); }
        cuuint64_t dims[2] = {(cuuint64_t)d_pad, (cuuint64_t)n};
        cuuint64_t strides[1] = {(cuuint64_t)d_pad * 2};
        cuuint32_t box[2] = {(cuuint32_t)hwer::kKBlock, (cuuint32_t)hwer::kTileItems};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = enc(&ix->tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(shadow_bf16_dev), dims,
                         strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            delete ix;
            char buf[96];
            snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
            return fail(HWER_E_CUDA, buf);
        }
        // Visit tiles in a scrambled order so early rounds sample the whole catalogue even when rows are
        // sorted (clustered catalogues): tile t -> (t * mul) mod n_tiles with gcd(mul, n_tiles) = 1.
        long long mul = (long long)((double)ix->n_tiles * 0.6180339887) | 1;
        while (mul > 1 && gcd_ll(mul, ix->n_tiles) != 1) mul -= 2;
        if (mul < 1 || ix->n_tiles < 8) mul = 1;
        ix->tile_mul = mul;
    }
    // [0] largest candidate-list demand seen by any kernel of the pending calls, [1] a wait on a peer GPU timed out
    if (cudaMalloc(&ix->needed_dev, 2 * sizeof(unsigned int)) != cudaSuccess ||
        cudaMallocHost(&ix->needed_host, 2 * sizeof(unsigned int)) != cudaSuccess) {
        delete ix;
        return fail(HWER_E_NOMEM, "hwer_index_create: allocation failed");
    }
    cudaMemset(ix->needed_dev, 0, 2 * sizeof(unsigned int));
    ix->needed_host[0] = ix->needed_host[1] = 0;
    if (const char* e = getenv("HWER_FIRST_ROWS")) ix->env_first_rows = atoll(e);
    if (const char* e = getenv("HWER_GROWTH")) ix->env_growth = atoi(e);
    if (const char* e = getenv("HWER_LATE_ROWS")) ix->env_late_rows = atoll(e);
    if (const char* e = getenv("HWER_DISABLE")) ix->env_disable = atoi(e);
    if (const char* e = getenv("HWER_NARROW_TILES")) ix->env_narrow_tiles = atoll(e);
    *out = ix;
    return HWER_OK;
}

int hwer_index_destroy(hwer_index_t* ix) {
    if (!ix) return HWER_OK;
    cudaSetDevice(ix->device);
    cudaDeviceSynchronize();
    if (ix->cand) cudaFree(ix->cand);
    if (ix->cnt) cudaFree(ix->cnt);
    if (ix->thr) cudaFree(ix->thr);
    if (ix->margin) cudaFree(ix->margin);
    if (ix->floor) cudaFree(ix->floor);
    if (ix->ovf) cudaFree(ix->ovf);
    if (ix->qmax) cudaFree(ix->qmax);
    if (ix->spill) cudaFree(ix->spill);
    if (ix->spill_cnt) cudaFree(ix->spill_cnt);
    if (ix->needed_dev) cudaFree(ix->needed_dev);
    if (ix->needed_host) cudaFreeHost(ix->needed_host);
    for (auto& e : ix->ev) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
    delete ix;
    return HWER_OK;
}

}  // extern "C"

namespace {

// The whole single-GPU pipeline (rounds of filter + select, then final).  With `peer` set, final_kernel stores
// each query's result into the exchange buffer of the GPU that merges it instead of the local out arrays.
int topk_impl(hwer_index_t* ix, const float* queries_dev, int32_t B, int32_t k, int32_t mode, uint32_t cap,
              int64_t idx_offset, int64_t* out_idx_dev, float* out_score_dev, double* out_score64_dev,
              const hwer::PeerDst* peer, const hwer::ExchangeView* xv, unsigned int call_epoch, void* stream_v) {
    if (!ix || B < 0 || k <= 0 || (B > 0 && (!queries_dev || (!peer && (!out_idx_dev || !out_score_dev)))))
        return fail(HWER_E_INVALID, "hwer_topk: bad argument");
    if (mode != HWER_MODE_EXACT && mode != HWER_MODE_BF16) return fail(HWER_E_INVALID, "hwer_topk: unknown mode");
    if (mode == HWER_MODE_BF16 && !ix->use_tc)
        return fail(HWER_E_INVALID, "hwer_topk: bf16 mode needs a bf16 shadow table (d_pad <= 256)");
    if ((long long)k > ix->n) return fail(HWER_E_K_TOO_LARGE, "hwer_topk: k exceeds the number of rows in the index");
    if (B == 0) return HWER_OK;
    cudaStream_t stream = (cudaStream_t)stream_v;
    HWER_CUDA(cudaSetDevice(ix->device));

    const int world_share = (xv && xv->world > 1) ? xv->world : 1;
    Schedule sch;
    int rc = make_schedule(ix, B, k, cap, world_share, &sch);
    if (rc) return rc;
    // bound the candidate workspace to ~1 GiB by chunking the query batch
    size_t chunk = ((size_t)1 << 30) / ((size_t)sch.cap * 8);
    chunk = chunk / 256 * 256;
    if (chunk < 256) chunk = 256;
    if (chunk > (size_t)B) chunk = B;
    rc = ensure_workspace(ix, chunk, sch.cap);
    if (rc) return rc;
    rc = ensure_spill(ix, (int)chunk, k, (sch.growth + world_share - 1) / world_share);
    if (rc) return rc;
    ix->last_cap = sch.cap;

    const bool exact = mode == HWER_MODE_EXACT;
    const float eps_rel = ix->use_tc ? hwer::bf16_score_eps_rel(ix->d_pad) : hwer::f32_score_eps_rel(ix->d);
    const float margin_factor = 2.0f * eps_rel * ix->max_norm * 1.0001f;
    const long long T = ix->n_tiles;
    // Sharded catalogue: every rank walks the SAME round schedule (that of the largest shard) because the rounds
    // exchange thresholds; a rank whose shard is a tile shorter simply scores an empty range in the last round.
    const bool share_thr = xv && xv->world > 1;
    long long T_sched = T;
    if (share_thr && xv->sched_rows > 0) T_sched = (xv->sched_rows + hwer::kTileItems - 1) / hwer::kTileItems;
    if (T_sched < T) T_sched = T;
    const int k_share = share_thr ? (k + xv->world - 1) / xv->world : k;
    int chunk_idx = 0;

    for (long long q0 = 0; q0 < B; q0 += (long long)chunk, ++chunk_idx) {
        const int Bc = (int)(((long long)B - q0) < (long long)chunk ? ((long long)B - q0) : (long long)chunk);
        const float* Q = queries_dev + (size_t)q0 * ix->d;
        // the CUDA-core path scores in fp32: its (tiny) margin keeps even bf16-less indexes exact
        const float* margin = (exact || !ix->use_tc) ? ix->margin : nullptr;
        // one launch: margin / floor / overflow guard per query + reset of its count, overflow mark and threshold
        HWER_CUDA(hwer::launch_query_margin(Q, Bc, ix->d, margin_factor, ix->max_norm, ix->margin, ix->floor, ix->qmax,
                                            ix->cnt, ix->ovf, ix->thr, stream));
        ix->other_launches += 2;   // set-up + final
        long long seen = 0;
        int round = 0;
        while (seen < T_sched) {
            long long take = round == 0 ? sch.first_tiles : seen * (seen >= sch.late_tiles ? 1 : sch.growth);
            long long end_s = seen + take;
            if (end_s > T_sched || (T_sched - end_s) * 4 < take) end_s = T_sched;   // absorb a short remainder ...
            if (round == 0 && end_s * hwer::kTileItems > (long long)sch.cap)        // ... unless the dense round would
                end_s = seen + take < T_sched ? seen + take : T_sched;              // outgrow the candidate lists
            const long long end = end_s < T ? end_s : T;                            // this shard's share of the round
            const long long seen_l = seen < T ? seen : T;
            const bool timed = prof_begin(ix, stream);
            if (ix->use_tc) {
                hwer::FilterParams p;
                memset(&p, 0, sizeof p);
                p.queries = Q; p.B = Bc; p.d = ix->d; p.kb = ix->d_pad / 64;
                int nq = (Bc + 31) / 32 * 32;     // whole 32-column epilogue chunks
                int nq_max = ix->d_pad > 192 ? 128 : hwer::kMaxNQ;
                // A/B knob (HWER_NARROW_TILES = t): rounds of fewer than t tiles use 128-query blocks, i.e. four
                // accumulator stages instead of two -- more room to absorb hit-handling latency where hits are dense
                if (ix->env_narrow_tiles > 0 && round > 0 && (end - seen_l) < ix->env_narrow_tiles) nq_max = 128;
                if (nq > nq_max) nq = nq_max;
                p.nq = nq; p.nqb = (Bc + nq - 1) / nq;
                // bf16 mode returns the tensor-core scores themselves: they must be products of the RN-bf16 query
                // (not of a per-round rescaled one), so it keeps the bias MMA
                p.thr = ix->thr; p.floor = ix->floor; p.qmax = ix->qmax; p.force_bias = (ix->env_disable & 1) | (exact ? 0 : 1); p.cand = ix->cand; p.cnt = ix->cnt; p.cap = sch.cap;
                p.n_items = ix->n; p.tile_begin = (int)seen_l; p.tile_end = (int)end;
                p.tile_mul = ix->tile_mul; p.tile_mod = T;
                p.dense = round == 0 ? 1 : 0;      // open threshold: positional writes, no atomics
                p.spill = ix->spill; p.spill_cnt = ix->spill_cnt; p.spill_cap = ix->spill_cap; p.spill_ctas = ix->num_sms;
                p.fused_extract = (ix->env_disable & 8) ? 1 : 0;
                p.no_extra_ctas = (ix->env_disable & 32) ? 0 : 1;
                HWER_CUDA(hwer::launch_filter_tc(ix->tmap, p, ix->num_sms, stream));
                if (round > 0 && !p.fused_extract) ix->other_launches += 1;     // spill_extract_kernel behind the filter
            } else {
                long long rb = seen_l * hwer::kTileItems, re = end * hwer::kTileItems;
                if (re > ix->n) re = ix->n;
                HWER_CUDA(hwer::launch_filter_simt(ix->table, ix->n, ix->d, Q, Bc, ix->thr, ix->cand, ix->cnt, sch.cap,
                                                   rb, re, ix->num_sms, stream));
            }
            prof_end(ix, stream, timed);
            const int fixed = (ix->use_tc && round == 0) ? (int)((end - seen_l) * hwer::kTileItems) : -1;
            const bool timed_sel = prof_begin(ix, stream, kStageSelect);
            hwer::SelExchange sx;
            memset(&sx, 0, sizeof sx);
            if (share_thr) {
                // each shard publishes its ceil(k/G)-th best score so far to every peer; the min over shards bounds the
                // global k-th best from below (DESIGN.md "Multi-GPU"), so all shards filter with ~1/G the hits.
                // Publish, wait and compaction happen inside the one select launch, query by query.
                sx.mode = hwer::kSelShared;
                sx.world = xv->world; sx.rank = xv->rank; sx.b_cap = xv->b_cap; sx.k_share = k_share; sx.q0 = q0;
                sx.epoch = (call_epoch * 64u + (unsigned int)chunk_idx) * 64u + (unsigned int)round + 1u;
                sx.flags = xv->flags[xv->rank];
                for (int r = 0; r < xv->world; ++r) sx.thr_x[r] = xv->thr_x[r];
            }
            HWER_CUDA(hwer::launch_select_compact(ix->cand, ix->cnt, sch.cap, Bc, k, fixed, margin, ix->thr,
                                                  ix->needed_dev, ix->ovf, share_thr ? &sx : nullptr,
                                                  (ix->env_disable & 16) ? 1 : 0, stream));
            prof_end(ix, stream, timed_sel);
            ix->filter_launches += 1;
            ix->other_launches += 1;
            seen = end_s;
            ++round;
        }
        hwer::PeerDst pd;
        if (peer) { pd = *peer; pd.q0 = q0; }
        // small batches keep the 512-thread shape: with a handful of queries the latency of ONE list is the stage's time
        const bool timed_fin = prof_begin(ix, stream, kStageFinal);
        HWER_CUDA(hwer::launch_final(ix->cand, ix->cnt, sch.cap, Bc, k, exact ? 1 : 0, ix->table, ix->d, Q, idx_offset,
                                     peer ? nullptr : (long long*)out_idx_dev + (size_t)q0 * k,
                                     peer ? nullptr : out_score_dev + (size_t)q0 * k,
                                     (!peer && out_score64_dev) ? out_score64_dev + (size_t)q0 * k : nullptr,
                                     ix->needed_dev, ix->ovf, peer ? &pd : nullptr,
                                     (ix->env_disable & 2) || Bc < 256 ? 0 : ((ix->env_disable & 4) ? 1 : 2), stream));
        prof_end(ix, stream, timed_fin);
    }
    return HWER_OK;
}

}  // namespace

struct hwer_exchange {
    hwer::ExchangeView v;
    unsigned int epoch = 0;
    int device = 0;
    bool share_thresholds = true;
};

namespace {

size_t align256(size_t x) { return (x + 255) / 256 * 256; }

struct ExchangeLayout {
    size_t flags, xs, xi, out_idx, out_score64, out_score, thr_x, total;
    int q_cap;
};

ExchangeLayout exchange_layout(int world, int b_cap, int k_cap) {
    ExchangeLayout L;
    L.q_cap = (b_cap + world - 1) / world;
    const size_t x = (size_t)world * L.q_cap * k_cap * 8, o = (size_t)b_cap * k_cap;
    size_t off = 0;
    L.flags = off; off += 256;
    L.xs = off; off += align256(x);
    L.xi = off; off += align256(x);
    L.out_idx = off; off += align256(o * 8);
    L.out_score64 = off; off += align256(o * 8);
    L.out_score = off; off += align256(o * 4);
    L.thr_x = off; off += align256((size_t)world * b_cap * sizeof(unsigned long long));
    L.total = off;
    return L;
}

}  // namespace

extern "C" {

int hwer_topk(hwer_index_t* ix, const float* queries_dev, int32_t B, int32_t k, int32_t mode, uint32_t cap,
              int64_t idx_offset, int64_t* out_idx_dev, float* out_score_dev, double* out_score64_dev, void* stream_v) {
    return topk_impl(ix, queries_dev, B, k, mode, cap, idx_offset, out_idx_dev, out_score_dev, out_score64_dev, nullptr,
                     nullptr, 0u, stream_v);
}

int64_t hwer_exchange_bytes(int32_t world, int32_t b_cap, int32_t k_cap) {
    if (world < 1 || world > hwer::kMaxPeers || b_cap < 1 || k_cap < 1) return -1;
    return (int64_t)exchange_layout(world, b_cap, k_cap).total;
}

int hwer_peer_alloc(int64_t bytes, void** dev_ptr_out, unsigned char* handle_out) {
    if (bytes <= 0 || !dev_ptr_out || !handle_out) return fail(HWER_E_INVALID, "hwer_peer_alloc: bad argument");
    void* p = nullptr;
    if (cudaMalloc(&p, (size_t)bytes) != cudaSuccess) {
        cudaGetLastError();
        return fail(HWER_E_NOMEM, "hwer_peer_alloc: cudaMalloc failed");
    }
    HWER_CUDA(cudaMemset(p, 0, (size_t)bytes));
    HWER_CUDA(cudaDeviceSynchronize());
    cudaIpcMemHandle_t h;
    static_assert(sizeof(h) == HWER_IPC_HANDLE_BYTES, "IPC handle size");
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) { cudaFree(p); return fail_cuda(e, "cudaIpcGetMemHandle"); }
    memcpy(handle_out, &h, sizeof h);
    *dev_ptr_out = p;
    return HWER_OK;
}

int hwer_peer_open(const unsigned char* handle, void** dev_ptr_out) {
    if (!handle || !dev_ptr_out) return fail(HWER_E_INVALID, "hwer_peer_open: bad argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof h);
    void* p = nullptr;
    HWER_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    *dev_ptr_out = p;
    return HWER_OK;
}

int hwer_peer_close(void* dev_ptr) {
    if (dev_ptr) HWER_CUDA(cudaIpcCloseMemHandle(dev_ptr));
    return HWER_OK;
}

int hwer_peer_free(void* dev_ptr) {
    if (dev_ptr) HWER_CUDA(cudaFree(dev_ptr));
    return HWER_OK;
}

int hwer_exchange_create(hwer_exchange_t** out, int32_t world, int32_t rank, int32_t b_cap, int32_t k_cap,
                         void* const* bases, int32_t device) {
    if (!out || world < 1 || world > hwer::kMaxPeers || rank < 0 || rank >= world || b_cap < 1 || k_cap < 1 || !bases)
        return fail(HWER_E_INVALID, "hwer_exchange_create: bad argument");
    for (int r = 0; r < world; ++r)
        if (!bases[r]) return fail(HWER_E_INVALID, "hwer_exchange_create: null peer buffer");
    hwer_exchange* x = new hwer_exchange();
    const ExchangeLayout L = exchange_layout(world, b_cap, k_cap);
    memset(&x->v, 0, sizeof x->v);
    x->v.world = world; x->v.rank = rank; x->v.q_cap = L.q_cap; x->v.k_cap = k_cap; x->v.b_cap = b_cap;
    for (int r = 0; r < world; ++r) {
        unsigned char* b = static_cast<unsigned char*>(bases[r]);
        x->v.flags[r] = reinterpret_cast<unsigned int*>(b + L.flags);
        x->v.xs[r] = reinterpret_cast<double*>(b + L.xs);
        x->v.xi[r] = reinterpret_cast<long long*>(b + L.xi);
        x->v.out_idx[r] = reinterpret_cast<long long*>(b + L.out_idx);
        x->v.out_score64[r] = reinterpret_cast<double*>(b + L.out_score64);
        x->v.out_score[r] = reinterpret_cast<float*>(b + L.out_score);
        x->v.thr_x[r] = reinterpret_cast<unsigned long long*>(b + L.thr_x);
    }
    x->device = device;
    cudaError_t pe = hwer::exchange_preload();
    if (pe != cudaSuccess) { delete x; return fail_cuda(pe, "hwer_exchange_create: loading the exchange kernels"); }
    *out = x;
    return HWER_OK;
}

int hwer_exchange_configure(hwer_exchange_t* x, int64_t largest_shard_rows, int32_t share_thresholds) {
    if (!x || largest_shard_rows < 0) return fail(HWER_E_INVALID, "hwer_exchange_configure: bad argument");
    x->v.sched_rows = largest_shard_rows;
    x->share_thresholds = share_thresholds != 0;
    return HWER_OK;
}

int hwer_exchange_destroy(hwer_exchange_t* x) {
    delete x;
    return HWER_OK;
}

int hwer_topk_sharded(hwer_index_t* ix, hwer_exchange_t* x, const float* queries_dev, int32_t B, int32_t k, int32_t mode,
                      uint32_t cap, int64_t idx_offset, int64_t* out_idx_dev, float* out_score_dev,
                      double* out_score64_dev, int32_t phases, void* stream_v) {
    if (!ix || !x || B < 1 || k < 1 || !queries_dev || !out_idx_dev || !out_score_dev || !(phases & 7))
        return fail(HWER_E_INVALID, "hwer_topk_sharded: bad argument");
    if (B > x->v.b_cap || k > x->v.k_cap) return fail(HWER_E_INVALID, "hwer_topk_sharded: batch or k exceeds the exchange buffers");
    cudaStream_t stream = (cudaStream_t)stream_v;
    HWER_CUDA(cudaSetDevice(ix->device));
    if (phases & HWER_PHASE_SEARCH) ++x->epoch;     // every rank calls in lockstep, so they agree on it
    const unsigned int epoch = x->epoch;
    x->v.q_per_owner = (B + x->v.world - 1) / x->v.world;
    const int owned = (phases & HWER_PHASE_OWNED) ? 1 : 0;
    if (!(phases & HWER_PHASE_SEARCH)) {
        if (phases & HWER_PHASE_MERGE) HWER_CUDA(hwer::launch_exchange_merge(x->v, B, k, epoch, owned, stream));
        if (phases & HWER_PHASE_COLLECT)
            HWER_CUDA(hwer::launch_exchange_collect(x->v, B, k, epoch, (long long*)out_idx_dev, out_score_dev,
                                                    out_score64_dev, ix->needed_dev, owned, stream));
        return HWER_OK;
    }
    hwer::PeerDst pd;
    memset(&pd, 0, sizeof pd);
    pd.world = x->v.world; pd.rank = x->v.rank; pd.q_per_owner = x->v.q_per_owner; pd.q_cap = x->v.q_cap; pd.k_cap = x->v.k_cap;
    for (int r = 0; r < x->v.world; ++r) { pd.xs[r] = x->v.xs[r]; pd.xi[r] = x->v.xi[r]; }
    int rc = topk_impl(ix, queries_dev, B, k, mode, cap, idx_offset, nullptr, nullptr, nullptr, &pd,
                       // tiny batches are launch-bound, not hit-bound: the two extra launches per round do not pay
                       (x->share_thresholds && B > 16) ? &x->v : nullptr, epoch, stream_v);
    if (rc) return rc;
    const bool timed_x = prof_begin(ix, stream, kStageExchange);
    HWER_CUDA(hwer::launch_exchange_signal(x->v, 0, epoch, ix->needed_dev, stream));
    if (phases & HWER_PHASE_MERGE) HWER_CUDA(hwer::launch_exchange_merge(x->v, B, k, epoch, owned, stream));
    if (phases & HWER_PHASE_COLLECT)
        HWER_CUDA(hwer::launch_exchange_collect(x->v, B, k, epoch, (long long*)out_idx_dev, out_score_dev,
                                                out_score64_dev, ix->needed_dev, owned, stream));
    prof_end(ix, stream, timed_x);
    ix->other_launches += 3;
    return HWER_OK;
}

int hwer_exchange_error(hwer_exchange_t* x, void* stream_v) {
    if (!x) return fail(HWER_E_INVALID, "hwer_exchange_error: null exchange");
    HWER_CUDA(cudaSetDevice(x->device));
    HWER_CUDA(cudaStreamSynchronize((cudaStream_t)stream_v));
    unsigned int err = 0;
    HWER_CUDA(cudaMemcpy(&err, x->v.flags[x->v.rank] + 17, sizeof err, cudaMemcpyDeviceToHost));
    if (err) return fail(HWER_E_PEER, "hwer_topk_sharded: timed out waiting for a peer GPU");
    return HWER_OK;
}

int hwer_topk_exhaustive(hwer_index_t* ix, const float* queries_dev, int32_t B, int32_t k, int64_t idx_offset,
                         int64_t* out_idx_dev, float* out_score_dev, double* out_score64_dev, void* stream_v) {
    if (!ix || B < 0 || k <= 0 || (B > 0 && (!queries_dev || !out_idx_dev || !out_score_dev)))
        return fail(HWER_E_INVALID, "hwer_topk_exhaustive: bad argument");
    if ((long long)k > ix->n) return fail(HWER_E_K_TOO_LARGE, "hwer_topk_exhaustive: k exceeds the number of rows in the index");
    if (B == 0) return HWER_OK;
    cudaStream_t stream = (cudaStream_t)stream_v;
    HWER_CUDA(cudaSetDevice(ix->device));
    const size_t bytes = hwer::bruteforce_scratch_bytes(ix->n);
    void* scratch = nullptr;
    if (cudaMallocAsync(&scratch, bytes, stream) != cudaSuccess) {
        cudaGetLastError();
        return fail(HWER_E_NOMEM, "hwer_topk_exhaustive: cannot allocate the sort scratch");
    }
    cudaError_t e = cudaSuccess;
    for (int q = 0; q < B && e == cudaSuccess; ++q)
        e = hwer::launch_bruteforce_topk(ix->table, ix->n, ix->d, queries_dev + (size_t)q * ix->d, k, idx_offset,
                                         (long long*)out_idx_dev + (size_t)q * k, out_score_dev + (size_t)q * k,
                                         out_score64_dev ? out_score64_dev + (size_t)q * k : nullptr, scratch, bytes,
                                         stream);
    cudaFreeAsync(scratch, stream);
    if (e != cudaSuccess) return fail_cuda(e, "hwer_topk_exhaustive");
    return HWER_OK;
}

int hwer_topk_finish(hwer_index_t* ix, void* stream_v, uint32_t* needed_cap) {
    if (!ix) return fail(HWER_E_INVALID, "hwer_topk_finish: null index");
    cudaStream_t stream = (cudaStream_t)stream_v;
    HWER_CUDA(cudaSetDevice(ix->device));
    HWER_CUDA(cudaMemcpyAsync(ix->needed_host, ix->needed_dev, 2 * sizeof(unsigned int), cudaMemcpyDeviceToHost, stream));
    HWER_CUDA(cudaMemsetAsync(ix->needed_dev, 0, 2 * sizeof(unsigned int), stream));
    HWER_CUDA(cudaStreamSynchronize(stream));
    const unsigned int need = ix->needed_host[0];
    if (needed_cap) *needed_cap = need;
    if (ix->needed_host[1]) return fail(HWER_E_PEER, "hwer_topk_sharded: timed out waiting for a peer GPU");
    if (need > ix->last_cap) {
        char buf[160];
        snprintf(buf, sizeof buf, "hwer_topk: candidate lists overflowed (cap %u, needed %u): re-run with a larger cap",
                 ix->last_cap, need);
        return fail(HWER_E_OVERFLOW, buf);
    }
    return HWER_OK;
}

int hwer_profile(hwer_index_t* ix, int enable) {
    if (!ix) return fail(HWER_E_INVALID, "hwer_profile: null index");
    ix->prof = enable != 0;
    ix->ev_used = 0;
    ix->filter_launches = ix->other_launches = 0;
    return HWER_OK;
}

int hwer_profile_read(hwer_index_t* ix, void* stream_v, double* filter_ms, int64_t* filter_launches,
                      int64_t* other_launches) {
    if (!ix) return fail(HWER_E_INVALID, "hwer_profile_read: null index");
    HWER_CUDA(cudaSetDevice(ix->device));
    HWER_CUDA(cudaStreamSynchronize((cudaStream_t)stream_v));
    double ms = 0.0;
    for (size_t i = 0; i < ix->ev_used; ++i) {
        if (ix->ev_stage[i] != kStageFilter) continue;
        float t = 0.f;
        HWER_CUDA(cudaEventElapsedTime(&t, ix->ev[i].first, ix->ev[i].second));
        ms += t;
    }
    if (filter_ms) *filter_ms = ms;
    if (filter_launches) *filter_launches = ix->filter_launches;
    if (other_launches) *other_launches = ix->other_launches;
    ix->ev_used = 0;
    ix->filter_launches = ix->other_launches = 0;
    return HWER_OK;
}

int hwer_profile_launches(hwer_index_t* ix, void* stream_v, double* out_ms, int32_t cap, int32_t* n_out) {
    if (!ix || !n_out || cap < 0 || (cap > 0 && !out_ms)) return fail(HWER_E_INVALID, "hwer_profile_launches: bad argument");
    HWER_CUDA(cudaSetDevice(ix->device));
    HWER_CUDA(cudaStreamSynchronize((cudaStream_t)stream_v));
    int32_t n = 0;
    for (size_t i = 0; i < ix->ev_used; ++i) {
        if (ix->ev_stage[i] != kStageFilter) continue;
        if (n < cap) {
            float t = 0.f;
            HWER_CUDA(cudaEventElapsedTime(&t, ix->ev[i].first, ix->ev[i].second));
            out_ms[n] = t;
        }
        ++n;
    }
    *n_out = n;
    return HWER_OK;
}

int hwer_profile_stages(hwer_index_t* ix, void* stream_v, double* out4_ms) {
    if (!ix || !out4_ms) return fail(HWER_E_INVALID, "hwer_profile_stages: bad argument");
    HWER_CUDA(cudaSetDevice(ix->device));
    HWER_CUDA(cudaStreamSynchronize((cudaStream_t)stream_v));
    for (int s = 0; s < kStages; ++s) out4_ms[s] = 0.0;
    for (size_t i = 0; i < ix->ev_used; ++i) {
        float t = 0.f;
        HWER_CUDA(cudaEventElapsedTime(&t, ix->ev[i].first, ix->ev[i].second));
        out4_ms[ix->ev_stage[i]] += t;
    }
    return HWER_OK;
}

int hwer_debug_scores(hwer_index_t* ix, const float* queries_dev, int32_t B, float* out_dev, int64_t ld, void* stream) {
    if (!ix || !ix->use_tc || !queries_dev || !out_dev || B <= 0 || ld < B)
        return fail(HWER_E_INVALID, "hwer_debug_scores: bad argument");
    HWER_CUDA(cudaSetDevice(ix->device));
    hwer::FilterParams p;
    memset(&p, 0, sizeof p);
    p.queries = queries_dev; p.B = B; p.d = ix->d; p.kb = ix->d_pad / 64;
    int nq = (B + 31) / 32 * 32;
    const int nq_max = ix->d_pad > 192 ? 128 : hwer::kMaxNQ;
    if (nq > nq_max) nq = nq_max;
    p.nq = nq; p.nqb = (B + nq - 1) / nq;
    p.n_items = ix->n; p.tile_begin = 0; p.tile_end = (int)ix->n_tiles;
    p.tile_mul = ix->tile_mul; p.tile_mod = ix->n_tiles;
    p.dump = out_dev; p.dump_ld = ld;
    HWER_CUDA(hwer::launch_filter_tc(ix->tmap, p, ix->num_sms, (cudaStream_t)stream));
    return HWER_OK;
}

int hwer_merge_topk(const double* scores_dev, const int64_t* idx_dev, int32_t G, int32_t B, int32_t k,
                    int64_t* out_idx_dev, float* out_score_dev, double* out_score64_dev, void* stream) {
    if (!scores_dev || !idx_dev || G <= 0 || B < 0 || k <= 0 || !out_idx_dev || !out_score_dev)
        return fail(HWER_E_INVALID, "hwer_merge_topk: bad argument");
    HWER_CUDA(hwer::launch_merge(scores_dev, (const long long*)idx_dev, G, B, k, (long long*)out_idx_dev,
                                 out_score_dev, out_score64_dev, (cudaStream_t)stream));
    return HWER_OK;
}

int hwer_pair_score(const float* table_dev, int64_t n, int32_t d, const int64_t* src_dev, const int64_t* dst_dev,
                    int64_t P, float* out_dev, void* stream) {
    if (!table_dev || n <= 0 || d <= 0 || P < 0 || (P > 0 && (!src_dev || !dst_dev || !out_dev)))
        return fail(HWER_E_INVALID, "hwer_pair_score: bad argument");
    HWER_CUDA(hwer::launch_pair_score(table_dev, n, d, (const long long*)src_dev, (const long long*)dst_dev, P,
                                      out_dev, (cudaStream_t)stream));
    return HWER_OK;
}

int hwer_compose_queries(const float* table_dev, int64_t n, int32_t d, const int64_t* anchor_rows_dev,
                         const int64_t* pos_ptr_dev, const int64_t* pos_rows_dev, const int64_t* neg_ptr_dev,
                         const int64_t* neg_rows_dev, int32_t B, float* out_dev, void* stream) {
    if (!table_dev || n <= 0 || d <= 0 || d > 1024 || B < 0 || (B > 0 && (!anchor_rows_dev || !out_dev)) ||
        (pos_ptr_dev && !pos_rows_dev) || (neg_ptr_dev && !neg_rows_dev))
        return fail(HWER_E_INVALID, "hwer_compose_queries: bad argument (d <= 1024)");
    HWER_CUDA(hwer::launch_compose_queries(table_dev, n, d, (const long long*)anchor_rows_dev,
                                           (const long long*)pos_ptr_dev, (const long long*)pos_rows_dev,
                                           (const long long*)neg_ptr_dev, (const long long*)neg_rows_dev, B, out_dev,
                                           (cudaStream_t)stream));
    return HWER_OK;
}

int hwer_average_embeddings(const float* table_dev, int64_t n, int32_t d, const int64_t* ptr_dev,
                            const int64_t* rows_dev, int32_t L, float* out_dev, void* stream) {
    if (!table_dev || n <= 0 || d <= 0 || d > 1024 || L < 0 || (L > 0 && (!ptr_dev || !rows_dev || !out_dev)))
        return fail(HWER_E_INVALID, "hwer_average_embeddings: bad argument (d <= 1024)");
    HWER_CUDA(hwer::launch_average_embeddings(table_dev, n, d, (const long long*)ptr_dev, (const long long*)rows_dev, L,
                                              out_dev, (cudaStream_t)stream));
    return HWER_OK;
}

int hwer_gather_rows(const float* table_dev, int64_t n, int32_t d, const int64_t* rows_dev, int64_t P, float* out_dev,
                     void* stream) {
    if (!table_dev || n <= 0 || d <= 0 || P < 0 || (P > 0 && (!rows_dev || !out_dev)))
        return fail(HWER_E_INVALID, "hwer_gather_rows: bad argument");
    HWER_CUDA(hwer::launch_gather_rows(table_dev, n, d, (const long long*)rows_dev, P, out_dev, (cudaStream_t)stream));
    return HWER_OK;
}

int hwer_map_rows(const int64_t* rows_dev, int64_t count, const int64_t* row_map_dev, int64_t offset, int64_t* out_dev,
                  void* stream) {
    if (count < 0 || (count > 0 && (!rows_dev || !out_dev))) return fail(HWER_E_INVALID, "hwer_map_rows: bad argument");
    HWER_CUDA(hwer::launch_map_rows((const long long*)rows_dev, count, (const long long*)row_map_dev, offset,
                                    (long long*)out_dev, (cudaStream_t)stream));
    return HWER_OK;
}

int hwer_rerank(const float* table_dev, int64_t n, int32_t d, const int64_t* rows_dev, const int64_t* row_map_dev,
                int32_t B, int32_t k, int32_t convention, const int64_t* anchor_rows_dev, const float* queries_dev,
                const float* given_dev, int64_t* out_rows_dev, double* out_score_dev, void* stream) {
    if (!table_dev || n <= 0 || d <= 0 || B < 0 || k <= 0 || k > 8192 ||
        (B > 0 && (!rows_dev || !out_rows_dev || !out_score_dev)))
        return fail(HWER_E_INVALID, "hwer_rerank: bad argument (0 < k <= 8192)");
    if (convention < HWER_SCORE_PAIR || convention > HWER_SCORE_EUCLID)
        return fail(HWER_E_INVALID, "hwer_rerank: unknown score convention");
    if (B > 0 && ((convention == HWER_SCORE_PAIR && !anchor_rows_dev) || (convention == HWER_SCORE_GIVEN && !given_dev) ||
                  ((convention == HWER_SCORE_DIST || convention == HWER_SCORE_EUCLID) && !queries_dev)))
        return fail(HWER_E_INVALID, "hwer_rerank: the input of this score convention is missing");
    HWER_CUDA(hwer::launch_rerank(table_dev, n, d, (const long long*)rows_dev, (const long long*)row_map_dev, B, k,
                                  convention, (const long long*)anchor_rows_dev, queries_dev, given_dev,
                                  (long long*)out_rows_dev, out_score_dev, (cudaStream_t)stream));
    return HWER_OK;
}

int hwer_hit_rank_metrics(const float* scores_dev, int32_t U, int32_t M, int32_t topn, double* out2_dev,
                          int32_t* rank_dev, void* stream_v) {
    if (!scores_dev || U <= 0 || M < 0 || topn <= 0 || !out2_dev)
        return fail(HWER_E_INVALID, "hwer_hit_rank_metrics: bad argument");
    cudaStream_t stream = (cudaStream_t)stream_v;
    int* rank = rank_dev;
    if (!rank) HWER_CUDA(cudaMallocAsync(&rank, sizeof(int) * (size_t)U, stream));
    cudaError_t e = hwer::launch_hit_rank(scores_dev, U, M + 1, topn, rank, out2_dev, stream);
    if (!rank_dev) cudaFreeAsync(rank, stream);
    if (e != cudaSuccess) return fail_cuda(e, "hwer_hit_rank_metrics");
    return HWER_OK;
}

int64_t hwer_ncf_param_count(int32_t F, int32_t depth) {
    if (F <= 0 || depth < 1 || depth > 16) return -1;
    return hwer::ncf_param_count(F, depth);
}

int hwer_ncf_score(const float* h_dev, int64_t n_rows, int32_t F, int32_t depth, const float* params_dev,
                   const int64_t* src_dev, const int64_t* dst_dev, int64_t P, float* out_dev, void* stream_v) {
    if (!h_dev || n_rows <= 0 || F <= 0 || F % 4 || depth < 1 || depth > 16 || !params_dev || P < 0 ||
        (P > 0 && (!src_dev || !dst_dev || !out_dev)))
        return fail(HWER_E_INVALID, "hwer_ncf_score: bad argument (F must be a positive multiple of 4)");
    if (P == 0) return HWER_OK;
    cudaStream_t stream = (cudaStream_t)stream_v;
    static const int no_tc = getenv("HWER_NCF_FFMA") ? atoi(getenv("HWER_NCF_FFMA")) : 0;     // A/B knob, read once
    if (hwer::ncf_tc_supported(F, depth) && !no_tc) {
        // tensor-core path (split-bf16, ncf_tc.cu): serving widths F = 64, 128, 256
        int dev = 0, sms = 0;
        HWER_CUDA(cudaGetDevice(&dev));
        HWER_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        long long chunk = hwer::ncf_tc_chunk(sms);
        const long long p_pad = (P + 127) / 128 * 128;
        if (chunk > p_pad) chunk = p_pad;
        void* scratch = nullptr;
        HWER_CUDA(cudaMallocAsync(&scratch, hwer::ncf_tc_scratch_elems(F, depth, chunk) * 2, stream));
        cudaError_t e = hwer::launch_ncf_score_tc(h_dev, n_rows, F, depth, params_dev, (const long long*)src_dev,
                                                  (const long long*)dst_dev, P, out_dev, scratch, chunk, sms, stream);
        cudaFreeAsync(scratch, stream);
        if (e != cudaSuccess) return fail_cuda(e, "hwer_ncf_score (tensor-core path)");
        return HWER_OK;
    }
    const long long chunk = P < 32768 ? P : 32768;
    float *ws0 = nullptr, *ws1 = nullptr;
    HWER_CUDA(cudaMallocAsync(&ws0, sizeof(float) * (size_t)chunk * 4 * F, stream));
    HWER_CUDA(cudaMallocAsync(&ws1, sizeof(float) * (size_t)chunk * 2 * F, stream));
    cudaError_t e = hwer::launch_ncf_score(h_dev, n_rows, F, depth, params_dev, (const long long*)src_dev,
                                           (const long long*)dst_dev, P, out_dev, ws0, ws1, chunk, stream);
    cudaFreeAsync(ws0, stream);
    cudaFreeAsync(ws1, stream);
    if (e != cudaSuccess) return fail_cuda(e, "hwer_ncf_score");
    return HWER_OK;
}

int hwer_gcn_infer(const float* node_emb_dev, const float* content_dev, int64_t n, int32_t C, int32_t F, int32_t layers,
                   const float* proj_w_dev, const float* proj_b_dev, const float* ln_g_dev, const float* ln_b_dev,
                   const int64_t* const* nbr_ptr_dev, const int64_t* const* nbr_idx_dev, const float* fc0_w_dev,
                   const float* fc0_b_dev, const float* fc1_w_dev, const float* fc1_b_dev, float* previous_dev, float ema,
                   float* out_dev, void* stream_v) {
    if (!node_emb_dev || !content_dev || !proj_w_dev || !proj_b_dev || !ln_g_dev || !ln_b_dev || !nbr_ptr_dev ||
        !nbr_idx_dev || !fc0_w_dev || !fc0_b_dev || !fc1_w_dev || !fc1_b_dev || !out_dev || n < 0 || layers < 1 ||
        layers > 8 || C < 4 || (C & 3) || F < 4 || (F & 3))
        return fail(HWER_E_INVALID, "hwer_gcn_infer: bad argument (C and F must be positive multiples of 4, 1 <= layers <= 8)");
    for (int i = 0; i < layers; ++i)
        if (!nbr_ptr_dev[i] || !nbr_idx_dev[i]) return fail(HWER_E_INVALID, "hwer_gcn_infer: missing neighbour list");
    if (n == 0) return HWER_OK;
    cudaStream_t stream = (cudaStream_t)stream_v;
    long long chunk = 1 << 18;                                  // nodes per pass of the dense layers
    if (chunk > n) chunk = n;
    float* scratch = nullptr;
    HWER_CUDA(cudaMallocAsync(&scratch, sizeof(float) * hwer::gcn_infer_scratch_floats(n, F, layers, chunk), stream));
    cudaError_t e = hwer::launch_gcn_infer(node_emb_dev, content_dev, C, proj_w_dev, proj_b_dev, ln_g_dev, ln_b_dev, n, F,
                                           layers, (const long long* const*)nbr_ptr_dev,
                                           (const long long* const*)nbr_idx_dev, fc0_w_dev, fc0_b_dev, fc1_w_dev,
                                           fc1_b_dev, previous_dev, ema, out_dev, scratch, chunk, stream);
    cudaFreeAsync(scratch, stream);
    if (e != cudaSuccess) return fail_cuda(e, "hwer_gcn_infer");
    return HWER_OK;
}

int hwer_eval_metrics(const int64_t* topk_dev, int32_t U, int32_t kret, const int64_t* train_ptr_dev,
                      const int64_t* train_idx_dev, const int64_t* val_ptr_dev, const int64_t* val_idx_dev,
                      const float* val_rel_dev, const int32_t* cutoffs_dev, int32_t n_cut, int64_t n_items,
                      double* out_dev, double* per_user_dev, void* stream_v) {
    if (!topk_dev || U <= 0 || kret <= 0 || !train_ptr_dev || !val_ptr_dev || !cutoffs_dev || n_cut <= 0 ||
        n_cut > 8 || n_items <= 0 || !out_dev)
        return fail(HWER_E_INVALID, "hwer_eval_metrics: bad argument");
    cudaStream_t stream = (cudaStream_t)stream_v;
    const int M = 3 * n_cut + 1;
    double* per_user = per_user_dev;
    if (!per_user) HWER_CUDA(cudaMallocAsync(&per_user, sizeof(double) * (size_t)U * M, stream));
    unsigned int* bitmap = nullptr;
    const size_t words = (size_t)((n_items + 31) / 32);
    HWER_CUDA(cudaMallocAsync(&bitmap, sizeof(unsigned int) * words, stream));
    HWER_CUDA(cudaMemsetAsync(bitmap, 0, sizeof(unsigned int) * words, stream));
    HWER_CUDA(hwer::launch_eval((const long long*)topk_dev, U, kret, (const long long*)train_ptr_dev,
                                (const long long*)train_idx_dev, (const long long*)val_ptr_dev,
                                (const long long*)val_idx_dev, val_rel_dev, cutoffs_dev, n_cut, n_items, per_user,
                                bitmap, stream));
    HWER_CUDA(hwer::launch_eval_reduce(per_user, U, M, (const long long*)val_ptr_dev, bitmap, n_items, n_cut, out_dev,
                                       stream));
    HWER_CUDA(cudaFreeAsync(bitmap, stream));
    if (!per_user_dev) HWER_CUDA(cudaFreeAsync(per_user, stream));
    return HWER_OK;
}

int hwer_link_metrics(const float* scores_dev, const uint8_t* labels_dev, int64_t P, float threshold, double* out8_dev,
                      void* stream) {
    if (!scores_dev || !labels_dev || !out8_dev || P <= 0 || P >= (1LL << 31) - 1)
        return fail(HWER_E_INVALID, "hwer_link_metrics: bad argument (0 < P < 2^31 - 1)");
    HWER_CUDA(hwer::launch_link_metrics(scores_dev, labels_dev, P, threshold, out8_dev, (cudaStream_t)stream));
    return HWER_OK;
}

}  // extern "C"
