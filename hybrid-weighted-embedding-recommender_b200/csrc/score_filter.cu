// Fused score-and-select, stage 1: score a range of catalogue rows against a
// resident block of queries on the tcgen05 tensor cores and admit only scores
// at or above the per-query threshold into per-query candidate lists.  The
// [items x queries] score matrix lives in TMEM only; it never reaches HBM.
//
// Replaces (together with select.cu) the exact k-NN search of the reference:
//   hwer/recommendation_base.py:78-83  MultiKNN.query -> sklearn KDTree.query
//   hwer/recommendation_base.py:171    find_closest_neighbours' candidate search
// On unit-norm rows Euclidean order == dot-product order,
// so the kernel scores raw dot products.
//
// Tile mapping (one CTA per SM, persistent):
//   UMMA M = 128 catalogue rows  (A operand: TMA-streamed bf16 tiles, K-major, 128B swizzle)
//   UMMA N = nq <= 256 queries   (B operand: resident in shared memory for the CTA's lifetime)
//   UMMA K = 16, d_pad/16 steps  (accumulators: fp32 in TMEM, acc_stages-deep)
// Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (+ TMEM owner),
// warps 2..17 = epilogue (tcgen05.ld -> release the accumulator -> max-of-8 test of score / thr -> 48-byte spill entry
// for a group of 8 columns with a hit; spill_extract_kernel behind the filter kernel appends the candidates).
#include "common.cuh"
#include "kernels.h"

namespace hwer {

namespace {

constexpr int kSlabBytes = kTileItems * 128;  // one 64-wide K block of an item tile: 16 KB
constexpr int kEpiWarps = 16;                 // four per TMEM lane quadrant (one SM scheduler each), splitting the
                                              // query columns: the epilogue is latency-, not issue-bound
constexpr int kThreadsTc = (2 + kEpiWarps) * 32;
constexpr int kColParts = kEpiWarps / 4;      // warps sharing a lane quadrant
constexpr int kBiasRowBytes = 32;             // K=16 bf16 per row of the threshold-MMA operands

enum : int { kModeFilter = 0, kModeDump = 1, kModeDense = 2 };

__device__ __forceinline__ void append_candidate(unsigned long long* cand, unsigned int* cnt, unsigned int cap,
                                                 int q, float s, uint32_t row) {
    unsigned int slot = atomicAdd(&cnt[q], 1u);
    if (slot < cap) cand[(size_t)q * cap + slot] = make_key(s, row);
}

// Hits leave the kernel as SPILL ENTRIES.  With two accumulator stages the slowest of a tile's 16 epilogue warps sets
// the pace of the tensor pipe, so what matters is the latency a hit adds inside a warp.  Walking a row's values,
// allocating a slot in the query's candidate list (a global atomic) and writing the key cost 300-400 cycles per hit
// however the appends were queued or deferred (v9-v12: the five hit-heavy rounds of a B = 4096 step took 2.3 ms for
// 0.7 ms of MMA work, profiles/README.md).  Instead, a lane that sees a non-negative (score - thr) in a group of 8
// columns stores the group's 8 raw values plus (row, first query of the group) -- 48 bytes, three fire-and-forget
// 128-bit stores -- into its thread's PRIVATE append buffer in global memory: no atomic, no shared counter, no
// branch per column, nothing to wait for.  spill_extract_kernel, launched right behind the filter kernel, walks
// the entries with one thread each and does the per-hit work at full memory-level parallelism.
// A thread whose buffer is full (heavily tied data) falls back to the blocking append below.
struct SpillCtx {
    uint4* mine;                  // this thread's buffer [spill_cap][3]
    uint32_t n;                   // entries written (may exceed spill_cap: the excess was appended directly)
};

// What the threshold MMA subtracts for a query: thr ~ hi + lo with both parts bf16 and hi + lo <= thr (lo rounds
// toward -inf), so folding it into the MMA only ever admits more; "no threshold yet" (-inf) becomes the query's
// floor.  Returns hi + lo and the two negated bf16 bit patterns of the bias operand.
__device__ __forceinline__ float thr_split(float t, uint32_t& nhi, uint32_t& nlo) {
    t = fminf(t, 3.0e38f);
    const __nv_bfloat16 hi = __float2bfloat16_rn(t);
    const float rem = t - __bfloat162float(hi);
    const uint32_t u = __float_as_uint(rem);
    uint32_t lo_bits = u & 0xFFFF0000u;
    if ((u & 0x80000000u) && (u & 0xFFFFu)) lo_bits += 0x10000u;     // negative: away from zero
    nhi = (uint32_t)(__bfloat16_as_ushort(hi) ^ 0x8000u);
    nlo = (lo_bits >> 16) ^ 0x8000u;
    return __bfloat162float(hi) + __uint_as_float(lo_bits);
}

__device__ __forceinline__ float query_threshold(const FilterParams& p, int q) {
    return q < p.B ? fmaxf(p.thr[q], p.floor[q]) : __int_as_float(0x7f800000);   // padded query: never admits
}

// Buffer full: blocking append of the hits of one 8-column group.  Out of line, rarely taken.
__device__ __noinline__ void append_group_direct(const FilterParams* p, uint4 a, uint4 b, uint32_t thr_addr,
                                                 uint32_t row, int q0, bool scaled) {
    const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll 1
    for (int j = 0; j < 8; ++j) {
        const float val = __uint_as_float(w[j]);          // score / thr (scaled block) or score - thr (biased block)
        if (scaled ? val >= 1.0f : val >= 0.0f) {
            float thr;
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(thr) : "r"(thr_addr + (uint32_t)j * 4u));
            const unsigned int slot = atomicAdd(&p->cnt[q0 + j], 1u);
            if (slot < p->cap) p->cand[(size_t)(q0 + j) * p->cap + slot] = make_key(scaled ? val * thr : val + thr, row);
        }
    }
}

// The filter rounds make "is any of these 8 scores admissible" a handful of 3-input ALU instructions, two ways.
//
// SCALED query blocks (every threshold of the block positive -- the normal case once round 0 has run): the query
// is staged as q / thr, so the accumulator holds score / thr and a score is admissible iff its word is >= 1.0f:
// the maximum of the eight words (FMNMX3; NaN operands are ignored, so NaN scores never admit) against 1.0f.
// No extra tensor-pipe work: a tile is exactly d_pad / 16 MMAs.
//
// BIASED query blocks (some threshold <= 0: tiny or anti-correlated catalogues, k close to n): one extra K=16 MMA
// multiplies a constant-ones slab with (-thr_hi, -thr_lo), the accumulator holds score - thr, and the test is "is
// any sign bit clear": the bitwise AND of the eight words keeps the sign bit only if all are negative.
__device__ __forceinline__ uint32_t and8(const uint32_t* v) {
    return (v[0] & v[1] & v[2]) & (v[3] & v[4] & v[5]) & (v[6] & v[7]);
}
__device__ __forceinline__ float max3f(float a, float b, float c) { return fmaxf(fmaxf(a, b), c); }
__device__ __forceinline__ float max8(const uint32_t* v) {
    const float m0 = max3f(__uint_as_float(v[0]), __uint_as_float(v[1]), __uint_as_float(v[2]));
    const float m1 = max3f(__uint_as_float(v[3]), __uint_as_float(v[4]), __uint_as_float(v[5]));
    return fmaxf(max3f(__uint_as_float(v[6]), __uint_as_float(v[7]), m0), m1);
}

// Can this query be filtered in the scaled domain?  thr > 0, and q / thr must stay far from overflow in bf16/fp32
// (an infinite operand would turn 0 * inf products into NaN scores, which never admit: a true neighbour lost).
__device__ __forceinline__ bool query_scalable(const FilterParams& p, int q, float thr) {
    return q >= p.B || (thr > 0.0f && p.qmax[q] < thr * 1.0e30f && !p.force_bias);
}

// 32 accumulator columns of one catalogue row (already out of TMEM, the accumulator stage already released)
// against the 32 queries they belong to.
template <int MODE, bool scaled>
__device__ __forceinline__ void epilogue_chunk(const uint32_t (&v)[32], const float* thr_s, int c0, int q_base,
                                               uint32_t row, bool row_ok, const FilterParams& p, SpillCtx& sp,
                                               uint32_t dense_pos) {
    if (MODE == kModeDump) {
        if (row_ok) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const int q = q_base + c0 + j;
                if (q < p.B) p.dump[(size_t)row * p.dump_ld + q] = __uint_as_float(v[j]);
            }
        }
        return;
    }
    if (MODE == kModeDense) {
        // round 0 (open threshold): every score is a candidate, slot = position in the round, no atomics;
        // for a fixed query the 32 lanes write 32 consecutive keys (256 B, coalesced)
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const int q = q_base + c0 + j;
            const float s = __uint_as_float(v[j]);
            if (q < p.B && dense_pos < p.cap) p.cand[(size_t)q * p.cap + dense_pos] = (row_ok && s == s) ? make_key(s, row) : 0ull;
        }
        return;
    }
    uint32_t grp[4];
    float mx[4];
    bool any;
    if (scaled) {
#pragma unroll
        for (int g = 0; g < 4; ++g) mx[g] = max8(&v[8 * g]);
        any = fmaxf(max3f(mx[0], mx[1], mx[2]), mx[3]) >= 1.0f;
    } else {
#pragma unroll
        for (int g = 0; g < 4; ++g) grp[g] = and8(&v[8 * g]);
        any = (int)((grp[0] & grp[1]) & (grp[2] & grp[3])) >= 0;
    }
    if (any && row_ok) {                           // this lane's row has an admissible score: rare
        if (sp.n + 4u <= (uint32_t)p.spill_cap) {
            // room for all four groups of the chunk (the normal case): no branch per group -- predicated stores and a
            // running entry pointer.  With two accumulator stages the slowest epilogue warp of a tile paces the tensor
            // pipe, so the LATENCY of this path (it was four BSSY / branch / BSYNC regions deep) is what a hit costs.
            uint4* e = sp.mine + (size_t)sp.n * 3;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const bool hit = scaled ? mx[g] >= 1.0f : (int)grp[g] >= 0;
                if (hit) {
                    e[0] = make_uint4(v[8 * g], v[8 * g + 1], v[8 * g + 2], v[8 * g + 3]);
                    e[1] = make_uint4(v[8 * g + 4], v[8 * g + 5], v[8 * g + 6], v[8 * g + 7]);
                    e[2] = make_uint4(row, (uint32_t)(q_base + c0 + 8 * g), 0u, 0u);
                }
                e += hit ? 3 : 0;
                sp.n += hit ? 1u : 0u;
            }
            return;
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            if (scaled ? mx[g] >= 1.0f : (int)grp[g] >= 0) {
                const uint4 a = make_uint4(v[8 * g], v[8 * g + 1], v[8 * g + 2], v[8 * g + 3]);
                const uint4 b = make_uint4(v[8 * g + 4], v[8 * g + 5], v[8 * g + 6], v[8 * g + 7]);
                if (sp.n < (uint32_t)p.spill_cap) {
                    uint4* e = sp.mine + (size_t)sp.n * 3;
                    e[0] = a;
                    e[1] = b;
                    e[2] = make_uint4(row, (uint32_t)(q_base + c0 + 8 * g), 0u, 0u);
                } else {
                    append_group_direct(&p, a, b, smem_u32(thr_s + c0 + 8 * g), row, q_base + c0 + 8 * g, scaled);
                }
                ++sp.n;
            }
        }
    }
}

// Which query blocks and which tiles a CTA of the filter grid walks.  Regular CTAs: `slots` CTAs share one query
// block and stride over the tiles [tile_begin, tile_split).  When the SM count is not a multiple of the number of
// query blocks (148 SMs, 16 blocks of a 4096-query batch: 9 slots = 144 CTAs), the `extra` left-over CTAs take the
// tail [tile_split, tile_end) instead, each for every extra-th query block in turn -- sized on the host so that
// every CTA of the grid scores the same number of tiles.
struct CtaWork {
    int qb0, qb_step;          // query blocks qb0, qb0 + qb_step, ... < nqb
    int t_first, t_stride, t_end;
};
__device__ __forceinline__ CtaWork cta_work(const FilterParams& p, int bid) {
    CtaWork w;
    const int regular = p.slots * (p.nqb < p.qb_step ? p.nqb : p.qb_step);
    if (p.extra > 0 && bid >= regular) {
        w.qb0 = bid - regular; w.qb_step = p.extra;
        w.t_first = p.tile_split; w.t_stride = 1; w.t_end = p.tile_end;
    } else {
        w.qb0 = bid / p.slots; w.qb_step = p.qb_step;
        w.t_first = p.tile_begin + bid % p.slots; w.t_stride = p.slots; w.t_end = p.extra > 0 ? p.tile_split : p.tile_end;
    }
    return w;
}

// The per-hit work the filter kernel's epilogue does not do.  Block b owns the 512 spill buffers of filter CTA b,
// one thread per buffer.  All of a filter CTA's hits belong to the (usually one) query block it scored, so the
// block first COUNTS its hits per query in shared memory, reserves one contiguous range per query in the global
// candidate lists with a single atomicAdd, and then walks the entries again to write the keys: ~30x fewer global
// atomics than one per hit at B = 4096, and at small batches (all hits on a few dozen counters) the per-address
// serialisation in L2 that bounded the early rounds is gone.
__device__ __forceinline__ void load_entry(const uint4* e, uint4& a, uint4& b) {
    a = e[0];
    b = e[1];
}

constexpr int kExtractSub = 2;                                  // threads sharing one buffer (entries interleaved)
constexpr int kExtractThreads = kEpiWarps * 32 * kExtractSub;

__global__ void __launch_bounds__(kExtractThreads)
spill_extract_kernel(const __grid_constant__ FilterParams p) {
    __shared__ unsigned int cnt_s[kMaxNQ];      // hits per query of the block, then the write cursor
    __shared__ float thr_x[kMaxNQ];             // the query's threshold (scaled block) / what the bias MMA subtracted
    const int tid = threadIdx.x;
    const size_t buf = (size_t)blockIdx.x * (kEpiWarps * 32) + (tid / kExtractSub);
    unsigned int n = p.spill_cnt[buf];
    if (n > (unsigned int)p.spill_cap) n = (unsigned int)p.spill_cap;
    if (__syncthreads_or(n != 0u) == 0) return;                      // nothing spilled by this filter CTA
    const uint4* e0 = p.spill + buf * (size_t)p.spill_cap * 3;
    const int nq = p.nq;
    unsigned int cursor = (unsigned int)(tid % kExtractSub);         // this thread's entries: cursor, cursor + Sub, ...
    const CtaWork work = cta_work(p, (int)blockIdx.x);
    for (int qb = work.qb0; qb < p.nqb; qb += work.qb_step) {                   // the filter CTA's query blocks
        const int q_base = qb * nq;
        bool ok = true;                                                         // the filter CTA's own vote, recomputed
        for (int i = tid; i < nq; i += kExtractThreads) {
            cnt_s[i] = 0u;
            ok = ok && query_scalable(p, q_base + i, query_threshold(p, q_base + i));
        }
        const bool scaled = __syncthreads_and(ok ? 1 : 0) != 0;                 // words are score / thr, else score - thr
        const float admit = scaled ? 1.0f : 0.0f;
        unsigned int end = cursor;
        {                                                                       // pass 1: count
            uint4 a, b, m;
            if (end < n) { const uint4* e = e0 + (size_t)end * 3; a = e[0]; b = e[1]; m = e[2]; }
            while (end < n) {
                const int ql = (int)m.y - q_base;
                if (ql >= nq) break;                                            // a later query block's entry
                const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
                end += kExtractSub;
                if (end < n) { const uint4* e = e0 + (size_t)end * 3; a = e[0]; b = e[1]; m = e[2]; }   // next in flight
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (__uint_as_float(w[j]) >= admit) atomicAdd(&cnt_s[ql + j], 1u);
            }
        }
        __syncthreads();
        for (int i = tid; i < nq; i += kExtractThreads) {                       // one global atomic per query
            const unsigned int c = cnt_s[i];
            if (c) {
                uint32_t nhi, nlo;
                const float t = query_threshold(p, q_base + i);
                thr_x[i] = scaled ? t : thr_split(t, nhi, nlo);
                cnt_s[i] = atomicAdd(&p.cnt[q_base + i], c);
            }
        }
        __syncthreads();
        {                                                                       // pass 2: write the keys
            unsigned int i = cursor;
            uint4 a, b, m;
            if (i < end) { const uint4* e = e0 + (size_t)i * 3; a = e[0]; b = e[1]; m = e[2]; }
            while (i < end) {
                const int ql = (int)m.y - q_base;
                const uint32_t row = m.x;
                const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
                i += kExtractSub;
                if (i < end) { const uint4* e = e0 + (size_t)i * 3; a = e[0]; b = e[1]; m = e[2]; }
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float val = __uint_as_float(w[j]);                    // score / thr or score - thr (NaN stays NaN)
                    if (val >= admit) {
                        const unsigned int pos = atomicAdd(&cnt_s[ql + j], 1u);
                        // the key carries the score back in query units: one fp32 rounding (~6e-8 relative), far
                        // inside the margin
                        if (pos < p.cap)
                            p.cand[(size_t)(q_base + ql + j) * p.cap + pos] =
                                make_key(scaled ? val * thr_x[ql + j] : val + thr_x[ql + j], row);
                    }
                }
            }
        }
        cursor = end;
        __syncthreads();
    }
}

// The same per-hit work done by the filter CTA itself, at the end of each query block (FilterParams::fused_extract):
// every epilogue thread walks ITS OWN buffer (its entry count is still in a register, the entries are L2-hot), the
// CTA counts hits per query in shared memory, reserves one range per query with a single global atomic, and writes
// the keys.  No second kernel launch per round (launch gap + 37 us of mostly fixed cost at B = 4096), and the
// thresholds / scaled flag of the block are the ones already in shared memory.  All kThreadsTc threads call this.
__device__ __forceinline__ void extract_own_spill(const FilterParams& p, const uint4* mine, uint32_t n_raw, int q_base,
                                                  int nq, const float* thr_s, bool scaled, unsigned int* cnt_s) {
    const int tid = threadIdx.x;
    const unsigned int n = n_raw < (uint32_t)p.spill_cap ? n_raw : (uint32_t)p.spill_cap;   // the excess went direct
    if (__syncthreads_or(n != 0u) == 0) return;                      // nothing spilled by this CTA for this block
    for (int i = tid; i < nq; i += kThreadsTc) cnt_s[i] = 0u;
    __syncthreads();
    const float admit = scaled ? 1.0f : 0.0f;
    {                                                                 // pass 1: count
        uint4 a, b, m;
        if (n) { a = mine[0]; b = mine[1]; m = mine[2]; }
        for (unsigned int e = 0; e < n;) {
            const int ql = (int)m.y - q_base;
            const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
            ++e;
            if (e < n) { const uint4* x = mine + (size_t)e * 3; a = x[0]; b = x[1]; m = x[2]; }      // next in flight
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (__uint_as_float(w[j]) >= admit) atomicAdd(&cnt_s[ql + j], 1u);
        }
    }
    __syncthreads();
    for (int i = tid; i < nq; i += kThreadsTc) {                      // one global atomic per query with hits
        const unsigned int c = cnt_s[i];
        if (c) cnt_s[i] = atomicAdd(&p.cnt[q_base + i], c);
    }
    __syncthreads();
    {                                                                 // pass 2: write the keys
        uint4 a, b, m;
        if (n) { a = mine[0]; b = mine[1]; m = mine[2]; }
        for (unsigned int e = 0; e < n;) {
            const int ql = (int)m.y - q_base;
            const uint32_t row = m.x;
            const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
            ++e;
            if (e < n) { const uint4* x = mine + (size_t)e * 3; a = x[0]; b = x[1]; m = x[2]; }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float val = __uint_as_float(w[j]);              // score / thr or score - thr (NaN stays NaN)
                if (val >= admit) {
                    const unsigned int pos = atomicAdd(&cnt_s[ql + j], 1u);
                    if (pos < p.cap)
                        p.cand[(size_t)(q_base + ql + j) * p.cap + pos] =
                            make_key(scaled ? val * thr_s[ql + j] : val + thr_s[ql + j], row);
                }
            }
        }
    }
    __syncthreads();                                                  // cnt_s / thr_s are rewritten by the next block
}

template <int MODE, int KB>
// 18 warps = 5 on one of the SM's four sub-partitions, each with a 16 K-entry register file: 96 registers per thread
// is the ceiling (16384 / (5 * 32) = 102, allocated in eights), which is why the tile loop spills a little
__global__ void __launch_bounds__(kThreadsTc, 1)
score_filter_tc_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ FilterParams p) {
    extern __shared__ uint8_t smem_raw[];
    // 128B-swizzled operands need 1024-byte aligned slabs.  Pointer arithmetic on the __shared__ array (not a
    // round trip through an integer) keeps the shared address space visible to the compiler: LDS/STS, not LD/ST.
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);

    const int nq = p.nq;                         // multiple of 32
    const int q_slab = nq * 128;                 // bytes of one K block of the query operand
    constexpr int stage_bytes = KB * kSlabBytes;

    uint8_t* q_smem = smem;
    uint8_t* item_smem = q_smem + KB * q_slab;   // KB*q_slab is a multiple of 1024 (nq % 8 == 0)
    uint8_t* ones_smem = item_smem + (size_t)p.stages * stage_bytes;   // [128 rows x K16] constant (1, 1, 0, ...)
    uint8_t* bias_smem = ones_smem + kBiasRowBytes * kTileItems;       // [nq rows x K16] (-thr_hi, -thr_lo, 0, ...)
    float* thr_s = reinterpret_cast<float*>(bias_smem + kBiasRowBytes * kMaxNQ);
    unsigned int* cnt_s = reinterpret_cast<unsigned int*>(thr_s + kMaxNQ);      // fused extract: hits per query
    uint64_t* bars = reinterpret_cast<uint64_t*>(cnt_s + kMaxNQ);
    uint64_t* full_bar = bars;                            // [stages]   TMA -> MMA
    uint64_t* empty_bar = bars + p.stages;                // [stages]   MMA -> TMA
    uint64_t* tfull_bar = bars + 2 * p.stages;            // [acc]      MMA -> epilogue
    uint64_t* tempty_bar = tfull_bar + p.acc_stages;      // [acc]      epilogue -> MMA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + p.acc_stages);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    // epilogue warps that own at least one 32-column chunk of this query-block width (narrow blocks: fewer warps)
    const int parts_active = nq / 32 < kColParts ? nq / 32 : kColParts;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap);
        for (int i = 0; i < p.stages; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < p.acc_stages; ++i) {
            mbar_init(&tfull_bar[i], 1);
            mbar_init(&tempty_bar[i], 4 * parts_active);   // one arrival per participating epilogue warp
        }
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
        tmem_relinquish();
    }
    if (MODE == kModeFilter) {
        // constant A operand of the threshold MMA: K-major, no swizzle, 8-row core matrices of 128 B;
        // row r, columns 0..7 live at (r/8)*256 + (r%8)*16, columns 8..15 (all zero) 128 B further
        for (int i = threadIdx.x; i < kTileItems * 2; i += kThreadsTc) {
            const int r = i >> 1, half = i & 1;
            const uint4 val = half ? make_uint4(0u, 0u, 0u, 0u) : make_uint4(0x3F803F80u, 0u, 0u, 0u);   // bf16 (1, 1)
            *reinterpret_cast<uint4*>(ones_smem + (r >> 3) * 256 + half * 128 + (r & 7) * 16) = val;
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    const CtaWork work = cta_work(p, (int)blockIdx.x);
    const int qb0 = work.qb0;
    const uint32_t idesc = umma_idesc_bf16_f32(kTileItems, (uint32_t)nq);
    const uint64_t policy = p.stream_once ? l2_policy_evict_first() : l2_policy_evict_last();

    // Tiles this CTA walks per query block, and the visiting order (32-bit: tile counts fit easily).
    const int t_first = work.t_first;
    const int my_tiles = t_first < work.t_end ? (work.t_end - t_first + work.t_stride - 1) / work.t_stride : 0;
    const uint32_t phys0 = (uint32_t)(((long long)t_first * p.tile_mul) % p.tile_mod);
    const uint32_t tile_mod = (uint32_t)p.tile_mod;
    const uint32_t tile_step = work.t_stride == 1 ? (uint32_t)(p.tile_mul % p.tile_mod) : (uint32_t)p.tile_step;
    // Ring positions persist across query blocks; every role keeps private copies (so they can live in uniform
    // registers where the role is warp-converged) re-derived from this count at the top of each block.
    uint32_t tiles_done = 0;
    // this thread's private spill buffer (epilogue threads of the filter rounds)
    SpillCtx sp;
    sp.n = 0u;
    sp.mine = nullptr;
    if (MODE == kModeFilter && warp >= 2)
        sp.mine = p.spill + ((size_t)blockIdx.x * (kEpiWarps * 32) + (threadIdx.x - 64)) * (size_t)p.spill_cap * 3;

    for (int qb = qb0; qb < p.nqb; qb += work.qb_step) {
        const int q_base = qb * nq;
        // ---- thresholds of the block, and how it will be filtered (scaled queries or the bias MMA) ----
        bool scaled = false;
        if (MODE == kModeFilter) {
            bool ok = true;
            for (int i = threadIdx.x; i < nq; i += kThreadsTc) {
                const float t = query_threshold(p, q_base + i);
                thr_s[i] = t;
                ok = ok && query_scalable(p, q_base + i, t);
            }
            scaled = __syncthreads_and(ok ? 1 : 0) != 0;           // also publishes thr_s to the staging loop below
            if (!scaled) {
                for (int i = threadIdx.x; i < nq; i += kThreadsTc) {
                    uint32_t nhi, nlo;
                    thr_s[i] = thr_split(thr_s[i], nhi, nlo);      // what the bias MMA subtracts
                    uint8_t* dst = bias_smem + (i >> 3) * 256 + (i & 7) * 16;
                    *reinterpret_cast<uint4*>(dst) = make_uint4(nhi | (nlo << 16), 0u, 0u, 0u);
                    *reinterpret_cast<uint4*>(dst + 128) = make_uint4(0u, 0u, 0u, 0u);
                }
            }
        } else {
            for (int i = threadIdx.x; i < nq; i += kThreadsTc) thr_s[i] = 0.0f;
        }
        // ---- stage the query block: fp32 (x 1 / thr in a scaled block) -> bf16 (RN), K-major, 128B swizzle ----
        {
            constexpr int chunks_per_row = KB * 8;   // 16-byte chunks (8 bf16) per query row
            const bool q_vec = (p.d & 7) == 0 && (reinterpret_cast<uintptr_t>(p.queries) & 15u) == 0;
            for (int i = threadIdx.x; i < nq * chunks_per_row; i += kThreadsTc) {
                const int r = i / chunks_per_row;
                const int c = i - r * chunks_per_row;
                const int q = q_base + r;
                // two fp32 roundings (reciprocal, product): 2^-23 relative, inside the margin's slack (kernels.h)
                const float scale = (scaled && q < p.B) ? 1.0f / thr_s[r] : 1.0f;
                float f[8];
                if (q_vec && q < p.B && c * 8 < p.d) {        // whole 8-column chunk inside the row: two 128-bit loads
                    const float4* src = reinterpret_cast<const float4*>(p.queries + (size_t)q * p.d + c * 8);
                    const float4 lo4 = __ldg(src), hi4 = __ldg(src + 1);
                    f[0] = lo4.x * scale; f[1] = lo4.y * scale; f[2] = lo4.z * scale; f[3] = lo4.w * scale;
                    f[4] = hi4.x * scale; f[5] = hi4.y * scale; f[6] = hi4.z * scale; f[7] = hi4.w * scale;
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int col = c * 8 + j;
                        f[j] = (q < p.B && col < p.d) ? __ldg(p.queries + (size_t)q * p.d + col) * scale : 0.0f;
                    }
                }
                uint4 pk;
                __nv_bfloat162 b0 = __floats2bfloat162_rn(f[0], f[1]);
                __nv_bfloat162 b1 = __floats2bfloat162_rn(f[2], f[3]);
                __nv_bfloat162 b2 = __floats2bfloat162_rn(f[4], f[5]);
                __nv_bfloat162 b3 = __floats2bfloat162_rn(f[6], f[7]);
                pk.x = *reinterpret_cast<uint32_t*>(&b0);
                pk.y = *reinterpret_cast<uint32_t*>(&b1);
                pk.z = *reinterpret_cast<uint32_t*>(&b2);
                pk.w = *reinterpret_cast<uint32_t*>(&b3);
                const int kblk = c >> 3;
                const int cc = (c & 7) ^ (r & 7);
                *reinterpret_cast<uint4*>(q_smem + (size_t)kblk * q_slab + r * 128 + cc * 16) = pk;
            }
            fence_proxy_async_smem();   // generic-proxy writes -> visible to the tensor-core (async) proxy
        }
        __syncthreads();

        if (warp == 0) {
            // ===================== TMA producer =====================
            if (lane == 0) {
                uint32_t stage = tiles_done % (uint32_t)p.stages, phase = (tiles_done / (uint32_t)p.stages) & 1u;
                uint32_t phys = phys0;
                for (int it = 0; it < my_tiles; ++it) {
                    mbar_wait(&empty_bar[stage], phase ^ 1u);
                    mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)stage_bytes);
                    uint8_t* dst = item_smem + (size_t)stage * stage_bytes;
                    const int tile_row = (int)(phys * (uint32_t)kTileItems);
                    phys += tile_step;                         // (slots * tile_mul) % tile_mod, from the host
                    if (phys >= tile_mod) phys -= tile_mod;
#pragma unroll
                    for (int k = 0; k < KB; ++k)
                        tma_load_2d(dst + k * kSlabBytes, &tmap, k * kKBlock, tile_row, &full_bar[stage], policy);
                    if (++stage == (uint32_t)p.stages) { stage = 0; phase ^= 1u; }
                }
            }
            __syncwarp();
        } else if (warp == 1) {
            // ===================== MMA issuer =====================
            // The warp walks the loop converged (ring positions and descriptor halves stay in uniform registers);
            // one elected lane issues a tile's MMAs and commits inside a single branch.  The K loop is unrolled at
            // compile time: a tile is KB*4 (+1 threshold) tcgen05.mma whose descriptors differ by constants.
            const bool leader = elect_one();
            const uint32_t sw_hi = (uint32_t)(umma_desc_k_sw128(0) >> 32);
            const uint32_t ns_hi = (uint32_t)(umma_desc_k_noswizzle(0) >> 32);
            const uint32_t q_lo = (uint32_t)umma_desc_k_sw128(smem_u32(q_smem));
            const uint32_t i_lo = (uint32_t)umma_desc_k_sw128(smem_u32(item_smem));
            const uint32_t ones_lo = (uint32_t)umma_desc_k_noswizzle(smem_u32(ones_smem));
            const uint32_t bias_lo = (uint32_t)umma_desc_k_noswizzle(smem_u32(bias_smem));
            constexpr uint32_t stage_step = (uint32_t)stage_bytes >> 4;     // descriptor address units are 16 bytes
            const uint32_t q_step = (uint32_t)q_slab >> 4;
            uint32_t stage = tiles_done % (uint32_t)p.stages, phase = (tiles_done / (uint32_t)p.stages) & 1u;
            uint32_t acc = tiles_done % (uint32_t)p.acc_stages, acc_phase = (tiles_done / (uint32_t)p.acc_stages) & 1u;
            for (int it = 0; it < my_tiles; ++it) {
                mbar_wait(&tempty_bar[acc], acc_phase ^ 1u);
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after_sync();
                if (leader) {
                    const uint32_t d_tmem = tmem_base + acc * (uint32_t)nq;
                    const uint32_t a_lo = i_lo + stage * stage_step;
#pragma unroll
                    for (int k = 0; k < KB; ++k) {
#pragma unroll
                        for (int s = 0; s < kKBlock / 16; ++s)
                            umma_bf16_lohi(d_tmem, a_lo + k * (kSlabBytes >> 4) + 2 * s, sw_hi,
                                           q_lo + k * q_step + 2 * s, sw_hi, idesc, (k | s) ? 1u : 0u);
                    }
                    if (MODE == kModeFilter && !scaled)     // biased block: accumulator -= thr (see and8 / max8)
                        umma_bf16_lohi(d_tmem, ones_lo, ns_hi, bias_lo, ns_hi, idesc, 1u);
                    umma_commit(&empty_bar[stage]);   // smem stage reusable once these MMAs retire
                    umma_commit(&tfull_bar[acc]);     // accumulator ready for the epilogue
                }
                __syncwarp();
                if (++stage == (uint32_t)p.stages) { stage = 0; phase ^= 1u; }
                if (++acc == (uint32_t)p.acc_stages) { acc = 0; acc_phase ^= 1u; }
            }
        } else if (((warp - 2) >> 2) < parts_active) {
            // ===================== epilogue (up to 16 warps) =====================
            // Per tile a warp owns 32 TMEM lanes (catalogue rows) x one or two 32-column chunks (queries).  It
            // pulls them into registers, hands the accumulator stage straight back to the MMA warp, and only then
            // looks at the values: the tensor pipe never waits for compare or hit handling.
            const int e = warp - 2;
            const uint32_t quad = (uint32_t)warp & 3u;   // TMEM lane quadrant this warp may read
            const int part = e >> 2;                      // its chunks: columns 32*part and 32*part + 128
            const bool two = 32 * part + 128 < nq;        // warp-uniform
            uint32_t acc = tiles_done % (uint32_t)p.acc_stages, acc_phase = (tiles_done / (uint32_t)p.acc_stages) & 1u;
            uint32_t phys = phys0;
            const uint32_t lane_row = quad * 32u + (uint32_t)lane;
            const uint32_t n_items = (uint32_t)p.n_items;
            const uint32_t tbase = tmem_base + ((quad * 32u) << 16) + 32u * (uint32_t)part;
            const int c0 = 32 * part;
            for (int it = 0; it < my_tiles; ++it) {
                mbar_wait(&tfull_bar[acc], acc_phase);
                tc_fence_after_sync();
                // ONE 32-column chunk is in registers at a time (r2): with both chunks live (64 value registers) the
                // loop state spilled to local memory, and in hit-dense rounds those reloads queued behind the hit
                // stores in the LSU -- the epilogue, not the tensor pipe, paced the round (ncu, round 3: 36 % tensor
                // active, the hot stalls were LDL results).  The accumulator stage is released after the SECOND
                // load, ~250 cycles later than before: still far inside the 1024 cycles the other stage's MMAs take.
                uint32_t v[32];
                const uint32_t taddr = tbase + acc * (uint32_t)nq;
                const uint32_t row = phys * (uint32_t)kTileItems + lane_row;
                phys += tile_step;
                if (phys >= tile_mod) phys -= tile_mod;
                const bool row_ok = row < n_items;
                const uint32_t dense_pos = (uint32_t)(t_first - p.tile_begin + it * work.t_stride) * (uint32_t)kTileItems + lane_row;
                tmem_ld_x32(taddr, v);
                tmem_ld_wait();
                if (two) {
                    if (scaled) epilogue_chunk<MODE, true>(v, thr_s, c0, q_base, row, row_ok, p, sp, dense_pos);
                    else epilogue_chunk<MODE, false>(v, thr_s, c0, q_base, row, row_ok, p, sp, dense_pos);
                    tmem_ld_x32(taddr + 128u, v);
                    tmem_ld_wait();
                }
                tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty_bar[acc]);        // values are in registers: release the stage
                if (++acc == (uint32_t)p.acc_stages) { acc = 0; acc_phase ^= 1u; }
                const int cc = two ? c0 + 128 : c0;
                if (scaled) epilogue_chunk<MODE, true>(v, thr_s, cc, q_base, row, row_ok, p, sp, dense_pos);
                else epilogue_chunk<MODE, false>(v, thr_s, cc, q_base, row, row_ok, p, sp, dense_pos);
            }
        }
        tiles_done += (uint32_t)my_tiles;
        // All MMAs reading this query block have retired once every epilogue warp is here.
        __syncthreads();
        if (MODE == kModeFilter && p.fused_extract) {
            extract_own_spill(p, sp.mine, warp >= 2 ? sp.n : 0u, q_base, nq, thr_s, scaled, cnt_s);
            sp.n = 0u;                                                // the buffer starts over for the next block
        }
    }

    if (MODE == kModeFilter && warp >= 2 && !p.fused_extract)
        p.spill_cnt[(size_t)blockIdx.x * (kEpiWarps * 32) + (threadIdx.x - 64)] = sp.n;
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

// ----------------------------------------------------------------------------
// Generic CUDA-core filter over the fp32 table (any d).  One thread per
// catalogue row, queries staged through shared memory eight at a time.
// ----------------------------------------------------------------------------
constexpr int kSimtQ = 8;
constexpr int kSimtThreads = 256;

__global__ void __launch_bounds__(kSimtThreads)
score_filter_simt_kernel(const float* __restrict__ table, long long n_items, int d, const float* __restrict__ queries,
                         int B, const float* __restrict__ thr, unsigned long long* cand, unsigned int* cnt,
                         unsigned int cap, long long row_begin, long long row_end) {
    extern __shared__ float qs[];   // [kSimtQ][d]
    __shared__ float ts[kSimtQ];
    for (long long base = row_begin + (long long)blockIdx.x * kSimtThreads; base < row_end;
         base += (long long)gridDim.x * kSimtThreads) {
        const long long row = base + threadIdx.x;
        const bool ok = row < row_end && row < n_items;
        const float* x = table + (size_t)(ok ? row : row_begin) * d;
        for (int q0 = 0; q0 < B; q0 += kSimtQ) {
            __syncthreads();
            for (int i = threadIdx.x; i < kSimtQ * d; i += kSimtThreads) {
                const int q = q0 + i / d;
                qs[i] = (q < B) ? queries[(size_t)q * d + (i % d)] : 0.0f;
            }
            if (threadIdx.x < kSimtQ)
                ts[threadIdx.x] = (q0 + threadIdx.x < B) ? thr[q0 + threadIdx.x] : __int_as_float(0x7f800000);
            __syncthreads();
            float acc[kSimtQ];
#pragma unroll
            for (int j = 0; j < kSimtQ; ++j) acc[j] = 0.0f;
            for (int c = 0; c < d; ++c) {
                const float xv = __ldg(x + c);
#pragma unroll
                for (int j = 0; j < kSimtQ; ++j) acc[j] = fmaf(xv, qs[j * d + c], acc[j]);
            }
            if (ok) {
#pragma unroll
                for (int j = 0; j < kSimtQ; ++j)
                    if (acc[j] >= ts[j]) append_candidate(cand, cnt, cap, q0 + j, acc[j], (uint32_t)row);
            }
        }
    }
}

// Per-query set-up of one search, one launch: the admission margin, the finite stand-in for "-inf", the overflow
// guard of the scaled filter -- and the reset of the query's candidate count, overflow mark and threshold (these
// were two memsets and a fill kernel: three more launches on a step that is ~15 launches long at small batches).
__global__ void query_margin_kernel(const float* __restrict__ queries, int B, int d, float factor, float max_norm,
                                    float* __restrict__ margin, float* __restrict__ floor, float* __restrict__ qmax,
                                    unsigned int* __restrict__ cnt, unsigned int* __restrict__ ovf,
                                    float* __restrict__ thr) {
    const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= B) return;
    float s = 0.0f, mx = 0.0f;
    for (int c = lane_id(); c < d; c += 32) {
        const float v = queries[(size_t)q * d + c];
        s = fmaf(v, v, s);
        mx = fmaxf(mx, fabsf(v));
    }
    s = warp_sum(s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane_id() == 0) {
        qmax[q] = mx;
        const float nrm = sqrtf(s) * 1.0001f;       // rounded up: the bounds must never be under-estimated
        if (margin) margin[q] = factor * nrm;
        // no score of this query can be below -|q| * max|x| (minus the bf16 slack): a finite "-inf"
        floor[q] = -(1.01f * nrm * max_norm + 1e-30f);
        cnt[q] = 0u;
        ovf[q] = 0u;
        thr[q] = __int_as_float(0xff800000);
    }
}

__global__ void fill_f32_kernel(float* p, long long n, float v) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// One instantiation per (mode, K blocks); the opt-in shared-memory ceiling is a per-function, per-process setting.
template <int MODE, int KB>
cudaError_t launch_tc_one(int grid, size_t smem, const CUtensorMap& tmap, const FilterParams& p, cudaStream_t stream) {
    // the attribute is per device (a process may build indexes on several GPUs): remember it per device ordinal
    static bool attr_done[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !attr_done[dev]) {
        cudaError_t e = cudaFuncSetAttribute(score_filter_tc_kernel<MODE, KB>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) attr_done[dev] = true;
    }
    score_filter_tc_kernel<MODE, KB><<<grid, kThreadsTc, smem, stream>>>(tmap, p);
    return cudaGetLastError();
}

template <int MODE>
cudaError_t launch_tc_mode(int kb, int grid, size_t smem, const CUtensorMap& tmap, const FilterParams& p,
                           cudaStream_t stream) {
    switch (kb) {
        case 1: return launch_tc_one<MODE, 1>(grid, smem, tmap, p, stream);
        case 2: return launch_tc_one<MODE, 2>(grid, smem, tmap, p, stream);
        case 3: return launch_tc_one<MODE, 3>(grid, smem, tmap, p, stream);
        case 4: return launch_tc_one<MODE, 4>(grid, smem, tmap, p, stream);
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_tc_variant(int mode, int kb, int grid, size_t smem, const CUtensorMap& tmap, const FilterParams& p,
                              cudaStream_t stream) {
    switch (mode) {
        case kModeDump: return launch_tc_mode<kModeDump>(kb, grid, smem, tmap, p, stream);
        case kModeDense: return launch_tc_mode<kModeDense>(kb, grid, smem, tmap, p, stream);
        default: return launch_tc_mode<kModeFilter>(kb, grid, smem, tmap, p, stream);
    }
}

}  // namespace

size_t filter_tc_smem_bytes(int nq, int kb, int stages) {
    return 1024 + (size_t)kb * nq * 128 + (size_t)stages * kb * kSlabBytes + 2 * kMaxNQ * sizeof(float) +
           (size_t)kBiasRowBytes * (kTileItems + kMaxNQ) +
           (2 * stages + 8) * sizeof(uint64_t) + 16 + kEpiWarps * sizeof(unsigned int);
}

int filter_tc_pick_stages(int nq, int kb) {
    int stages = 8;
    while (stages > 0 && filter_tc_smem_bytes(nq, kb, stages) > (size_t)kSmemBudget) --stages;
    return stages;
}

cudaError_t launch_filter_tc(const CUtensorMap& tmap, FilterParams p, int num_sms, cudaStream_t stream) {
    if (p.tile_end <= p.tile_begin || p.B <= 0) return cudaSuccess;
    p.stages = filter_tc_pick_stages(p.nq, p.kb);
    if (p.stages < 2) return cudaErrorInvalidConfiguration;
    p.acc_stages = 512 / p.nq;
    if (p.acc_stages > 4) p.acc_stages = 4;
    int cols = p.acc_stages * p.nq, pow2 = 32;
    while (pow2 < cols) pow2 <<= 1;
    p.tmem_cols = pow2;
    const int ntiles = p.tile_end - p.tile_begin;
    int grid;
    p.extra = 0;
    p.tile_split = p.tile_end;
    if (p.nqb <= num_sms) {
        p.slots = num_sms / p.nqb;
        if (p.slots > ntiles) p.slots = ntiles;
        p.qb_step = p.nqb;
        grid = p.slots * p.nqb;
        // left-over SMs (148 - 9 * 16 = 4 at B = 4096): give them the tail of the round, a quarter of the query
        // blocks each, sized so that all CTAs score the same number of tiles: (split - begin) / slots ==
        // ceil(nqb / extra) * (end - split)
        const int extra = num_sms - grid;
        if (extra > 0 && p.nqb > 1 && !p.no_extra_ctas && p.slots == num_sms / p.nqb) {
            const int per_extra = (p.nqb + extra - 1) / extra;                      // query blocks per extra CTA
            const int tail = (int)((long long)ntiles / ((long long)p.slots * per_extra + 1));
            if (tail >= 8) {
                p.extra = extra;
                p.tile_split = p.tile_end - tail;
                grid += extra;
            }
        }
    } else {
        p.slots = 1;
        p.qb_step = num_sms;
        grid = num_sms;
    }
    p.stream_once = (p.nqb == 1) ? 1 : 0;
    p.tile_step = ((long long)p.slots * p.tile_mul) % p.tile_mod;
    const size_t smem = filter_tc_smem_bytes(p.nq, p.kb, p.stages);
    const int mode = p.dump ? kModeDump : (p.dense ? kModeDense : kModeFilter);
    if (mode == kModeFilter && (!p.spill || !p.spill_cnt || p.spill_cap < 1 || grid > p.spill_ctas))
        return cudaErrorInvalidValue;
    cudaError_t e = launch_tc_variant(mode, p.kb, grid, smem, tmap, p, stream);
    if (e != cudaSuccess || mode != kModeFilter || p.fused_extract) return e;
    spill_extract_kernel<<<grid, kExtractThreads, 0, stream>>>(p);
    return cudaGetLastError();
}

size_t filter_tc_spill_entries(int num_sms, int spill_cap) { return (size_t)num_sms * kEpiWarps * 32 * spill_cap; }
size_t filter_tc_spill_buffers(int num_sms) { return (size_t)num_sms * kEpiWarps * 32; }

cudaError_t launch_filter_simt(const float* table, long long n_items, int d, const float* queries, int B,
                               const float* thr, unsigned long long* cand, unsigned int* cnt, unsigned int cap,
                               long long row_begin, long long row_end, int num_sms, cudaStream_t stream) {
    if (row_end <= row_begin || B <= 0) return cudaSuccess;
    long long blocks = (row_end - row_begin + kSimtThreads - 1) / kSimtThreads;
    long long maxb = (long long)num_sms * 8;
    int grid = (int)(blocks < maxb ? blocks : maxb);
    size_t smem = (size_t)kSimtQ * d * sizeof(float);
    if (smem > 48 * 1024) {      // only widths past 1536 columns need the opt-in ceiling
        cudaError_t e = cudaFuncSetAttribute(score_filter_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    score_filter_simt_kernel<<<grid, kSimtThreads, smem, stream>>>(table, n_items, d, queries, B, thr, cand, cnt, cap,
                                                                   row_begin, row_end);
    return cudaGetLastError();
}

cudaError_t launch_query_margin(const float* queries, int B, int d, float factor, float max_norm, float* margin,
                                float* floor, float* qmax, unsigned int* cnt, unsigned int* ovf, float* thr,
                                cudaStream_t stream) {
    if (B <= 0) return cudaSuccess;
    query_margin_kernel<<<(B + 7) / 8, 256, 0, stream>>>(queries, B, d, factor, max_norm, margin, floor, qmax, cnt, ovf, thr);
    return cudaGetLastError();
}

cudaError_t launch_fill_f32(float* p, long long n, float v, cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    fill_f32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(p, n, v);
    return cudaGetLastError();
}

}  // namespace hwer
