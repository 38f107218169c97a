// Fused cosine score + segmented per-document top-k, fp32 FFMA parity mode (sm_100a).
//
// Replaces Retriever._get_similarities (reference src/_modules.py:1978-1997) and the per-document
// torch.topk (src/_modules.py:2015-2016) -- see include/rdv.h for the contract.
//
// Shape of the work: one question per document, so this is a batch of ragged GEMVs: 0.5 flop/byte,
// HBM-bound, and for the small batches the reference actually runs (tens of MB) latency-bound.  The
// batch is cut on the host into row tiles that never cross a document (rdv_tile_desc, 32 bytes each).
//
// Two kernels share the descriptor, the arithmetic and the selection epilogue:
//
//  * score_topk_tma_kernel (default for d in {128,256,384,512,768,1024}): persistent, one block per SM,
//    each block owns a contiguous run of tiles.  A producer warp streams the tiles (and the tile's
//    question vector) into a shared-memory ring with 1-D bulk async copies (cp.async.bulk -> UBLKCP,
//    completion on an mbarrier), up to ~190 KB in flight per SM independent of occupancy; eight
//    consumer warps take one row each out of shared memory (conflict-free LDS.128), produce the dot
//    product and the squared norm from the same registers, and release the stage.  Every byte of E is read
//    from HBM exactly once.
//  * score_topk_ldg_kernel (any d % 4 == 0): one block per tile, a warp per row, ROWS rows in flight
//    per warp as independent 128-bit no-allocate loads.
//
// Epilogue (both): every similarity is written (the reference returns the full vector,
// src/_modules.py:2176-2180); a device-scope row counter per document tells which block finished the
// document, and that block runs the selection out of L2 / shared memory: k rounds of a block-wide arg-max
// over packed (score, ~index) keys, which makes "descending score, lowest index first" one u64 compare.
#include "rdv_common.cuh"

namespace rdv {

constexpr int kScoreThreads = 256;            // 8 compute warps
constexpr int kScoreWarps = kScoreThreads / 32;
constexpr int kMaxCacheFloats = 8192;         // selection pass caches up to this many scores in smem (32 KB)
constexpr int kTmaThreads = kScoreThreads + 32;   // + 1 producer warp
constexpr int kTmaMaxStages = 16;
constexpr int kTmaRingBytes = 176 * 1024;

struct ScoreParams {
    const rdv_tile_desc* tiles;
    const int64_t* row_off;
    const float* q;
    int32_t B, d, k, total_tiles, cache_floats;
    int32_t tile_rows, stages;                // TMA kernel: rows per stage (max), ring depth
    float* sims;
    int32_t* topk_idx;
    float* topk_val;
    int32_t* topk_cnt;
    int32_t* doc_done;
};

struct BlockSync {       // whole block
    __device__ __forceinline__ void operator()() const { __syncthreads(); }
};
struct ConsumerSync {    // the 256 consumer threads of the TMA kernel (named barrier 1)
    __device__ __forceinline__ void operator()() const { asm volatile("bar.sync 1, 256;" ::: "memory"); }
};

__device__ __forceinline__ float cosine(float dot, float ss_e, float ss_q) {
    // reference: dot / (||e|| * ||q|| + 1e-8), all fp32, IEEE sqrt and divide
    return __fdiv_rn(dot, __fadd_rn(__fmul_rn(__fsqrt_rn(ss_e), __fsqrt_rn(ss_q)), 1e-8f));
}

// Selection of the k best (score desc, index asc) among n scores of one document by 256 threads.
template <class Sync>
__device__ void select_topk(const ScoreParams& p, int b, const float* __restrict__ src, int n,
                            float* cache, unsigned long long* s_red, Sync sync) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool cached = n <= p.cache_floats;
    if (cached) {
        for (int i = tid; i < n; i += kScoreThreads) cache[i] = __ldcg(src + i);
        sync();
    }
    const int k_min = n < p.k ? n : p.k;
    unsigned long long prev = 0;
    for (int r = 0; r < k_min; ++r) {
        unsigned long long best = 0;   // every real key is > 0 (order_key(-inf) = 0x007FFFFF)
        if (cached) {
#pragma unroll 4
            for (int i = tid; i < n; i += kScoreThreads) {
                unsigned long long key = pack_key(cache[i], (uint32_t)i);
                if ((r == 0 || key < prev) && key > best) best = key;
            }
        } else {
#pragma unroll 4
            for (int i = tid; i < n; i += kScoreThreads) {
                unsigned long long key = pack_key(__ldcg(src + i), (uint32_t)i);
                if ((r == 0 || key < prev) && key > best) best = key;
            }
        }
        best = warp_max_u64(best);
        if (lane == 0) s_red[warp] = best;
        sync();
        unsigned long long win = s_red[0];
#pragma unroll
        for (int w = 1; w < kScoreWarps; ++w) win = s_red[w] > win ? s_red[w] : win;
        sync();
        if (tid == 0) {
            uint32_t idx = key_index(win);
            p.topk_idx[(size_t)b * p.k + r] = (int32_t)idx;
            p.topk_val[(size_t)b * p.k + r] = cached ? cache[idx] : __ldcg(src + idx);
        }
        prev = win;
    }
    for (int r = k_min + tid; r < p.k; r += kScoreThreads) {
        p.topk_idx[(size_t)b * p.k + r] = -1;
        p.topk_val[(size_t)b * p.k + r] = -INFINITY;
    }
    if (tid == 0) {
        p.topk_cnt[b] = k_min;
        p.doc_done[b] = 0;   // leave the workspace zeroed for the next call
    }
}

// documents with no chunks own no tile: their (empty) results are written by block 0
__device__ __forceinline__ void write_empty_docs(const ScoreParams& p, int tid, int nthreads) {
    for (int b = tid; b < p.B; b += nthreads) {
        if (p.row_off[b + 1] == p.row_off[b]) {
            p.topk_cnt[b] = 0;
            for (int r = 0; r < p.k; ++r) {
                p.topk_idx[(size_t)b * p.k + r] = -1;
                p.topk_val[(size_t)b * p.k + r] = -INFINITY;
            }
        }
    }
}

// `rows` more rows of document b are written: publish, and if that completes the document run its
// selection.  Called by all 256 compute threads.
template <class Sync>
__device__ __forceinline__ void publish_rows(const ScoreParams& p, int b, int rows, int doc_rows, float* cache,
                                             unsigned long long* s_red, int* s_last, Sync sync) {
    __threadfence();
    sync();
    if (threadIdx.x == 0) {
        const int prev = atomicAdd(p.doc_done + b, rows);
        *s_last = (prev + rows == doc_rows);
        __threadfence();
    }
    sync();
    if (*s_last) select_topk(p, b, p.sims + p.row_off[b], doc_rows, cache, s_red, sync);
}

template <int VPL>
__device__ __forceinline__ void fma_row(const float4 (&e)[VPL], const float4 (&q)[VPL], float& dot, float& ss) {
    dot = 0.f; ss = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        dot = fmaf(e[i].x, q[i].x, dot); ss = fmaf(e[i].x, e[i].x, ss);
        dot = fmaf(e[i].y, q[i].y, dot); ss = fmaf(e[i].y, e[i].y, ss);
        dot = fmaf(e[i].z, q[i].z, dot); ss = fmaf(e[i].z, e[i].z, ss);
        dot = fmaf(e[i].w, q[i].w, dot); ss = fmaf(e[i].w, e[i].w, ss);
    }
}

template <int VPL>
__device__ __forceinline__ float sumsq(const float4 (&q)[VPL]) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        s = fmaf(q[i].x, q[i].x, s); s = fmaf(q[i].y, q[i].y, s);
        s = fmaf(q[i].z, q[i].z, s); s = fmaf(q[i].w, q[i].w, s);
    }
    return warp_sum(s);
}

// =====================================================================================================
// LDG kernel: one block per tile.  VPL > 0: d == 128 * VPL, everything in registers.  VPL == 0: any d % 4 == 0.
// =====================================================================================================
template <int VPL, int ROWS>
__global__ void __launch_bounds__(kScoreThreads) score_topk_ldg_kernel(const ScoreParams p) {
    extern __shared__ float4 smem_dyn[];
    __shared__ unsigned long long s_red[kScoreWarps];
    __shared__ int s_last;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int d4 = p.d >> 2;
    float* cache = reinterpret_cast<float*>(smem_dyn);

    if (blockIdx.x == 0) write_empty_docs(p, tid, kScoreThreads);
    if ((int)blockIdx.x >= p.total_tiles) return;

    const rdv_tile_desc t = p.tiles[blockIdx.x];          // one broadcast 32-byte load, no search
    const float4* __restrict__ E = reinterpret_cast<const float4*>(t.src);
    const float4* __restrict__ Q = reinterpret_cast<const float4*>(p.q) + (size_t)t.doc * d4;
    float* __restrict__ out = p.sims + t.sims_off;

    if constexpr (VPL > 0) {
        float4 qv[VPL];
#pragma unroll
        for (int i = 0; i < VPL; ++i) qv[i] = __ldg(Q + lane + 32 * i);   // in flight together with the rows
        float ss_q = 0.f;
        bool have_q = false;
        for (int r = warp * ROWS; r < t.rows; r += kScoreWarps * ROWS) {
            float4 ev[ROWS][VPL];
#pragma unroll
            for (int j = 0; j < ROWS; ++j) {
                const bool ok = r + j < t.rows;
                const float4* src = E + (size_t)(ok ? r + j : 0) * d4 + lane;
#pragma unroll
                for (int i = 0; i < VPL; ++i)
                    ev[j][i] = ok ? ldg_stream(src + 32 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            if (!have_q) { ss_q = sumsq<VPL>(qv); have_q = true; }
            float mine = 0.f;
#pragma unroll
            for (int j = 0; j < ROWS; ++j) {
                float dot, ss;
                fma_row<VPL>(ev[j], qv, dot, ss);
                const float sim = cosine(warp_sum(dot), warp_sum(ss), ss_q);
                if (lane == j) mine = sim;
            }
            if (lane < ROWS && r + lane < t.rows) out[r + lane] = mine;
        }
    } else {
        float ss_q = 0.f;
        for (int i = lane; i < d4; i += 32) {
            const float4 v = __ldg(Q + i);
            ss_q = fmaf(v.x, v.x, ss_q); ss_q = fmaf(v.y, v.y, ss_q);
            ss_q = fmaf(v.z, v.z, ss_q); ss_q = fmaf(v.w, v.w, ss_q);
        }
        ss_q = warp_sum(ss_q);
        for (int r = warp * ROWS; r < t.rows; r += kScoreWarps * ROWS) {
            float dot[ROWS], ss[ROWS];
#pragma unroll
            for (int j = 0; j < ROWS; ++j) { dot[j] = 0.f; ss[j] = 0.f; }
#pragma unroll 2
            for (int i = lane; i < d4; i += 32) {
                const float4 qv = __ldg(Q + i);
#pragma unroll
                for (int j = 0; j < ROWS; ++j) {
                    const bool ok = r + j < t.rows;
                    const float4 e = ok ? ldg_stream(E + (size_t)(r + j) * d4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
                    dot[j] = fmaf(e.x, qv.x, dot[j]); ss[j] = fmaf(e.x, e.x, ss[j]);
                    dot[j] = fmaf(e.y, qv.y, dot[j]); ss[j] = fmaf(e.y, e.y, ss[j]);
                    dot[j] = fmaf(e.z, qv.z, dot[j]); ss[j] = fmaf(e.z, e.z, ss[j]);
                    dot[j] = fmaf(e.w, qv.w, dot[j]); ss[j] = fmaf(e.w, e.w, ss[j]);
                }
            }
            float mine = 0.f;
#pragma unroll
            for (int j = 0; j < ROWS; ++j) {
                const float sim = cosine(warp_sum(dot[j]), warp_sum(ss[j]), ss_q);
                if (lane == j) mine = sim;
            }
            if (lane < ROWS && r + lane < t.rows) out[r + lane] = mine;
        }
    }
    publish_rows(p, t.doc, t.rows, t.doc_rows, cache, s_red, &s_last, BlockSync());
}

// =====================================================================================================
// TMA kernel: persistent, producer warp + shared-memory ring of bulk async copies.
// =====================================================================================================
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!done);
}
// 1-D bulk async copy global -> shared, completion (bytes) signalled on an mbarrier.  SASS: UBLKCP.
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

template <int VPL>
__global__ void __launch_bounds__(kTmaThreads, 1) score_topk_tma_kernel(const ScoreParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t s_full[kTmaMaxStages];
    __shared__ __align__(8) uint64_t s_empty[kTmaMaxStages];
    __shared__ rdv_tile_desc s_desc[kTmaMaxStages];
    __shared__ unsigned long long s_red[kScoreWarps];
    __shared__ int s_last;

    constexpr int D4 = 32 * VPL;                       // float4 per row
    constexpr uint32_t kRowBytes = D4 * 16;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int S = p.stages;
    const uint32_t stage_bytes = (uint32_t)(p.tile_rows + 1) * kRowBytes;   // [question row][tile rows]
    float* cache = reinterpret_cast<float*>(smem_raw + (size_t)S * stage_bytes);

    // contiguous run of tiles for this block
    const int G = gridDim.x;
    const int t0 = (int)((long long)blockIdx.x * p.total_tiles / G);
    const int t1 = (int)((long long)(blockIdx.x + 1) * p.total_tiles / G);

    if (tid == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(&s_full[s], 1); mbar_init(&s_empty[s], kScoreWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == kScoreWarps) {
        // ===== producer warp: descriptors are fetched 32 at a time, lane 0 issues the copies =====
        int stage = 0; uint32_t phase = 0;
        for (int base = t0; base < t1; base += 32) {
            rdv_tile_desc mine = {};
            if (base + lane < t1) mine = p.tiles[base + lane];
            const int cnt = min(32, t1 - base);
            for (int j = 0; j < cnt; ++j) {
                rdv_tile_desc t;
                t.src = reinterpret_cast<const void*>(__shfl_sync(0xffffffffu, (unsigned long long)mine.src, j));
                t.sims_off = __shfl_sync(0xffffffffu, mine.sims_off, j);
                t.rows = __shfl_sync(0xffffffffu, mine.rows, j);
                t.doc = __shfl_sync(0xffffffffu, mine.doc, j);
                t.doc_rows = __shfl_sync(0xffffffffu, mine.doc_rows, j);
                t.reserved = 0;
                if (lane == 0) {
                    mbar_wait(&s_empty[stage], phase ^ 1);
                    unsigned char* dst = smem_raw + (size_t)stage * stage_bytes;
                    s_desc[stage] = t;
                    const uint32_t bytes = (uint32_t)t.rows * kRowBytes;
                    mbar_expect_tx(&s_full[stage], bytes + kRowBytes);
                    bulk_g2s(dst, p.q + (size_t)t.doc * (D4 * 4), kRowBytes, &s_full[stage]);
                    bulk_g2s(dst + kRowBytes, t.src, bytes, &s_full[stage]);
                }
                if (++stage == S) { stage = 0; phase ^= 1; }
            }
        }
        return;
    }

    // ===== consumers: 8 warps, warp w takes rows w, w+8, ... of every tile =====
    if (blockIdx.x == 0) write_empty_docs(p, tid, kScoreThreads);
    int stage = 0; uint32_t phase = 0;
    int pend_doc = -1, pend_rows = 0, pend_n = 0;
    for (int t_idx = t0; t_idx < t1; ++t_idx) {
        mbar_wait(&s_full[stage], phase);
        const rdv_tile_desc t = s_desc[stage];
        if (t.doc != pend_doc) {
            if (pend_doc >= 0) publish_rows(p, pend_doc, pend_rows, pend_n, cache, s_red, &s_last, ConsumerSync());
            pend_doc = t.doc; pend_rows = 0; pend_n = t.doc_rows;
        }
        const float4* sq = reinterpret_cast<const float4*>(smem_raw + (size_t)stage * stage_bytes);
        const float4* se = sq + D4;
        float4 qv[VPL];
#pragma unroll
        for (int i = 0; i < VPL; ++i) qv[i] = sq[lane + 32 * i];
        const float ss_q = sumsq<VPL>(qv);
        float* __restrict__ out = p.sims + t.sims_off;
        for (int r = warp; r < t.rows; r += 2 * kScoreWarps) {
            const bool two = r + kScoreWarps < t.rows;
            float4 e0[VPL], e1[VPL];
#pragma unroll
            for (int i = 0; i < VPL; ++i) {
                e0[i] = se[(size_t)r * D4 + lane + 32 * i];
                e1[i] = two ? se[(size_t)(r + kScoreWarps) * D4 + lane + 32 * i] : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            float d0, s0, d1, s1;
            fma_row<VPL>(e0, qv, d0, s0);
            fma_row<VPL>(e1, qv, d1, s1);
            d0 = warp_sum(d0); s0 = warp_sum(s0); d1 = warp_sum(d1); s1 = warp_sum(s1);
            if (lane == 0) out[r] = cosine(d0, s0, ss_q);
            if (lane == 1 && two) out[r + kScoreWarps] = cosine(d1, s1, ss_q);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_empty[stage]);      // stage may be overwritten by the producer
        pend_rows += t.rows;
        if (++stage == S) { stage = 0; phase ^= 1; }
    }
    if (pend_doc >= 0) publish_rows(p, pend_doc, pend_rows, pend_n, cache, s_red, &s_last, ConsumerSync());
}

// ---- launch plumbing ------------------------------------------------------------------------------
template <int VPL, int ROWS>
static int launch_ldg(const ScoreParams& p, cudaStream_t stream) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(score_topk_ldg_kernel<VPL, ROWS>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(score_topk_ldg)");
        attr_set = true;
    }
    const int grid = p.total_tiles > 0 ? p.total_tiles : 1;
    const size_t smem = (size_t)p.cache_floats * sizeof(float) + 16;
    score_topk_ldg_kernel<VPL, ROWS><<<grid, kScoreThreads, smem, stream>>>(p);
    RDV_LAUNCH_CHECK("score_topk_ldg_kernel");
    return RDV_OK;
}

static int tma_stage_plan(int d, int tile_rows, int* stages) {
    const int stage_bytes = (tile_rows + 1) * d * 4;
    int s = kTmaRingBytes / stage_bytes;
    if (s > kTmaMaxStages) s = kTmaMaxStages;
    *stages = s;
    return stage_bytes;
}

template <int VPL>
static int launch_tma(ScoreParams p, cudaStream_t stream) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(score_topk_tma_kernel<VPL>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(score_topk_tma)");
        attr_set = true;
    }
    int stages = 0;
    const int stage_bytes = tma_stage_plan(p.d, p.tile_rows, &stages);
    RDV_REQUIRE(stages >= 2, RDV_E_LIMIT, "score_topk_f32: tile_rows=%d too large for the TMA ring at d=%d", p.tile_rows, p.d);
    p.stages = stages;
    const size_t smem = (size_t)stages * stage_bytes + (size_t)p.cache_floats * sizeof(float) + 16;
    int grid = sm_count();
    if (grid > p.total_tiles) grid = p.total_tiles > 0 ? p.total_tiles : 1;
    score_topk_tma_kernel<VPL><<<grid, kTmaThreads, smem, stream>>>(p);
    RDV_LAUNCH_CHECK("score_topk_tma_kernel");
    return RDV_OK;
}

static bool tma_supported(int d) {
    return d == 128 || d == 256 || d == 384 || d == 512 || d == 768 || d == 1024;
}

}  // namespace rdv

extern "C" int rdv_score_plan(int64_t total_rows, int32_t d, int32_t algo, int32_t* algo_out, int32_t* tile_rows) {
    using namespace rdv;
    RDV_REQUIRE(algo_out && tile_rows, RDV_E_INVALID, "score_plan: null output");
    RDV_REQUIRE(algo >= RDV_SCORE_AUTO && algo <= RDV_SCORE_TMA, RDV_E_INVALID, "score_plan: unknown algo %d", algo);
    // measured on B200 (profiles/): the LDG kernel is ahead at every size so far, so AUTO picks it; the TMA
    // kernel stays selectable (its consumers serialise on one tile at a time -- see DESIGN.md, next steps)
    if (algo == RDV_SCORE_AUTO) algo = RDV_SCORE_LDG;
    RDV_REQUIRE(algo != RDV_SCORE_TMA || tma_supported(d), RDV_E_INVALID,
                "score_plan: the TMA kernel supports d in {128,256,384,512,768,1024}, got %d", d);
    *algo_out = algo;
    if (algo == RDV_SCORE_TMA) {
        // ~24 KB stages: deep ring, and >= 8 tiles per SM on the small batches so the static split balances
        int rows = (24 * 1024) / (d * 4);
        rows = rows >= 16 ? 16 : 8;
        *tile_rows = rows;
    } else {
        const int64_t want_tiles = (int64_t)sm_count() * 8;
        int t = 128;
        while (t > 32 && total_rows / t < want_tiles) t >>= 1;
        *tile_rows = t;
    }
    return RDV_OK;
}

extern "C" int rdv_score_topk_f32(const rdv_tile_desc* d_tiles, int32_t total_tiles, int32_t tile_rows, int32_t algo,
                                  const int64_t* d_row_off, const float* d_q, int32_t B, int32_t d, int32_t k,
                                  int32_t max_rows, float* d_sims, int32_t* d_topk_idx, float* d_topk_val,
                                  int32_t* d_topk_cnt, int32_t* d_doc_done, void* stream) {
    using namespace rdv;
    RDV_REQUIRE(B >= 0 && total_tiles >= 0 && max_rows >= 0, RDV_E_INVALID, "score_topk_f32: negative size");
    if (B == 0) return RDV_OK;
    RDV_REQUIRE(d_row_off && d_q && d_topk_idx && d_topk_val && d_topk_cnt && d_doc_done, RDV_E_INVALID,
                "score_topk_f32: null pointer");
    RDV_REQUIRE((d_sims && d_tiles) || total_tiles == 0, RDV_E_INVALID, "score_topk_f32: null sims / tiles");
    RDV_REQUIRE(d >= 4 && d <= 8192 && (d & 3) == 0, RDV_E_INVALID,
                "score_topk_f32: d=%d must be a multiple of 4 in [4, 8192]", d);
    RDV_REQUIRE(k >= 1 && k <= 1024, RDV_E_LIMIT, "score_topk_f32: k=%d outside [1, 1024]", k);
    RDV_REQUIRE(tile_rows >= 1 && tile_rows <= 1024, RDV_E_INVALID, "score_topk_f32: tile_rows=%d outside [1, 1024]", tile_rows);
    RDV_REQUIRE(algo == RDV_SCORE_LDG || algo == RDV_SCORE_TMA, RDV_E_INVALID,
                "score_topk_f32: algo must be RDV_SCORE_LDG or RDV_SCORE_TMA (resolve AUTO with rdv_score_plan)");
    RDV_REQUIRE(aligned16(d_q) && aligned16(d_tiles), RDV_E_ALIGN, "score_topk_f32: q / tiles not 16-byte aligned");

    ScoreParams p = {};
    p.tiles = d_tiles; p.row_off = d_row_off; p.q = d_q;
    p.B = B; p.d = d; p.k = k; p.total_tiles = total_tiles; p.tile_rows = tile_rows;
    p.cache_floats = max_rows < kMaxCacheFloats ? max_rows : kMaxCacheFloats;
    p.sims = d_sims; p.topk_idx = d_topk_idx; p.topk_val = d_topk_val; p.topk_cnt = d_topk_cnt;
    p.doc_done = d_doc_done;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (algo == RDV_SCORE_TMA) {
        switch (d) {
            case 128:  return launch_tma<1>(p, s);
            case 256:  return launch_tma<2>(p, s);
            case 384:  return launch_tma<3>(p, s);
            case 512:  return launch_tma<4>(p, s);
            case 768:  return launch_tma<6>(p, s);
            case 1024: return launch_tma<8>(p, s);
            default:
                set_error("score_topk_f32: the TMA kernel supports d in {128,256,384,512,768,1024}, got %d", d);
                return RDV_E_INVALID;
        }
    }
    switch (d) {
        case 128:  return launch_ldg<1, 8>(p, s);
        case 256:  return launch_ldg<2, 4>(p, s);
        case 384:  return launch_ldg<3, 4>(p, s);
        case 512:  return launch_ldg<4, 4>(p, s);
        case 768:  return launch_ldg<6, 2>(p, s);
        case 1024: return launch_ldg<8, 2>(p, s);
        default:   return launch_ldg<0, 2>(p, s);
    }
}

// ---------------------------------------------------------------------------------------------------
// Stand-alone segmented top-k over an existing score vector (the visual path: MaxSim scores -> top-k
// strips, reference src/_modules.py:2408).  One block per document, same selection as above.
// ---------------------------------------------------------------------------------------------------
namespace rdv {
__global__ void __launch_bounds__(kScoreThreads) topk_segments_kernel(const ScoreParams p) {
    extern __shared__ float4 smem_dyn[];
    __shared__ unsigned long long s_red[kScoreWarps];
    const int b = blockIdx.x;
    const int64_t r0 = p.row_off[b];
    const int n = (int)(p.row_off[b + 1] - r0);
    select_topk(p, b, p.sims + r0, n, reinterpret_cast<float*>(smem_dyn), s_red, BlockSync());
}
}  // namespace rdv

extern "C" int rdv_topk_segments_f32(const float* d_scores, const int64_t* d_row_off, int32_t B, int32_t k,
                                     int32_t max_rows, int32_t* d_topk_idx, float* d_topk_val,
                                     int32_t* d_topk_cnt, int32_t* d_doc_done, void* stream) {
    using namespace rdv;
    RDV_REQUIRE(B >= 0 && max_rows >= 0, RDV_E_INVALID, "topk_segments_f32: negative size");
    if (B == 0) return RDV_OK;
    RDV_REQUIRE(d_row_off && d_topk_idx && d_topk_val && d_topk_cnt && d_doc_done, RDV_E_INVALID,
                "topk_segments_f32: null pointer");
    RDV_REQUIRE(d_scores || max_rows == 0, RDV_E_INVALID, "topk_segments_f32: null scores");
    RDV_REQUIRE(k >= 1 && k <= 1024, RDV_E_LIMIT, "topk_segments_f32: k=%d outside [1, 1024]", k);
    ScoreParams p = {};
    p.row_off = d_row_off; p.B = B; p.k = k;
    p.cache_floats = max_rows < kMaxCacheFloats ? max_rows : kMaxCacheFloats;
    p.sims = const_cast<float*>(d_scores);
    p.topk_idx = d_topk_idx; p.topk_val = d_topk_val; p.topk_cnt = d_topk_cnt; p.doc_done = d_doc_done;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(topk_segments_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             64 * 1024);
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(topk_segments)");
        attr_set = true;
    }
    const size_t smem = (size_t)p.cache_floats * sizeof(float) + 16;
    topk_segments_kernel<<<B, kScoreThreads, smem, static_cast<cudaStream_t>(stream)>>>(p);
    RDV_LAUNCH_CHECK("topk_segments_kernel");
    return RDV_OK;
}
