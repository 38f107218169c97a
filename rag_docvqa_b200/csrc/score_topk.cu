// Fused cosine score + segmented per-document top-k, fp32 FFMA parity mode (sm_100a).
//
// Replaces Retriever._get_similarities (reference src/_modules.py:1978-1997) and the per-document
// torch.topk (src/_modules.py:2015-2016) -- see include/rdv.h for the contract.
//
// Shape of the work: one question per document, so this is a batch of ragged GEMVs: 0.5 flop/byte,
// HBM-bound, and for the small batches the reference actually runs (tens of MB) latency-bound.  The
// batch is cut on the host into row tiles that never cross a document (rdv_tile_desc, 32 bytes each).
//
// Kernels (all read every byte of E from HBM exactly once and write every similarity -- the reference
// returns the full vector, src/_modules.py:2176-2180):
//
//  * score_ldg_kernel (any d % 4 == 0): one block per tile, a warp per row, ROWS rows in flight per warp as
//    independent 128-bit no-allocate loads; the question vector is loaded in the same burst.  Scores only: the
//    selection runs in topk_segments_kernel / inside the gather kernel, so no device-scope fence or atomic sits on
//    the streaming path.  (Batches of short documents take retrieve_cluster.cu instead: one launch for score, top-k
//    and gather, the hand-over inside a thread-block cluster.  A one-launch variant of THIS kernel -- the block that
//    completes a document selects, told by a device-scope counter -- was measured at 9.7 us (score + top-k) / 14.0 us
//    (+ gather) against 6.9 + 5.8 us for two launches at C2, and 617 against 610 us at C3: every stage after the
//    streaming phase is a dependent global round trip, fence -> atomic -> re-read.  Removed.)
//  * score_tma_kernel (d in {128,256,384,512,768,1024}): persistent, one block per SM owning a contiguous
//    run of tiles; every warp owns a private shared-memory ring and ITS OWN mbarriers, requests its tiles with
//    1-D bulk async copies (cp.async.bulk -> SASS UBLKCP) and consumes them with conflict-free LDS.128 -- no
//    producer warp and no cross-warp synchronisation.  192 KB in flight per SM independent of occupancy and of
//    registers.  Measured on B200: 7.1 TB/s at C3 (595 us per batch; LDG kernel 593 us), 12.8 us at C2 (LDG 6.9 us: a 32 MB batch
//    is launch/latency-bound and the ring adds a descriptor -> copy -> wait chain), so AUTO stays on LDG.
//
// Selection (select.cuh): packed (score, ~index) u64 keys make "descending score, lowest index first" one integer
// compare; register-resident, warp-level extraction (redux.sync) for k <= 32, block-wide rounds above.
#include "select.cuh"

namespace rdv {

constexpr int kTmaThreads = kScoreThreads;        // 8 self-serving warps
constexpr int kTmaMaxStages = 4;                  // per consumer warp
constexpr int kTmaRingBytes = 192 * 1024;         // all warps

struct ScoreParams {
    const rdv_tile_desc* tiles;
    const int64_t* row_off;
    const float* q;
    int32_t B, d, total_tiles, reserved;
    int32_t tile_rows, stages;                // TMA kernel: rows per stage (max), ring depth per warp
    float* sims;
    SelectArgs sel;
};

__device__ __forceinline__ float cosine(float dot, float ss_e, float ss_q) {
    // reference: dot / (||e|| * ||q|| + 1e-8), all fp32, IEEE sqrt and divide
    return __fdiv_rn(dot, __fadd_rn(__fmul_rn(__fsqrt_rn(ss_e), __fsqrt_rn(ss_q)), 1e-8f));
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
template <int VPL, int ROWS, int MINB>
__global__ void __launch_bounds__(kScoreThreads, MINB) score_ldg_kernel(const ScoreParams p) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int d4 = p.d >> 2;

    pdl_launch_dependents();      // the selection / gather kernel may be scheduled behind this grid's last wave
    pdl_wait();                   // embeddings, questions and descriptors come from earlier work in the stream

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
}

// =====================================================================================================
// TMA kernel: persistent, per-warp self-service rings of bulk async copies.
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

// Self-service ring: every warp owns S stages and their mbarriers, issues its OWN bulk copies (lane 0) and consumes
// them; no producer warp, no cross-warp synchronisation.  A warp's tiles are li = warp, warp + 8, ... of the block's
// contiguous run; the first S are requested before anything is consumed, and a stage is refilled as soon as the
// warp has read it.  (The round-1 kernel had one producer lane feeding all eight rings: at ~1000 cycles per 12 KB
// copy it was the limiter -- 2.6 TB/s at C2, 3.2-4.2 TB/s at C3.)
template <int VPL>
__global__ void __launch_bounds__(kTmaThreads, 1) score_tma_kernel(const ScoreParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t s_full[kTmaThreads / 32][kTmaMaxStages];

    constexpr int D4 = 32 * VPL;                       // float4 per row
    constexpr uint32_t kRowBytes = D4 * 16;
    constexpr int NW = kTmaThreads / 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int S = p.stages;
    const uint32_t stage_bytes = (uint32_t)p.tile_rows * kRowBytes;

    pdl_launch_dependents();
    pdl_wait();

    // contiguous run of tiles for this block; local tile li belongs to warp li % NW
    const int G = gridDim.x;
    const int t0 = (int)((long long)blockIdx.x * p.total_tiles / G);
    const int t1 = (int)((long long)(blockIdx.x + 1) * p.total_tiles / G);
    const int ntiles = t1 - t0;
    const int mine_n = ntiles > warp ? (ntiles - warp + NW - 1) / NW : 0;      // tiles of this warp

    if (lane == 0) {
        for (int s = 0; s < S; ++s) mbar_init(&s_full[warp][s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    unsigned char* ring = smem_raw + (size_t)warp * S * stage_bytes;

    int cur_doc = -1;
    float4 qv[VPL];
    float ss_q = 0.f;
    rdv_tile_desc next = {};                             // descriptors: 32 at a time, one per lane, for issue and for use
    rdv_tile_desc mine = {};
    int issued = 0;
    auto issue = [&](int u) {                            // lane 0: request local tile u of this warp into stage u % S
        const int j = u & 31;
        const void* src = reinterpret_cast<const void*>(__shfl_sync(0xffffffffu, (unsigned long long)next.src, j));
        const int rows = __shfl_sync(0xffffffffu, next.rows, j);
        if (lane == 0) {
            const uint32_t bytes = (uint32_t)rows * kRowBytes;
            uint64_t* bar = &s_full[warp][u % S];
            mbar_expect_tx(bar, bytes);
            bulk_g2s(ring + (size_t)(u % S) * stage_bytes, src, bytes, bar);
        }
    };
    auto load_descs = [&](int ubase) {                   // descriptors of local tiles ubase .. ubase + 31
        rdv_tile_desc d = {};
        const int li = warp + NW * (ubase + lane);
        if (ubase + lane < mine_n) d = p.tiles[t0 + li];
        return d;
    };
    next = load_descs(0);
    for (; issued < mine_n && issued < S; ++issued) issue(issued);             // prologue (all within the first 32)

    for (int u = 0; u < mine_n; ++u) {
        if ((u & 31) == 0) mine = (u == 0) ? next : load_descs(u);
        const int j = u & 31;
        const int doc = __shfl_sync(0xffffffffu, mine.doc, j);
        const int rows = __shfl_sync(0xffffffffu, mine.rows, j);
        const long long sims_off = __shfl_sync(0xffffffffu, mine.sims_off, j);
        if (doc != cur_doc) {                              // question vector: issued before waiting on the copy
            const float4* Q = reinterpret_cast<const float4*>(p.q) + (size_t)doc * D4;
#pragma unroll
            for (int i = 0; i < VPL; ++i) qv[i] = __ldg(Q + lane + 32 * i);
            ss_q = sumsq<VPL>(qv);
            cur_doc = doc;
        }
        const int st = u % S;
        mbar_wait(&s_full[warp][st], (uint32_t)(u / S) & 1u);
        const float4* se = reinterpret_cast<const float4*>(ring + (size_t)st * stage_bytes);
        float* __restrict__ out = p.sims + sims_off;
        for (int r = 0; r < rows; r += 2) {
            const bool two = r + 1 < rows;
            float4 e0[VPL], e1[VPL];
#pragma unroll
            for (int i = 0; i < VPL; ++i) {
                e0[i] = se[(size_t)r * D4 + lane + 32 * i];
                e1[i] = two ? se[(size_t)(r + 1) * D4 + lane + 32 * i] : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            float d0, s0, d1, s1;
            fma_row<VPL>(e0, qv, d0, s0);
            fma_row<VPL>(e1, qv, d1, s1);
            d0 = warp_sum(d0); s0 = warp_sum(s0); d1 = warp_sum(d1); s1 = warp_sum(s1);
            if (lane == 0) out[r] = cosine(d0, s0, ss_q);
            if (lane == 1 && two) out[r + 1] = cosine(d1, s1, ss_q);
        }
        __syncwarp();                                      // every lane has consumed the stage
        if (issued < mine_n) {                             // refill it with this warp's tile u + S
            if ((issued & 31) == 0) next = load_descs(issued);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic reads before the async-proxy write
            issue(issued);
            ++issued;
        }
    }
}

// ---- launch plumbing ------------------------------------------------------------------------------
template <int VPL, int ROWS, int MINB>
static int launch_ldg(const ScoreParams& p, cudaStream_t stream) {
    if (p.total_tiles == 0) return RDV_OK;
    cudaError_t e = launch_pdl(kPdlStream, score_ldg_kernel<VPL, ROWS, MINB>, dim3(p.total_tiles), dim3(kScoreThreads), 0, stream, p);
    if (e != cudaSuccess) return cuda_fail(e, "score_ldg_kernel");
    return RDV_OK;
}

// rows per stage and stages per warp for the TMA kernel: 24 KB per warp, as ONE 24 KB stage for large batches
// (measured at C3: 7.05 TB/s, vs 6.5 TB/s with two 12 KB stages) and two 12 KB stages for small ones
static void tma_plan(int d, int* tile_rows, int* stages, int64_t total_rows = 0) {
    const bool large = total_rows * (int64_t)d * 4 >= (256ll << 20);
    int rows = ((large ? 24 : 12) * 1024) / (d * 4);
    rows = rows < 1 ? 1 : (rows > 16 ? 16 : rows);
    int s = (kTmaRingBytes / kScoreWarps) / (rows * d * 4);
    s = s > kTmaMaxStages ? kTmaMaxStages : s;
    *tile_rows = rows;
    *stages = s;
}

template <int VPL>
static int launch_tma(ScoreParams p, cudaStream_t stream) {
    RDV_ONCE_PER_DEVICE(cudaFuncSetAttribute(score_tma_kernel<VPL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             kTmaRingBytes + 1024), "cudaFuncSetAttribute(score_tma)");
    if (p.total_tiles == 0) return RDV_OK;
    int max_rows = 0, stages = 0;
    tma_plan(p.d, &max_rows, &stages);
    RDV_REQUIRE(p.tile_rows >= 1 && p.tile_rows <= 16, RDV_E_LIMIT, "score (TMA): tile_rows=%d outside [1, 16]", p.tile_rows);
    stages = (kTmaRingBytes / kScoreWarps) / (p.tile_rows * p.d * 4);
    if (stages > kTmaMaxStages) stages = kTmaMaxStages;
    RDV_REQUIRE(stages >= 1, RDV_E_LIMIT, "score (TMA): tile_rows=%d too large for the ring at d=%d", p.tile_rows, p.d);
    p.stages = stages;
    const size_t smem = (size_t)kScoreWarps * stages * p.tile_rows * p.d * 4 + 128;
    int grid = sm_count();
    if (grid > p.total_tiles) grid = p.total_tiles;
    score_tma_kernel<VPL><<<grid, kTmaThreads, smem, stream>>>(p);
    RDV_LAUNCH_CHECK("score_tma_kernel");
    return RDV_OK;
}

static bool tma_supported(int d) {
    return d == 128 || d == 256 || d == 384 || d == 512 || d == 768 || d == 1024;
}

static int launch_stream(const ScoreParams& p, int algo, cudaStream_t s) {
    if (algo == RDV_SCORE_TMA) {
        switch (p.d) {
            case 128:  return launch_tma<1>(p, s);
            case 256:  return launch_tma<2>(p, s);
            case 384:  return launch_tma<3>(p, s);
            case 512:  return launch_tma<4>(p, s);
            case 768:  return launch_tma<6>(p, s);
            case 1024: return launch_tma<8>(p, s);
            default:
                set_error("score: the TMA kernel supports d in {128,256,384,512,768,1024}, got %d", p.d);
                return RDV_E_INVALID;
        }
    }
    // Two shapes per width (measured with scripts/probe_stream.cu on B200).  Small batches (the plan gives them
    // <= 32-row tiles) are launch/latency-bound: fewer rows in flight per warp, <= 51 registers, 5 blocks per SM
    // so the whole batch is resident in ~1 wave.  Large batches keep more rows in flight per warp.
    const bool small = p.tile_rows <= 32;
    switch (p.d) {
        case 128:  return small ? launch_ldg<1, 4, 5>(p, s) : launch_ldg<1, 8, 3>(p, s);
        case 256:  return small ? launch_ldg<2, 2, 5>(p, s) : launch_ldg<2, 4, 3>(p, s);
        case 384:  return small ? launch_ldg<3, 2, 5>(p, s) : launch_ldg<3, 4, 3>(p, s);
        case 512:  return small ? launch_ldg<4, 2, 4>(p, s) : launch_ldg<4, 4, 2>(p, s);
        case 768:  return small ? launch_ldg<6, 1, 4>(p, s) : launch_ldg<6, 2, 2>(p, s);
        case 1024: return small ? launch_ldg<8, 1, 3>(p, s) : launch_ldg<8, 2, 2>(p, s);
        default:   return launch_ldg<0, 2, 2>(p, s);
    }
}

__global__ void __launch_bounds__(kScoreThreads) topk_segments_kernel(const int64_t* __restrict__ row_off,
                                                                      const float* __restrict__ scores,
                                                                      const SelectArgs sel) {
    extern __shared__ float4 smem_dyn[];
    __shared__ unsigned long long s_red[kScoreWarps];
    const int b = blockIdx.x;
    pdl_launch_dependents();
    pdl_wait();
    const int64_t r0 = row_off[b];
    const int n = (int)(row_off[b + 1] - r0);
    select_topk<40>(sel, b, scores + r0, n, reinterpret_cast<float*>(smem_dyn), s_red, BlockSync());
}

// Second level of pooled-patch retrieval, one block per document: the k best of the document's (strips x kk) per-strip
// candidates, mapped back to patch indices; the strips' own scores (a strip's rank-0 candidate = its best patch); the
// k_strips best strips.
__global__ void __launch_bounds__(kScoreThreads) pooled_doc_kernel(const int64_t* __restrict__ doc_strip_off, int L, int kk,
                                                                   const int32_t* __restrict__ l_idx,
                                                                   const float* __restrict__ l_val, SelectArgs patch,
                                                                   float* __restrict__ strip_scores, SelectArgs strips) {
    __shared__ unsigned long long s_red[kScoreWarps];
    const int b = blockIdx.x, tid = threadIdx.x;
    pdl_launch_dependents();
    pdl_wait();
    const int64_t s0 = doc_strip_off[b];
    const int ns = (int)(doc_strip_off[b + 1] - s0);
    for (int s = tid; s < ns; s += kScoreThreads) strip_scores[s0 + s] = __ldcg(l_val + (s0 + s) * kk);
    select_topk<16>(patch, b, l_val + s0 * kk, ns * kk, nullptr, s_red, BlockSync());
    __syncthreads();                               // the positions are written; the strip scores are visible to the block
    if (tid < patch.k) {
        const size_t o = (size_t)b * patch.k + tid;
        const int pos = patch.topk_idx[o];
        if (pos >= 0) patch.topk_idx[o] = (pos / kk) * L + __ldcg(l_idx + s0 * kk + pos);     // strip in the document, patch in the strip
    }
    select_topk<16>(strips, b, strip_scores + s0, ns, nullptr, s_red, BlockSync());
}

static int launch_segments(const float* scores, const int64_t* row_off, int B, const SelectArgs& sel, cudaStream_t s) {
    RDV_ONCE_PER_DEVICE(cudaFuncSetAttribute(topk_segments_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024),
                        "cudaFuncSetAttribute(topk_segments)");
    const size_t smem = (size_t)sel.cache_floats * sizeof(float) + 16;
    cudaError_t e = launch_pdl(kPdlSelect, topk_segments_kernel, dim3(B), dim3(kScoreThreads), smem, s, row_off, scores, sel);
    if (e != cudaSuccess) return cuda_fail(e, "topk_segments_kernel");
    return RDV_OK;
}

}  // namespace rdv

using namespace rdv;

extern "C" int rdv_score_plan(int64_t total_rows, int32_t d, int32_t algo, int32_t* algo_out, int32_t* tile_rows) {
    RDV_REQUIRE(algo_out && tile_rows, RDV_E_INVALID, "score_plan: null output");
    RDV_REQUIRE(algo >= RDV_SCORE_AUTO && algo <= RDV_SCORE_TMA, RDV_E_INVALID, "score_plan: unknown algo %d", algo);
    // measured on B200 (profiles/, scripts/probe_tma.py): the LDG streaming kernel leads at every size (C2 6.9 us vs
    // 12.8 us, C3 7.2 vs 7.05 TB/s for the self-service TMA ring), so AUTO picks it; RDV_SCORE_TMA stays selectable
    if (algo == RDV_SCORE_AUTO) algo = RDV_SCORE_LDG;
    RDV_REQUIRE(algo != RDV_SCORE_TMA || tma_supported(d), RDV_E_INVALID,
                "score_plan: the TMA kernel supports d in {128,256,384,512,768,1024}, got %d", d);
    *algo_out = algo;
    if (algo == RDV_SCORE_TMA) {
        int stages = 0;
        tma_plan(d, tile_rows, &stages, total_rows);
    } else {
        // one block per tile: >= ~8 tiles per SM on small batches so the hardware scheduler balances ragged docs
        const int64_t want_tiles = (int64_t)sm_count() * 8;
        int t = 128;
        while (t > 32 && total_rows / t < want_tiles) t >>= 1;
        *tile_rows = t;
    }
    return RDV_OK;
}

static int check_score_args(const char* who, const rdv_tile_desc* d_tiles, int32_t total_tiles, int32_t tile_rows,
                            int32_t algo, const float* d_q, int32_t B, int32_t d, const float* d_sims) {
    RDV_REQUIRE(B >= 0 && total_tiles >= 0, RDV_E_INVALID, "%s: negative size", who);
    RDV_REQUIRE(d_q, RDV_E_INVALID, "%s: null q", who);
    RDV_REQUIRE((d_sims && d_tiles) || total_tiles == 0, RDV_E_INVALID, "%s: null sims / tiles", who);
    RDV_REQUIRE(d >= 4 && d <= 8192 && (d & 3) == 0, RDV_E_INVALID, "%s: d=%d must be a multiple of 4 in [4, 8192]", who, d);
    RDV_REQUIRE(tile_rows >= 1 && tile_rows <= 1024, RDV_E_INVALID, "%s: tile_rows=%d outside [1, 1024]", who, tile_rows);
    RDV_REQUIRE(algo == RDV_SCORE_LDG || algo == RDV_SCORE_TMA, RDV_E_INVALID,
                "%s: algo must be RDV_SCORE_LDG or RDV_SCORE_TMA (resolve AUTO with rdv_score_plan)", who);
    RDV_REQUIRE(aligned16(d_q) && aligned16(d_tiles), RDV_E_ALIGN, "%s: q / tiles not 16-byte aligned", who);
    return RDV_OK;
}

extern "C" int rdv_score_f32(const rdv_tile_desc* d_tiles, int32_t total_tiles, int32_t tile_rows, int32_t algo,
                             const float* d_q, int32_t B, int32_t d, float* d_sims, void* stream) {
    int rc = check_score_args("score_f32", d_tiles, total_tiles, tile_rows, algo, d_q, B, d, d_sims);
    if (rc) return rc;
    if (B == 0 || total_tiles == 0) return RDV_OK;
    ScoreParams p = {};
    p.tiles = d_tiles; p.q = d_q; p.B = B; p.d = d; p.total_tiles = total_tiles; p.tile_rows = tile_rows;
    p.sims = d_sims;
    return launch_stream(p, algo, static_cast<cudaStream_t>(stream));
}

extern "C" int rdv_score_topk_f32(const rdv_tile_desc* d_tiles, int32_t total_tiles, int32_t tile_rows, int32_t algo,
                                  const int64_t* d_row_off, const float* d_q, int32_t B, int32_t d, int32_t k,
                                  int32_t max_rows, float* d_sims, int32_t* d_topk_idx, float* d_topk_val,
                                  int32_t* d_topk_cnt, void* stream) {
    int rc = check_score_args("score_topk_f32", d_tiles, total_tiles, tile_rows, algo, d_q, B, d, d_sims);
    if (rc) return rc;
    RDV_REQUIRE(max_rows >= 0, RDV_E_INVALID, "score_topk_f32: negative size");
    if (B == 0) return RDV_OK;
    RDV_REQUIRE(d_row_off && d_topk_idx && d_topk_val && d_topk_cnt, RDV_E_INVALID, "score_topk_f32: null pointer");
    RDV_REQUIRE(k >= 1 && k <= 1024, RDV_E_LIMIT, "score_topk_f32: k=%d outside [1, 1024]", k);
    ScoreParams p = {};
    p.tiles = d_tiles; p.row_off = d_row_off; p.q = d_q;
    p.B = B; p.d = d; p.total_tiles = total_tiles; p.tile_rows = tile_rows; p.sims = d_sims;
    p.sel.k = k;
    p.sel.cache_floats = cache_floats_for(max_rows, k, 40 * kScoreThreads);
    p.sel.topk_idx = d_topk_idx; p.sel.topk_val = d_topk_val; p.sel.topk_cnt = d_topk_cnt; p.sel.doc_done = nullptr;
    p.sel.smem_idx = nullptr;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    rc = launch_stream(p, algo, s);
    if (rc) return rc;
    return launch_segments(d_sims, d_row_off, B, p.sel, s);
}

extern "C" int rdv_pooled_select_f32(const float* d_sims, const int64_t* d_strip_row_off, int64_t n_strips, int32_t L,
                                     const int64_t* d_doc_strip_off, int32_t B, int32_t max_strips, int32_t k, int32_t k_strips,
                                     int32_t* d_ws_idx, float* d_ws_val, int32_t* d_ws_cnt, int32_t* d_patch_idx,
                                     float* d_patch_val, int32_t* d_patch_cnt, float* d_strip_scores, int32_t* d_strip_idx,
                                     float* d_strip_val, int32_t* d_strip_cnt, void* stream) {
    RDV_REQUIRE(B >= 0 && n_strips >= 0 && n_strips < (1ll << 31) && L >= 1 && max_strips >= 0, RDV_E_INVALID, "pooled_select_f32: bad sizes");
    RDV_REQUIRE(k >= 1 && k <= kSelWarpK && k_strips >= 1 && k_strips <= kSelWarpK, RDV_E_LIMIT,
                "pooled_select_f32: k=%d / k_strips=%d outside [1, %d]", k, k_strips, kSelWarpK);
    const int kk = k < L ? k : L;
    RDV_REQUIRE((int64_t)max_strips * kk <= 16 * kScoreThreads, RDV_E_LIMIT,
                "pooled_select_f32: %d strips x %d candidates exceed the %d a block selects from registers", max_strips, kk,
                16 * kScoreThreads);
    if (B == 0) return RDV_OK;
    RDV_REQUIRE(d_doc_strip_off && d_patch_idx && d_patch_val && d_patch_cnt && d_strip_idx && d_strip_val && d_strip_cnt,
                RDV_E_INVALID, "pooled_select_f32: null pointer");
    RDV_REQUIRE(n_strips == 0 || (d_sims && d_strip_row_off && d_ws_idx && d_ws_val && d_ws_cnt && d_strip_scores), RDV_E_INVALID,
                "pooled_select_f32: null pointer");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (n_strips > 0) {
        SelectArgs l = {};
        l.k = kk; l.cache_floats = cache_floats_for(L, kk, 40 * kScoreThreads);
        l.topk_idx = d_ws_idx; l.topk_val = d_ws_val; l.topk_cnt = d_ws_cnt;
        int rc = launch_segments(d_sims, d_strip_row_off, (int)n_strips, l, s);          // k best patches of every strip
        if (rc) return rc;
    }
    SelectArgs p = {}, q = {};
    p.k = k; p.topk_idx = d_patch_idx; p.topk_val = d_patch_val; p.topk_cnt = d_patch_cnt;
    q.k = k_strips; q.topk_idx = d_strip_idx; q.topk_val = d_strip_val; q.topk_cnt = d_strip_cnt;
    cudaError_t e = launch_pdl(kPdlSelect, pooled_doc_kernel, dim3(B), dim3(kScoreThreads), 0, s, d_doc_strip_off, (int)L, kk,
                               (const int32_t*)d_ws_idx, (const float*)d_ws_val, p, d_strip_scores, q);
    if (e != cudaSuccess) return cuda_fail(e, "pooled_doc_kernel");
    return RDV_OK;
}

extern "C" int rdv_topk_segments_f32(const float* d_scores, const int64_t* d_row_off, int32_t B, int32_t k,
                                     int32_t max_rows, int32_t* d_topk_idx, float* d_topk_val,
                                     int32_t* d_topk_cnt, void* stream) {
    RDV_REQUIRE(B >= 0 && max_rows >= 0, RDV_E_INVALID, "topk_segments_f32: negative size");
    if (B == 0) return RDV_OK;
    RDV_REQUIRE(d_row_off && d_topk_idx && d_topk_val && d_topk_cnt, RDV_E_INVALID, "topk_segments_f32: null pointer");
    RDV_REQUIRE(d_scores || max_rows == 0, RDV_E_INVALID, "topk_segments_f32: null scores");
    RDV_REQUIRE(k >= 1 && k <= 1024, RDV_E_LIMIT, "topk_segments_f32: k=%d outside [1, 1024]", k);
    SelectArgs sel = {};
    sel.k = k; sel.cache_floats = cache_floats_for(max_rows, k, 40 * kScoreThreads);
    sel.topk_idx = d_topk_idx; sel.topk_val = d_topk_val; sel.topk_cnt = d_topk_cnt; sel.doc_done = nullptr;
    sel.smem_idx = nullptr;
    return launch_segments(d_scores, d_row_off, B, sel, static_cast<cudaStream_t>(stream));
}
