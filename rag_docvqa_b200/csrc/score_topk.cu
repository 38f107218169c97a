// Fused cosine score + segmented per-document top-k, fp32 FFMA parity mode (sm_100a).
//
// Replaces Retriever._get_similarities (reference src/_modules.py:1978-1997) and the per-document
// torch.topk (src/_modules.py:2015-2016) -- see include/rdv.h for the contract.
//
// Shape of the work: one question per document, so this is a batch of ragged GEMVs: 0.5 flop/byte,
// HBM-bound.  Design:
//   * a thread block owns one tile of `tile_rows` consecutive chunks of ONE document; the hardware
//     block scheduler balances the ragged documents (tiles are small when the batch is small);
//   * a warp owns whole rows: lane l reads float4 number l, l+32, ... of the row, so every load
//     instruction of a warp covers 512 contiguous bytes; ROWS rows are in flight per warp
//     (ROWS*VPL independent 128-bit loads per thread) and the row is read exactly ONCE: the dot product
//     and the squared norm come out of the same registers;
//   * the question vector is staged once per block in shared memory, then held in registers;
//   * every similarity is written (the reference returns the full vector, src/_modules.py:2176-2180);
//   * the block that completes a document's last tile (device-scope counter) runs the selection for that
//     document out of L2/shared memory: k rounds of a block-wide arg-max over packed (score, ~index)
//     keys, which makes "descending score, lowest index first" a single u64 comparison.
#include "rdv_common.cuh"

namespace rdv {

constexpr int kScoreThreads = 256;
constexpr int kScoreWarps = kScoreThreads / 32;
constexpr int kMaxSmemDocs = 1024;       // tile_off staged in shared memory when B <= this
constexpr int kMaxCacheFloats = 12288;   // selection pass caches up to this many scores (48 KB)

struct ScoreParams {
    const void* const* doc_ptr;
    const int64_t* row_off;
    const int32_t* tile_off;
    const float* q;
    int32_t B, d, k, tile_rows, total_tiles, cache_floats;
    float* sims;
    int32_t* topk_idx;
    float* topk_val;
    int32_t* topk_cnt;
    int32_t* doc_done;
};

// largest b in [0, B) with off[b] <= tile  (documents with zero tiles are skipped naturally)
__device__ __forceinline__ int find_doc(const int32_t* off, int B, int tile) {
    int lo = 0, hi = B;
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (off[mid] <= tile) lo = mid; else hi = mid;
    }
    return lo;
}

__device__ __forceinline__ float cosine(float dot, float ss_e, float ss_q) {
    // reference: dot / (||e|| * ||q|| + 1e-8), all fp32, IEEE sqrt and divide
    return __fdiv_rn(dot, __fadd_rn(__fmul_rn(__fsqrt_rn(ss_e), __fsqrt_rn(ss_q)), 1e-8f));
}

// Block-wide selection of the k best (score desc, index asc) among n scores of one document.
__device__ void select_topk(const ScoreParams& p, int b, const float* __restrict__ src, int n,
                            float* cache, unsigned long long* s_red) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool cached = n <= p.cache_floats;
    if (cached) {
        for (int i = tid; i < n; i += kScoreThreads) cache[i] = __ldcg(src + i);
        __syncthreads();
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
        __syncthreads();
        unsigned long long win = s_red[0];
#pragma unroll
        for (int w = 1; w < kScoreWarps; ++w) win = s_red[w] > win ? s_red[w] : win;
        __syncthreads();
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

// VPL > 0: d == 128 * VPL, question and rows fully in registers.  VPL == 0: any d % 4 == 0.
template <int VPL, int ROWS>
__global__ void __launch_bounds__(kScoreThreads) score_topk_f32_kernel(const ScoreParams p) {
    extern __shared__ float4 smem_dyn[];
    __shared__ int32_t s_tile_off[kMaxSmemDocs + 1];
    __shared__ unsigned long long s_red[kScoreWarps];
    __shared__ int s_last;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int d4 = p.d >> 2;
    float4* s_q = smem_dyn;
    float* cache = reinterpret_cast<float*>(smem_dyn + d4);

    // documents with no chunks own no tile: block 0 writes their (empty) results
    if (blockIdx.x == 0) {
        for (int b = tid; b < p.B; b += kScoreThreads) {
            if (p.row_off[b + 1] == p.row_off[b]) {
                p.topk_cnt[b] = 0;
                for (int r = 0; r < p.k; ++r) {
                    p.topk_idx[(size_t)b * p.k + r] = -1;
                    p.topk_val[(size_t)b * p.k + r] = -INFINITY;
                }
            }
        }
    }
    const int tile = blockIdx.x;
    if (tile >= p.total_tiles) return;

    const int32_t* toff = p.tile_off;
    if (p.B <= kMaxSmemDocs) {
        for (int i = tid; i <= p.B; i += kScoreThreads) s_tile_off[i] = p.tile_off[i];
        __syncthreads();
        toff = s_tile_off;
    }
    const int b = find_doc(toff, p.B, tile);
    const int tile_in_doc = tile - toff[b];
    const int tiles_in_doc = toff[b + 1] - toff[b];
    const int64_t r0 = p.row_off[b];
    const int n = (int)(p.row_off[b + 1] - r0);
    const float4* __restrict__ E = reinterpret_cast<const float4*>(p.doc_ptr[b]);
    const float4* __restrict__ Q = reinterpret_cast<const float4*>(p.q) + (size_t)b * d4;

    for (int i = tid; i < d4; i += kScoreThreads) s_q[i] = Q[i];
    __syncthreads();

    const int rows_per_warp = p.tile_rows / kScoreWarps;
    const int row_base = tile_in_doc * p.tile_rows + warp * rows_per_warp;
    float* __restrict__ out = p.sims + r0;

    if constexpr (VPL > 0) {
        float4 qv[VPL];
        float ss_q = 0.f;
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            qv[i] = s_q[lane + 32 * i];
            ss_q = fmaf(qv[i].x, qv[i].x, ss_q); ss_q = fmaf(qv[i].y, qv[i].y, ss_q);
            ss_q = fmaf(qv[i].z, qv[i].z, ss_q); ss_q = fmaf(qv[i].w, qv[i].w, ss_q);
        }
        ss_q = warp_sum(ss_q);
        for (int r = 0; r < rows_per_warp; r += ROWS) {
            float4 ev[ROWS][VPL];
#pragma unroll
            for (int j = 0; j < ROWS; ++j) {
                const int row = row_base + r + j;
                const bool ok = (r + j < rows_per_warp) && (row < n);
                const float4* src = E + (size_t)(ok ? row : 0) * d4 + lane;
#pragma unroll
                for (int i = 0; i < VPL; ++i)
                    ev[j][i] = ok ? ldg_stream(src + 32 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            float mine = 0.f;
#pragma unroll
            for (int j = 0; j < ROWS; ++j) {
                float dot = 0.f, ss = 0.f;
#pragma unroll
                for (int i = 0; i < VPL; ++i) {
                    dot = fmaf(ev[j][i].x, qv[i].x, dot); ss = fmaf(ev[j][i].x, ev[j][i].x, ss);
                    dot = fmaf(ev[j][i].y, qv[i].y, dot); ss = fmaf(ev[j][i].y, ev[j][i].y, ss);
                    dot = fmaf(ev[j][i].z, qv[i].z, dot); ss = fmaf(ev[j][i].z, ev[j][i].z, ss);
                    dot = fmaf(ev[j][i].w, qv[i].w, dot); ss = fmaf(ev[j][i].w, ev[j][i].w, ss);
                }
                dot = warp_sum(dot);
                ss = warp_sum(ss);
                const float sim = cosine(dot, ss, ss_q);
                if (lane == j) mine = sim;
            }
            const int row = row_base + r + lane;
            if (lane < ROWS && r + lane < rows_per_warp && row < n) out[row] = mine;
        }
    } else {
        float ss_q = 0.f;
        for (int i = lane; i < d4; i += 32) {
            const float4 v = s_q[i];
            ss_q = fmaf(v.x, v.x, ss_q); ss_q = fmaf(v.y, v.y, ss_q);
            ss_q = fmaf(v.z, v.z, ss_q); ss_q = fmaf(v.w, v.w, ss_q);
        }
        ss_q = warp_sum(ss_q);
        for (int r = 0; r < rows_per_warp; r += ROWS) {
            float dot[ROWS], ss[ROWS];
            const float4* src[ROWS];
            bool ok[ROWS];
#pragma unroll
            for (int j = 0; j < ROWS; ++j) {
                const int row = row_base + r + j;
                ok[j] = (r + j < rows_per_warp) && (row < n);
                src[j] = E + (size_t)(ok[j] ? row : 0) * d4;
                dot[j] = 0.f; ss[j] = 0.f;
            }
#pragma unroll 2
            for (int i = lane; i < d4; i += 32) {
                const float4 qv = s_q[i];
#pragma unroll
                for (int j = 0; j < ROWS; ++j) {
                    const float4 e = ok[j] ? ldg_stream(src[j] + i) : make_float4(0.f, 0.f, 0.f, 0.f);
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
            const int row = row_base + r + lane;
            if (lane < ROWS && r + lane < rows_per_warp && row < n) out[row] = mine;
        }
    }

    // last-tile-done: the block that finishes the document runs its selection
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const int prev = atomicAdd(p.doc_done + b, 1);
        s_last = (prev == tiles_in_doc - 1);
        __threadfence();
    }
    __syncthreads();
    if (s_last) select_topk(p, b, out, n, cache, s_red);
}

template <int VPL, int ROWS>
static int launch_score(const ScoreParams& p, size_t smem, cudaStream_t stream) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(score_topk_f32_kernel<VPL, ROWS>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(score_topk_f32)");
        attr_set = true;
    }
    const int grid = p.total_tiles > 0 ? p.total_tiles : 1;
    score_topk_f32_kernel<VPL, ROWS><<<grid, kScoreThreads, smem, stream>>>(p);
    RDV_LAUNCH_CHECK("score_topk_f32_kernel");
    return RDV_OK;
}

}  // namespace rdv

extern "C" int32_t rdv_score_tile_rows(int64_t total_rows, int32_t d) {
    (void)d;
    const int64_t want_tiles = (int64_t)rdv::sm_count() * 8;
    for (int t = 128; t > 8; t >>= 1)
        if (total_rows / t >= want_tiles) return t;
    return 8;
}

extern "C" int rdv_score_topk_f32(const void* const* d_doc_ptr, const int64_t* d_row_off,
                                  const int32_t* d_tile_off, const float* d_q, int32_t B, int32_t d,
                                  int32_t k, int32_t tile_rows, int32_t total_tiles, int32_t max_rows,
                                  float* d_sims, int32_t* d_topk_idx, float* d_topk_val,
                                  int32_t* d_topk_cnt, int32_t* d_doc_done, void* stream) {
    using namespace rdv;
    RDV_REQUIRE(B >= 0 && total_tiles >= 0 && max_rows >= 0, RDV_E_INVALID, "score_topk_f32: negative size");
    if (B == 0) return RDV_OK;
    RDV_REQUIRE(d_doc_ptr && d_row_off && d_tile_off && d_q && d_topk_idx && d_topk_val && d_topk_cnt &&
                d_doc_done, RDV_E_INVALID, "score_topk_f32: null pointer");
    RDV_REQUIRE(d_sims || total_tiles == 0, RDV_E_INVALID, "score_topk_f32: null sims");
    RDV_REQUIRE(d >= 4 && d <= 8192 && (d & 3) == 0, RDV_E_INVALID,
                "score_topk_f32: d=%d must be a multiple of 4 in [4, 8192]", d);
    RDV_REQUIRE(k >= 1 && k <= 1024, RDV_E_LIMIT, "score_topk_f32: k=%d outside [1, 1024]", k);
    RDV_REQUIRE(tile_rows >= 8 && tile_rows <= 256 && (tile_rows & 7) == 0, RDV_E_INVALID,
                "score_topk_f32: tile_rows=%d must be a multiple of 8 in [8, 256]", tile_rows);
    RDV_REQUIRE(aligned16(d_q), RDV_E_ALIGN, "score_topk_f32: q not 16-byte aligned");

    ScoreParams p;
    p.doc_ptr = d_doc_ptr; p.row_off = d_row_off; p.tile_off = d_tile_off; p.q = d_q;
    p.B = B; p.d = d; p.k = k; p.tile_rows = tile_rows; p.total_tiles = total_tiles;
    p.cache_floats = max_rows < kMaxCacheFloats ? max_rows : kMaxCacheFloats;
    p.sims = d_sims; p.topk_idx = d_topk_idx; p.topk_val = d_topk_val; p.topk_cnt = d_topk_cnt;
    p.doc_done = d_doc_done;
    const size_t smem = (size_t)d * sizeof(float) + (size_t)p.cache_floats * sizeof(float);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    switch (d) {
        case 128:  return launch_score<1, 8>(p, smem, s);
        case 256:  return launch_score<2, 4>(p, smem, s);
        case 384:  return launch_score<3, 4>(p, smem, s);
        case 512:  return launch_score<4, 4>(p, smem, s);
        case 768:  return launch_score<6, 2>(p, smem, s);
        case 1024: return launch_score<8, 2>(p, smem, s);
        default:   return launch_score<0, 2>(p, smem, s);
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
    select_topk(p, b, p.sims + r0, n, reinterpret_cast<float*>(smem_dyn), s_red);
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
