// Block-wide top-k selection shared by the score kernels and the gather kernel (sm_100a).
#pragma once
#include "rdv_common.cuh"

namespace rdv {

constexpr int kScoreThreads = 256;            // 8 compute warps
constexpr int kScoreWarps = kScoreThreads / 32;
constexpr int kMaxCacheFloats = 8192;         // selection pass caches up to this many scores in smem (32 KB)

struct SelectArgs {
    int32_t k, cache_floats;
    int32_t* topk_idx;      // (B, k)
    float* topk_val;        // (B, k)
    int32_t* topk_cnt;      // (B)
    int32_t* doc_done;      // (B) workspace reset to 0 by the selecting block (may be null)
    int32_t* smem_idx;      // optional shared-memory copy of the winners (k_min entries) for a fused consumer
};

struct BlockSync {       // whole block
    __device__ __forceinline__ void operator()() const { __syncthreads(); }
};
struct ConsumerSync {    // the 256 consumer threads of the TMA kernel (named barrier 1)
    __device__ __forceinline__ void operator()() const { asm volatile("bar.sync 1, 256;" ::: "memory"); }
};

// Selection of the k best (score desc, index asc) among n scores of one document by 256 threads.
template <class Sync>
__device__ void select_topk(const SelectArgs& p, int b, const float* __restrict__ src, int n,
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
            if (p.smem_idx) p.smem_idx[r] = (int32_t)idx;
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
        if (p.doc_done) p.doc_done[b] = 0;   // leave the workspace zeroed for the next call
    }
}

}  // namespace rdv
