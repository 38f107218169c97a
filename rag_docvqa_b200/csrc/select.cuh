// Block-wide top-k selection shared by the score kernels and the gather kernel (sm_100a).
//
// Ordering everywhere: packed (order(score) << 32 | ~index) u64 keys, so "descending score, lowest index
// first, NaN greatest, -0 == +0" is one unsigned compare (rdv_common.cuh).
//
// k <= 32 (every reference config: chunk_num 5 / 10 / 20): register-resident, warp-level.  Thread t owns
// the strided slice {i : i % 256 == t} of the document's scores as keys in registers (n <= 10240: every
// reference workload) or as a private column of the shared-memory cache (larger n).  Each warp extracts the k best of its 32 slices
// with k rounds of a two-instruction warp arg-max (redux.sync on the high and low key halves) -- no
// block barrier inside the rounds -- then ONE barrier, and warp 0 merges the 8 x k candidates the same way.
// k > 32: k rounds of a block-wide arg-max (two barriers per round).
#pragma once
#include "rdv_common.cuh"

namespace rdv {

constexpr int kScoreThreads = 256;            // 8 compute warps
constexpr int kScoreWarps = kScoreThreads / 32;
constexpr int kMaxCacheFloats = 8192;         // selection pass caches up to this many scores in smem (32 KB).  A 64 KB cache was
                                              // measured: the selection alone 48 -> 37 us at C3, but the step 640 -> 730 us (the larger
                                              // shared-memory carve-out delays the streaming kernel that follows), so it stays at 32 KB
constexpr int kSelWarpK = 32;                 // largest k of the warp-level path

struct SelectArgs {
    int32_t k, cache_floats;
    int32_t* topk_idx;      // (B, k)
    float* topk_val;        // (B, k)
    int32_t* topk_cnt;      // (B)
    int32_t* doc_done;      // (B) workspace reset to 0 by the selecting block (may be null)
    int32_t* smem_idx;      // optional shared-memory copy of the winners (k_min entries) for a fused consumer
};

// shared-memory cache (floats) a selection launch needs: none when every document is selected out of registers
// (k <= kSelWarpK is the common case; larger k always uses the cache)
__host__ __device__ __forceinline__ int cache_floats_for(int max_rows, int k, int reg_rows) {
    if (k <= kSelWarpK && max_rows <= reg_rows) return 0;
    return max_rows < kMaxCacheFloats ? max_rows : kMaxCacheFloats;
}

struct BlockSync {       // whole block
    __device__ __forceinline__ void operator()() const { __syncthreads(); }
};

// warp-wide max of u64 keys: redux.sync.max.u32 on the high halves, then on the low halves of the lanes
// that hold the winning high half.  All-zero (no key) stays zero.
__device__ __forceinline__ unsigned long long warp_max_key(unsigned long long v) {
    const uint32_t hi = static_cast<uint32_t>(v >> 32), lo = static_cast<uint32_t>(v);
    const uint32_t mhi = __reduce_max_sync(0xffffffffu, hi);
    const uint32_t mlo = __reduce_max_sync(0xffffffffu, hi == mhi ? lo : 0u);
    return (static_cast<unsigned long long>(mhi) << 32) | mlo;
}

// k rounds over NREG register-resident keys per lane; lane r keeps the r-th winner.
template <int NREG>
__device__ __forceinline__ unsigned long long warp_rounds(unsigned long long (&key)[NREG], int rounds, int lane) {
    unsigned long long mine = 0;
    for (int r = 0; r < rounds; ++r) {
        unsigned long long best = key[0];
#pragma unroll
        for (int j = 1; j < NREG; ++j) best = key[j] > best ? key[j] : best;
        const unsigned long long win = warp_max_key(best);
#pragma unroll
        for (int j = 0; j < NREG; ++j) key[j] = key[j] == win ? 0ull : key[j];   // keys are unique: removes one
        if (lane == r) mine = win;
    }
    return mine;
}

template <int NREG>
__device__ __forceinline__ unsigned long long warp_topk_regs(const float* __restrict__ src, int n, int rounds, int tid, int lane) {
    unsigned long long key[NREG];
#pragma unroll
    for (int j = 0; j < NREG; ++j) {
        const int i = tid + j * kScoreThreads;
        key[j] = i < n ? pack_key(__ldcg(src + i), (uint32_t)i) : 0ull;   // every real key is > 0
    }
    return warp_rounds<NREG>(key, rounds, lane);
}

// Selection of the k best (score desc, index asc) among n scores of one document by 256 threads.
// Ends with the winners written (global + optional smem_idx); the caller synchronises before reading smem_idx.
// MAXREG: most keys a thread keeps in registers (16 or 40).  40 (n <= 10240, 80 key registers) is for the stand-alone
// selection kernel; kernels that do other work around the selection (gather, fused score) stay at 16 so that their
// register count -- and with it the occupancy of the whole kernel -- does not follow the selection's worst case.
template <int MAXREG, class Sync>
__device__ void select_topk(const SelectArgs& p, int b, const float* __restrict__ src, int n,
                            float* cache, unsigned long long* s_red, Sync sync) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int k_min = n < p.k ? n : p.k;
    if (p.k <= kSelWarpK) {
        __shared__ unsigned long long s_cand[kScoreWarps * kSelWarpK];
        unsigned long long mine;
        if (n <= 4 * kScoreThreads) {
            mine = warp_topk_regs<4>(src, n, k_min, tid, lane);
        } else if (n <= 16 * kScoreThreads) {
            mine = warp_topk_regs<16>(src, n, k_min, tid, lane);
        } else if (MAXREG > 16 && n <= MAXREG * kScoreThreads) {
            mine = warp_topk_regs<(MAXREG > 16 ? MAXREG : 16)>(src, n, k_min, tid, lane);
        } else {
            // private column of the cache (thread t only ever touches i % 256 == t: no barrier needed)
            const bool cached = n <= p.cache_floats;
            if (cached)
                for (int i = tid; i < n; i += kScoreThreads) cache[i] = __ldcg(src + i);
            unsigned long long prev = 0;
            mine = 0;
            for (int r = 0; r < k_min; ++r) {
                unsigned long long best = 0;
#pragma unroll 4
                for (int i = tid; i < n; i += kScoreThreads) {
                    const unsigned long long key = pack_key(cached ? cache[i] : __ldcg(src + i), (uint32_t)i);
                    if ((r == 0 || key < prev) && key > best) best = key;
                }
                prev = warp_max_key(best);
                if (lane == r) mine = prev;
                if (prev == 0) break;            // this warp's slices are exhausted
            }
        }
        if (lane < k_min) s_cand[warp * kSelWarpK + lane] = mine;
        sync();
        if (warp == 0) {
            // merge: candidate (w, r) of 8 warps x k_min rounds; lane l takes r = l, one key per warp
            unsigned long long key[kScoreWarps];
#pragma unroll
            for (int w = 0; w < kScoreWarps; ++w) key[w] = lane < k_min ? s_cand[w * kSelWarpK + lane] : 0ull;
            const unsigned long long win = warp_rounds<kScoreWarps>(key, k_min, lane);
            if (lane < k_min) {
                const uint32_t idx = key_index(win);
                p.topk_idx[(size_t)b * p.k + lane] = (int32_t)idx;
                if (p.smem_idx) p.smem_idx[lane] = (int32_t)idx;
                p.topk_val[(size_t)b * p.k + lane] = __ldcg(src + idx);
            }
        }
    } else {
        const bool cached = n <= p.cache_floats;
        if (cached) {
            for (int i = tid; i < n; i += kScoreThreads) cache[i] = __ldcg(src + i);
            sync();
        }
        unsigned long long prev = 0;
        for (int r = 0; r < k_min; ++r) {
            unsigned long long best = 0;   // every real key is > 0 (order_key(-inf) = 0x007FFFFF)
#pragma unroll 4
            for (int i = tid; i < n; i += kScoreThreads) {
                const unsigned long long key = pack_key(cached ? cache[i] : __ldcg(src + i), (uint32_t)i);
                if ((r == 0 || key < prev) && key > best) best = key;
            }
            best = warp_max_key(best);
            if (lane == 0) s_red[warp] = best;
            sync();
            unsigned long long win = s_red[0];
#pragma unroll
            for (int w = 1; w < kScoreWarps; ++w) win = s_red[w] > win ? s_red[w] : win;
            sync();
            if (tid == 0) {
                const uint32_t idx = key_index(win);
                p.topk_idx[(size_t)b * p.k + r] = (int32_t)idx;
                if (p.smem_idx) p.smem_idx[r] = (int32_t)idx;
                p.topk_val[(size_t)b * p.k + r] = cached ? cache[idx] : __ldcg(src + idx);
            }
            prev = win;
        }
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
