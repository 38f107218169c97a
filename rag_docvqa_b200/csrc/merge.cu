// Final merge of per-shard top-k candidates (corpus mode, BASELINE.json configs[4]) -- sm_100a.
//
// No reference counterpart: the reference is single-GPU.  Each rank scores its row shard and keeps a
// local top-k per query with GLOBAL row ids; after an all-gather every rank holds (Q, m = world*k)
// candidates.  This kernel selects the k best per query by (score desc, global id asc) -- the same
// ordering the single-GPU kernel uses -- so sharded == unsharded bit for bit.
// One warp per query; m is small (<= 4096), candidates are held in registers.
#include "rdv_common.cuh"

namespace rdv {



struct Cand {
    uint32_t key;     // order_key(score)
    long long idx;    // global id, < 0 = empty slot
};

__device__ __forceinline__ bool better(const Cand& a, const Cand& b) {   // a strictly better than b
    if (a.idx < 0) return false;
    if (b.idx < 0) return true;
    return a.key > b.key || (a.key == b.key && a.idx < b.idx);
}

// Candidates of query q: `parts` lists of k_in entries; list r starts at cand_val + r * part_stride_val + q * k_in (idx
// likewise).  parts == 1 is the plain (Q, m) matrix; parts == world is the all-gather receive buffer as NCCL fills it
// (rank-major), read in place -- no transpose copy between the collective and the merge.
__global__ void __launch_bounds__(128) topk_merge_kernel(const float* __restrict__ cand_val,
                                                         const int64_t* __restrict__ cand_idx, int Q, int parts, int k_in,
                                                         long long part_stride_val, long long part_stride_idx,
                                                         int k, float* __restrict__ out_val,
                                                         int64_t* __restrict__ out_idx) {
    const int lane = threadIdx.x & 31;
    const int q = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (q >= Q) return;
    const int m = parts * k_in;
    const float* cv = cand_val + (size_t)q * k_in;
    const int64_t* ci = cand_idx + (size_t)q * k_in;
    Cand prev;
    prev.key = 0xFFFFFFFFu; prev.idx = -1;    // "nothing selected yet"
    bool have_prev = false;
    for (int r = 0; r < k; ++r) {
        // best candidate strictly worse than prev (candidates are unique in (key, idx) unless ids repeat)
        Cand best; best.key = 0; best.idx = -1;
        float best_val = -INFINITY;
        for (int i = lane; i < m; i += 32) {
            const int part = i / k_in, j = i - part * k_in;
            Cand c; c.idx = ci[(size_t)part * part_stride_idx + j];
            const float v = cv[(size_t)part * part_stride_val + j];
            c.key = order_key(v);
            if (c.idx < 0) continue;
            if (have_prev && !better(prev, c)) continue;   // already emitted (or equal to prev)
            if (better(c, best)) { best = c; best_val = v; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            Cand other;
            other.key = __shfl_xor_sync(0xffffffffu, best.key, o);
            other.idx = __shfl_xor_sync(0xffffffffu, best.idx, o);
            const float ov = __shfl_xor_sync(0xffffffffu, best_val, o);
            if (better(other, best)) { best = other; best_val = ov; }
        }
        if (lane == 0) {
            out_idx[(size_t)q * k + r] = best.idx;
            out_val[(size_t)q * k + r] = best.idx >= 0 ? best_val : -INFINITY;
        }
        if (best.idx < 0) {           // ran out of candidates: pad the rest
            for (int rr = r + 1 + lane; rr < k; rr += 32) {
                out_idx[(size_t)q * k + rr] = -1;
                out_val[(size_t)q * k + rr] = -INFINITY;
            }
            break;
        }
        prev = best;
        have_prev = true;
    }
}

}  // namespace rdv

extern "C" int rdv_topk_merge(const float* d_cand_val, const int64_t* d_cand_idx, int32_t Q, int32_t m, int32_t k,
                              float* d_out_val, int64_t* d_out_idx, void* stream) {
    using namespace rdv;
    RDV_REQUIRE(Q >= 0 && m >= 0 && k >= 1, RDV_E_INVALID, "topk_merge: bad sizes Q=%d m=%d k=%d", Q, m, k);
    if (Q == 0) return RDV_OK;
    RDV_REQUIRE(d_cand_val && d_cand_idx && d_out_val && d_out_idx, RDV_E_INVALID, "topk_merge: null pointer");
    const int blocks = (Q + 3) / 4;
    topk_merge_kernel<<<blocks, 128, 0, static_cast<cudaStream_t>(stream)>>>(d_cand_val, d_cand_idx, Q, 1, m, 0, 0, k,
                                                                            d_out_val, d_out_idx);
    RDV_LAUNCH_CHECK("topk_merge_kernel");
    return RDV_OK;
}

extern "C" int rdv_topk_merge_parts(const float* d_cand_val, const int64_t* d_cand_idx, int32_t Q, int32_t parts, int32_t k_in,
                                    int64_t part_stride_val, int64_t part_stride_idx, int32_t k, float* d_out_val,
                                    int64_t* d_out_idx, void* stream) {
    using namespace rdv;
    RDV_REQUIRE(Q >= 0 && parts >= 1 && k_in >= 0 && k >= 1 && part_stride_val >= 0 && part_stride_idx >= 0, RDV_E_INVALID,
                "topk_merge_parts: bad sizes Q=%d parts=%d k_in=%d k=%d", Q, parts, k_in, k);
    RDV_REQUIRE((int64_t)parts * k_in < (1ll << 31), RDV_E_LIMIT, "topk_merge_parts: too many candidates per query");
    if (Q == 0) return RDV_OK;
    RDV_REQUIRE(d_cand_val && d_cand_idx && d_out_val && d_out_idx, RDV_E_INVALID, "topk_merge_parts: null pointer");
    const int blocks = (Q + 3) / 4;
    topk_merge_kernel<<<blocks, 128, 0, static_cast<cudaStream_t>(stream)>>>(d_cand_val, d_cand_idx, Q, parts, k_in,
                                                                            part_stride_val, part_stride_idx, k, d_out_val,
                                                                            d_out_idx);
    RDV_LAUNCH_CHECK("topk_merge_kernel");
    return RDV_OK;
}
