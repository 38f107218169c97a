// Maximum of consecutive fixed-length groups of a score vector -- sm_100a.
//
// Pooled-patch visual retrieval (BASELINE.json configs[3] as north_star words it): every patch vector of a page strip is
// scored against the pooled question (Retriever._get_similarities' formula, src/_modules.py:1990-1993); a strip's score is
// the best of its patches (torch.max over the strip: NaN propagates), and the strips are then ranked by torch.topk
// (src/_modules.py:2408).  One block per group; the scores were just written by the score kernel (L2-resident).
#include "rdv_common.cuh"

namespace rdv {

__global__ void __launch_bounds__(256) group_max_kernel(const float* __restrict__ scores, int group_len,
                                                        float* __restrict__ out) {
    const float* g = scores + (size_t)blockIdx.x * group_len;
    float best = -INFINITY;
    bool nan = false;
    for (int i = threadIdx.x; i < group_len; i += 256) {
        const float v = __ldcg(g + i);
        nan |= (v != v);
        best = fmaxf(best, v);                       // fmaxf drops NaN: tracked separately
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        best = fmaxf(best, __shfl_xor_sync(0xffffffffu, best, o));
        nan |= (bool)__shfl_xor_sync(0xffffffffu, (int)nan, o);
    }
    __shared__ float s_best[8];
    __shared__ int s_nan[8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { s_best[warp] = best; s_nan[warp] = nan; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) { best = fmaxf(best, s_best[w]); nan |= (bool)s_nan[w]; }
        out[blockIdx.x] = nan ? __int_as_float(0x7fc00000) : best;
    }
}

}  // namespace rdv

extern "C" int rdv_group_max_f32(const float* d_scores, int64_t n_groups, int32_t group_len, float* d_out, void* stream) {
    using namespace rdv;
    RDV_REQUIRE(n_groups >= 0 && n_groups < (1ll << 31) && group_len >= 1, RDV_E_INVALID, "group_max_f32: bad sizes");
    if (n_groups == 0) return RDV_OK;
    RDV_REQUIRE(d_scores && d_out, RDV_E_INVALID, "group_max_f32: null pointer");
    group_max_kernel<<<(unsigned)n_groups, 256, 0, static_cast<cudaStream_t>(stream)>>>(d_scores, group_len, d_out);
    RDV_LAUNCH_CHECK("group_max_kernel");
    return RDV_OK;
}
