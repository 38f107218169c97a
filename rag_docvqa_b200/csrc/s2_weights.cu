// S2Chunker: the pairwise weight matrices of a page's layout regions (second half) -- sm_100a.
//
//   spatial[i][j]  = 1 / (1 + || centroid_i - centroid_j ||)          src/_modules.py:1755-1773
//   semantic[i][j] = cosine_similarity(embeddings)[i][j]              src/_modules.py:1775-1788 (sklearn, float32)
//   combined       = (spatial + semantic) / 2                         src/_modules.py:1790-1802
//                    (cluster_mode "spatial": semantic = spatial, so combined == spatial exactly)
//
// The reference fills the spatial matrix with a Python double loop per page (two np.array + np.linalg.norm per
// entry, ~5 us each: 4.5 ms for a page of 30 regions).  Here ONE launch covers every page of a batch: a warp owns one
// (page, i, j) entry at a time.  The spatial part is float64 with individually rounded operations in the reference's
// order, INCLUDING the one fused multiply-add numpy's ddot (OpenBLAS on an FMA machine) uses for the squared length
// of the 2-vector: sqrt(fma(dy, dy, dx * dx)) -- bit-exact against the reference as run in the build container.  The
// semantic part mirrors sklearn: rows scaled by 1 / ||row|| (zero rows left alone) in float32, then the dot product;
// only the summation order differs from the BLAS sgemm (tolerance 1e-6 in the tests).
#include "rdv_common.cuh"

namespace rdv {

constexpr int kS2Warps = 8;

__device__ __forceinline__ int page_of_entry(const int64_t* __restrict__ out_off, int P, long long e) {
    int lo = 0, hi = P;                       // last p with out_off[p] <= e
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (out_off[mid] <= e) lo = mid; else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(kS2Warps * 32) s2_weights_kernel(
    const double* __restrict__ node_box, const int32_t* __restrict__ page_node_off, int P,
    const float* __restrict__ emb, int d, int what, const int64_t* __restrict__ out_off, double* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const long long total = out_off[P];
    const long long stride = (long long)gridDim.x * kS2Warps;
    for (long long e = (long long)blockIdx.x * kS2Warps + (threadIdx.x >> 5); e < total; e += stride) {
        const int p = page_of_entry(out_off, P, e);
        const int n0 = page_node_off[p], n = page_node_off[p + 1] - n0;
        const int local = (int)(e - out_off[p]);
        const int i = local / n, j = local - i * n;
        const double2* bi = reinterpret_cast<const double2*>(node_box + (size_t)(n0 + i) * 4);
        const double2* bj = reinterpret_cast<const double2*>(node_box + (size_t)(n0 + j) * 4);
        const double2 i0 = __ldg(bi), i1 = __ldg(bi + 1), j0 = __ldg(bj), j1 = __ldg(bj + 1);
        const double cix = __ddiv_rn(__dadd_rn(i0.x, i1.x), 2.0), ciy = __ddiv_rn(__dadd_rn(i0.y, i1.y), 2.0);
        const double cjx = __ddiv_rn(__dadd_rn(j0.x, j1.x), 2.0), cjy = __ddiv_rn(__dadd_rn(j0.y, j1.y), 2.0);
        const double dx = __dsub_rn(cix, cjx), dy = __dsub_rn(ciy, cjy);
        const double dist = __dsqrt_rn(__fma_rn(dy, dy, __dmul_rn(dx, dx)));
        const double spatial = __ddiv_rn(1.0, __dadd_rn(1.0, dist));
        double semantic = spatial;
        if (emb != nullptr && what != RDV_S2_SPATIAL) {
            const float* __restrict__ xi = emb + (size_t)(n0 + i) * d;
            const float* __restrict__ xj = emb + (size_t)(n0 + j) * d;
            float si = 0.f, sj = 0.f;
            for (int t = lane; t < d; t += 32) {
                const float a = __ldg(xi + t), b = __ldg(xj + t);
                si = fmaf(a, a, si); sj = fmaf(b, b, sj);
            }
            si = warp_sum(si); sj = warp_sum(sj);
            float ni = sqrtf(si), nj = sqrtf(sj);
            if (ni == 0.f) ni = 1.f;                                   // sklearn normalize: zero rows are left alone
            if (nj == 0.f) nj = 1.f;
            float dot = 0.f;
            for (int t = lane; t < d; t += 32) dot = fmaf(__fdiv_rn(__ldg(xi + t), ni), __fdiv_rn(__ldg(xj + t), nj), dot);
            semantic = (double)warp_sum(dot);
        }
        if (lane == 0)
            out[e] = what == RDV_S2_SPATIAL ? spatial : what == RDV_S2_SEMANTIC ? semantic : __ddiv_rn(__dadd_rn(spatial, semantic), 2.0);
    }
}

}  // namespace rdv

extern "C" int rdv_s2_weights(const double* d_node_box, const int32_t* d_page_node_off, int32_t P, const float* d_emb,
                              int32_t d, int32_t what, const int64_t* d_out_off, int64_t total_entries, double* d_out, void* stream) {
    using namespace rdv;
    RDV_REQUIRE(P >= 0 && total_entries >= 0, RDV_E_INVALID, "s2_weights: negative size");
    if (P == 0 || total_entries == 0) return RDV_OK;
    RDV_REQUIRE(d_node_box && d_page_node_off && d_out_off && d_out, RDV_E_INVALID, "s2_weights: null pointer");
    RDV_REQUIRE(what == RDV_S2_COMBINED || what == RDV_S2_SPATIAL || what == RDV_S2_SEMANTIC, RDV_E_INVALID,
                "s2_weights: what=%d is not RDV_S2_COMBINED / SPATIAL / SEMANTIC", what);
    RDV_REQUIRE(d_emb != nullptr || what != RDV_S2_SEMANTIC, RDV_E_INVALID, "s2_weights: semantic weights need embeddings");
    RDV_REQUIRE(d_emb == nullptr || (d >= 1 && d <= 65536), RDV_E_INVALID, "s2_weights: d=%d outside [1, 65536]", d);
    RDV_REQUIRE(aligned16(d_node_box), RDV_E_ALIGN, "s2_weights: boxes must be 16-byte aligned");
    long long blocks = (total_entries + kS2Warps - 1) / kS2Warps;
    const long long cap = (long long)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    s2_weights_kernel<<<(int)blocks, kS2Warps * 32, 0, static_cast<cudaStream_t>(stream)>>>(
        d_node_box, d_page_node_off, P, d_emb, d, what, d_out_off, d_out);
    RDV_LAUNCH_CHECK("s2_weights_kernel");
    return RDV_OK;
}
