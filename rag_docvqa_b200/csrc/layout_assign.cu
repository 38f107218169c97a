// Chunker: which OCR words lie inside which layout box -- sm_100a.
//
//   containment_ratio(word box, layout box) > 0.5     src/utils.py:328-341, src/_modules.py:1023-1033
//
// The reference walks words x layout boxes of every page in Python (O(words * boxes) float arithmetic per page, the
// dominant cost of Chunker.get_chunks with a layout model).  Here a warp owns 32 consecutive words of one page:
// each lane keeps its word box in registers, the page's layout boxes (already in the reference's order: sorted by
// (xmin, ymin), src/_modules.py:1006-1018) are broadcast loads, and one ballot per (word group, layout box) yields 32
// membership bits.  float64 with individually rounded operations in the reference's order -- the decision `> 0.5`
// is bit-exact against Python floats.  Per word the label of the LAST containing box is kept (:1030 overwrites).
#include "rdv_common.cuh"

namespace rdv {

constexpr int kLayWarps = 8;

__global__ void __launch_bounds__(kLayWarps * 32) layout_assign_kernel(
    const double* __restrict__ word_box, const int32_t* __restrict__ page_word_off,
    const double* __restrict__ lay_box, const int32_t* __restrict__ lay_label, const int32_t* __restrict__ page_lay_off,
    const int32_t* __restrict__ group_page, const int32_t* __restrict__ page_group_off, int n_groups,
    int default_label, const int64_t* __restrict__ bits_off, uint32_t* __restrict__ bits,
    int32_t* __restrict__ word_label) {
    const int lane = threadIdx.x & 31;
    const int grp = blockIdx.x * kLayWarps + (threadIdx.x >> 5);
    if (grp >= n_groups) return;
    const int p = group_page[grp];
    const int sub = grp - page_group_off[p];                  // 32-word group inside the page
    const int w0 = page_word_off[p], w1 = page_word_off[p + 1];
    const int w = w0 + sub * 32 + lane;
    const bool live = w < w1;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    if (live) {
        const double2* wb = reinterpret_cast<const double2*>(word_box + (size_t)w * 4);
        const double2 a = wb[0], b = wb[1];
        s0 = a.x; s1 = a.y; s2 = b.x; s3 = b.y;
    }
    const double small_area = __dmul_rn(__dsub_rn(s2, s0), __dsub_rn(s3, s1));
    int label = default_label;
    const int g0 = page_lay_off[p], g1 = page_lay_off[p + 1];
    for (int g = g0; g < g1; ++g) {
        const double2* lb = reinterpret_cast<const double2*>(lay_box + (size_t)g * 4);
        const double2 a = __ldg(lb), b = __ldg(lb + 1);
        const double x1 = s0 > a.x ? s0 : a.x, y1 = s1 > a.y ? s1 : a.y;       // max(small, large)
        const double x2 = s2 < b.x ? s2 : b.x, y2 = s3 < b.y ? s3 : b.y;       // min(small, large)
        const double dw = __dsub_rn(x2, x1), dh = __dsub_rn(y2, y1);
        const double iw = dw > 0.0 ? dw : 0.0, ih = dh > 0.0 ? dh : 0.0;        // max(0, .)
        const double inter = __dmul_rn(iw, ih);
        const double ratio = small_area > 0.0 ? __ddiv_rn(inter, small_area) : 0.0;
        const bool inside = live && ratio > 0.5;
        const uint32_t m = __ballot_sync(0xffffffffu, inside);
        if (lane == 0) bits[bits_off[g] + sub] = m;
        if (inside) label = lay_label[g];
    }
    if (live) word_label[w] = label;
}

}  // namespace rdv

extern "C" int rdv_layout_assign(const double* d_word_box, const int32_t* d_page_word_off, const double* d_lay_box,
                                 const int32_t* d_lay_label, const int32_t* d_page_lay_off, const int32_t* d_group_page,
                                 const int32_t* d_page_group_off, int32_t n_groups, int32_t default_label,
                                 const int64_t* d_bits_off, uint32_t* d_bits, int32_t* d_word_label, void* stream) {
    using namespace rdv;
    RDV_REQUIRE(n_groups >= 0, RDV_E_INVALID, "layout_assign: negative n_groups");
    if (n_groups == 0) return RDV_OK;
    RDV_REQUIRE(d_word_box && d_page_word_off && d_lay_box && d_lay_label && d_page_lay_off && d_group_page &&
                d_page_group_off && d_bits_off && d_bits && d_word_label, RDV_E_INVALID, "layout_assign: null pointer");
    RDV_REQUIRE(aligned16(d_word_box) && aligned16(d_lay_box), RDV_E_ALIGN, "layout_assign: boxes must be 16-byte aligned");
    const int blocks = (n_groups + kLayWarps - 1) / kLayWarps;
    layout_assign_kernel<<<blocks, kLayWarps * 32, 0, static_cast<cudaStream_t>(stream)>>>(
        d_word_box, d_page_word_off, d_lay_box, d_lay_label, d_page_lay_off, d_group_page, d_page_group_off, n_groups,
        default_label, d_bits_off, d_bits, d_word_label);
    RDV_LAUNCH_CHECK("layout_assign_kernel");
    return RDV_OK;
}
