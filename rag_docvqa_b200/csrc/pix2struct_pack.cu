// Retrieved image crops -> Pix2Struct flattened patches, on the device (sm_100a).  SURVEY.md section 8 row a12,
// Pix2Struct half.
//
// Replaces, for page images resident in HBM, what the reference's image processor does on the host with the crops
// VisualRetriever returns (src/RAGPix2Struct.py:221 -> src/custom_pix2struct_processor.py):
//   normalize                                :175-196   (x - mean) / max(std, 1/sqrt(#elements)), whole-image statistics
//   extract_flattened_patches_single         :33-95     rows x cols patch grid from the patch budget, bilinear
//                                                        anti-aliased resize (torch F.interpolate, ATen's separable
//                                                        two-pass kernel: horizontal, then vertical), 16 x 16 patches
//                                                        flattened pixel-major / channel-minor behind (row id, col id)
//   extract_multi_image_flattened_patches    :97-132    equal budget per image, row ids continue across images,
//                                                        zero padding to max_total_patches
//   attention mask                           :225       row sum != 0
// The per-image plan (rows, cols, offsets: a few float64 operations per image) is computed by the caller, which knows
// the crop rectangles (the visual decode is host-side integer work, src/_modules.py:2386-2450); the kernels do the
// pixel work.  fp32 throughout, as the reference; tolerance against torch's CPU kernel is stated in the tests.
#include "rdv_common.cuh"

namespace rdv {
namespace p2s {

constexpr int kThreads = 256;
constexpr int kMaxTaps = 64;          // 2 * ceil(support) + 1 of the anti-aliasing filter (support = down-scale factor)

struct Params {
    rdv_pagestore ps;
    rdv_p2s_args a;
};

// ---- kernel 1: exact byte sums of every crop (kStatSlices blocks per image) -----------------------------------
// The statistics are integers (sum of bytes, sum of squared bytes), so the reduction order does not matter: rows are
// dealt to warps across kStatSlices blocks, each row is read as aligned 32-bit words (dp4a sums four bytes and their
// squares in one instruction each), and every block adds its part to the image's two u64 accumulators.
constexpr int kStatSlices = 16;

struct StatAcc { unsigned long long s, ss; };     // per image, zeroed by the entry point

__global__ void __launch_bounds__(kThreads) p2s_stats_kernel(const Params P) {
    const rdv_p2s_img& im = P.a.images[blockIdx.y];
    const int tid = threadIdx.x, lane = tid & 31;
    const unsigned char* page = P.ps.pixels + P.ps.page_off[im.page];
    const int W = P.ps.page_wh[2 * im.page], H = P.ps.page_wh[2 * im.page + 1];
    const int xa = max(im.x0, 0), xb = min(im.x1, W);
    const int ya = max(im.y0, 0), yb = min(im.y1, H);
    unsigned long long s = 0, ss = 0;
    for (int y = ya + blockIdx.x * (kThreads / 32) + (tid >> 5); y < yb; y += kStatSlices * (kThreads / 32)) {
        const unsigned char* p0 = page + ((size_t)y * W + xa) * 3;
        const unsigned char* p1 = page + ((size_t)y * W + xb) * 3;
        if (p0 >= p1) continue;
        const unsigned char* q0 = reinterpret_cast<const unsigned char*>((reinterpret_cast<uintptr_t>(p0) + 3) & ~uintptr_t(3));
        const unsigned char* q1 = reinterpret_cast<const unsigned char*>(reinterpret_cast<uintptr_t>(p1) & ~uintptr_t(3));
        unsigned rs = 0, rss = 0;                 // one row: < 2^32 (8192 pixels * 3 * 255^2 = 1.6e9)
        if (q0 >= q1) {                           // too short for an aligned word
            for (const unsigned char* p = p0 + lane; p < p1; p += 32) { const unsigned v = __ldg(p); rs += v; rss += v * v; }
        } else {
            if (p0 + lane < q0) { const unsigned v = __ldg(p0 + lane); rs += v; rss += v * v; }        // head: < 4 bytes
            if (q1 + lane < p1) { const unsigned v = __ldg(q1 + lane); rs += v; rss += v * v; }        // tail: < 4 bytes
            const unsigned* w0 = reinterpret_cast<const unsigned*>(q0);
            const int nw = (int)((q1 - q0) >> 2);
            for (int i = lane; i < nw; i += 32) {
                const unsigned v = __ldg(w0 + i);
                rs = __dp4a(v, 0x01010101u, rs);
                rss = __dp4a(v, v, rss);
            }
        }
        s += rs; ss += rss;
    }
    __shared__ unsigned long long sh_s[kThreads / 32], sh_ss[kThreads / 32];
    for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); ss += __shfl_xor_sync(0xffffffffu, ss, o); }
    if (lane == 0) { sh_s[tid >> 5] = s; sh_ss[tid >> 5] = ss; }
    __syncthreads();
    if (tid == 0) {
        for (int i = 1; i < kThreads / 32; ++i) { s += sh_s[i]; ss += sh_ss[i]; }
        StatAcc* acc = reinterpret_cast<StatAcc*>(P.a.stats) + blockIdx.y;
        if (s) atomicAdd(&acc->s, s);
        if (ss) atomicAdd(&acc->ss, ss);
    }
}

// mean and adjusted std of image `img` from its byte sums (:175-196), and the 256 possible normalised pixel values
// ((v - mean) / adj in fp32, the reference's own two operations) as a shared-memory table: the resize reads bytes.
__device__ __forceinline__ void normalised_lut(const Params& P, int img, float* lut /* [256] shared */) {
    const rdv_p2s_img& im = P.a.images[img];
    const StatAcc acc = reinterpret_cast<const StatAcc*>(P.a.stats)[img];
    const double n = (double)(im.x1 - im.x0) * (double)(im.y1 - im.y0) * 3.0;
    const double mean = (double)acc.s / n;
    double var = (double)acc.ss / n - mean * mean;
    if (var < 0.0) var = 0.0;
    const float std32 = (float)sqrt(var);
    const double floor_ = 1.0 / sqrt(n);                           // max(std, 1.0 / math.sqrt(np.prod(shape)))  (:188)
    const float adj = P.a.do_normalize ? (((double)std32 > floor_) ? std32 : (float)floor_) : 1.f;
    const float mean32 = P.a.do_normalize ? (float)mean : 0.f;
    for (int v = threadIdx.x; v < 256; v += blockDim.x) lut[v] = __fdiv_rn(__fsub_rn((float)v, mean32), adj);
    __syncthreads();
}

// ATen's anti-aliased linear weights for output position i (upsample_bilinear2d_aa, align_corners = False), fp32
struct Taps { int first, count; float w[kMaxTaps]; };
__device__ __forceinline__ void aa_taps(int i, int in_size, int out_size, Taps& t) {
    const float scale = (float)in_size / (float)out_size;
    const float support = scale >= 1.0f ? scale : 1.0f;            // interp_size * 0.5 * scale, interp_size = 2
    const float invscale = scale >= 1.0f ? 1.0f / scale : 1.0f;
    const float center = scale * ((float)i + 0.5f);
    int xmin = (int)(center - support + 0.5f);
    if (xmin < 0) xmin = 0;
    int xsize = (int)(center + support + 0.5f);
    if (xsize > in_size) xsize = in_size;
    xsize -= xmin;
    if (xsize > kMaxTaps) xsize = kMaxTaps;
    float total = 0.f;
    for (int j = 0; j < xsize; ++j) {
        float x = ((float)(j + xmin) - center + 0.5f) * invscale;
        x = fabsf(x);
        const float wgt = x < 1.0f ? 1.0f - x : 0.0f;
        t.w[j] = wgt;
        total += wgt;
    }
    if (total != 0.f) for (int j = 0; j < xsize; ++j) t.w[j] /= total;
    t.first = xmin;
    t.count = xsize;
}

// ---- kernel 2: horizontal pass (normalised crop rows -> temp rows of the resized width) ---------------------
// grid (resized-width tiles, row chunks, images): a thread owns one output column for kRowsH consecutive rows, so its
// taps are computed once and the grid fills the machine (8 documents x 5 crops: 3 x 28 x 40 blocks, was 3 x 40).
constexpr int kRowsH = 8;

__global__ void __launch_bounds__(kThreads) p2s_resize_h_kernel(const Params P) {
    __shared__ float lut[256];
    const rdv_p2s_img& im = P.a.images[blockIdx.z];
    const int rw = im.cols * P.a.patch, w = im.x1 - im.x0, h = im.y1 - im.y0;
    const int ybeg = blockIdx.y * kRowsH;
    if (ybeg >= h) return;                                          // whole block: no barrier is skipped by a part of it
    normalised_lut(P, blockIdx.z, lut);
    const int X = blockIdx.x * kThreads + threadIdx.x;
    if (X >= rw) return;
    const unsigned char* page = P.ps.pixels + P.ps.page_off[im.page];
    const int W = P.ps.page_wh[2 * im.page], H = P.ps.page_wh[2 * im.page + 1];
    Taps t;
    aa_taps(X, w, rw, t);
    float* temp = P.a.temp + im.temp_off;
    const int yend = min(h, ybeg + kRowsH);
    const int sx0 = im.x0 + t.first;
    for (int y = ybeg; y < yend; ++y) {
        const int sy = im.y0 + y;
        const bool row_in = sy >= 0 && sy < H;
        const unsigned char* row = page + (size_t)(row_in ? sy : 0) * W * 3;
        float acc[3] = {0.f, 0.f, 0.f};
        for (int j = 0; j < t.count; ++j) {
            const int sx = sx0 + j;
            const bool in = row_in && sx >= 0 && sx < W;           // black outside the page (PIL crop), then normalised (:220)
            const unsigned char* px = row + (size_t)(in ? sx : 0) * 3;
            const float wj = t.w[j];
            acc[0] = fmaf(lut[in ? __ldg(px) : 0], wj, acc[0]);
            acc[1] = fmaf(lut[in ? __ldg(px + 1) : 0], wj, acc[1]);
            acc[2] = fmaf(lut[in ? __ldg(px + 2) : 0], wj, acc[2]);
        }
        float* o = temp + ((size_t)y * rw + X) * 3;
        o[0] = acc[0]; o[1] = acc[1]; o[2] = acc[2];
    }
}

// ---- kernel 3: vertical pass, written straight into the flattened-patch layout ------------------------------
__global__ void __launch_bounds__(kThreads) p2s_resize_v_kernel(const Params P) {
    const rdv_p2s_img& im = P.a.images[blockIdx.y];
    const int ps = P.a.patch, rw = im.cols * ps, rh = im.rows * ps, h = im.y1 - im.y0;
    const int idx = blockIdx.x * kThreads + threadIdx.x;            // (Y, X) of the resized image
    if (idx >= rw * rh) return;
    const int Y = idx / rw, X = idx - Y * rw;
    const int r = Y / ps, py = Y - r * ps, cc = X / ps, px = X - cc * ps;
    const int pidx = r * im.cols + cc;
    if (pidx >= im.kept) return;                                    // result[:max_patches]  (:95)
    Taps t;
    aa_taps(Y, h, rh, t);
    const float* temp = P.a.temp + im.temp_off;
    float acc[3] = {0.f, 0.f, 0.f};
    for (int j = 0; j < t.count; ++j) {
        const float* s = temp + ((size_t)(t.first + j) * rw + X) * 3;
        acc[0] = fmaf(s[0], t.w[j], acc[0]); acc[1] = fmaf(s[1], t.w[j], acc[1]); acc[2] = fmaf(s[2], t.w[j], acc[2]);
    }
    const int depth = 2 + ps * ps * 3;
    float* row = P.a.out + ((size_t)im.doc * P.a.max_total + im.out_start + pidx) * depth;
    float* f = row + 2 + (py * ps + px) * 3;
    f[0] = acc[0]; f[1] = acc[1]; f[2] = acc[2];
    if (py == 0 && px == 0) { row[0] = (float)(r + 1 + im.row_offset); row[1] = (float)(cc + 1); }   // :81-85
}

// ---- kernel 4: zero padding + attention mask (one warp per output row) --------------------------------------
__global__ void __launch_bounds__(kThreads) p2s_finish_kernel(const Params P, int n_docs) {
    const int warp = (blockIdx.x * kThreads + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= n_docs * P.a.max_total) return;
    const int b = warp / P.a.max_total, rowi = warp - b * P.a.max_total;
    const int depth = 2 + P.a.patch * P.a.patch * 3;
    float* row = P.a.out + (size_t)warp * depth;
    float sum = 0.f;
    if (rowi >= P.a.doc_total[b]) {
        for (int i = lane; i < depth; i += 32) row[i] = 0.f;        // padding (:125-129)
    } else {
        for (int i = lane; i < depth; i += 32) sum += row[i];
        sum = warp_sum(sum);
    }
    if (lane == 0) P.a.mask[warp] = sum != 0.f ? 1.f : 0.f;         // (flattened.sum(-1) != 0)  (:225)
}

}  // namespace p2s
}  // namespace rdv

extern "C" int rdv_pix2struct_patches(const rdv_pagestore* ps, const rdv_p2s_args* args, void* stream) {
    using namespace rdv;
    RDV_REQUIRE(ps && args, RDV_E_INVALID, "pix2struct_patches: null struct");
    RDV_REQUIRE(args->n_docs >= 0 && args->n_images >= 0, RDV_E_INVALID, "pix2struct_patches: negative size");
    if (args->n_docs == 0) return RDV_OK;
    RDV_REQUIRE(args->patch >= 1 && args->patch <= 64 && args->max_total >= 1, RDV_E_INVALID, "pix2struct_patches: bad patch / max_total");
    RDV_REQUIRE(ps->page_wh && ps->page_off && ps->pixels, RDV_E_INVALID, "pix2struct_patches: page store has a null array");
    RDV_REQUIRE(args->out && args->mask && args->doc_total && (args->n_images == 0 || (args->images && args->stats && args->temp)),
                RDV_E_INVALID, "pix2struct_patches: args has a null array");
    RDV_REQUIRE(args->max_rw >= 0 && args->max_rwh >= 0 && args->max_h >= 0 && args->max_h <= 65535 * p2s::kRowsH &&
                args->n_images <= 65535, RDV_E_INVALID, "pix2struct_patches: bad launch bounds");
    p2s::Params P;
    P.ps = *ps;
    P.a = *args;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (args->n_images > 0) {
        cudaError_t me = cudaMemsetAsync(args->stats, 0, (size_t)args->n_images * sizeof(p2s::StatAcc), s);
        if (me != cudaSuccess) return cuda_fail(me, "cudaMemsetAsync(p2s stats)");
        p2s::p2s_stats_kernel<<<dim3(p2s::kStatSlices, args->n_images), p2s::kThreads, 0, s>>>(P);
        RDV_LAUNCH_CHECK("p2s_stats_kernel");
        p2s::p2s_resize_h_kernel<<<dim3((args->max_rw + p2s::kThreads - 1) / p2s::kThreads,
                                        (args->max_h + p2s::kRowsH - 1) / p2s::kRowsH, args->n_images), p2s::kThreads, 0, s>>>(P);
        RDV_LAUNCH_CHECK("p2s_resize_h_kernel");
        p2s::p2s_resize_v_kernel<<<dim3((args->max_rwh + p2s::kThreads - 1) / p2s::kThreads, args->n_images), p2s::kThreads, 0, s>>>(P);
        RDV_LAUNCH_CHECK("p2s_resize_v_kernel");
    }
    const long long warps = (long long)args->n_docs * args->max_total;
    p2s::p2s_finish_kernel<<<(unsigned)((warps * 32 + p2s::kThreads - 1) / p2s::kThreads), p2s::kThreads, 0, s>>>(P, args->n_docs);
    RDV_LAUNCH_CHECK("p2s_finish_kernel");
    return RDV_OK;
}
