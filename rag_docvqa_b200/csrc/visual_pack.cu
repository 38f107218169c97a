// Retrieved patches -> the generator's visual input, on the device (sm_100a).  SURVEY.md section 8f rank 1.
//
// Replaces, for page images resident in HBM, what the reference does on the host after retrieval when
// page_retrieval == "concat" (every shipped config):
//   page.crop(rect) per hit                     src/_modules.py:2102-2121   (a9; 63 % of the reference's retrieve)
//   concatenate_patches(patches, mode="grid")   src/utils.py:180-231        strip packing into one RGB canvas
//   feature extractor resize to S x S           src/_modules.py:133         PIL.Image.resize (Pillow Resample.c, 8 bpc)
//   rescale 1/255, (x - mean) / std             (HF image processor)        fp32 CHW pixel_values
// The canvas is never materialised: the horizontal resampling pass reads the page pixels through the patch
// layout, one virtual canvas row at a time staged in shared memory; its uint8 output (rows x S x 3) is the only
// intermediate, exactly as in Pillow, whose two passes round to uint8 in between.  Integer / fixed-point work:
// bit-exact against Pillow (the oracle restates Resample.c and is pinned to the installed Pillow).
#include "rdv_common.cuh"

namespace rdv {
namespace vp {

constexpr int kThreads = 256;
constexpr int kMaxK = 64;
constexpr int kPrecisionBits = 32 - 8 - 2;
constexpr int kMaxRowBytes = 24 * 1024;       // staged canvas row: up to 8192 pixels wide

struct Params {
    rdv_pagestore ps;
    rdv_visual_args a;
};

// ---- Pillow's filters and coefficient tables, evaluated with individually rounded double operations so the
// ---- device reproduces the host compiler's (non-contracted) arithmetic -----------------------------------
__device__ __forceinline__ double pil_filter(int kind, double x) {
    if (x < 0.0) x = -x;
    if (kind == 2) return x < 1.0 ? __dsub_rn(1.0, x) : 0.0;                        // bilinear
    if (x < 1.0)                                                                      // bicubic, a = -0.5
        return __dadd_rn(__dmul_rn(__dmul_rn(__dsub_rn(__dmul_rn(1.5, x), 2.5), x), x), 1.0);
    if (x < 2.0)
        return __dmul_rn(__dsub_rn(__dmul_rn(__dadd_rn(__dmul_rn(__dsub_rn(x, 5.0), x), 8.0), x), 4.0), -0.5);
    return 0.0;
}

struct Axis { double scale, support, ss; int ksize; };
__device__ __forceinline__ Axis axis_of(int in_size, int out_size, int kind) {
    Axis ax;
    ax.scale = __ddiv_rn((double)(float)in_size, (double)out_size);
    const double filterscale = ax.scale < 1.0 ? 1.0 : ax.scale;
    ax.support = __dmul_rn(kind == 2 ? 1.0 : 2.0, filterscale);
    ax.ss = __ddiv_rn(1.0, filterscale);
    ax.ksize = (int)ceil(ax.support) * 2 + 1;
    return ax;
}

// one output position: bounds (first, count) and fixed-point weights w[0 .. count)
__device__ void pil_coeffs_at(const Axis& ax, int in_size, int kind, int xx, int* first, int* count, int* w, int cap) {
    const double center = __dadd_rn(0.0, __dmul_rn((double)xx + 0.5, ax.scale));
    int xmin = (int)__dadd_rn(__dsub_rn(center, ax.support), 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)__dadd_rn(__dadd_rn(center, ax.support), 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    if (xmax > cap) xmax = cap;                                   // capacity is checked by the caller (status)
    double ww = 0.0;
    for (int x = 0; x < xmax; ++x)
        ww = __dadd_rn(ww, pil_filter(kind, __dmul_rn(__dadd_rn(__dsub_rn((double)(x + xmin), center), 0.5), ax.ss)));
    for (int x = 0; x < xmax; ++x) {
        double v = pil_filter(kind, __dmul_rn(__dadd_rn(__dsub_rn((double)(x + xmin), center), 0.5), ax.ss));
        if (ww != 0.0) v = __ddiv_rn(v, ww);
        const double scaled = __dmul_rn(v, (double)(1 << kPrecisionBits));
        w[x] = v < 0.0 ? (int)__dadd_rn(-0.5, scaled) : (int)__dadd_rn(0.5, scaled);
    }
    *first = xmin;
    *count = xmax;
}

// per-document header in the layout workspace
enum { kHdrGridW = 0, kHdrGridH, kHdrRowFirst, kHdrRows, kHdrKsizeH, kHdrKsizeV, kHdrN, kHdrPad, kHdrInts };

__device__ __forceinline__ int* doc_header(const rdv_visual_args& a, int b) { return a.layout + (size_t)b * (kHdrInts + 4 * a.k); }
__device__ __forceinline__ int* doc_patches(const rdv_visual_args& a, int b) { return doc_header(a, b) + kHdrInts; }

// ---- kernel 1: strip packing + coefficient tables (one block per document) -----------------------------
__global__ void __launch_bounds__(kThreads) visual_layout_kernel(const Params P) {
    const rdv_visual_args& a = P.a;
    const int b = blockIdx.x, tid = threadIdx.x;
    __shared__ int s_hdr[kHdrInts];
    int* hdr = doc_header(a, b);
    int* pat = doc_patches(a, b);
    const int S = a.out_size;
    if (tid == 0) {
        const int n = min(a.hit_cnt[b], a.k);
        long long area = 0;
        int gw = 0;
        for (int i = 0; i < n; ++i) {
            const int32_t* r = a.hit_rect + ((size_t)b * a.k + i) * 4;
            const int w = r[2] - r[0], h = r[3] - r[1];
            area += (long long)w * h;                                         // src/utils.py:183
            gw = max(gw, w);                                                   // :185
        }
        int status = 0, gh;
        if (n == 0) { gw = 5; gh = 5; }                                        // blank image, :193-195
        else if (gw <= 0) { gh = 0; status = 2; }                              // the reference divides by zero here
        else gh = (int)((double)area / (double)gw);                            // int(total_area / grid_width), :186
        if (gh <= 0 && status == 0) status = 2;                                // empty canvas: PIL cannot resize it
        int x_off = 0, y_off = 0, row_h = 0;
        for (int i = 0; i < n; ++i) {                                          // :223-230
            const int32_t* r = a.hit_rect + ((size_t)b * a.k + i) * 4;
            const int w = r[2] - r[0], h = r[3] - r[1];
            if (x_off + w > gw) { x_off = 0; y_off += row_h; row_h = 0; }
            pat[4 * i] = x_off; pat[4 * i + 1] = y_off; pat[4 * i + 2] = w; pat[4 * i + 3] = h;
            x_off += w;
            row_h = max(row_h, h);
        }
        s_hdr[kHdrGridW] = gw; s_hdr[kHdrGridH] = gh; s_hdr[kHdrN] = n;
        if (status == 0) {
            const Axis ah = axis_of(gw, S, a.filter), av = axis_of(gh, S, a.filter);
            s_hdr[kHdrKsizeH] = ah.ksize; s_hdr[kHdrKsizeV] = av.ksize;
            if (ah.ksize > a.ksize_cap_h || av.ksize > a.ksize_cap_v || gw * 3 > kMaxRowBytes) status = 1;
        }
        s_hdr[kHdrPad] = status;
    }
    __syncthreads();
    const int gw = s_hdr[kHdrGridW], gh = s_hdr[kHdrGridH];
    if (s_hdr[kHdrPad] == 0) {
        const Axis ah = axis_of(gw, S, a.filter), av = axis_of(gh, S, a.filter);
        int* ch = a.coeff_h + (size_t)b * S * (a.ksize_cap_h + 2);
        int* cv = a.coeff_v + (size_t)b * S * (a.ksize_cap_v + 2);
        for (int xx = tid; xx < S; xx += kThreads) {
            int* row = ch + (size_t)xx * (a.ksize_cap_h + 2);
            pil_coeffs_at(ah, gw, a.filter, xx, row, row + 1, row + 2, a.ksize_cap_h);
            row = cv + (size_t)xx * (a.ksize_cap_v + 2);
            pil_coeffs_at(av, gh, a.filter, xx, row, row + 1, row + 2, a.ksize_cap_v);
        }
        __syncthreads();
        if (tid == 0) {
            // rows of the canvas the vertical pass reads (ImagingResample: ybox_first .. ybox_last)
            const int first = cv[0];
            const int* last_row = cv + (size_t)(S - 1) * (a.ksize_cap_v + 2);
            const int rows = last_row[0] + last_row[1] - first;
            s_hdr[kHdrRowFirst] = first; s_hdr[kHdrRows] = rows;
            if (rows > a.rows_cap) s_hdr[kHdrPad] = 1;
        }
        __syncthreads();
    }
    if (tid < kHdrInts) hdr[tid] = s_hdr[tid];
    if (tid == 0) a.status[b] = s_hdr[kHdrPad];
}

// ---- kernel 2: horizontal pass, canvas rows read through the layout ---------------------------------------
// One block = one document x kGroups groups of R consecutive canvas rows (R <= 8, as many as fit the row budget).
// The R virtual rows of a group are staged in shared memory (black, then the row segments of every patch that covers
// them: warp w stages row w), the document's coefficient table sits beside them for the whole block, and a thread owns
// one (output x, channel) for all R rows, so a weight is fetched once per tap and applied R times.
// The kernel is bound by instruction issue (ncu: 81 % issue-active, 1.06 G warp instructions for a C2 batch of which
// the taps are a fifth), so what pays is fewer instructions around the taps.  Measured and rejected: rows staged as
// RGBX words with one 32-bit shared load per tap (1 LDS + 3 extractions + 3 IMAD for 3 channel-taps instead of
// 3 LDS.U8 + 3 IMAD: as many instructions, lower occupancy; 1.83 ms per C2 batch against 1.61).
constexpr int kRowBudget = 64 * 1024;          // most bytes of staged rows
constexpr int kCoeffBudget = 32 * 1024;        // most bytes of the coefficient table kept in shared memory (else L1/L2)
constexpr int kMaxRows = 8;
constexpr int kGroups = 8;                     // row groups per block: the coefficient table is loaded once for all of them

__host__ __device__ __forceinline__ int row_pitch(int gw) { return (gw * 3 + 15) & ~15; }
// shared memory the launch reserves for rows: 8 rows of the widest possible canvas, within the budget -- sized by
// the batch, not by the worst case, so several blocks share an SM (C2: 20 KB rows + 17 KB table -> 6 blocks)
__host__ __device__ __forceinline__ int row_bytes_for(int max_w) {
    const long long want = (long long)kMaxRows * row_pitch(max_w > 0 ? max_w : 1);
    const int one = row_pitch(max_w > 0 ? max_w : 1);
    return (int)(want <= kRowBudget ? want : (one > kRowBudget ? one : kRowBudget));
}
__host__ __device__ __forceinline__ int rows_per_block(int gw, int row_bytes) {
    int r = row_bytes / row_pitch(gw);
    return r < 1 ? 1 : (r > kMaxRows ? kMaxRows : r);
}

// n bytes global -> shared by ONE WARP, any alignment on either side: destination words are assembled from two aligned
// source words with a byte permute (one 32-bit store per 4 bytes instead of four byte loads and stores).  The source
// buffer's size is a multiple of 4, so the aligned word holding the last byte is readable.
__device__ __forceinline__ void copy_row_bytes(unsigned char* dst, const unsigned char* __restrict__ src, int n, int lane) {
    const int head = min(n, (int)((4u - (unsigned)(uintptr_t)dst) & 3u));
    if (lane < head) dst[lane] = __ldg(src + lane);
    const int words = (n - head) >> 2;
    uint32_t* dw = reinterpret_cast<uint32_t*>(dst + head);
    const unsigned char* s0 = src + head;
    const unsigned sh = (unsigned)(uintptr_t)s0 & 3u;
    const uint32_t* sw = reinterpret_cast<const uint32_t*>(s0 - sh);
    const unsigned sel = 0x3210u + 0x1111u * sh;
    // four iterations' loads are issued before the first store: a warp stages its row alone, so without the unrolling
    // every iteration waited a full global-memory round trip (22 % of the kernel's stall samples)
    if (sh == 0) {
#pragma unroll 4
        for (int w = lane; w < words; w += 32) dw[w] = __ldg(sw + w);
    } else {
#pragma unroll 4
        for (int w = lane; w < words; w += 32) dw[w] = __byte_perm(__ldg(sw + w), __ldg(sw + w + 1), sel);
    }
    const int done = head + (words << 2);
    if (lane < n - done) dst[done + lane] = __ldg(src + done + lane);
}

// one (output x, channel) for the rows of a group.  ALL = the group has all kMaxRows rows (every group of a document but
// its last): the tap loop is straight-line -- per tap one weight and, per row, one LDS.U8 + one IMAD.  With a per-row
// `r < nr` test inside the loop the compiler emitted a branch and a recomputed shared-memory base per (tap, row): 8
// instructions instead of 2, 80 % of the kernel's 1.06 G warp instructions (ncu, profiles/r1e_ncu_summary.md).
template <bool ALL>
__device__ __forceinline__ void h_column(const unsigned char* __restrict__ px, int pitch, const int* __restrict__ wts, int cnt,
                                         int nr, unsigned char* __restrict__ out, int out_pitch) {
    int acc[kMaxRows];
#pragma unroll
    for (int r = 0; r < kMaxRows; ++r) acc[r] = 1 << (kPrecisionBits - 1);
    for (int t = 0; t < cnt; ++t) {
        const int wgt = wts[t];
#pragma unroll
        for (int r = 0; r < kMaxRows; ++r)
            if (ALL || r < nr) acc[r] += (int)px[r * pitch + t * 3] * wgt;
    }
#pragma unroll
    for (int r = 0; r < kMaxRows; ++r)
        if (ALL || r < nr) out[(size_t)r * out_pitch] = (unsigned char)min(max(acc[r] >> kPrecisionBits, 0), 255);
}

__global__ void __launch_bounds__(kThreads) visual_resize_h_kernel(const Params P) {
    const rdv_pagestore& ps = P.ps;
    const rdv_visual_args& a = P.a;
    const int b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int* hdr = doc_header(a, b);
    if (hdr[kHdrPad] != 0) return;
    const int rows = hdr[kHdrRows];
    const int gw = hdr[kHdrGridW], gh = hdr[kHdrGridH], n = hdr[kHdrN], row_first = hdr[kHdrRowFirst];
    const int row_bytes = row_bytes_for(a.max_page_w);
    const int R = rows_per_block(gw, row_bytes);
    if (blockIdx.x * (R * kGroups) >= rows) return;
    extern __shared__ __align__(16) unsigned char s_dyn[];
    __shared__ int s_pat[kMaxK * 4];
    __shared__ int s_page[kMaxK];
    __shared__ int s_src[kMaxK * 2];
    const int S = a.out_size, cap = a.ksize_cap_h + 2, ksz = hdr[kHdrKsizeH] + 2;
    const int pitch = row_pitch(gw);
    unsigned char* s_rows = s_dyn;                                             // R x pitch
    int* s_k = reinterpret_cast<int*>(s_dyn + row_bytes);                      // S x ksz (first, count, weights), if it fits
    const bool k_in_smem = (size_t)S * ksz * 4 <= (size_t)min(kCoeffBudget, S * cap * 4);
    const int* pat = doc_patches(a, b);
    for (int i = tid; i < n; i += kThreads) {
        s_pat[4 * i] = pat[4 * i]; s_pat[4 * i + 1] = pat[4 * i + 1]; s_pat[4 * i + 2] = pat[4 * i + 2]; s_pat[4 * i + 3] = pat[4 * i + 3];
        s_page[i] = ps.doc_page_off[b] + a.hit_page[(size_t)b * a.k + i];
        s_src[2 * i] = a.hit_rect[((size_t)b * a.k + i) * 4]; s_src[2 * i + 1] = a.hit_rect[((size_t)b * a.k + i) * 4 + 1];
    }
    // the document's coefficient table: loaded ONCE per block and used for kGroups row groups (it is 17 KB at C2 --
    // reloading it for every 8 rows was a seventh of the kernel's instructions); a warp copies whole table rows
    const int* ch = a.coeff_h + (size_t)b * S * cap;
    if (k_in_smem)
    {
        if (ksz <= 32) {                     // the usual case (bicubic down to a ninth): one load per lane and table row
#pragma unroll 4
            for (int xx = warp; xx < S; xx += kThreads / 32)
                if (lane < ksz) s_k[xx * ksz + lane] = __ldg(ch + (size_t)xx * cap + lane);
        } else {
            for (int xx = warp; xx < S; xx += kThreads / 32)
                for (int j = lane; j < ksz; j += 32) s_k[xx * ksz + j] = __ldg(ch + (size_t)xx * cap + j);
        }
    }
    for (int g = 0; g < kGroups; ++g) {
        const int r_begin = (blockIdx.x * kGroups + g) * R;
        if (r_begin >= rows) break;
        const int nr = min(R, rows - r_begin);
        __syncthreads();                                    // the previous group's taps are done (g = 0: s_pat is visible)
        // Warp w owns row w of the group (R <= 8 rows, 8 warps): it blackens the row, then pastes the row segment of
        // every patch that covers it (pastes clip at the canvas border, crops are black outside their page) -- no block
        // barrier between the two, no per-row call overhead, no divisions.
        if (warp < nr) {
            uint4* zrow = reinterpret_cast<uint4*>(s_rows + (size_t)warp * pitch);
            for (int i = lane; i < pitch / 16; i += 32) zrow[i] = make_uint4(0u, 0u, 0u, 0u);
            __syncwarp();
            const int y = row_first + r_begin + warp;
            for (int i = 0; i < n; ++i) {
                const int dx = s_pat[4 * i], dy = s_pat[4 * i + 1], w = s_pat[4 * i + 2], h = s_pat[4 * i + 3];
                if (y < dy || y >= min(dy + h, gh)) continue;
                const int pg = s_page[i];
                const int W = ps.page_wh[2 * pg], H = ps.page_wh[2 * pg + 1];
                const int sx0 = s_src[2 * i], sy0 = s_src[2 * i + 1];
                const int cw = min(w, gw - dx);
                const int x_lo = max(0, -sx0), x_hi = min(cw, W - sx0);         // patch columns that exist in the page
                const int sy = sy0 + (y - dy);
                if (cw <= 0 || x_lo >= x_hi || sy < 0 || sy >= H) continue;
                const unsigned char* src = ps.pixels + ps.page_off[pg] + ((size_t)sy * W + sx0 + x_lo) * 3;
                copy_row_bytes(s_rows + (size_t)warp * pitch + (dx + x_lo) * 3, src, (x_hi - x_lo) * 3, lane);
                __syncwarp();                                                    // a later patch may overwrite these bytes
            }
        }
        __syncthreads();
        unsigned char* temp = a.temp + ((size_t)b * a.rows_cap + r_begin) * S * 3;
        for (int o = tid; o < S * 3; o += kThreads) {
            const int xx = o / 3, c = o - xx * 3;
            const int* k = k_in_smem ? s_k + xx * ksz : ch + (size_t)xx * cap;
            const unsigned char* px = s_rows + k[0] * 3 + c;
            if (nr == kMaxRows) h_column<true>(px, pitch, k + 2, k[1], nr, temp + o, S * 3);
            else h_column<false>(px, pitch, k + 2, k[1], nr, temp + o, S * 3);
        }
    }
}

// ---- kernel 3: vertical pass + rescale / normalise ----------------------------------------------------------
__global__ void __launch_bounds__(kThreads) visual_resize_v_kernel(const Params P) {
    const rdv_visual_args& a = P.a;
    const int b = blockIdx.y, yy = blockIdx.x, tid = threadIdx.x;
    const int* hdr = doc_header(a, b);
    const int S = a.out_size;
    unsigned char* out = a.out_u8 + ((size_t)b * S + yy) * S * 3;
    const bool bad = hdr[kHdrPad] != 0;
    const int cap = a.ksize_cap_v + 2;
    const int* k = a.coeff_v + ((size_t)b * S + yy) * cap;
    const int first = bad ? 0 : __ldg(k) - hdr[kHdrRowFirst], cnt = bad ? 0 : __ldg(k + 1);
    const unsigned char* temp = a.temp + (size_t)b * a.rows_cap * S * 3;
    if (!bad && (S * 3) % 4 == 0 && ((uintptr_t)temp & 3u) == 0 && ((uintptr_t)out & 3u) == 0) {
        // the usual case (S = 224): a thread owns FOUR CONSECUTIVE bytes of the output row, so a tap is one aligned 32-bit
        // load of the intermediate row (coalesced across the warp), three byte extractions and four IMADs, with no
        // predicate inside the loop (the strided version below spends 16 issue slots per tap on 2.6 live outputs)
        const int stride = S * 3 / 4;
        const uint32_t* rows32 = reinterpret_cast<const uint32_t*>(temp + (size_t)first * S * 3);
        for (int q = tid; q < stride; q += kThreads) {
            int a0 = 1 << (kPrecisionBits - 1), a1 = a0, a2 = a0, a3 = a0;
            const uint32_t* col = rows32 + q;
#pragma unroll 4
            for (int t = 0; t < cnt; ++t) {
                const int wgt = __ldg(k + 2 + t);
                const uint32_t p = __ldg(col + (size_t)t * stride);
                a0 += (int)(p & 0xFFu) * wgt;
                a1 += (int)__byte_perm(p, 0u, 0x4441u) * wgt;
                a2 += (int)__byte_perm(p, 0u, 0x4442u) * wgt;
                a3 += (int)(p >> 24) * wgt;
            }
            const int v[4] = {min(max(a0 >> kPrecisionBits, 0), 255), min(max(a1 >> kPrecisionBits, 0), 255),
                              min(max(a2 >> kPrecisionBits, 0), 255), min(max(a3 >> kPrecisionBits, 0), 255)};
            reinterpret_cast<uint32_t*>(out)[q] = (uint32_t)v[0] | ((uint32_t)v[1] << 8) | ((uint32_t)v[2] << 16) | ((uint32_t)v[3] << 24);
            if (a.out_px) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int o = 4 * q + j, xx = o / 3, c = o - xx * 3;
                    const float f = __fdiv_rn(__fsub_rn(__fmul_rn((float)v[j], 1.0f / 255.0f), a.mean[c]), a.std[c]);
                    a.out_px[(((size_t)b * 3 + c) * S + yy) * S + xx] = f;
                }
            }
        }
        return;
    }
    // a thread owns up to kVOut outputs of the row (o, o + 256, ...), so each weight is fetched once per tap for all of them
    constexpr int kVOut = 4;
    for (int o0 = tid; o0 < S * 3; o0 += kThreads * kVOut) {
        int acc[kVOut];
#pragma unroll
        for (int u = 0; u < kVOut; ++u) acc[u] = 1 << (kPrecisionBits - 1);
        if (!bad) {
            const unsigned char* col = temp + (size_t)first * S * 3 + o0;
            for (int t = 0; t < cnt; ++t) {
                const int wgt = __ldg(k + 2 + t);
#pragma unroll
                for (int u = 0; u < kVOut; ++u)
                    if (o0 + u * kThreads < S * 3) acc[u] += (int)__ldg(col + (size_t)t * S * 3 + u * kThreads) * wgt;
            }
        }
#pragma unroll
        for (int u = 0; u < kVOut; ++u) {
            const int o = o0 + u * kThreads;
            if (o >= S * 3) continue;
            const int v = bad ? 0 : min(max(acc[u] >> kPrecisionBits, 0), 255);
            out[o] = (unsigned char)v;
            if (a.out_px) {
                const int xx = o / 3, c = o - xx * 3;
                const float f = __fdiv_rn(__fsub_rn(__fmul_rn((float)v, 1.0f / 255.0f), a.mean[c]), a.std[c]);
                a.out_px[(((size_t)b * 3 + c) * S + yy) * S + xx] = f;
            }
        }
    }
}

}  // namespace vp
}  // namespace rdv

extern "C" int rdv_visual_pack(const rdv_pagestore* ps, const rdv_visual_args* args, void* stream) {
    using namespace rdv;
    RDV_REQUIRE(ps && args, RDV_E_INVALID, "visual_pack: null struct");
    RDV_REQUIRE(ps->B >= 0, RDV_E_INVALID, "visual_pack: negative B");
    if (ps->B == 0) return RDV_OK;
    RDV_REQUIRE(args->k >= 1 && args->k <= vp::kMaxK, RDV_E_LIMIT, "visual_pack: k=%d outside [1, %d]", args->k, vp::kMaxK);
    RDV_REQUIRE(args->out_size >= 1 && args->out_size <= 4096, RDV_E_LIMIT, "visual_pack: out_size=%d outside [1, 4096]", args->out_size);
    RDV_REQUIRE(args->filter == 2 || args->filter == 3, RDV_E_INVALID, "visual_pack: filter must be 2 (bilinear) or 3 (bicubic)");
    RDV_REQUIRE(ps->doc_page_off && ps->page_wh && ps->page_off && ps->pixels, RDV_E_INVALID, "visual_pack: page store has a null array");
    RDV_REQUIRE(args->hit_page && args->hit_rect && args->hit_cnt && args->layout && args->coeff_h && args->coeff_v && args->temp &&
                args->out_u8 && args->status, RDV_E_INVALID, "visual_pack: args has a null array");
    RDV_REQUIRE(args->ksize_cap_h >= 3 && args->ksize_cap_v >= 3 && args->rows_cap >= 1, RDV_E_INVALID, "visual_pack: bad capacities");
    vp::Params P;
    P.ps = *ps;
    P.a = *args;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    RDV_ONCE_PER_DEVICE(cudaFuncSetAttribute(vp::visual_resize_h_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             vp::kRowBudget + vp::kCoeffBudget), "cudaFuncSetAttribute(visual_resize_h)");
    vp::visual_layout_kernel<<<ps->B, vp::kThreads, 0, s>>>(P);
    RDV_LAUNCH_CHECK("visual_layout_kernel");
    // the grid covers the row capacity at the smallest rows-per-block any document can have; blocks past a document's
    // own row count exit at once
    RDV_REQUIRE(args->max_page_w >= 1 && vp::row_pitch(args->max_page_w) <= vp::kMaxRowBytes, RDV_E_LIMIT,
                "visual_pack: max_page_w=%d outside [1, %d]", args->max_page_w, vp::kMaxRowBytes / 3);
    const int row_bytes = vp::row_bytes_for(args->max_page_w);
    const int r_min = vp::rows_per_block(args->max_page_w, row_bytes);            // a canvas is at most a page wide
    const int row_blocks = (args->rows_cap + r_min * vp::kGroups - 1) / (r_min * vp::kGroups);
    const long long coeff_bytes = (long long)args->out_size * (args->ksize_cap_h + 2) * 4;
    const size_t smem_h = (size_t)row_bytes + (size_t)(coeff_bytes < vp::kCoeffBudget ? coeff_bytes : vp::kCoeffBudget);
    vp::visual_resize_h_kernel<<<dim3(row_blocks, ps->B), vp::kThreads, smem_h, s>>>(P);
    RDV_LAUNCH_CHECK("visual_resize_h_kernel");
    vp::visual_resize_v_kernel<<<dim3(args->out_size, ps->B), vp::kThreads, 0, s>>>(P);
    RDV_LAUNCH_CHECK("visual_resize_v_kernel");
    return RDV_OK;
}
