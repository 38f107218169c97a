// librdv: error plumbing and device queries shared by every entry point.
#include "rdv_common.cuh"

#include <stdlib.h>
#include <string.h>

namespace rdv {

static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t err, const char* what) {
    set_error("%s: %s (%s)", what, cudaGetErrorString(err), cudaGetErrorName(err));
    return RDV_E_CUDA;
}

int carveout_pct() {
    static const int pct = [] {
        const char* v = getenv("RDV_CARVEOUT");
        return v ? atoi(v) : -1;
    }();
    return pct;
}

int pdl_mask() {
    static const int mask = [] {
        const char* v = getenv("RDV_PDL");
        return v ? atoi(v) : kPdlStream;
    }();
    return mask;
}

int sm_count() {
    static thread_local int cached_dev = -1, cached_sms = 148;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != cached_dev) {
        int sms = 0;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && sms > 0) {
            cached_sms = sms;
            cached_dev = dev;
        }
    }
    return cached_sms;
}

}  // namespace rdv

extern "C" int rdv_abi_version(void) { return 14; }

extern "C" const char* rdv_last_error(void) { return rdv::g_error; }

extern "C" int rdv_device_info(int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return rdv::cuda_fail(e, "cudaGetDevice");
    int sms = 0, major = 0, minor = 0;
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
    if (e != cudaSuccess) return rdv::cuda_fail(e, "cudaDeviceGetAttribute");
    if (sm_count) *sm_count = sms;
    if (cc_major) *cc_major = major;
    if (cc_minor) *cc_minor = minor;
    return RDV_OK;
}

// ---- host-side staging helpers (no kernels): the per-batch bookkeeping of the Python layer in C ----------------
extern "C" int64_t rdv_count_tiles(const int64_t* rows, int32_t B, int32_t tile_rows) {
    if (!rows || B < 0 || tile_rows < 1) return -1;
    int64_t t = 0;
    for (int32_t b = 0; b < B; ++b) {
        if (rows[b] < 0) return -1;
        t += (rows[b] + tile_rows - 1) / tile_rows;
    }
    return t;
}

extern "C" int rdv_build_doc_table(const void* const* d_docs, const int64_t* rows, int32_t B, int32_t d, int32_t tile_rows,
                                   int64_t* h_row_off, rdv_tile_desc* h_tiles, int64_t n_tiles, int32_t* max_rows) {
    using namespace rdv;
    RDV_REQUIRE(B >= 0 && d >= 4 && (d & 3) == 0 && tile_rows >= 1, RDV_E_INVALID, "build_doc_table: bad B / d / tile_rows");
    RDV_REQUIRE((d_docs && rows && h_row_off) || B == 0, RDV_E_INVALID, "build_doc_table: null pointer");
    RDV_REQUIRE(h_tiles || n_tiles == 0, RDV_E_INVALID, "build_doc_table: null tiles");
    int64_t off = 0, t = 0;
    int32_t mx = 0;
    if (h_row_off) h_row_off[0] = 0;
    for (int32_t b = 0; b < B; ++b) {
        const int64_t n = rows[b];
        RDV_REQUIRE(n >= 0 && n < (1ll << 31), RDV_E_LIMIT, "build_doc_table: document %d has %lld rows", b, (long long)n);
        RDV_REQUIRE(n == 0 || (d_docs[b] && aligned16(d_docs[b])), RDV_E_ALIGN, "build_doc_table: document %d is null / not 16-byte aligned", b);
        const char* base = static_cast<const char*>(d_docs[b]);
        for (int64_t r = 0; r < n; r += tile_rows) {
            RDV_REQUIRE(t < n_tiles, RDV_E_INVALID, "build_doc_table: more tiles than the %lld provided", (long long)n_tiles);
            rdv_tile_desc& td = h_tiles[t++];
            td.src = base + (size_t)r * (size_t)d * 4;
            td.sims_off = off + r;
            td.rows = (int32_t)(n - r < tile_rows ? n - r : tile_rows);
            td.doc = b;
            td.doc_rows = (int32_t)n;
            td.reserved = 0;
        }
        off += n;
        h_row_off[b + 1] = off;
        if (n > mx) mx = (int32_t)n;
    }
    RDV_REQUIRE(t == n_tiles, RDV_E_INVALID, "build_doc_table: %lld tiles built, %lld expected", (long long)t, (long long)n_tiles);
    if (max_rows) *max_rows = mx;
    return RDV_OK;
}

extern "C" int rdv_upload_docs_f32(const void* const* h_docs, const int64_t* rows, int32_t B, int32_t d, float* d_packed,
                                   void* stream) {
    using namespace rdv;
    RDV_REQUIRE(B >= 0 && d >= 1, RDV_E_INVALID, "upload_docs_f32: bad B / d");
    RDV_REQUIRE((h_docs && rows) || B == 0, RDV_E_INVALID, "upload_docs_f32: null pointer");
    int64_t off = 0;
    for (int32_t b = 0; b < B; ++b) {
        const int64_t n = rows[b];
        RDV_REQUIRE(n >= 0, RDV_E_INVALID, "upload_docs_f32: negative size");
        if (n == 0) continue;
        RDV_REQUIRE(h_docs[b] && d_packed, RDV_E_INVALID, "upload_docs_f32: null buffer");
        cudaError_t e = cudaMemcpyAsync(d_packed + (size_t)off * d, h_docs[b], (size_t)n * d * sizeof(float),
                                        cudaMemcpyHostToDevice, static_cast<cudaStream_t>(stream));
        if (e != cudaSuccess) return cuda_fail(e, "cudaMemcpyAsync (upload_docs_f32)");
        off += n;
    }
    return RDV_OK;
}
