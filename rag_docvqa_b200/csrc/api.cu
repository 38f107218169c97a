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

extern "C" int rdv_abi_version(void) { return 19; }

extern "C" const char* rdv_last_error(void) { return rdv::g_error; }

extern "C" int64_t rdv_struct_size(const char* name) {
    if (!name) return -1;
#define RDV_SIZE_OF(T) if (strcmp(name, #T) == 0) return (int64_t)sizeof(T)
    RDV_SIZE_OF(rdv_tile_desc); RDV_SIZE_OF(rdv_cta_desc); RDV_SIZE_OF(rdv_small_layout); RDV_SIZE_OF(rdv_chunk_rec);
    RDV_SIZE_OF(rdv_tok_rec); RDV_SIZE_OF(rdv_docstore); RDV_SIZE_OF(rdv_gather_args); RDV_SIZE_OF(rdv_pagestore);
    RDV_SIZE_OF(rdv_visual_args); RDV_SIZE_OF(rdv_p2s_img); RDV_SIZE_OF(rdv_p2s_args); RDV_SIZE_OF(rdv_vt5_embed_tables);
#undef RDV_SIZE_OF
    return -1;
}

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

// ---- small host batches: the whole Retriever.retrieve device round trip in ONE call ----------------------------
// (C1 = 1 page x 30 chunks x 384-d is 46 KB: every framework-level call on the way -- table building, two copies, a
// launch, a synchronise -- costs more than the work.  The layout / pack halves are pure host code, testable without a GPU.)
static inline int64_t round_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

extern "C" int rdv_small_batch_layout(const int64_t* rows, int32_t B, int32_t d, int32_t k, rdv_small_layout* lay) {
    using namespace rdv;
    RDV_REQUIRE(lay && (rows || B == 0), RDV_E_INVALID, "small_batch_layout: null pointer");
    RDV_REQUIRE(B >= 1 && d >= 4 && d <= 8192 && (d & 3) == 0, RDV_E_INVALID, "small_batch_layout: bad B=%d / d=%d", B, d);
    RDV_REQUIRE(k >= 1 && k <= 1024, RDV_E_LIMIT, "small_batch_layout: k=%d outside [1, 1024]", k);
    int64_t total = 0, mx = 0;
    for (int32_t b = 0; b < B; ++b) {
        RDV_REQUIRE(rows[b] >= 0 && rows[b] < (1ll << 31), RDV_E_LIMIT, "small_batch_layout: document %d has %lld rows", b,
                    (long long)rows[b]);
        total += rows[b];
        if (rows[b] > mx) mx = rows[b];
    }
    RDV_REQUIRE(total * (int64_t)d * 4 <= (1ll << 30), RDV_E_LIMIT, "small_batch_layout: %lld rows is not a small batch",
                (long long)total);
    int32_t algo = 0, tile_rows = 0, use_cluster = 0;
    int rc = rdv_score_plan(total, d, RDV_SCORE_AUTO, &algo, &tile_rows);
    if (rc) return rc;
    rc = rdv_retrieve_plan(total, (int32_t)mx, B, d, k, &use_cluster);
    if (rc) return rc;
    int32_t cluster = 0, slice_rows = 0;
    int64_t n_ctas = 0;
    if (use_cluster) {
        rc = rdv_cluster_plan(rows, B, d, k, 0, &cluster, &slice_rows, &n_ctas);
        if (rc) return rc;
        if (n_ctas == 0) use_cluster = 0;
    }
    if (use_cluster) algo = RDV_SMALL_CLUSTER;
    const int64_t T = rdv_count_tiles(rows, B, tile_rows);
    RDV_REQUIRE(T >= 0 && T < (1ll << 31), RDV_E_LIMIT, "small_batch_layout: too many tiles");
    memset(lay, 0, sizeof(*lay));
    lay->algo = algo; lay->tile_rows = tile_rows; lay->n_tiles = (int32_t)T; lay->max_rows = (int32_t)mx;
    lay->total_rows = total;
    lay->o_tiles = round_up(8 * ((int64_t)B + 1), 32);
    lay->cluster = use_cluster ? cluster : 0;
    lay->slice_rows = use_cluster ? slice_rows : 0;
    lay->n_ctas = use_cluster ? n_ctas : 0;
    lay->o_ctas = lay->o_tiles + 32 * (T > 0 ? T : 1);
    lay->o_q = lay->o_ctas + 32 * lay->n_ctas;
    lay->o_emb = lay->o_q + (int64_t)B * d * 4;
    lay->in_bytes = lay->o_emb + total * (int64_t)d * 4;
    lay->o_idx = total * 4;
    lay->o_cnt = lay->o_idx + (int64_t)B * k * 4;
    lay->read_bytes = lay->o_cnt + (int64_t)B * 4;
    lay->o_val = lay->read_bytes;
    lay->out_bytes = lay->o_val + (int64_t)B * k * 4;
    return RDV_OK;
}

extern "C" int rdv_small_batch_pack(const void* const* h_docs, const int64_t* rows, int32_t B, int32_t d, const float* h_q,
                                    const rdv_small_layout* lay, void* h_blob, const void* d_blob) {
    using namespace rdv;
    RDV_REQUIRE(h_docs && rows && h_q && lay && h_blob && d_blob, RDV_E_INVALID, "small_batch_pack: null pointer");
    RDV_REQUIRE(B >= 1 && B <= 4096, RDV_E_LIMIT, "small_batch_pack: B=%d outside [1, 4096]", B);
    RDV_REQUIRE(aligned16(h_blob) && aligned16(d_blob), RDV_E_ALIGN, "small_batch_pack: blobs must be 16-byte aligned");
    char* hb = static_cast<char*>(h_blob);
    const char* db = static_cast<const char*>(d_blob);
    const void* d_docs[4096];
    int64_t off = 0;
    const size_t row_bytes = (size_t)d * 4;
    for (int32_t b = 0; b < B; ++b) {
        const int64_t n = rows[b];
        RDV_REQUIRE(n >= 0 && (n == 0 || h_docs[b]), RDV_E_INVALID, "small_batch_pack: document %d is null", b);
        d_docs[b] = n ? db + lay->o_emb + (size_t)off * row_bytes : nullptr;
        if (n) memcpy(hb + lay->o_emb + (size_t)off * row_bytes, h_docs[b], (size_t)n * row_bytes);
        off += n;
    }
    RDV_REQUIRE(off == lay->total_rows, RDV_E_INVALID, "small_batch_pack: layout is for %lld rows, batch has %lld",
                (long long)lay->total_rows, (long long)off);
    memcpy(hb + lay->o_q, h_q, (size_t)B * row_bytes);
    if (lay->n_ctas > 0) {                                                       // the cluster kernel's view of the batch
        int rc = rdv_build_cluster_table(d_docs, rows, B, d, (int32_t)lay->cluster, (int32_t)lay->slice_rows, reinterpret_cast<rdv_cta_desc*>(hb + lay->o_ctas), lay->n_ctas);
        if (rc) return rc;
    }
    int32_t mx = 0;
    return rdv_build_doc_table(d_docs, rows, B, d, lay->tile_rows, reinterpret_cast<int64_t*>(hb),
                               reinterpret_cast<rdv_tile_desc*>(hb + lay->o_tiles), lay->n_tiles, &mx);
}

extern "C" int rdv_retrieve_small_f32(const void* const* h_docs, const int64_t* rows, int32_t B, int32_t d, int32_t k,
                                      const float* h_q, void* h_blob, void* d_blob, int64_t blob_bytes, void* d_out,
                                      int64_t d_out_bytes, void* h_out, int64_t h_out_bytes, rdv_small_layout* lay,
                                      void* stream) {
    using namespace rdv;
    int rc = rdv_small_batch_layout(rows, B, d, k, lay);
    if (rc) return rc;
    if (lay->in_bytes > blob_bytes || lay->out_bytes > d_out_bytes || lay->read_bytes > h_out_bytes)
        return RDV_SMALL_GROW;                                   // nothing touched: the caller grows its buffers and calls again
    RDV_REQUIRE(d_out && h_out, RDV_E_INVALID, "retrieve_small_f32: null pointer");
    rc = rdv_small_batch_pack(h_docs, rows, B, d, h_q, lay, h_blob, d_blob);
    if (rc) return rc;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    cudaError_t e = cudaMemcpyAsync(d_blob, h_blob, (size_t)lay->in_bytes, cudaMemcpyHostToDevice, s);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemcpyAsync (retrieve_small_f32, upload)");
    char* in = static_cast<char*>(d_blob);
    char* out = static_cast<char*>(d_out);
    if (lay->algo == RDV_SMALL_CLUSTER)
        rc = rdv_score_topk_cluster_f32(reinterpret_cast<const rdv_cta_desc*>(in + lay->o_ctas), lay->n_ctas, (int32_t)lay->cluster,
                                        reinterpret_cast<const float*>(in + lay->o_q), B, d, k, lay->max_rows,
                                        reinterpret_cast<float*>(out), reinterpret_cast<int32_t*>(out + lay->o_idx),
                                        reinterpret_cast<float*>(out + lay->o_val), reinterpret_cast<int32_t*>(out + lay->o_cnt),
                                        stream);
    else
        rc = rdv_score_topk_f32(reinterpret_cast<const rdv_tile_desc*>(in + lay->o_tiles), lay->n_tiles, lay->tile_rows, lay->algo,
                                reinterpret_cast<const int64_t*>(in), reinterpret_cast<const float*>(in + lay->o_q), B, d, k,
                                lay->max_rows, reinterpret_cast<float*>(out), reinterpret_cast<int32_t*>(out + lay->o_idx),
                                reinterpret_cast<float*>(out + lay->o_val), reinterpret_cast<int32_t*>(out + lay->o_cnt), stream);
    if (rc) return rc;
    e = cudaMemcpyAsync(h_out, d_out, (size_t)lay->read_bytes, cudaMemcpyDeviceToHost, s);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemcpyAsync (retrieve_small_f32, read-back)");
    e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) return cuda_fail(e, "cudaStreamSynchronize (retrieve_small_f32)");
    return RDV_OK;
}
