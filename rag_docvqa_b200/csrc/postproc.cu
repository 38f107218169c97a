// What consumes the top-k before generation -- sm_100a.
//
//   Reranker.rerank post-processing   src/_modules.py:1579-1595   argsort of the cross-encoder scores (descending),
//                                                                 threshold filter, max / min clamp -> the index list
//                                                                 applied to the candidates and all their companions
//   majorpage / weightmajorpage        src/RAGVT5.py:455-475       per-document weighted vote over the hits' pages
//
// Both are tiny segmented reductions over <= 64 hits per document: one warp per document, everything in registers
// and a few hundred bytes of shared memory.  Integer / IEEE arithmetic in the reference's own order, so the
// results are bit-exact against the oracle (the cross-encoder itself is a model and out of scope: its scores
// are the input).
#include "rdv_common.cuh"

namespace rdv {

constexpr int kPostMaxK = 64;
constexpr int kPostWarps = 4;

// total order of np.argsort: NaN greatest, -0 == +0
__device__ __forceinline__ unsigned long long order_key64(double v) {
    unsigned long long u = (unsigned long long)__double_as_longlong(v);
    if (v != v) return ~0ull;
    if (u == 0x8000000000000000ull) u = 0ull;
    return (u & 0x8000000000000000ull) ? ~u : (u | 0x8000000000000000ull);
}

// a float converts to double exactly and monotonically, so one double path serves both score types
template <typename T>
__global__ void __launch_bounds__(kPostWarps * 32) rerank_order_kernel(
    const T* __restrict__ scores, const int32_t* __restrict__ cnt, int B, int k, double thresh, int max_num,
    int min_num, int32_t* __restrict__ order, int32_t* __restrict__ out_cnt, T* __restrict__ out_scores) {
    __shared__ unsigned long long s_key[kPostWarps][kPostMaxK];
    __shared__ unsigned char s_pass[kPostWarps][kPostMaxK];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.x * kPostWarps + warp;
    if (b >= B) return;
    const int n = min(cnt ? cnt[b] : k, k);
    const T* s = scores + (size_t)b * k;
    int n_pass = 0;
    for (int i = lane; i < kPostMaxK; i += 32) {
        const bool live = i < n;
        const double v = live ? (double)s[i] : 0.0;
        const bool pass = live && v >= thresh;               // NaN never passes (src/_modules.py:1585)
        s_key[warp][i] = live ? order_key64(v) : 0ull;
        s_pass[warp][i] = pass;
        n_pass += __popc(__ballot_sync(0xffffffffu, pass));
    }
    __syncwarp();
    // the three outcomes of :1586-1590
    const bool fallback = !(n_pass > max_num) && n_pass < min_num;
    const int n_out = n_pass > max_num ? max_num : (fallback ? min(min_num, n) : n_pass);
    for (int i = lane; i < k; i += 32) {                      // padding first: the scatter below overwrites [0, n_out)
        order[(size_t)b * k + i] = -1;
        if (out_scores) out_scores[(size_t)b * k + i] = (T)0;
    }
    __syncwarp();
    for (int i = lane; i < n; i += 32) {
        const unsigned long long ki = s_key[warp][i];
        int rank = 0, rank_pass = 0;                          // entries sorted before i: all, and those that pass
        for (int j = 0; j < n; ++j) {
            const unsigned long long kj = s_key[warp][j];
            const bool before = kj > ki || (kj == ki && j > i);   // descending; equal scores: higher index first
            rank += before;
            rank_pass += before && s_pass[warp][j];
        }
        const int pos = fallback ? rank : (s_pass[warp][i] ? rank_pass : kPostMaxK);
        if (pos < n_out) {
            order[(size_t)b * k + pos] = i;
            if (out_scores) out_scores[(size_t)b * k + pos] = s[i];
        }
    }
    if (lane == 0) out_cnt[b] = n_out;
}

// CPython's set of small non-negative ints (Objects/setobject.c; hash(i) = i): open addressing, 9 linear probes,
// then the perturbed jump.  Empty slots hold -1.
// returns the slot holding `v`, inserting it if absent (*inserted = true)
__device__ int set_find_or_insert(int* table, unsigned mask, int v, bool* inserted) {
    unsigned perturb = (unsigned)v;
    unsigned i = (unsigned)v & mask;
    while (true) {
        const int probes = (i + 9u <= mask) ? 9 : 0;          // LINEAR_PROBES
        for (int j = 0; j <= probes; ++j) {
            const int e = table[i + j];
            if (e < 0) { table[i + j] = v; *inserted = true; return (int)(i + j); }
            if (e == v) { *inserted = false; return (int)(i + j); }
        }
        perturb >>= 5;                                         // PERTURB_SHIFT
        i = (i * 5u + 1u + perturb) & mask;
    }
}

constexpr int kSetSlots = 128;    // 64 distinct pages: the table grows 8 -> 32 -> 128 (4 x used when fill*5 >= mask*3)

__global__ void __launch_bounds__(kPostWarps * 32) page_vote_kernel(
    const int32_t* __restrict__ hit_page, const int32_t* __restrict__ hit_cnt, const float* __restrict__ sims,
    const int64_t* __restrict__ row_off, int B, int k, int weighted, int legacy_promotion,
    int32_t* __restrict__ major, double* __restrict__ major_weight) {
    __shared__ int s_table[kPostWarps][2][kSetSlots];
    __shared__ double s_acc[kPostWarps][kSetSlots];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.x * kPostWarps + warp;
    if (b >= B) return;
    const int64_t c0 = row_off[b];
    const int n_doc = (int)(row_off[b + 1] - c0);
    const int n_hits = min(min(hit_cnt[b], k), n_doc);        // zip(pages, weights) stops at the shorter one
    const bool f32_acc = weighted && !legacy_promotion;       // NEP 50: int 0 + float32 stays float32
    // ---- sum(w): sequential, in chunk order, starting from the Python int 0 (src/RAGVT5.py:462) ----
    // A chain of n_doc dependent adds is the whole cost of the kernel for long documents (10 k chunks: 156 us measured),
    // so the float64 sum first asks whether the order can matter at all: the inputs are float32 (24-bit significands);
    // if the span from the largest input's exponent down to the smallest non-zero input's last bit, plus log2(n) carry
    // bits, fits the 53-bit significand, EVERY partial sum in ANY order is exact, and a lane-strided sum + warp tree
    // gives the sequential result bit for bit.  Otherwise (tiny or non-finite values, float32 accumulation) the
    // sequential chain runs: broadcasts unrolled ahead of it, tail padded with +0.0 (x + 0.0 == x: a sum that starts
    // at +0 is never -0).
    double total = 0.0;
    float total32 = 0.0f;
    bool done = false;
    if (weighted && legacy_promotion) {
        int emax = 0, emin = 255;
        bool odd = false;                                       // NaN / Inf / subnormal: leave it to the chain
#pragma unroll 8
        for (int i = lane; i < n_doc; i += 32) {                // 8 loads in flight: one warp walks the document alone
            const unsigned bits = __float_as_uint(__ldg(sims + c0 + i));
            const unsigned e = (bits >> 23) & 0xFFu;
            const bool zero = (bits & 0x7FFFFFFFu) == 0u;
            odd |= e == 255u || (e == 0u && !zero);
            emax = max(emax, zero ? 0 : (int)e);
            emin = min(emin, zero ? 255 : (int)e);
        }
        emax = __reduce_max_sync(0xffffffffu, emax);
        emin = __reduce_min_sync(0xffffffffu, emin);
        odd = __any_sync(0xffffffffu, odd);
        // in units of the smallest input's last bit every |partial sum| < n * 2^(emax - emin + 24) <= 2^53: an exact integer
        const int carry = n_doc > 1 ? 32 - __clz(n_doc - 1) : 0;      // ceil(log2(n_doc))
        if (!odd && (emin > emax || emax - emin + 24 + carry <= 53)) {
            double part = 0.0;
#pragma unroll 8
            for (int i = lane; i < n_doc; i += 32) part += (double)__ldg(sims + c0 + i);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
            total = part;
            done = true;
        }
    }
    if (weighted && !done) {
        for (int base = 0; base < n_doc; base += 32) {
            const float v = base + lane < n_doc ? __ldg(sims + c0 + base + lane) : 0.0f;
            float x[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) x[j] = __shfl_sync(0xffffffffu, v, j);
            if (legacy_promotion) {
#pragma unroll
                for (int j = 0; j < 32; ++j) total = __dadd_rn(total, (double)x[j]);
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) total32 = __fadd_rn(total32, x[j]);
            }
        }
    }
    if (lane != 0) return;
    if (n_hits == 0) {                                         // :471-473
        major[b] = 0;
        if (major_weight) major_weight[b] = 0.0;
        return;
    }
    // ---- list(set(page_indices_b)) (:466): iteration order = slot order of CPython's table ----
    int* table = s_table[warp][0];
    int* spare = s_table[warp][1];
    unsigned mask = 7;
    int used = 0;
    for (int i = 0; i < 8; ++i) table[i] = -1;
    const int32_t* pages = hit_page + (size_t)b * k;
    for (int j = 0; j < n_hits; ++j) {
        bool inserted;
        set_find_or_insert(table, mask, pages[j], &inserted);
        if (!inserted) continue;
        ++used;
        if ((unsigned)used * 5u >= mask * 3u) {               // set_table_resize(so, used * 4)
            unsigned size = 8;
            while (size <= (unsigned)used * 4u) size <<= 1;
            for (unsigned i = 0; i < size; ++i) spare[i] = -1;
            for (unsigned i = 0; i <= mask; ++i)
                if (table[i] >= 0) { bool ins; set_find_or_insert(spare, size - 1, table[i], &ins); }
            int* t = table; table = spare; spare = t;
            mask = size - 1;
        }
    }
    // ---- page_weights[page] += weight, in hit order (:468-469) ----
    double* acc = s_acc[warp];
    for (unsigned i = 0; i <= mask; ++i) acc[i] = 0.0;
    const float denom32 = legacy_promotion ? (float)total : total32;   // array / scalar: float32 either way
    const double uniform = 1.0 / (double)n_doc;               // np.ones(n) / sum(ones) (:459, :463)
    for (int j = 0; j < n_hits; ++j) {
        bool ins;
        const int slot = set_find_or_insert(table, mask, pages[j], &ins);
        if (!weighted) {
            acc[slot] = __dadd_rn(acc[slot], uniform);
        } else {
            const float w = __fdiv_rn(__ldg(sims + c0 + j), denom32);   // weight of CHUNK j, as the reference zips
            acc[slot] = f32_acc ? (double)__fadd_rn((float)acc[slot], w) : __dadd_rn(acc[slot], (double)w);
        }
    }
    // ---- max(page_weights, key=page_weights.get) (:474): first maximum in set order ----
    int best = -1;
    for (unsigned i = 0; i <= mask; ++i) {
        if (table[i] < 0) continue;
        if (best < 0 || acc[i] > acc[best]) best = (int)i;
    }
    major[b] = table[best];
    if (major_weight) major_weight[b] = acc[best];
}

}  // namespace rdv

extern "C" int rdv_rerank_order(const void* d_scores, int32_t scores_f64, const int32_t* d_cnt, int32_t B, int32_t k,
                                double filter_thresh, int32_t max_num, int32_t min_num, int32_t* d_order,
                                int32_t* d_out_cnt, void* d_out_scores, void* stream) {
    using namespace rdv;
    RDV_REQUIRE(B >= 0 && k >= 1 && k <= kPostMaxK, RDV_E_LIMIT, "rerank_order: B=%d, k=%d outside [1, %d]", B, k, kPostMaxK);
    RDV_REQUIRE(max_num >= 0 && min_num >= 0, RDV_E_INVALID, "rerank_order: negative max_num / min_num");
    if (B == 0) return RDV_OK;
    RDV_REQUIRE(d_scores && d_order && d_out_cnt, RDV_E_INVALID, "rerank_order: null pointer");
    const int blocks = (B + kPostWarps - 1) / kPostWarps;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (scores_f64)
        rerank_order_kernel<double><<<blocks, kPostWarps * 32, 0, st>>>(static_cast<const double*>(d_scores), d_cnt, B, k,
            filter_thresh, max_num, min_num, d_order, d_out_cnt, static_cast<double*>(d_out_scores));
    else
        rerank_order_kernel<float><<<blocks, kPostWarps * 32, 0, st>>>(static_cast<const float*>(d_scores), d_cnt, B, k,
            filter_thresh, max_num, min_num, d_order, d_out_cnt, static_cast<float*>(d_out_scores));
    RDV_LAUNCH_CHECK("rerank_order_kernel");
    return RDV_OK;
}

extern "C" int rdv_page_vote(const int32_t* d_hit_page, const int32_t* d_hit_cnt, const float* d_sims,
                             const int64_t* d_row_off, int32_t B, int32_t k, int32_t weighted, int32_t legacy_promotion,
                             int32_t* d_major, double* d_major_weight, void* stream) {
    using namespace rdv;
    RDV_REQUIRE(B >= 0 && k >= 1 && k <= kPostMaxK, RDV_E_LIMIT, "page_vote: B=%d, k=%d outside [1, %d]", B, k, kPostMaxK);
    if (B == 0) return RDV_OK;
    RDV_REQUIRE(d_hit_page && d_hit_cnt && d_row_off && d_major && (!weighted || d_sims), RDV_E_INVALID,
                "page_vote: null pointer");
    const int blocks = (B + kPostWarps - 1) / kPostWarps;
    page_vote_kernel<<<blocks, kPostWarps * 32, 0, static_cast<cudaStream_t>(stream)>>>(
        d_hit_page, d_hit_cnt, d_sims, d_row_off, B, k, weighted, legacy_promotion, d_major, d_major_weight);
    RDV_LAUNCH_CHECK("page_vote_kernel");
    return RDV_OK;
}
