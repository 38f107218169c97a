// Gather of the retrieved chunks into the generator's input tensors -- sm_100a.
//
// Replaces, for a pre-tokenised document store, the Python that follows torch.topk in the reference:
//   Retriever._get_top_k           src/_modules.py:2014-2100   hit -> page-list range (+- surroundings),
//                                                              minus words already emitted by better hits
//   Chunker.compact_chunks         src/_modules.py:1102-1132   bbox = min/max over the emitted word boxes
//   crop rectangle                 src/_modules.py:2108-2119   int(bbox * page size), order fix
//   reorder_chunks                 src/_modules.py:2129-2142   stable sort by (page, ymin, xmin)
//   flatten (+ separator)          src/utils.py:233-253
//   VT5.prepare_inputs_for_vqa     src/VT5.py:141-185          prompt ids | word ids (box*1000 truncated,
//                                                              repeated per sub-token) | EOS | padding
// One thread block per document; everything is index arithmetic over CSR arrays (int32 / f64), driven
// by the top-k kernel's device output -- no host round trip between scoring and the generator input.
// Byte/integer work: bit-exact against the oracle.
#include "select.cuh"

namespace rdv {

constexpr int kGatherThreads = 256;
constexpr int kGatherMaxK = 64;
constexpr int kMaxFresh = 24;     // fresh sub-intervals of one hit after removing better hits' ranges
constexpr int kSmemSegs = 4;      // word segments per hit kept in shared memory (more spill to the global ws)

struct GatherParams {
    rdv_docstore ds;
    rdv_gather_args a;
};

struct Interval { int lo, hi; };

// SURR = false: include_surroundings == 0 (every shipped config).  The neighbour-window code (interval subtraction,
// page walks, word-box pass) is compiled out, which matters here: a block runs its code exactly once, so the
// instruction fetch of the un-taken paths' surroundings shows up as `no_instructions` stalls (17 % in profiles/).
template <bool SURR>
__global__ void __launch_bounds__(kGatherThreads) gather_vt5_kernel(const GatherParams P) {
    const rdv_docstore& ds = P.ds;
    const rdv_gather_args& a = P.a;
    const int b = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int k = a.k;

    __shared__ int s_chunk[kGatherMaxK];       // global chunk id of hit i
    __shared__ int s_page[kGatherMaxK];
    __shared__ int s_label[kGatherMaxK];
    __shared__ int s_lo[kGatherMaxK], s_hi[kGatherMaxK];
    __shared__ int s_nseg[kGatherMaxK], s_ntok[kGatherMaxK], s_nwords[kGatherMaxK];
    __shared__ double s_bbox[kGatherMaxK][4];
    __shared__ int s_order[kGatherMaxK];       // output position r -> hit i
    __shared__ int s_start[kGatherMaxK + 1];   // output token offset of ordered hit r (after its separator)
    __shared__ int s_total;
    __shared__ int s_overflow;
    __shared__ int s_hit[kGatherMaxK];         // chunk index (within the document) of hit i
    __shared__ int s_seg[kGatherMaxK][kSmemSegs][2];   // the first segments of every hit (global ws holds all)
    __shared__ int s_seg_tok[kGatherMaxK][kSmemSegs][2];   // their token ranges [begin, end)

    pdl_launch_dependents();   // the next batch's score kernel may be scheduled while this one gathers
    pdl_wait();                // similarities / hits come from the preceding kernel in the stream
    const int64_t c0 = ds.chunk_off[b];
    const int n_doc = (int)(ds.chunk_off[b + 1] - c0);
    const int p0 = a.prompt_off[b], plen = a.prompt_off[b + 1] - p0;   // independent of the hits: issue early
    // no neighbours: a hit is exactly its own chunk, so its bbox is the chunk's precomputed bbox and phase C
    // (a dependent pass over the word boxes) disappears
    const int surroundings = SURR ? a.include_surroundings : 0;
    const bool own_bbox = !SURR && ds.chunk_bbox != nullptr;
    const int page0 = (ds.doc_page_off && ds.page_wh) ? ds.doc_page_off[b] : -1;
    if (a.sims) {
        // fused selection: this block owns document b, so the top-k needs no cross-block traffic at all
        extern __shared__ float4 smem_dyn[];
        __shared__ unsigned long long s_red[kScoreWarps];
        SelectArgs sel;
        sel.k = k; sel.cache_floats = cache_floats_for(a.max_rows, k, 16 * kScoreThreads);
        sel.topk_idx = a.topk_idx; sel.topk_val = a.topk_val; sel.topk_cnt = a.topk_cnt; sel.doc_done = nullptr;
        sel.smem_idx = s_hit;
        select_topk<16>(sel, b, a.sims + c0, n_doc, reinterpret_cast<float*>(smem_dyn), s_red, BlockSync());
        __syncthreads();
    } else if (tid < k) {
        s_hit[tid] = a.topk_idx[(size_t)b * k + tid];
    }
    const int cnt = a.sims ? min(k, n_doc) : a.topk_cnt[b];
    if (!a.sims) __syncthreads();
    int32_t* seg_ws = a.seg_ws + ((size_t)b * k) * (2 * a.max_seg);
    if (tid == 0) s_overflow = 0;

    // ---- A: raw page-list interval of every hit -------------------------------------------------
    rdv_chunk_rec rec = {};
    if (tid < cnt) {
        const int gc = (int)(c0 + s_hit[tid]);
        rec = ds.chunk_rec[gc];                                 // one 32-byte record: no dependent hops
        if (own_bbox) {                                         // independent of rec: both loads in flight together
            const double2* cb = reinterpret_cast<const double2*>(ds.chunk_bbox + (size_t)gc * 4);
            const double2 lo2 = cb[0], hi2 = cb[1];
            s_bbox[tid][0] = lo2.x; s_bbox[tid][1] = lo2.y; s_bbox[tid][2] = hi2.x; s_bbox[tid][3] = hi2.y;
        }
        const int start = rec.page_start;
        const int nw = rec.word_end - rec.word_begin;
        s_chunk[tid] = gc;
        s_page[tid] = rec.page;
        s_label[tid] = rec.label;
        if (!SURR) {
            s_lo[tid] = start; s_hi[tid] = start + nw;         // no neighbours: the page length is not needed
        } else {
            const int last = ds.page_chunks[ds.run_end[gc] - 1];
            const int page_len = ds.chunk_page_start[last] + (ds.chunk_word_off[last + 1] - ds.chunk_word_off[last]);
            s_lo[tid] = max(0, start - surroundings);
            s_hi[tid] = min(page_len, start + nw + surroundings);
        }
        if (!a.reorder_chunks) s_order[tid] = tid;
    }
    if (SURR) __syncthreads();                                  // phase B reads the other hits' intervals

    // ---- B: fresh sub-intervals (minus better hits on the same page) -> global word segments ------
    if constexpr (!SURR) {
        if (tid < cnt) {
            // ranges of distinct chunks are disjoint in the page word list: the hit is exactly its own words
            const int wb = rec.word_begin, we = rec.word_end;
            s_seg[tid][0][0] = wb; s_seg[tid][0][1] = we;
            s_seg_tok[tid][0][0] = rec.tok_begin; s_seg_tok[tid][0][1] = rec.tok_end;
            s_nseg[tid] = we > wb ? 1 : 0;
            s_nwords[tid] = we - wb;
            s_ntok[tid] = rec.tok_end - rec.tok_begin;
        }
    } else if (tid < cnt) {
        Interval fresh[kMaxFresh];
        int nf = 1;
        bool overflow = false;
        fresh[0].lo = s_lo[tid]; fresh[0].hi = s_hi[tid];
        for (int j = 0; j < tid && nf > 0; ++j) {
            if (s_page[j] != s_page[tid]) continue;
            const int cl = s_lo[j], ch = s_hi[j];
            int out = 0;
            Interval next[kMaxFresh];
            for (int f = 0; f < nf; ++f) {
                const int lo = fresh[f].lo, hi = fresh[f].hi;
                if (ch <= lo || cl >= hi) { if (out < kMaxFresh) next[out++] = fresh[f]; else overflow = true; continue; }
                if (lo < cl) { if (out < kMaxFresh) { next[out].lo = lo; next[out].hi = cl; ++out; } else overflow = true; }
                if (ch < hi) { if (out < kMaxFresh) { next[out].lo = ch; next[out].hi = hi; ++out; } else overflow = true; }
            }
            nf = out;
            for (int f = 0; f < nf; ++f) fresh[f] = next[f];
        }
        // walk the page's chunks (ordered by position) and cut the fresh intervals at chunk borders;
        // adjacent pieces that are contiguous in the global word array are merged back
        const int gc = s_chunk[tid];
        const int rb = ds.run_begin[gc], re = ds.run_end[gc];
        int nseg = 0, ntok = 0, nwords = 0;
        int* segs = seg_ws + (size_t)tid * (2 * a.max_seg);
        for (int f = 0; f < nf; ++f) {
            const int lo = fresh[f].lo, hi = fresh[f].hi;
            if (lo >= hi) continue;
            // first slot whose chunk ends after lo
            int s0 = rb, s1 = re;
            while (s0 < s1) {
                const int mid = (s0 + s1) >> 1;
                const int cc = ds.page_chunks[mid];
                const int ce = ds.chunk_page_start[cc] + (ds.chunk_word_off[cc + 1] - ds.chunk_word_off[cc]);
                if (ce > lo) s1 = mid; else s0 = mid + 1;
            }
            for (int slot = s0; slot < re; ++slot) {
                const int cc = ds.page_chunks[slot];
                const int cs = ds.chunk_page_start[cc];
                if (cs >= hi) break;
                const int ce = cs + (ds.chunk_word_off[cc + 1] - ds.chunk_word_off[cc]);
                const int x0 = max(lo, cs), x1 = min(hi, ce);
                if (x0 >= x1) continue;
                const int wb = ds.chunk_word_off[cc] + (x0 - cs), we = ds.chunk_word_off[cc] + (x1 - cs);
                if (nseg > 0 && segs[2 * (nseg - 1) + 1] == wb) {
                    segs[2 * (nseg - 1) + 1] = we;
                } else if (nseg < a.max_seg) {
                    segs[2 * nseg] = wb; segs[2 * nseg + 1] = we; ++nseg;
                } else {
                    overflow = true;
                }
                nwords += x1 - x0;
            }
        }
        for (int sidx = 0; sidx < nseg; ++sidx) {
            const int tb = ds.word_tok_off[segs[2 * sidx]], te = ds.word_tok_off[segs[2 * sidx + 1]];
            ntok += te - tb;
            if (sidx < kSmemSegs) {
                s_seg[tid][sidx][0] = segs[2 * sidx]; s_seg[tid][sidx][1] = segs[2 * sidx + 1];
                s_seg_tok[tid][sidx][0] = tb; s_seg_tok[tid][sidx][1] = te;
            }
        }
        s_nseg[tid] = nseg; s_ntok[tid] = ntok; s_nwords[tid] = nwords;
        if (overflow) s_overflow = 1;
    }
    __syncthreads();

    // ---- C: bbox of the emitted words (one warp per hit), crop rectangle, labels, pages ---------
    for (int i = warp; i < cnt && !own_bbox; i += kGatherThreads / 32) {
        const int* segs = s_nseg[i] <= kSmemSegs ? &s_seg[i][0][0] : seg_ws + (size_t)i * (2 * a.max_seg);
        double x0 = INFINITY, y0 = INFINITY, x1 = -INFINITY, y1 = -INFINITY;
        for (int sidx = 0; sidx < s_nseg[i]; ++sidx) {
            for (int w = segs[2 * sidx] + lane; w < segs[2 * sidx + 1]; w += 32) {
                const double* bx = ds.word_box + (size_t)w * 4;
                x0 = fmin(x0, bx[0]); y0 = fmin(y0, bx[1]); x1 = fmax(x1, bx[2]); y1 = fmax(y1, bx[3]);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            x0 = fmin(x0, __shfl_xor_sync(0xffffffffu, x0, o)); y0 = fmin(y0, __shfl_xor_sync(0xffffffffu, y0, o));
            x1 = fmax(x1, __shfl_xor_sync(0xffffffffu, x1, o)); y1 = fmax(y1, __shfl_xor_sync(0xffffffffu, y1, o));
        }
        if (lane == 0) {
            if (s_nwords[i] == 0) { x0 = 0.0; y0 = 0.0; x1 = 1.0; y1 = 1.0; }   // src/_modules.py:1126-1127
            s_bbox[i][0] = x0; s_bbox[i][1] = y0; s_bbox[i][2] = x1; s_bbox[i][3] = y1;
        }
    }
    if (!own_bbox) __syncthreads();

    // ---- D: output order (identity, or stable sort by (page, ymin, xmin)) -------------------------
    if (a.reorder_chunks) {
        if (tid < cnt) {
            int rank = 0;
            const int pg = s_page[tid];
            const double ky = s_bbox[tid][1], kx = s_bbox[tid][0];
            for (int j = 0; j < cnt; ++j) {
                if (j == tid) continue;
                const int pj = s_page[j];
                const double jy = s_bbox[j][1], jx = s_bbox[j][0];
                const bool less = pj < pg || (pj == pg && (jy < ky || (jy == ky && jx < kx)));
                const bool equal = pj == pg && jy == ky && jx == kx;
                if (less || (equal && j < tid)) ++rank;
            }
            s_order[rank] = tid;
        }
        __syncthreads();
    }

    // ---- D': the reranker's index list (src/_modules.py:1592-1595) applies to retrieve()'s OUTPUT order: positions of
    // the order above are permuted / dropped; the words of every hit (dedup against better hits) stay as retrieved
    int n_out = cnt;
    if (a.emit_order) {
        n_out = min(a.emit_cnt[b], cnt);
        int v = 0;
        if (tid < n_out) {
            const int r = a.emit_order[(size_t)b * k + tid];
            v = s_order[min(max(r, 0), cnt - 1)];
        }
        __syncthreads();
        if (tid < n_out) s_order[tid] = v;
        __syncthreads();
    }

    // per-hit metadata, in OUTPUT order
    if (tid < k) {
        const size_t o = (size_t)b * k + tid;
        if (tid < n_out) {
            const int i = s_order[tid];
            const int gc = s_chunk[i];
            a.hit_chunk[o] = (int32_t)(gc - c0);
            a.hit_page[o] = s_page[i];
            a.hit_label[o] = s_label[i];
            a.hit_nwords[o] = s_nwords[i];
            double* bb = a.hit_bbox + o * 4;
            bb[0] = s_bbox[i][0]; bb[1] = s_bbox[i][1]; bb[2] = s_bbox[i][2]; bb[3] = s_bbox[i][3];
            int32_t* rc = a.hit_rect + o * 4;
            if (page0 >= 0) {
                const int pidx = page0 + s_page[i];
                const double W = (double)ds.page_wh[2 * pidx], H = (double)ds.page_wh[2 * pidx + 1];
                const int rx0 = (int)(bb[0] * W), ry0 = (int)(bb[1] * H);      // int() truncation
                const int rx1 = (int)(bb[2] * W), ry1 = (int)(bb[3] * H);
                rc[0] = min(rx0, rx1); rc[1] = min(ry0, ry1); rc[2] = max(rx0, rx1); rc[3] = max(ry0, ry1);
            } else {
                rc[0] = rc[1] = rc[2] = rc[3] = -1;
            }
        } else {
            a.hit_chunk[o] = -1; a.hit_page[o] = -1; a.hit_label[o] = -1; a.hit_nwords[o] = 0;
            for (int e = 0; e < 4; ++e) { a.hit_bbox[o * 4 + e] = 0.0; a.hit_rect[o * 4 + e] = -1; }
        }
    }

    // ---- E: token offsets of the ordered hits (warp 1, while warp 0 writes the per-hit metadata) --------
    if (tid == 32) {
        int pos = plen;
        for (int r = 0; r < n_out; ++r) {
            const int i = s_order[r];
            if (r > 0 && s_nwords[i] > 0) pos += a.n_sep;       // flatten(): separator before non-empty sublists
            s_start[r] = pos;
            pos += s_ntok[i];
        }
        s_start[n_out] = pos;
        s_total = pos;
        a.full_len[b] = pos + 1;                                // + EOS, before truncation (src/VT5.py:170)
        a.status[b] = s_overflow;
    }
    __syncthreads();

    // ---- F: emit ------------------------------------------------------------------------------------
    const int Lmax = a.max_len;
    const int body = min(s_total, Lmax - 1);                    // ids[:max_len-1] + [eos]  (src/VT5.py:166)
    int64_t* ids = a.out_ids + (size_t)b * Lmax;
    int64_t* box = a.out_boxes + (size_t)b * Lmax * 4;
    int64_t* msk = a.out_mask + (size_t)b * Lmax;
    int64_t* lab = a.out_labels ? a.out_labels + (size_t)b * Lmax : nullptr;
    for (int pos = tid; pos < Lmax; pos += kGatherThreads) {
        int64_t id = a.pad_id, bx0 = 0, bx1 = 0, bx2 = 0, bx3 = 0, m = 0, lb = 4;
        if (pos < body) {
            m = 1;
            if (pos < plen) {
                id = a.prompt_ids[p0 + pos]; bx2 = 1000; bx3 = 1000; lb = 4;      // prompt_box, prompt label
            } else {
                // ordered hit r with s_start[r] - sep <= pos < s_start[r+1]-sep(next)
                int r = 0;
                while (r + 1 < n_out) {
                    const int nxt = s_order[r + 1];
                    const int nxt_begin = s_start[r + 1] - ((s_nwords[nxt] > 0) ? a.n_sep : 0);
                    if (pos < nxt_begin) break;
                    ++r;
                }
                const int i = s_order[r];
                if (pos < s_start[r]) {                          // separator token: box 0, label 0
                    const int sep_begin = s_start[r] - a.n_sep;
                    id = a.sep_ids[pos - sep_begin]; lb = 0;
                } else {
                    int o = pos - s_start[r];
                    int t;
                    if (s_nseg[i] <= kSmemSegs) {                // token ranges of the segments are in shared memory
                        int sidx = 0, nt = s_seg_tok[i][0][1] - s_seg_tok[i][0][0];
                        while (o >= nt) { o -= nt; ++sidx; nt = s_seg_tok[i][sidx][1] - s_seg_tok[i][sidx][0]; }
                        t = s_seg_tok[i][sidx][0] + o;
                    } else {
                        const int* segs = seg_ws + (size_t)i * (2 * a.max_seg);
                        int sidx = 0, wb = segs[0], we = segs[1];
                        int nt = ds.word_tok_off[we] - ds.word_tok_off[wb];
                        while (o >= nt) { o -= nt; ++sidx; wb = segs[2 * sidx]; we = segs[2 * sidx + 1];
                                          nt = ds.word_tok_off[we] - ds.word_tok_off[wb]; }
                        t = ds.word_tok_off[wb] + o;
                    }
                    if (ds.tok_rec) {
                        // one 32-byte record per token: id + its word's box already multiplied by 1000 and truncated
                        const int4* tr = reinterpret_cast<const int4*>(ds.tok_rec + t);
                        const int4 r0 = tr[0], r1 = tr[1];
                        id = r0.x; bx0 = r0.z; bx1 = r0.w; bx2 = r1.x; bx3 = r1.y;
                    } else {
                        id = ds.tok_ids[t];
                        const double* wbx = ds.word_box + (size_t)ds.tok_word[t] * 4;   // token -> its word's box
                        bx0 = (int64_t)(wbx[0] * 1000.0); bx1 = (int64_t)(wbx[1] * 1000.0);   // f64 -> i64 truncation
                        bx2 = (int64_t)(wbx[2] * 1000.0); bx3 = (int64_t)(wbx[3] * 1000.0);
                    }
                    lb = s_label[i];
                }
            }
        } else if (pos == body) {
            id = a.eos_id; m = 1; lb = 4;                        // EOS: box 0, label 4
        }
        ids[pos] = id; msk[pos] = m;
        longlong2* bo = reinterpret_cast<longlong2*>(box + 4 * (size_t)pos);      // (B, L, 4) int64: 32-byte aligned
        bo[0] = make_longlong2(bx0, bx1); bo[1] = make_longlong2(bx2, bx3);
        if (lab) lab[pos] = lb;
    }
}

}  // namespace rdv

extern "C" int rdv_gather_vt5_inputs(const rdv_docstore* ds, const rdv_gather_args* args, void* stream) {
    using namespace rdv;
    RDV_REQUIRE(ds && args, RDV_E_INVALID, "gather_vt5_inputs: null struct");
    RDV_REQUIRE(ds->B >= 0, RDV_E_INVALID, "gather_vt5_inputs: negative B");
    if (ds->B == 0) return RDV_OK;
    RDV_REQUIRE(args->k >= 1 && args->k <= kGatherMaxK, RDV_E_LIMIT, "gather_vt5_inputs: k=%d outside [1, %d]",
                args->k, kGatherMaxK);
    RDV_REQUIRE(args->max_len >= 2 && args->max_seg >= 1 && args->n_sep >= 0 && args->include_surroundings >= 0,
                RDV_E_INVALID, "gather_vt5_inputs: bad max_len / max_seg / n_sep / include_surroundings");
    RDV_REQUIRE(ds->chunk_rec && ds->chunk_off && ds->chunk_word_off && ds->word_tok_off && ds->tok_ids && ds->word_box &&
                ds->tok_word && ds->chunk_label && ds->chunk_page && ds->chunk_page_start && ds->page_chunks && ds->run_begin &&
                ds->run_end, RDV_E_INVALID, "gather_vt5_inputs: docstore has a null array");
    RDV_REQUIRE(args->topk_idx && args->topk_cnt && args->prompt_off && args->prompt_ids && args->seg_ws &&
                args->out_ids && args->out_boxes && args->out_mask && args->full_len && args->status &&
                args->hit_chunk && args->hit_page && args->hit_label && args->hit_nwords && args->hit_bbox &&
                args->hit_rect && (args->n_sep == 0 || args->sep_ids) && (!args->emit_order || args->emit_cnt), RDV_E_INVALID,
                "gather_vt5_inputs: args has a null array");
    RDV_REQUIRE(aligned16(args->out_boxes) && aligned16(ds->chunk_bbox) && aligned16(ds->tok_rec), RDV_E_ALIGN,
                "gather_vt5_inputs: out_boxes / chunk_bbox / tok_rec must be 16-byte aligned");
    RDV_REQUIRE(!args->sims || (args->topk_val && args->max_rows >= 0), RDV_E_INVALID,
                "gather_vt5_inputs: fused selection needs topk_val and max_rows");
    GatherParams P;
    P.ds = *ds;
    P.a = *args;
    size_t smem = 0;
    if (args->sims) {
        RDV_ONCE_PER_DEVICE(cudaFuncSetAttribute(gather_vt5_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024),
                            "cudaFuncSetAttribute(gather_vt5<false>)");
        RDV_ONCE_PER_DEVICE(cudaFuncSetAttribute(gather_vt5_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024),
                            "cudaFuncSetAttribute(gather_vt5<true>)");
        smem = (size_t)(cache_floats_for(args->max_rows, args->k, 16 * kScoreThreads)) * sizeof(float) + 16;
    }
    cudaError_t le = args->include_surroundings != 0
        ? launch_pdl(kPdlSelect, gather_vt5_kernel<true>, dim3(ds->B), dim3(kGatherThreads), smem, static_cast<cudaStream_t>(stream), P)
        : launch_pdl(kPdlSelect, gather_vt5_kernel<false>, dim3(ds->B), dim3(kGatherThreads), smem, static_cast<cudaStream_t>(stream), P);
    if (le != cudaSuccess) return cuda_fail(le, "gather_vt5_kernel");
    return RDV_OK;
}
