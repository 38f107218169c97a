// Gather of ONE document's retrieved chunks into the generator's input tensors -- device code shared by the
// stand-alone gather kernel (gather.cu) and the one-launch retrieval kernel (score_topk.cu).  sm_100a.
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
// One thread block (256 threads) per document; everything is index arithmetic over CSR arrays (int32 / f64).
// Byte/integer work: bit-exact against the oracle.
#pragma once
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

int rdv_gather_check_args(const rdv_docstore* ds, const rdv_gather_args* args);   // gather.cu

// Per-block scratch of gather_document (~9 KB).  The caller owns it, so a kernel that does other work before the
// gather can overlay it with its own scratch.
struct GatherSmem {
    int chunk[kGatherMaxK];       // global chunk id of hit i
    int page[kGatherMaxK];
    int label[kGatherMaxK];
    int lo[kGatherMaxK], hi[kGatherMaxK];
    int nseg[kGatherMaxK], ntok[kGatherMaxK], nwords[kGatherMaxK];
    double bbox[kGatherMaxK][4];
    int order[kGatherMaxK];       // output position r -> hit i
    int start[kGatherMaxK + 1];   // output token offset of ordered hit r (after its separator)
    int total;
    int overflow;
    int hit[kGatherMaxK];         // chunk index (within the document) of hit i: filled by the caller
    int seg[kGatherMaxK][kSmemSegs][2];       // the first segments of every hit (global ws holds all)
    int seg_tok[kGatherMaxK][kSmemSegs][2];   // their token ranges [begin, end)
};

// The caller has put the document's hits (chunk indices within the document, rank order) into S.hit[0 .. cnt) and
// synchronised the block; all kGatherThreads threads of the block call this.
// Optional, for a caller that already holds them (shared memory): c0_known = the document's first global chunk (-1: read
// ds.chunk_off[b]); pre_rec / pre_bbox = the hits' chunk records / chunk_bbox entries, indexed by hit; pre_doc =
// {doc_page_off[b], prompt_off[b], prompt_off[b + 1]} -- every one of them a dependent global load less on the way to
// the packed tensors.
template <bool SURR>
__device__ __forceinline__ void gather_document(const rdv_docstore& ds, const rdv_gather_args& a, const int b, const int cnt,
                                                GatherSmem& S, const int64_t c0_known = -1,
                                                const rdv_chunk_rec* pre_rec = nullptr, const double* pre_bbox = nullptr,
                                                const int* pre_doc = nullptr) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int k = a.k;
    const int64_t c0 = c0_known >= 0 ? c0_known : ds.chunk_off[b];
    const int p0 = pre_doc ? pre_doc[1] : a.prompt_off[b];             // independent of the hits: issue early
    const int plen = (pre_doc ? pre_doc[2] : a.prompt_off[b + 1]) - p0;
    // no neighbours: a hit is exactly its own chunk, so its bbox is the chunk's precomputed bbox and phase C
    // (a dependent pass over the word boxes) disappears
    const int surroundings = SURR ? a.include_surroundings : 0;
    const bool own_bbox = !SURR && ds.chunk_bbox != nullptr;
    const int page0 = (ds.doc_page_off && ds.page_wh) ? (pre_doc ? pre_doc[0] : ds.doc_page_off[b]) : -1;
    int32_t* seg_ws = a.seg_ws + ((size_t)b * k) * (2 * a.max_seg);
    if (tid == 0) S.overflow = 0;

    // ---- A: raw page-list interval of every hit -------------------------------------------------
    rdv_chunk_rec rec = {};
    if (tid < cnt) {
        const int gc = (int)(c0 + S.hit[tid]);
        rec = pre_rec ? pre_rec[tid] : ds.chunk_rec[gc];        // one 32-byte record: no dependent hops
        if (own_bbox) {                                         // independent of rec: both loads in flight together
            const double2* cb = pre_bbox ? reinterpret_cast<const double2*>(pre_bbox + (size_t)tid * 4)
                                         : reinterpret_cast<const double2*>(ds.chunk_bbox + (size_t)gc * 4);
            const double2 lo2 = cb[0], hi2 = cb[1];
            S.bbox[tid][0] = lo2.x; S.bbox[tid][1] = lo2.y; S.bbox[tid][2] = hi2.x; S.bbox[tid][3] = hi2.y;
        }
        const int start = rec.page_start;
        const int nw = rec.word_end - rec.word_begin;
        S.chunk[tid] = gc;
        S.page[tid] = rec.page;
        S.label[tid] = rec.label;
        if (!SURR) {
            S.lo[tid] = start; S.hi[tid] = start + nw;         // no neighbours: the page length is not needed
        } else {
            const int last = ds.page_chunks[ds.run_end[gc] - 1];
            const int page_len = ds.chunk_page_start[last] + (ds.chunk_word_off[last + 1] - ds.chunk_word_off[last]);
            S.lo[tid] = max(0, start - surroundings);
            S.hi[tid] = min(page_len, start + nw + surroundings);
        }
        if (!a.reorder_chunks) S.order[tid] = tid;
    }
    if (SURR) __syncthreads();                                  // phase B reads the other hits' intervals

    // ---- B: fresh sub-intervals (minus better hits on the same page) -> global word segments ------
    if constexpr (!SURR) {
        if (tid < cnt) {
            // ranges of distinct chunks are disjoint in the page word list: the hit is exactly its own words
            const int wb = rec.word_begin, we = rec.word_end;
            S.seg[tid][0][0] = wb; S.seg[tid][0][1] = we;
            S.seg_tok[tid][0][0] = rec.tok_begin; S.seg_tok[tid][0][1] = rec.tok_end;
            S.nseg[tid] = we > wb ? 1 : 0;
            S.nwords[tid] = we - wb;
            S.ntok[tid] = rec.tok_end - rec.tok_begin;
        }
    } else if (tid < cnt) {
        Interval fresh[kMaxFresh];
        int nf = 1;
        bool overflow = false;
        fresh[0].lo = S.lo[tid]; fresh[0].hi = S.hi[tid];
        for (int j = 0; j < tid && nf > 0; ++j) {
            if (S.page[j] != S.page[tid]) continue;
            const int cl = S.lo[j], ch = S.hi[j];
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
        const int gc = S.chunk[tid];
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
                S.seg[tid][sidx][0] = segs[2 * sidx]; S.seg[tid][sidx][1] = segs[2 * sidx + 1];
                S.seg_tok[tid][sidx][0] = tb; S.seg_tok[tid][sidx][1] = te;
            }
        }
        S.nseg[tid] = nseg; S.ntok[tid] = ntok; S.nwords[tid] = nwords;
        if (overflow) S.overflow = 1;
    }
    __syncthreads();

    // ---- C: bbox of the emitted words (one warp per hit), crop rectangle, labels, pages ---------
    for (int i = warp; i < cnt && !own_bbox; i += kGatherThreads / 32) {
        const int* segs = S.nseg[i] <= kSmemSegs ? &S.seg[i][0][0] : seg_ws + (size_t)i * (2 * a.max_seg);
        double x0 = INFINITY, y0 = INFINITY, x1 = -INFINITY, y1 = -INFINITY;
        for (int sidx = 0; sidx < S.nseg[i]; ++sidx) {
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
            if (S.nwords[i] == 0) { x0 = 0.0; y0 = 0.0; x1 = 1.0; y1 = 1.0; }   // src/_modules.py:1126-1127
            S.bbox[i][0] = x0; S.bbox[i][1] = y0; S.bbox[i][2] = x1; S.bbox[i][3] = y1;
        }
    }
    if (!own_bbox) __syncthreads();

    // ---- D: output order (identity, or stable sort by (page, ymin, xmin)) -------------------------
    if (a.reorder_chunks) {
        if (tid < cnt) {
            int rank = 0;
            const int pg = S.page[tid];
            const double ky = S.bbox[tid][1], kx = S.bbox[tid][0];
            for (int j = 0; j < cnt; ++j) {
                if (j == tid) continue;
                const int pj = S.page[j];
                const double jy = S.bbox[j][1], jx = S.bbox[j][0];
                const bool less = pj < pg || (pj == pg && (jy < ky || (jy == ky && jx < kx)));
                const bool equal = pj == pg && jy == ky && jx == kx;
                if (less || (equal && j < tid)) ++rank;
            }
            S.order[rank] = tid;
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
            v = S.order[min(max(r, 0), cnt - 1)];
        }
        __syncthreads();
        if (tid < n_out) S.order[tid] = v;
        __syncthreads();
    }

    // per-hit metadata, in OUTPUT order.  The page size a crop rectangle needs is one more dependent load: it is ISSUED
    // here and only used after the emission below, so it never holds up the block barrier in between.
    int page_w = 0, page_h = 0;
    if (tid < k) {
        const size_t o = (size_t)b * k + tid;
        if (tid < n_out) {
            const int i = S.order[tid];
            const int gc = S.chunk[i];
            if (page0 >= 0) {
                const int pidx = page0 + S.page[i];
                page_w = ds.page_wh[2 * pidx]; page_h = ds.page_wh[2 * pidx + 1];
            }
            a.hit_chunk[o] = (int32_t)(gc - c0);
            a.hit_page[o] = S.page[i];
            a.hit_label[o] = S.label[i];
            a.hit_nwords[o] = S.nwords[i];
            double* bb = a.hit_bbox + o * 4;
            bb[0] = S.bbox[i][0]; bb[1] = S.bbox[i][1]; bb[2] = S.bbox[i][2]; bb[3] = S.bbox[i][3];
        } else {
            a.hit_chunk[o] = -1; a.hit_page[o] = -1; a.hit_label[o] = -1; a.hit_nwords[o] = 0;
            for (int e = 0; e < 4; ++e) { a.hit_bbox[o * 4 + e] = 0.0; a.hit_rect[o * 4 + e] = -1; }
        }
    }

    // ---- E: token offsets of the ordered hits (warp 1, while warp 0 writes the per-hit metadata) --------
    if (tid == 32) {
        int pos = plen;
        for (int r = 0; r < n_out; ++r) {
            const int i = S.order[r];
            if (r > 0 && S.nwords[i] > 0) pos += a.n_sep;       // flatten(): separator before non-empty sublists
            S.start[r] = pos;
            pos += S.ntok[i];
        }
        S.start[n_out] = pos;
        S.total = pos;
        a.full_len[b] = pos + 1;                                // + EOS, before truncation (src/VT5.py:170)
        a.status[b] = S.overflow;
    }
    __syncthreads();

    // ---- F: emit ------------------------------------------------------------------------------------
    const int Lmax = a.max_len;
    const int body = min(S.total, Lmax - 1);                    // ids[:max_len-1] + [eos]  (src/VT5.py:166)
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
                // ordered hit r with S.start[r] - sep <= pos < S.start[r+1]-sep(next)
                int r = 0;
                while (r + 1 < n_out) {
                    const int nxt = S.order[r + 1];
                    const int nxt_begin = S.start[r + 1] - ((S.nwords[nxt] > 0) ? a.n_sep : 0);
                    if (pos < nxt_begin) break;
                    ++r;
                }
                const int i = S.order[r];
                if (pos < S.start[r]) {                          // separator token: box 0, label 0
                    const int sep_begin = S.start[r] - a.n_sep;
                    id = a.sep_ids[pos - sep_begin]; lb = 0;
                } else {
                    int o = pos - S.start[r];
                    int t;
                    if (S.nseg[i] <= kSmemSegs) {                // token ranges of the segments are in shared memory
                        int sidx = 0, nt = S.seg_tok[i][0][1] - S.seg_tok[i][0][0];
                        while (o >= nt) { o -= nt; ++sidx; nt = S.seg_tok[i][sidx][1] - S.seg_tok[i][sidx][0]; }
                        t = S.seg_tok[i][sidx][0] + o;
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
                    lb = S.label[i];
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

    // crop rectangle of every hit (src/_modules.py:2108-2119): int() truncation of bbox * page size, then the order fix
    if (tid < n_out) {
        const int i = S.order[tid];
        int32_t* rc = a.hit_rect + ((size_t)b * k + tid) * 4;
        if (page0 >= 0) {
            const double W = (double)page_w, H = (double)page_h;
            const int rx0 = (int)(S.bbox[i][0] * W), ry0 = (int)(S.bbox[i][1] * H);
            const int rx1 = (int)(S.bbox[i][2] * W), ry1 = (int)(S.bbox[i][3] * H);
            rc[0] = min(rx0, rx1); rc[1] = min(ry0, ry1); rc[2] = max(rx0, rx1); rc[3] = max(ry0, ry1);
        } else {
            rc[0] = rc[1] = rc[2] = rc[3] = -1;
        }
    }
}

}  // namespace rdv
