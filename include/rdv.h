/*
 * rdv.h -- C ABI of librdv.so: the B200 (sm_100a) retrieval hot path of RAG-DocVQA.
 *
 * The reference (Pikurrot/RAG-DocVQA) is 100% Python and has no FFI of its own: the "operator API"
 * this library sits behind is four Python call sites:
 *     Retriever.retrieve            src/_modules.py:2155-2180
 *     VisualRetriever.retrieve      src/_modules.py:2453-2464
 *     mean_pooling                  src/_model_utils.py:49-61
 *     late_interaction              src/utils.py:442-458
 * Each entry point below names the reference lines whose arithmetic it replaces.  The reference-side
 * binding a maintainer would add is the ctypes stub in INTEGRATION.md (rag_docvqa_b200/_lib.py is that
 * stub, grown into the drop-in classes).
 *
 * Conventions
 *   - plain pointers and sizes only; every `d_` pointer is DEVICE memory of the current CUDA device,
 *     owned by the caller; nothing is allocated, freed or synchronised inside the library (one exception, stated
 *     at its declaration: rdv_retrieve_small_f32 is a whole host-to-host call and ends with a stream synchronise).
 *   - `stream` is a cudaStream_t / CUstream passed as void* (NULL = legacy default stream).
 *   - return value: 0 = launched, <0 = rejected before launch (RDV_E_*); rdv_last_error() describes it.
 *     There is no CPU fallback: a build without the CUDA kernels does not exist.
 *   - thread-safe as long as concurrent calls use distinct output / workspace buffers.
 */
#ifndef RDV_H_
#define RDV_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RDV_API __attribute__((visibility("default")))

#define RDV_OK 0
#define RDV_E_INVALID (-1)   /* bad argument (null pointer, negative size, unsupported shape) */
#define RDV_E_ALIGN (-2)     /* pointer / row pitch not aligned as the kernel requires        */
#define RDV_E_CUDA (-3)      /* CUDA runtime reported an error at launch                      */
#define RDV_E_LIMIT (-4)     /* size beyond what the kernel supports (see the function)       */

/* ABI version: bumped whenever a signature below changes. */
RDV_API int rdv_abi_version(void);

/* Thread-local text of the last error returned on this thread ("" if none). */
RDV_API const char* rdv_last_error(void);

/* sizeof() of a struct of this header by name ("rdv_gather_args", ...), -1 for an unknown name: a binding checks its own
 * mirror of the layout against it (tests/test_abi.py does for the ctypes one). */
RDV_API int64_t rdv_struct_size(const char* name);

/* SM count / compute capability of the current device (as the launch heuristics see it). */
RDV_API int rdv_device_info(int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor);

/* ---------------------------------------------------------------------------------------------
 * Fused cosine score + segmented per-document top-k, fp32 (the parity mode).
 *
 * Replaces Retriever._get_similarities (src/_modules.py:1978-1997):
 *     sim[b][i] = dot(E_b[i], q_b) / (||E_b[i]|| * ||q_b|| + 1e-8)        (fp32, eps on the product)
 * and the per-document torch.topk (src/_modules.py:2015-2016; also :2408 for the visual path) with
 * k_b = min(k, n_b), descending score, ties broken by LOWEST index, NaN greatest, -0 == +0.
 *
 * Documents are ragged and live in separate allocations (BiEncoder.batch_forward returns one tensor
 * per document, src/_modules.py:1415-1416).  The caller cuts the batch into row tiles that never cross
 * a document and passes one 32-byte descriptor per tile (a packed CSR matrix is the special case
 * src = base + row * d * 4):
 *   rdv_tile_desc.src       first row of the tile: `rows` consecutive rows of d fp32, 16-byte aligned
 *   rdv_tile_desc.sims_off  index of that row in d_sims (global chunk number)
 *   rdv_tile_desc.rows      1 .. tile_rows
 *   rdv_tile_desc.doc       document (= question) the tile belongs to
 *   rdv_tile_desc.doc_rows  n_b of that document
 * Tiles of one document must be contiguous in d_tiles and ordered by sims_off.
 *
 *   d_row_off   [B+1]  exclusive prefix sum of n_b (int64)
 *   d_q         (B,d)  fp32 question embeddings, 16-byte aligned
 *   d_sims      [N]    out: every similarity, chunk order (the 9th output of Retriever.retrieve)
 *   d_topk_idx  (B,k)  out: int32 chunk index within the document, rank order, -1 padded
 *   d_topk_val  (B,k)  out: fp32 score of each hit (-inf padded)
 *   d_topk_cnt  [B]    out: k_b
 *   tile_rows          the maximum rdv_tile_desc.rows used (from rdv_score_plan)
 *   algo               RDV_SCORE_TMA (d in {128,256,384,512,768,1024}) or RDV_SCORE_LDG (any d % 4 == 0); see the
 *                      defines below.  Both launch the streaming score kernel followed by the per-document selection
 *                      kernel (two launches, no device-scope fence on the streaming path).  Batches of short
 *                      documents have a ONE-launch path: rdv_score_topk_cluster_f32 below.
 *   max_rows           max_b n_b (sizes the shared-memory cache of the selection pass)
 * Requirements: d % 4 == 0, 4 <= d <= 8192, 1 <= k <= 1024, B >= 0.
 * ------------------------------------------------------------------------------------------- */
typedef struct rdv_tile_desc {
    const void* src;
    int64_t sims_off;
    int32_t rows;
    int32_t doc;
    int32_t doc_rows;
    int32_t reserved;
} rdv_tile_desc;

#define RDV_SCORE_AUTO 0
#define RDV_SCORE_LDG 1        /* one block per tile, 128-bit loads; selection in a second kernel        */
#define RDV_SCORE_TMA 2        /* persistent, bulk-async-copy (TMA) rings in shared memory; selection in a second kernel */

/* Picks the kernel (AUTO -> RDV_SCORE_LDG, the faster one on B200 at every measured size) and the tile
 * height for a batch of total_rows rows. */
RDV_API int rdv_score_plan(int64_t total_rows, int32_t d, int32_t algo, int32_t* algo_out, int32_t* tile_rows);

RDV_API int rdv_score_topk_f32(const rdv_tile_desc* d_tiles, int32_t total_tiles, int32_t tile_rows, int32_t algo,
                               const int64_t* d_row_off, const float* d_q, int32_t B, int32_t d, int32_t k,
                               int32_t max_rows, float* d_sims, int32_t* d_topk_idx, float* d_topk_val,
                               int32_t* d_topk_cnt, void* stream);

/* ---------------------------------------------------------------------------------------------
 * The same contract in ONE launch, for batches of short documents: row tiles packed into thread-block CLUSTERS.
 *
 * The unit of work is a slice of <= 32 rows per CTA (the streaming kernel's parallelism); the host packs the CTAs into
 * clusters of `cluster` CTAs (rdv_cluster_plan) so that all CTAs of a document sit in ONE cluster (a document of n rows
 * takes min(max(ceil(n / slice_rows), 1), cluster) CTAs and splits its rows evenly over them).  Every CTA reduces its
 * slice to the k best (score, index) keys and pushes them into the shared memory of its document's first CTA
 * (distributed shared memory); one cluster barrier later that CTA merges the candidates.  No global-memory handshake
 * (fence / atomic / re-read of the scores) sits between scoring and selection; same values and ordering as
 * rdv_score_topk_f32.
 *   rdv_cta_desc          one 32-byte descriptor per CTA: src = the DOCUMENT's (doc_rows, d) fp32 matrix, sims_off = index
 *                         of its first similarity in d_sims (= its first global chunk number), doc = its index, part /
 *                         nparts = this CTA's place among the document's CTAs (CTA `part` takes rows part * per .. of the
 *                         document, per = ceil(doc_rows / nparts)); nparts == 0 marks a padding CTA.  The CTAs of a document are consecutive within a cluster.
 *   rdv_cluster_table_size   number of descriptors (a multiple of the cluster size) for a batch; -1 on bad input
 *   rdv_build_cluster_table  fills HOST descriptors (best-fit decreasing packing); the caller uploads them
 *   max_rows                 max_b n_b
 * Requirements: k <= rdv_cluster_max_k() (32), max_rows <= rdv_cluster_max_rows(cluster), d % 4 == 0, fewer than 2^31
 * rows; `cluster` the size the table was built for; outside that: RDV_E_LIMIT, run rdv_score_topk_f32.  rdv_retrieve_plan says
 * whether this path is the faster one.
 * ------------------------------------------------------------------------------------------- */
typedef struct rdv_cta_desc {
    const void* src;
    int32_t sims_off;
    int32_t doc_rows;
    int32_t doc;
    int16_t part;
    int16_t nparts;
    int32_t reserved[2];
} rdv_cta_desc;

/* The plan for a batch (no launch): `cluster` = 16 (non-portable cluster size) where the current device places 16 CTAs of
 * the kernel variant such a batch runs in one GPC, else the portable 8; `slice_rows` = the most rows a CTA takes of a document
 * of at most cluster CTAs (32, 40, ... 64): the smallest for which ALL clusters of the batch are resident at once (a second
 * wave of clusters would start only when whole clusters retire); `n_ctas` = descriptors the table needs, 0 when the batch is
 * outside the kernels' limits.  The occupancy calculator is asked once per device and variant; without a device: 8 / 32.
 * RDV_CLUSTER_SIZE=8|16 and RDV_CLUSTER_SLICE=32..64 in the environment force the two (measurement knobs). */
RDV_API int rdv_cluster_plan(const int64_t* rows, int32_t B, int32_t d, int32_t k, int32_t with_gather, int32_t* cluster,
                             int32_t* slice_rows, int64_t* n_ctas);
RDV_API int32_t rdv_cluster_max_rows(int32_t cluster);     /* 4 slices of 32 rows per CTA: 1024 (clusters of 8), 2048 (16) */
RDV_API int32_t rdv_cluster_max_k(void);
/* *use_cluster = 1 when the one-launch cluster kernels apply to the batch (limits above) AND are the measured better choice
 * on B200.  As measured in round 2 they are not (C2, one dependent chain: the whole step 14.4-16.3 us against 13.3 us for two
 * launches; score + top-k alone 10.3-10.7 us), so this answers 1 only for batches of at most 1 MB (one launch less in a call
 * that is all fixed cost) unless RDV_CLUSTER=1 / 0 in the environment forces it; the kernels stay selectable and are covered by
 * the parity tests. */
RDV_API int rdv_retrieve_plan(int64_t total_rows, int32_t max_rows, int32_t B, int32_t d, int32_t k, int32_t* use_cluster);
RDV_API int64_t rdv_cluster_table_size(const int64_t* rows, int32_t B, int32_t cluster, int32_t slice_rows);
RDV_API int rdv_build_cluster_table(const void* const* d_docs, const int64_t* rows, int32_t B, int32_t d, int32_t cluster,
                                    int32_t slice_rows, rdv_cta_desc* h_ctas, int64_t n_ctas);
/* Diagnostics of the cluster kernels (never on in the product): while d_trace is non-NULL every CTA writes %globaltimer
 * (ns) stamps to d_trace[cta * 8 + s]: s = 0 started, 1 rows streamed, 2 cluster complete, 3 candidates pushed, 4 (merging
 * CTA) candidates here, 5 merged, 6 gathered.  The buffer holds 8 x n_ctas uint64; NULL switches it off. */
RDV_API int rdv_debug_trace(void* d_trace);
RDV_API int rdv_score_topk_cluster_f32(const rdv_cta_desc* d_ctas, int64_t n_ctas, int32_t cluster, const float* d_q, int32_t B,
                                       int32_t d, int32_t k, int32_t max_rows, float* d_sims, int32_t* d_topk_idx,
                                       float* d_topk_val, int32_t* d_topk_cnt, void* stream);

/* Host-side staging helpers (no kernel is launched): the per-batch bookkeeping of the caller, in C.
 *   rdv_count_tiles      number of row tiles a ragged batch cuts into (sum of ceil(rows[b] / tile_rows)); -1 on bad input.
 *   rdv_build_doc_table  fills HOST buffers h_row_off[B+1] and h_tiles[n_tiles] (n_tiles = rdv_count_tiles) from the
 *                        B device pointers d_docs[b] (rows[b] x d fp32 each, NULL allowed when rows[b] == 0); the caller
 *                        uploads them (one pinned copy) and passes the device copies to the score entry points.
 *   rdv_upload_docs_f32  one cudaMemcpyAsync per document: HOST matrices h_docs[b] (rows[b] x d fp32) -> consecutive rows
 *                        of the device buffer d_packed (sum rows x d).  Replaces B framework-level copies. */
RDV_API int64_t rdv_count_tiles(const int64_t* rows, int32_t B, int32_t tile_rows);
RDV_API int rdv_build_doc_table(const void* const* d_docs, const int64_t* rows, int32_t B, int32_t d, int32_t tile_rows,
                                int64_t* h_row_off, rdv_tile_desc* h_tiles, int64_t n_tiles, int32_t* max_rows);
RDV_API int rdv_upload_docs_f32(const void* const* h_docs, const int64_t* rows, int32_t B, int32_t d, float* d_packed,
                                void* stream);

/* Small host batches: the device round trip of Retriever.retrieve (src/_modules.py:2155-2180: _get_similarities
 * :1978-1997 + torch.topk :2015-2016) in ONE call, for batches whose fixed costs exceed their work (C1: 1 page x 30
 * chunks x 384-d = 46 KB; the reference runs this case on the CPU).
 *   upload blob  (pinned host h_blob, device twin d_blob):  row_off[B+1] i64 | pad | tiles[n_tiles] | ctas[n_ctas] | questions (B,d) | rows
 *   result blob  (device d_out; the first read_bytes are copied to pinned h_out):  sims[total] | idx (B,k) | cnt[B] || val (B,k)
 * rdv_small_batch_layout   offsets and sizes of both blobs + the launch plan (pure host arithmetic).
 * rdv_small_batch_pack     fills h_blob: copies the B host matrices h_docs[b] (rows[b] x d fp32, contiguous) and the
 *                          questions h_q (B,d), and builds offsets + tile descriptors that point into d_blob (pure host code).
 * rdv_retrieve_small_f32   layout + pack + ONE cudaMemcpyAsync up + ONE launch (rdv_score_topk_cluster_f32; two launches,
 *                          rdv_score_topk_f32, outside the cluster kernel's limits) + ONE cudaMemcpyAsync back +
 *                          cudaStreamSynchronize: on return h_out holds similarities, hits and counts (same values and
 *                          ordering as rdv_score_topk_f32).  Returns RDV_SMALL_GROW (> 0), with *lay filled and nothing
 *                          else touched, when a buffer is smaller than the layout needs.  This entry point synchronises
 *                          the stream (it is the whole call). */
typedef struct rdv_small_layout {
    int32_t algo, tile_rows, n_tiles, max_rows;             /* algo: RDV_SCORE_LDG / _TMA, or RDV_SMALL_CLUSTER */
    int64_t total_rows, n_ctas;                             /* n_ctas: cluster-kernel descriptors (0 when not taken) */
    int64_t cluster, slice_rows;                            /* ... and the plan they were packed for (rdv_cluster_plan) */
    int64_t o_tiles, o_ctas, o_q, o_emb, in_bytes;          /* upload blob (row_off at 0) */
    int64_t o_idx, o_cnt, read_bytes, o_val, out_bytes;     /* result blob (sims at 0)    */
} rdv_small_layout;
#define RDV_SMALL_GROW 1
#define RDV_SMALL_CLUSTER 3    /* rdv_small_layout.algo: the batch takes the one-launch cluster kernel */
RDV_API int rdv_small_batch_layout(const int64_t* rows, int32_t B, int32_t d, int32_t k, rdv_small_layout* lay);
RDV_API int rdv_small_batch_pack(const void* const* h_docs, const int64_t* rows, int32_t B, int32_t d, const float* h_q,
                                 const rdv_small_layout* lay, void* h_blob, const void* d_blob);
RDV_API int rdv_retrieve_small_f32(const void* const* h_docs, const int64_t* rows, int32_t B, int32_t d, int32_t k,
                                   const float* h_q, void* h_blob, void* d_blob, int64_t blob_bytes, void* d_out,
                                   int64_t d_out_bytes, void* h_out, int64_t h_out_bytes, rdv_small_layout* lay,
                                   void* stream);

/* Scores only (the streaming half of rdv_score_topk_f32): writes d_sims.  Used when the selection runs
 * elsewhere (rdv_topk_segments_f32, or inside rdv_gather_vt5_inputs).  algo: RDV_SCORE_LDG or RDV_SCORE_TMA. */
RDV_API int rdv_score_f32(const rdv_tile_desc* d_tiles, int32_t total_tiles, int32_t tile_rows, int32_t algo,
                          const float* d_q, int32_t B, int32_t d, float* d_sims, void* stream);

/* Stand-alone segmented top-k over an existing score vector (same ordering rules as above); the
 * selection half of rdv_score_topk_f32, and the visual path's top-k over MaxSim scores (torch.topk at
 * src/_modules.py:2408).  d_scores [N] fp32, d_row_off [B+1] int64; outputs as in rdv_score_topk_f32. */
RDV_API int rdv_topk_segments_f32(const float* d_scores, const int64_t* d_row_off, int32_t B, int32_t k,
                                  int32_t max_rows, int32_t* d_topk_idx, float* d_topk_val,
                                  int32_t* d_topk_cnt, void* stream);




/* Selection half of pooled-patch visual retrieval (BASELINE.json configs[3] in north_star's wording): d_sims holds the cosine
 * of EVERY patch vector (rdv_score_f32 on the (strips * L, d) matrices; src/_modules.py:1990-1993), strip s owning
 * d_sims[d_strip_row_off[s] .. d_strip_row_off[s + 1]) (L scores), document b the strips d_doc_strip_off[b] ..
 * d_doc_strip_off[b + 1].  Two launches: the min(k, L) best patches of every strip (register-resident, one block per strip),
 * then per document the k best of its strips' candidates as flat patch indices (strip in the document) * L + patch, a
 * strip's score = its best patch (torch.max: NaN greatest) and the k_strips best strips (torch.topk, src/_modules.py:2408;
 * same ordering rules as rdv_score_topk_f32).  Workspaces d_ws_idx / d_ws_val (n_strips, min(k, L)) and d_ws_cnt (n_strips);
 * outputs d_patch_* (B, k), d_strip_scores [n_strips], d_strip_* (B, k_strips).
 * Requirements: k, k_strips <= 32, max_strips * min(k, L) <= 4096. */
RDV_API int rdv_pooled_select_f32(const float* d_sims, const int64_t* d_strip_row_off, int64_t n_strips, int32_t L,
                                  const int64_t* d_doc_strip_off, int32_t B, int32_t max_strips, int32_t k, int32_t k_strips,
                                  int32_t* d_ws_idx, float* d_ws_val, int32_t* d_ws_cnt, int32_t* d_patch_idx,
                                  float* d_patch_val, int32_t* d_patch_cnt, float* d_strip_scores, int32_t* d_strip_idx,
                                  float* d_strip_val, int32_t* d_strip_cnt, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Masked mean pooling of encoder token outputs (+ optional fused L2 normalisation / bf16 copy).
 *
 * Replaces mean_pooling (src/_model_utils.py:49-61; call site src/_modules.py:1474):
 *     out[i,:] = sum_t embs[i,t,:] * mask[i,t] / max(sum_t mask[i,t], 1e-9)
 *   d_embs (n,L,d) fp32 contiguous, 16-byte aligned;  d_mask (n,L) int64 (HF tokenizer attention mask)
 *   normalise != 0: out row is divided by max(||row||, 1e-12) (F.normalize, as src/utils.py:445 does for
 *                   the visual path; sentence-transformers' Normalize module for BGE)
 *   d_out (n,d) fp32 or NULL; d_out_bf16 (n,d) bf16 or NULL (corpus shards); d_out_norm (n,) or NULL:
 *   L2 norm of the pooled row before normalisation.
 * Tokens with mask == 0 are not read.  Requirements: d % 4 == 0, d <= 8192, L <= 4096.
 * ------------------------------------------------------------------------------------------- */
RDV_API int rdv_mean_pool_f32(const float* d_embs, const int64_t* d_mask, int32_t n, int32_t L, int32_t d,
                              int32_t normalise, float* d_out, void* d_out_bf16, float* d_out_norm,
                              void* stream);

/* ---------------------------------------------------------------------------------------------
 * MaxSim late interaction, fp32 parity mode.
 *
 * Replaces late_interaction (src/utils.py:442-458; call site src/_modules.py:2202):
 *     score[n] = sum_i max_j < Q[i]/max(||Q[i]||,1e-12) , P[n][j]/max(||P[n][j]||,1e-12) >
 *   rdv_row_inv_norm_f32: inv[r] = 1 / max(||x[r]||, 1e-12) for a (rows, d) fp32 matrix (F.normalize's
 *                         denominator, src/utils.py:445-446); run it on Q and on P first.
 *   rdv_maxsim_f32: d_q (Lq,d), d_p (n,Lp,d) fp32 16-byte aligned; d_partial (n * rdv_maxsim_tiles_i(Lq))
 *                   fp32 workspace; d_counter (n) int32 workspace, zero on entry, left zero; d_out (n).
 * The (Lq x Lp) similarity matrix is never materialised.  Requirements: d % 4 == 0, Lp >= 1, n <= 65535.
 * ------------------------------------------------------------------------------------------- */
RDV_API int rdv_row_inv_norm_f32(const float* d_x, int64_t rows, int32_t d, float* d_inv, void* stream);
RDV_API int rdv_maxsim_f32(const float* d_q, const float* d_p, const float* d_inv_q, const float* d_inv_p,
                           int32_t n, int32_t Lq, int32_t Lp, int32_t d, float* d_partial,
                           int32_t* d_counter, float* d_out, void* stream);
RDV_API int32_t rdv_maxsim_tiles_i(int32_t Lq);

/* ---------------------------------------------------------------------------------------------
 * Merge of per-shard top-k candidates (corpus mode, BASELINE.json configs[4]; no reference
 * counterpart -- the reference is single-GPU).  (Q, m) candidates with GLOBAL ids (id < 0 = empty)
 * -> (Q, k) by (score desc, id asc), NaN greatest: the ordering of rdv_score_topk_f32, so a sharded
 * search returns exactly what an unsharded one would.
 * ------------------------------------------------------------------------------------------- */
RDV_API int rdv_topk_merge(const float* d_cand_val, const int64_t* d_cand_idx, int32_t Q, int32_t m, int32_t k,
                           float* d_out_val, int64_t* d_out_idx, void* stream);
/* The same merge over `parts` candidate lists that live `part_stride_*` ELEMENTS apart: list r of query q is
 * d_cand_val[r * part_stride_val + q * k_in .. + k_in) (ids likewise).  With parts = world this reads the receive buffer
 * of the all-gather exactly as NCCL fills it (rank-major), so no transpose copy sits between the collective and the
 * merge; parts = 1 is rdv_topk_merge. */
RDV_API int rdv_topk_merge_parts(const float* d_cand_val, const int64_t* d_cand_idx, int32_t Q, int32_t parts, int32_t k_in,
                                 int64_t part_stride_val, int64_t part_stride_idx, int32_t k, float* d_out_val,
                                 int64_t* d_out_idx, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Gather of the retrieved chunks into the VT5 generator's input tensors.
 *
 * Replaces, for a pre-tokenised batch of documents, everything between torch.topk and the generator's
 * embedding lookup: Retriever._get_top_k (src/_modules.py:2014-2100: page-list range of each hit
 * +- include_surroundings, minus words already emitted by better hits), Chunker.compact_chunks
 * (src/_modules.py:1102-1132: bbox), the crop rectangle (src/_modules.py:2108-2119), reorder_chunks
 * (src/_modules.py:2129-2142), flatten with separator (src/utils.py:233-253) and the id/box/mask/label
 * packing of VT5.prepare_inputs_for_vqa (src/VT5.py:141-185).
 *
 * rdv_docstore: CSR view of B documents, all arrays in DEVICE memory (N chunks, W words, T tokens):
 *   chunk_rec[N]            rdv_chunk_rec: the per-chunk fields below packed so a hit costs one 32-byte load
 *   chunk_off[B+1] i64      first global chunk of each document (== d_row_off of the score kernel)
 *   chunk_word_off[N+1]     first global word of each chunk        word_tok_off[W+1]  first token of each word
 *   tok_ids[T]              token ids (the tokenizer's ids without EOS, src/VT5.py:160)
 *   tok_word[T]             global word index of each token
 *   word_box[W*4] f64       word boxes, 0..1, as the Python floats the reference multiplies by 1000
 *   chunk_label[N], chunk_page[N]
 *   page_chunks[N]          GLOBAL chunk ids grouped by (document, page), chunk order inside a page
 *   run_begin[N], run_end[N]  per chunk: the [begin, end) slice of page_chunks holding its page
 *   chunk_page_start[N]     position of the chunk's first word in its page's word list (src/_modules.py:2044)
 *   doc_page_off[B+1], page_wh[P*2]  (optional) page sizes in pixels for the crop rectangles
 *   tok_rec[T]              (optional) rdv_tok_rec: token id + its word's box as emitted, int(box * 1000)
 *                           (src/VT5.py:162,174: float64 product truncated toward zero) -- one 32-byte load per
 *                           output token instead of tok_ids -> tok_word -> word_box; requires |box * 1000| < 2^31
 *   chunk_bbox[N*4] f64     (optional) bbox of each chunk's own words (src/_modules.py:1102-1132; [0,0,1,1] for an
 *                           empty chunk): with include_surroundings == 0 a hit IS its chunk, so no pass over
 *                           the word boxes is needed
 * rdv_gather_args: hits (the score kernel's d_topk_idx / d_topk_cnt, row pitch k), options, prompt ids
 *   (without EOS, src/VT5.py:147-148), separator ids, and the outputs:
 *   out_ids/out_mask/out_labels (B,max_len) i64, out_boxes (B,max_len,4) i64  (out_labels may be NULL)
 *   full_len[B]   len(input_ids)+1 before truncation: longest_seq = min(max_b full_len, max_len) (src/VT5.py:170)
 *   status[B]     0 ok, 1 = segment workspace overflow (raise max_seg)
 *   hit_* (B,k)   per hit in OUTPUT order: chunk index (-1 pad), page, label, #emitted words,
 *                 bbox (f64 x4; [0,0,1,1] when no word was emitted) and crop rectangle (int32 x4; -1 without pages)
 *   seg_ws        int32 workspace of B*k*max_seg*2
 *   sims / topk_val / max_rows   optional fused selection: when `sims` is non-NULL the kernel first selects the
 *                 top-k of sims[chunk_off[b] .. chunk_off[b+1]) itself (same ordering as rdv_score_topk_f32)
 *                 and WRITES topk_idx / topk_val / topk_cnt, so a step is rdv_score_f32 + this kernel.
 *   emit_order (B,k) / emit_cnt[B]   optional (NULL = everything, in retrieve()'s order): the reranker's index list
 *                 (rdv_rerank_order; src/_modules.py:1592-1595).  Output position r then holds the hit that retrieve()
 *                 -- after reorder_chunks -- would have put at position emit_order[b][r], and only emit_cnt[b] hits
 *                 are emitted; every hit keeps the words it was retrieved with (dedup against better hits).
 * Requirements: 1 <= k <= 64.
 * ------------------------------------------------------------------------------------------- */
typedef struct rdv_chunk_rec {   /* everything the gather needs about one chunk, in one 32-byte load */
    int32_t word_begin, word_end;   /* = chunk_word_off[c], chunk_word_off[c+1]           */
    int32_t tok_begin, tok_end;     /* = word_tok_off[word_begin], word_tok_off[word_end] */
    int32_t page, label, page_start, reserved;
} rdv_chunk_rec;

typedef struct rdv_tok_rec {     /* everything the emission needs about one token, in one 32-byte load */
    int32_t id;                  /* = tok_ids[t]                                    */
    int32_t word;                /* = tok_word[t]                                   */
    int32_t box[4];              /* = (int64)(word_box[word][e] * 1000.0)           */
    int32_t reserved[2];
} rdv_tok_rec;

typedef struct rdv_docstore {
    int32_t B;
    int32_t reserved;
    const rdv_chunk_rec* chunk_rec;
    const int64_t* chunk_off;
    const int32_t* chunk_word_off;
    const int32_t* word_tok_off;
    const int32_t* tok_ids;
    const int32_t* tok_word;
    const double* word_box;
    const int32_t* chunk_label;
    const int32_t* chunk_page;
    const int32_t* chunk_page_start;
    const int32_t* page_chunks;
    const int32_t* run_begin;
    const int32_t* run_end;
    const int32_t* doc_page_off;
    const int32_t* page_wh;
    const rdv_tok_rec* tok_rec;
    const double* chunk_bbox;
} rdv_docstore;

typedef struct rdv_gather_args {
    int32_t* topk_idx;
    int32_t* topk_cnt;
    int32_t k;
    int32_t include_surroundings;
    int32_t reorder_chunks;
    int32_t n_sep;
    const int32_t* prompt_off;
    const int32_t* prompt_ids;
    const int32_t* sep_ids;
    int32_t eos_id;
    int32_t pad_id;
    int32_t max_len;
    int32_t max_seg;
    int32_t* seg_ws;
    int64_t* out_ids;
    int64_t* out_boxes;
    int64_t* out_mask;
    int64_t* out_labels;
    int32_t* full_len;
    int32_t* status;
    int32_t* hit_chunk;
    int32_t* hit_page;
    int32_t* hit_label;
    int32_t* hit_nwords;
    double* hit_bbox;
    int32_t* hit_rect;
    const float* sims;
    float* topk_val;
    int32_t max_rows;
    int32_t reserved;
    const int32_t* emit_order;
    const int32_t* emit_cnt;
} rdv_gather_args;

RDV_API int rdv_gather_vt5_inputs(const rdv_docstore* ds, const rdv_gather_args* args, void* stream);

/* The whole retrieval step of RAGVT5 in ONE launch: Retriever._get_similarities (src/_modules.py:1978-1997), the
 * per-document torch.topk (:2015-2016), Retriever._get_top_k (:2014-2121) and the id / box / mask packing of
 * VT5.prepare_inputs_for_vqa (src/VT5.py:141-185): rdv_score_topk_cluster_f32 whose merging CTA goes straight on to
 * gather its document (rdv_gather_vt5_inputs' body).  Every CTA also stages the 32-byte chunk records / bounding boxes of
 * its rows into shared memory while the rows stream (cp.async) and pushes those of its candidates along with the keys,
 * so after the streaming phase the step has ONE dependent global load left (the winners' token records) before the
 * packed tensors are written.
 *   d_ctas / n_ctas / d_q / d / max_rows / d_sims   as rdv_score_topk_cluster_f32 (B = ds->B; document b's sims_off must
 *                   equal ds->chunk_off[b]: the store and the embeddings describe the same chunks)
 *   ds, args        as rdv_gather_vt5_inputs; args->topk_idx / topk_val / topk_cnt are WRITTEN, args->sims is ignored
 * Requirements: the cluster kernel's limits, include_surroundings == 0, emit_order == NULL (otherwise RDV_E_LIMIT /
 * RDV_E_INVALID: run rdv_score_f32 + rdv_gather_vt5_inputs). */
RDV_API int rdv_retrieve_vt5_f32(const rdv_cta_desc* d_ctas, int64_t n_ctas, int32_t cluster, const float* d_q, int32_t d,
                                 int32_t max_rows, float* d_sims, const rdv_docstore* ds, const rdv_gather_args* args,
                                 void* stream);


/* ---------------------------------------------------------------------------------------------
 * The generator's input embeddings from the gather's tensors (last clause): the tail of
 * VT5.prepare_inputs_for_vqa (src/VT5.py:194-204)
 *     input_embeds = language_backbone.shared(ids) + spatial_embedding(boxes) [+ layout_embedding(labels) * scale]
 * with SpatialEmbeddings.forward (src/_modules.py:70-86; inference, dropout = identity):
 *     Linear(LayerNorm(x_emb[l] + y_emb[u] + x_emb[r] + y_emb[b])).
 * LayerNorm's centring and the Linear are linear in the summed rows, so the (tokens, D) x (D, D) GEMM of the reference is
 * folded into per-coordinate tables once per model (rdv_vt5_embed_tables_build, fp64 accumulation):
 *     xw / yw (n_pos, D)   ((row - mean(row)) * ln_weight) W^T          c (D) = ln_bias W^T + lin_bias
 *     gxx / gxy / gyy (n_pos, n_pos)   dot products of the centred rows: var(sum) * D is ten of their entries
 * and rdv_vt5_input_embeds_f32 is ONE pass: per token four table rows, ten scalars, the token's embedding row, one row out
 * (fp32; agrees with the reference's fp32 modules to ~1e-6 relative, the tolerance tests/test_vt5_embed_gpu.py states).
 *   d_ids (B, L) int64 or NULL (spatial embedding only), d_boxes (B, L, 4) int64, d_labels (B, L) int64 or NULL; row b of
 *   each starts b * ld tokens into its buffer (ld >= L: the gather's (B, max_len) buffers trimmed to the longest row);
 *   d_out (B, out_ld, D) with out_ld >= L: the tokens of row b go to rows [b * out_ld, b * out_ld + L) -- out_ld = L is the
 *   contiguous (B, L, D); out_ld = L + n_visual_tokens writes straight into the buffer the reference builds with torch.cat
 *   (src/VT5.py:205), whose tail the caller fills with the generator's visual tokens.  *d_bad (may be NULL) gets bit 0 / 1 / 2 set when a box coordinate / token id / layout
 *   label is outside its table (torch.nn.Embedding raises; here the entry is clamped and the flag tells the host).
 *   d_work: two int32 of scratch (the chunk counter of the persistent kernel), ZERO before the first launch; the kernel
 *   leaves them zero.  Launches that may run concurrently (different streams) need their own. */
typedef struct rdv_vt5_embed_tables {
    int32_t D, n_pos;
    const float *xw, *yw, *gxx, *gxy, *gyy, *c;
    float eps;                /* LayerNorm eps (CustomT5Config.layer_norm_eps, src/_modules.py:45) */
    int32_t V;                /* rows of `shared` */
    const float* shared;      /* (V, D) language_backbone.shared.weight or NULL */
    const float* layout;      /* (n_labels, D) layout_embedding.weight or NULL */
    int32_t n_labels;
    float layout_scale;       /* layout_embedding_scale (src/VT5.py:35) */
} rdv_vt5_embed_tables;

/* d_ws_means: 2 * n_pos doubles of scratch.  D: a multiple of 4, <= 1024.  d_lin_bias may be NULL. */
RDV_API int rdv_vt5_embed_tables_build(const float* d_x_emb, const float* d_y_emb, int32_t n_pos, int32_t D,
                                       const float* d_ln_weight, const float* d_ln_bias, const float* d_lin_weight,
                                       const float* d_lin_bias, double* d_ws_means, float* d_xw, float* d_yw,
                                       float* d_gxx, float* d_gxy, float* d_gyy, float* d_c, void* stream);
RDV_API int rdv_vt5_input_embeds_f32(const rdv_vt5_embed_tables* t, const int64_t* d_ids, const int64_t* d_boxes,
                                     const int64_t* d_labels, int32_t B, int32_t L, int64_t ld, float* d_out,
                                     int64_t out_ld, int32_t* d_bad, int32_t* d_work, void* stream);


/* ---------------------------------------------------------------------------------------------
 * What consumes the hits before generation.
 *
 * rdv_rerank_order: Reranker.rerank after the cross-encoder call (src/_modules.py:1579-1595).  d_scores (B,k) float32
 *   (scores_f64 = 0: CrossEncoder.predict) or float64 (scores_f64 = 1: FlagLLMReranker.compute_score), d_cnt[B]
 *   candidates per document (NULL = k).  np.argsort(scores)[::-1] (equal scores: higher index first, NaN first), keep
 *   scores >= filter_thresh (compared in float64), more than max_num -> the first max_num, fewer than min_num ->
 *   the first min_num of the unfiltered order.  d_order (B,k) int32, -1 padded; d_out_cnt[B]; d_out_scores (B,k)
 *   optional, the scores in the new order (same type as d_scores).  1 <= k <= 64.
 * rdv_page_vote: major_page_indices of RAGVT5.forward (src/RAGVT5.py:455-475).  d_hit_page (B,k) / d_hit_cnt[B]:
 *   top_k_page_indices; d_sims / d_row_off[B+1]: every similarity of the batch (the score kernel's output).
 *   weighted = 0 (majorpage): every hit weighs 1 / n_b.  weighted = 1 (weightmajorpage): hit j weighs
 *   sims[b][j] / sum(sims[b]) -- the similarity of CHUNK j, as the reference's zip pairs them -- with the sum
 *   taken sequentially in chunk order.  legacy_promotion = 1: numpy 1.x scalar promotion (the reference pins
 *   1.26.4): sum and per-page accumulation in float64; 0: NEP 50 (numpy >= 2): float32.  Ties between pages go to
 *   the first page in CPython's iteration order of set(pages) (Objects/setobject.c, reproduced slot for slot);
 *   no hits -> page 0.  d_major[B] int32; d_major_weight[B] float64 optional (the winning page's weight).
 * ------------------------------------------------------------------------------------------- */
RDV_API int rdv_rerank_order(const void* d_scores, int32_t scores_f64, const int32_t* d_cnt, int32_t B, int32_t k,
                             double filter_thresh, int32_t max_num, int32_t min_num, int32_t* d_order,
                             int32_t* d_out_cnt, void* d_out_scores, void* stream);
RDV_API int rdv_page_vote(const int32_t* d_hit_page, const int32_t* d_hit_cnt, const float* d_sims,
                          const int64_t* d_row_off, int32_t B, int32_t k, int32_t weighted, int32_t legacy_promotion,
                          int32_t* d_major, double* d_major_weight, void* stream);


/* ---------------------------------------------------------------------------------------------
 * Chunker: OCR words -> layout boxes.
 *
 * rdv_layout_assign: containment_ratio(word, layout box) > 0.5 (src/utils.py:328-341) for every word x layout box of
 *   every page (src/_modules.py:1023-1033).  P pages, W words, G layout boxes, all arrays in device memory:
 *   d_word_box (W,4) f64 / d_page_word_off[P+1]; d_lay_box (G,4) f64, d_lay_label[G] / d_page_lay_off[P+1], the boxes
 *   of a page in the order the reference visits them (sorted by (xmin, ymin), :1006-1018).  Words are processed in
 *   groups of 32 consecutive words of one page: d_page_group_off[P+1] (prefix sums of ceil(words / 32)),
 *   d_group_page[n_groups] (page of each group).  Outputs: d_bits -- for layout box g a row of ceil(words of its
 *   page / 32) uint32 starting at d_bits_off[g] (bit l of word j = word 32*j + l of the page is inside);
 *   d_word_label[W] -- the label of the last containing box in visiting order, else default_label (:1003, :1030).
 *   float64, every operation rounded individually in the reference's order: the decision is bit-exact against Python
 *   floats (boxes given as ints are exact below 2^26).
 * ------------------------------------------------------------------------------------------- */
RDV_API int rdv_layout_assign(const double* d_word_box, const int32_t* d_page_word_off, const double* d_lay_box,
                              const int32_t* d_lay_label, const int32_t* d_page_lay_off, const int32_t* d_group_page,
                              const int32_t* d_page_group_off, int32_t n_groups, int32_t default_label,
                              const int64_t* d_bits_off, uint32_t* d_bits, int32_t* d_word_label, void* stream);

/* rdv_s2_weights: S2Chunker's pairwise weight matrices of the layout regions ("nodes") of every page of a batch in one
 *   launch (src/_modules.py:1755-1802).  P pages, N nodes; page p owns nodes
 *   d_page_node_off[p] .. d_page_node_off[p+1] and the n_p x n_p row-major matrix at d_out[d_out_off[p]] (d_out_off[P+1]
 *   = prefix sums of n_p^2, total_entries = d_out_off[P]):
 *     spatial[i][j]  = 1 / (1 + ||centroid_i - centroid_j||)   (_spatial_weights_calculation :1755-1773), float64, each
 *                      operation rounded as numpy does here (sqrt(fma(dy, dy, dx*dx)): OpenBLAS ddot on an FMA CPU)
 *     semantic[i][j] = sklearn cosine_similarity of the rows of d_emb (N, d) fp32 (_semantic_weights_calculation
 *                      :1775-1788): rows divided by their norm (zero rows untouched) in fp32, then the dot product
 *     out            = (spatial + semantic) / 2                 (_combined_weights :1790-1802)
 *   d_emb == NULL is cluster_mode "spatial" (semantic = spatial, so out == spatial bit for bit).  `what` selects the
 *   matrix written: RDV_S2_COMBINED, or one of its two terms (RDV_S2_SPATIAL, RDV_S2_SEMANTIC -- the latter needs d_emb).
 *   d_node_box (N,4) f64 xyxy, 16-byte aligned. */
#define RDV_S2_COMBINED 0
#define RDV_S2_SPATIAL 1
#define RDV_S2_SEMANTIC 2
RDV_API int rdv_s2_weights(const double* d_node_box, const int32_t* d_page_node_off, int32_t P, const float* d_emb,
                           int32_t d, int32_t what, const int64_t* d_out_off, int64_t total_entries, double* d_out,
                           void* stream);


/* ---------------------------------------------------------------------------------------------
 * Retrieved patches -> the generator's visual input.
 *
 * Replaces, for page images resident on the device, the host work that follows retrieval when
 * page_retrieval == "concat" (every shipped config; src/RAGVT5.py:378): page.crop(rect) per hit
 * (src/_modules.py:2102-2121), concatenate_patches(patches, mode="grid") (src/utils.py:180-231: strip packing
 * in hit order on a canvas max-width x int(total area / max width), pastes clipped to the canvas, 5 x 5 blank
 * image when there is no hit) and the feature extractor's resize to out_size x out_size (src/_modules.py:133;
 * PIL.Image.resize = Pillow's two-pass 8-bit resampler: horizontal pass rounded to uint8, then vertical
 * pass; filter 2 = BILINEAR, 3 = BICUBIC), then pixel_values = (u8 * (1/255) - mean) / std as fp32 CHW.
 * uint8 results are bit-exact against Pillow.
 *
 * rdv_pagestore: B documents; page (b, p) is entry doc_page_off[b] + p of page_wh[P*2] (width, height) and
 *   page_off[P] (int64 byte offset into `pixels` of its RGB rows, tightly packed H x W x 3 uint8; several
 *   entries may share one image).
 * rdv_visual_args: hit_page / hit_rect / hit_cnt = the gather kernel's per-hit outputs in output order
 *   ((B,k) page index within the document, (B,k,4) crop rectangle x0,y0,x1,y1, (B) number of hits);
 *   workspaces layout (B * (8 + 4k) int32), coeff_h / coeff_v (B * out_size * (ksize_cap + 2) int32),
 *   temp (B * rows_cap * out_size * 3 bytes); outputs out_u8 (B, S, S, 3), out_px (B, 3, S, S) fp32 or NULL,
 *   status[B]: 0 ok, 1 = a capacity (ksize_cap_h / ksize_cap_v / rows_cap / 8192-pixel canvas width) is too
 *   small, 2 = degenerate canvas (zero width or height: the reference raises there).  Capacities for a batch
 *   whose pages are at most Wmax x Hmax: ksize_cap = 2 * ceil(support * max(extent / out_size, 1)) + 1 with
 *   support 1 (bilinear) or 2 (bicubic) and extent = Wmax (h) or k * Hmax (v); rows_cap = k * Hmax.
 * Requirements: 1 <= k <= 64.
 * ------------------------------------------------------------------------------------------- */
typedef struct rdv_pagestore {
    int32_t B;
    int32_t reserved;
    const int32_t* doc_page_off;
    const int32_t* page_wh;
    const int64_t* page_off;
    const uint8_t* pixels;
} rdv_pagestore;

typedef struct rdv_visual_args {
    const int32_t* hit_page;
    const int32_t* hit_rect;
    const int32_t* hit_cnt;
    int32_t k;
    int32_t out_size;
    int32_t filter;
    int32_t ksize_cap_h;
    int32_t ksize_cap_v;
    int32_t rows_cap;
    int32_t max_page_w;          /* widest page of the batch, 1..8192 (sizes the launch and its shared memory) */
    int32_t reserved;
    float mean[3];
    float std[3];
    int32_t* layout;
    int32_t* coeff_h;
    int32_t* coeff_v;
    uint8_t* temp;
    uint8_t* out_u8;
    float* out_px;
    int32_t* status;
} rdv_visual_args;

RDV_API int rdv_visual_pack(const rdv_pagestore* ps, const rdv_visual_args* args, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Retrieved image crops -> Pix2Struct flattened patches (SURVEY.md section 8 row a12, Pix2Struct half).
 *
 * Replaces what src/custom_pix2struct_processor.py does with the crops VisualRetriever returns (call site
 * src/RAGPix2Struct.py:221): per-image normalize (:175-196), extract_flattened_patches_single (:33-95: patch grid from
 * the patch budget, anti-aliased bilinear resize = torch F.interpolate(antialias=True), patch x patch patches flattened
 * pixel-major / channel-minor behind (row id, column id)), extract_multi_image_flattened_patches (:97-132: equal budget
 * per image, row ids continue across images, zero padding) and the attention mask (:225).  The header text that
 * render_header draws onto the first image (:214) is out of scope: an image is a crop of a page in the store.
 *
 * rdv_p2s_img (one per image, HOST-planned and uploaded; `page` indexes the store's page_wh / page_off entries):
 *   crop rectangle x0,y0,x1,y1 (black outside the page), patch grid rows x cols, `kept` = min(rows*cols, budget),
 *   out_start = first output row of the image within its document, row_offset = row ids already used by the
 *   document's earlier images, temp_off = offset (floats) of its horizontally resized rows in `temp`
 *   ((y1-y0) * cols*patch * 3 floats).
 * rdv_p2s_args: images[n_images], stats (16 bytes per image: workspace for the exact byte sums, zeroed by the call),
 *   temp, doc_total[n_docs] = output rows used by each document, out (n_docs, max_total, 2 + patch*patch*3) fp32,
 *   mask (n_docs, max_total) fp32; max_rw / max_rwh / max_h = largest resized width / resized width*height / crop
 *   height over the images (launch bounds; n_images <= 65535).
 * fp32 like the reference; results agree with torch's CPU kernel to float rounding (tests state the tolerance).
 * ------------------------------------------------------------------------------------------- */
typedef struct rdv_p2s_img {
    int32_t doc, page;
    int32_t x0, y0, x1, y1;
    int32_t rows, cols, kept, out_start, row_offset, reserved;
    int64_t temp_off;
} rdv_p2s_img;

typedef struct rdv_p2s_args {
    const rdv_p2s_img* images;
    int32_t n_images, n_docs;
    int32_t max_total, patch;
    int32_t do_normalize, max_rw;
    int64_t max_rwh;
    void* stats;
    float* temp;
    const int32_t* doc_total;
    float* out;
    float* mask;
    int32_t max_h;
    int32_t reserved;
} rdv_p2s_args;

RDV_API int rdv_pix2struct_patches(const rdv_pagestore* ps, const rdv_p2s_args* args, void* stream);

/* ---------------------------------------------------------------------------------------------
 * bf16 tensor-core scoring (tcgen05.mma, accumulators in TMEM, operands staged by TMA).
 *
 * rdv_rows_to_bf16      fp32 (rows, d) -> bf16 copy, optionally L2-normalised first (F.normalize,
 *                       src/utils.py:445-446); d_inv_norm (optional) = 1 / ||bf16 row||.
 * rdv_bf16_inv_norm     inverse row norms of a bf16 matrix (built once per corpus shard).
 *
 * rdv_corpus_score_topk_bf16 -- corpus mode (BASELINE.json configs[4]): n_questions x n_rows cosine
 *   (formula of src/_modules.py:1990-1993 on every pair) with the per-question top-k folded into the
 *   accumulator epilogue; the score matrix is never written.  Row chunks (`groups`, from
 *   rdv_corpus_groups) are scored independently; outputs are per-question CANDIDATE lists
 *   d_cand_val / d_cand_idx (n_questions, groups * rdv_tc_candidates_per_group()) with global ids
 *   id_offset + row (empty slots: id -1), ready for rdv_topk_merge (locally, and again after the
 *   all-gather across shards).  d_part_val / d_part_idx: workspace of
 *   groups * ceil(n_questions / rdv_tc_tile_m()) * rdv_tc_tile_m() * rdv_tc_candidates_per_group().
 *   bf16 mode is reported as recall@k against the fp32 result, not as bit parity.
 *   Requirements: d % 8 == 0, 1 <= k <= 16, n_rows < 2^31.
 *
 * rdv_maxsim_bf16_tc -- fast mode of late_interaction (src/utils.py:442-458): d_qn_bf16 (Lq, d) and
 *   d_pn_bf16 (n, Lp, d) are the L2-NORMALISED bf16 copies (rdv_rows_to_bf16 with normalise = 1);
 *   d_partial: workspace of n * ceil(Lq / rdv_tc_tile_m()) floats; d_out (n).
 * ------------------------------------------------------------------------------------------- */
RDV_API int rdv_rows_to_bf16(const float* d_x, int64_t rows, int32_t d, int32_t normalise, void* d_out_bf16,
                             float* d_inv_norm, void* stream);
RDV_API int rdv_bf16_inv_norm(const void* d_x_bf16, int64_t rows, int32_t d, float* d_inv_norm, void* stream);
RDV_API int32_t rdv_corpus_groups(int64_t n_rows, int32_t n_questions);
RDV_API int32_t rdv_tc_tile_m(void);
RDV_API int32_t rdv_tc_candidates_per_group(void);
RDV_API int rdv_corpus_score_topk_bf16(const void* d_e_bf16, const float* d_e_inv_norm, int64_t n_rows, int32_t d,
                                       const void* d_q_bf16, const float* d_q_inv_norm, int32_t n_questions,
                                       int32_t k, int64_t id_offset, int32_t groups, float* d_part_val,
                                       int32_t* d_part_idx, float* d_cand_val, int64_t* d_cand_idx, void* stream);
RDV_API int rdv_maxsim_bf16_tc(const void* d_qn_bf16, const void* d_pn_bf16, int32_t n, int32_t Lq, int32_t Lp,
                               int32_t d, float* d_partial, float* d_out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * fp32-grade MaxSim on the tensor pipe (3xTF32): the parity mode of late_interaction (src/utils.py:442-458)
 * at tensor-core speed.  Every fp32 value is split exactly into hi = tf32(x) and lo = x - hi;
 *     q . p = q_hi.p_hi + q_hi.p_lo + q_lo.p_hi   (+ q_lo.p_lo ~ 2^-22 |q||p|, dropped)
 * runs as three tcgen05 kind::tf32 MMAs accumulating in fp32 in TMEM (the dominant product in its own
 * accumulator).  Measured against float64 on B200: every cosine comes out ~d * 2.1e-9 relative low (1.6e-6 at
 * d = 768, 4.4e-6 at d = 2048: the tensor core's accumulator rounds toward zero at each step), inside the
 * north_star bar of 1e-5 relative; rdv_maxsim_f32 (~1e-7) is the strict fp32 alternative.
 *   rdv_rows_split_tf32   x (rows, d) fp32 -> hi, lo (rows, d) fp32; normalise != 0: x is first divided by
 *                         max(||x||, 1e-12) (F.normalize, src/utils.py:445-446).
 *   rdv_maxsim_tf32x3_tc  d_q_hi/lo (Lq, d), d_p_hi/lo (n, Lp, d): the NORMALISED, split operands;
 *                         d_partial: workspace of n * ceil(Lq / rdv_tc_tile_m()) floats; d_out (n).
 * Requirements: d % 4 == 0.
 * ------------------------------------------------------------------------------------------- */
RDV_API int rdv_rows_split_tf32(const float* d_x, int64_t rows, int32_t d, int32_t normalise, float* d_hi, float* d_lo,
                                void* stream);
RDV_API int rdv_maxsim_tf32x3_tc(const float* d_q_hi, const float* d_q_lo, const float* d_p_hi, const float* d_p_lo,
                                 int32_t n, int32_t Lq, int32_t Lp, int32_t d, float* d_partial, float* d_out,
                                 void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RDV_H_ */
