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
 *     owned by the caller; nothing is allocated, freed or synchronised inside the library.
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
 * per document, src/_modules.py:1415-1416), so the kernel takes a table of row-block pointers
 * instead of one packed matrix; a packed CSR matrix is the special case ptr[b] = base + row_off[b]*d.
 *
 *   d_doc_ptr   [B]    device pointers to (n_b, d) fp32, row-major, row pitch = d floats; 16-byte aligned
 *   d_row_off   [B+1]  exclusive prefix sum of n_b (int64): document b writes sims[row_off[b] ..)
 *   d_tile_off  [B+1]  exclusive prefix sum of ceil(n_b / tile_rows) (int32)
 *   d_q         (B,d)  fp32 question embeddings, 16-byte aligned
 *   d_sims      [N]    out: every similarity, chunk order (the 9th output of Retriever.retrieve)
 *   d_topk_idx  (B,k)  out: int32 chunk index within the document, rank order, -1 padded
 *   d_topk_val  (B,k)  out: fp32 score of each hit (-inf padded)
 *   d_topk_cnt  [B]    out: k_b
 *   d_doc_done  [B]    int32 workspace, all zero on entry; the kernel leaves it all zero again
 *   tile_rows          rows per thread block: a multiple of 8 in [8, 256] (see rdv_score_tile_rows)
 *   max_rows           max_b n_b (sizes the shared-memory cache of the selection pass)
 * Requirements: d % 4 == 0, 4 <= d <= 8192, 1 <= k <= 1024, B >= 0.
 * ------------------------------------------------------------------------------------------- */
RDV_API int rdv_score_topk_f32(const void* const* d_doc_ptr, const int64_t* d_row_off,
                               const int32_t* d_tile_off, const float* d_q, int32_t B, int32_t d,
                               int32_t k, int32_t tile_rows, int32_t total_tiles, int32_t max_rows,
                               float* d_sims, int32_t* d_topk_idx, float* d_topk_val,
                               int32_t* d_topk_cnt, int32_t* d_doc_done, void* stream);

/* Launch heuristic for tile_rows given the batch's total row count (keeps >= ~8 tiles per SM for
 * small batches so the hardware scheduler can balance ragged documents). */
RDV_API int32_t rdv_score_tile_rows(int64_t total_rows, int32_t d);

#ifdef __cplusplus
}
#endif
#endif /* RDV_H_ */
