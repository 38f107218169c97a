// Gather of the retrieved chunks into the generator's input tensors -- sm_100a (entry point + stand-alone kernel).
// Stand-alone gather kernel: one thread block per document (gather.cuh holds the per-document body).
// With `sims` the block first selects the document's top-k itself (rdv_gather_vt5_inputs, include/rdv.h).
#include "gather.cuh"

namespace rdv {

// SURR = false: include_surroundings == 0 (every shipped config).  The neighbour-window code (interval subtraction,
// page walks, word-box pass) is compiled out, which matters here: a block runs its code exactly once, so the
// instruction fetch of the un-taken paths' surroundings shows up as `no_instructions` stalls (17 % in profiles/).
template <bool SURR>
__global__ void __launch_bounds__(kGatherThreads) gather_vt5_kernel(const GatherParams P) {
    const rdv_docstore& ds = P.ds;
    const rdv_gather_args& a = P.a;
    const int b = blockIdx.x;
    const int tid = threadIdx.x;
    const int k = a.k;
    __shared__ GatherSmem S;

    pdl_launch_dependents();   // the next batch's score kernel may be scheduled while this one gathers
    pdl_wait();                // similarities / hits come from the preceding kernel in the stream
    int cnt;
    if (a.sims) {
        // fused selection: this block owns document b, so the top-k needs no cross-block traffic at all
        extern __shared__ float4 smem_dyn[];
        __shared__ unsigned long long s_red[kScoreWarps];
        const int64_t c0 = ds.chunk_off[b];
        const int n_doc = (int)(ds.chunk_off[b + 1] - c0);
        SelectArgs sel;
        sel.k = k; sel.cache_floats = cache_floats_for(a.max_rows, k, 16 * kScoreThreads);
        sel.topk_idx = a.topk_idx; sel.topk_val = a.topk_val; sel.topk_cnt = a.topk_cnt; sel.doc_done = nullptr;
        sel.smem_idx = S.hit;
        select_topk<16>(sel, b, a.sims + c0, n_doc, reinterpret_cast<float*>(smem_dyn), s_red, BlockSync());
        cnt = min(k, n_doc);
    } else {
        if (tid < k) S.hit[tid] = a.topk_idx[(size_t)b * k + tid];
        cnt = a.topk_cnt[b];
    }
    __syncthreads();
    gather_document<SURR>(ds, a, b, cnt, S);
}

}  // namespace rdv

namespace rdv {
int rdv_gather_check_args(const rdv_docstore* ds, const rdv_gather_args* args) {
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
    return RDV_OK;
}
}  // namespace rdv

extern "C" int rdv_gather_vt5_inputs(const rdv_docstore* ds, const rdv_gather_args* args, void* stream) {
    using namespace rdv;
    RDV_REQUIRE(ds && args, RDV_E_INVALID, "gather_vt5_inputs: null struct");
    RDV_REQUIRE(ds->B >= 0, RDV_E_INVALID, "gather_vt5_inputs: negative B");
    if (ds->B == 0) return RDV_OK;
    int rc = rdv_gather_check_args(ds, args);
    if (rc) return rc;
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
