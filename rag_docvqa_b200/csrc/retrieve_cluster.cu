// One-launch per-document retrieval: row tiles packed into thread-block CLUSTERS, documents never straddling one -- sm_100a.
//
// Replaces, in ONE kernel launch and with no global-memory handshake between its stages:
//   Retriever._get_similarities     src/_modules.py:1978-1997   cosine of question b against every chunk of document b
//   torch.topk per document         src/_modules.py:2015-2016   k_b = min(k, n_b), (score desc, lowest index first)
//   Retriever._get_top_k + VT5.prepare_inputs_for_vqa (gather.cuh)   -- MODE 2 only
//
// Why.  A batch the reference really runs (C2: 64 questions x <= 600 chunks x 384-d = 32 MB) is 5 us of HBM time, so
// every dependent global-memory round trip (~0.6 us) after the streaming phase shows.  Measured on B200, one dependent
// chain: streaming score kernel 6.9 us + (block per document: re-read scores, select, gather) 5.9 us = 13.2 us; a
// one-launch kernel that hands per-tile candidates to a "last block" through global memory (fence, atomic, re-read)
// 14.0 us; a cluster of 8 CTAs per DOCUMENT 13.2 us for score + top-k alone (a 600-chunk document then has 64 warps
// and a chain of five dependent load rounds where the tile kernel gives it 152 warps and two).
//
// (A dependent round of row loads costs ~1.4 us under load; an empty launch 2.5 us -- scripts/probe_cluster_launch.cu:
// clusters launch as fast as plain CTAs.)
//
// Here the unit of work stays a slice of <= 32 rows per CTA -- two load rounds per warp, exactly the streaming kernel's
// parallelism -- and the HOST packs the CTAs into clusters of 16 so that all CTAs of a document sit in ONE cluster
// (rdv_build_cluster_table: best-fit decreasing; a document of n rows takes min(ceil(n / 32), 16) CTAs and splits its
// rows evenly over them).  A CTA reduces its slice to the k best (score, index) keys with one warp and PUSHES keys, values and the
// candidates' chunk records / bounding boxes (staged into shared memory by cp.async while the rows stream) into the
// shared memory of its document's first CTA (st.shared::cluster).  One cluster barrier later that CTA merges the
// candidates and goes straight to the winners' token records: after the streaming phase the step has ONE dependent global
// load left before the packed tensors are written.
//
// Every byte of E is read from HBM exactly once; every similarity is written (the reference returns them all).
#include "gather.cuh"

#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

namespace rdv {

constexpr int kClMax = 16;                      // CTAs per cluster (non-portable size; 8 where 16 is refused)
constexpr int kClTileRows = 32;                 // 8 warps x 2 rows in flight x 2 rounds
constexpr int kClMaxLocalTiles = 4;             // a CTA takes up to 4 x 32 rows of a long document

struct ClusterParams {
    const rdv_cta_desc* ctas;                   // [n_ctas]
    const float* q;
    int32_t n_ctas, d, k, reserved;
    float* sims;
    int32_t* topk_idx;
    float* topk_val;
    int32_t* topk_cnt;
    unsigned long long* trace;                  // diagnostics (rdv_debug_trace): 8 x %globaltimer stamps per CTA, or null
};

__device__ __forceinline__ void cl_stamp(const ClusterParams& p, int slot) {
    if (p.trace && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        p.trace[(size_t)blockIdx.x * 8 + slot] = t;
    }
}

__device__ __forceinline__ uint32_t cl_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cl_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cl_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
// address of `local` (a shared-memory object of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t cl_map(const void* local, uint32_t rank) {
    uint32_t a = (uint32_t)__cvta_generic_to_shared(local), r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank));
    return r;
}
__device__ __forceinline__ void cl_st_u64(uint32_t addr, unsigned long long v) {
    asm volatile("st.shared::cluster.u64 [%0], %1;" ::"r"(addr), "l"(v) : "memory");
}
__device__ __forceinline__ void cl_st_f32(uint32_t addr, float v) {
    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void cl_st_u4(uint32_t addr, uint4 v) {
    asm volatile("st.shared::cluster.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// st.async: a store into another CTA's shared memory that counts its bytes on an mbarrier THERE when the data has
// landed -- the receiver waits on its own mbarrier and needs no fence, no cluster barrier and no still-resident sender
__device__ __forceinline__ void cl_sta_u64(uint32_t addr, unsigned long long v, uint32_t bar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];" ::"r"(addr), "l"(v), "r"(bar) : "memory");
}
__device__ __forceinline__ void cl_sta_u32(uint32_t addr, uint32_t v, uint32_t bar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(addr), "r"(v), "r"(bar) : "memory");
}
__device__ __forceinline__ void cl_sta_u4(uint32_t addr, uint4 v, uint32_t bar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
                 ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(bar) : "memory");
}
__device__ __forceinline__ void cl_mbar_init_expect(uint64_t* bar, uint32_t bytes) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(a) : "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void cl_mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(bar);
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(a), "r"(parity) : "memory");
    } while (!done);
}
// asynchronous global -> shared copies (no register, no stall): records the CTA will only need after the streaming phase
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__device__ __forceinline__ float cl_cosine(float dot, float ss_e, float ss_q) {
    // reference: dot / (||e|| * ||q|| + 1e-8), all fp32, IEEE sqrt and divide
    return __fdiv_rn(dot, __fadd_rn(__fmul_rn(__fsqrt_rn(ss_e), __fsqrt_rn(ss_q)), 1e-8f));
}

// what the CTAs of a document push into its first CTA (KCAP = largest k of the instantiation)
template <int KCAP, int MODE>
struct Exchange {
    unsigned long long key[kClMax][KCAP];
    float val[kClMax][KCAP];
    uint4 rec[kClMax][MODE == 2 ? KCAP : 1][2];        // rdv_chunk_rec of each candidate (MODE 2)
    uint4 bbox[kClMax][MODE == 2 ? KCAP : 1][2];       // its chunk_bbox (4 doubles)
};

// VPL > 0: d == 128 * VPL (everything in registers); VPL == 0: any d % 4 == 0.
// NK: a CTA's slice has up to 32 * NK rows (documents of up to NK * cluster size * 32 rows).  MODE 1: score + top-k.
// MODE 2: + gather.
template <int VPL, int NK, int KCAP, int MINB, int MODE>
__global__ void __launch_bounds__(kScoreThreads, MINB)
retrieve_cluster_kernel(const __grid_constant__ ClusterParams p, const __grid_constant__ GatherParams G) {
    constexpr int ROWS = 2;
    constexpr uint32_t kSlotBytes = 8 + 4 + (MODE == 2 ? 64 : 0);     // what a lane pushes: key, value (, record, box)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t rank = cl_rank();
    const int d4 = p.d >> 2;

    extern __shared__ __align__(16) unsigned char smem_dyn[];
    // dynamic (38 KB at k = 32, beyond the static limit together with the rest); same offset in every CTA of the cluster.
    // The copy in a document's FIRST CTA is the one that gets filled.
    Exchange<KCAP, MODE>& s_ex = *reinterpret_cast<Exchange<KCAP, MODE>*>(smem_dyn);
    __shared__ float s_sims[NK * kClTileRows];
    __shared__ GatherSmem S;                               // MODE 2 only (unused objects are dropped by the compiler)
    __shared__ __align__(16) uint4 s_rec[KCAP][2];         // records / boxes of the winners, hit order
    __shared__ __align__(16) uint4 s_bbox[KCAP][2];
    __shared__ __align__(16) uint4 s_tile_rec[NK * kClTileRows][2];    // chunk records of this CTA's rows (cp.async)
    __shared__ __align__(16) uint4 s_tile_bbox[NK * kClTileRows][2];
    __shared__ __align__(16) int s_doc[4];                 // first CTA: doc_page_off[b], prompt_off[b], prompt_off[b + 1]
    __shared__ __align__(8) uint64_t s_bar;                // first CTA of a document: counts the candidates' bytes

    pdl_launch_dependents();
    pdl_wait();                   // embeddings, questions and descriptors come from earlier work in the stream
    cl_stamp(p, 0);
    const rdv_cta_desc c = p.ctas[blockIdx.x];             // one broadcast 32-byte load
    const int n = c.doc_rows, b = c.doc;
    const int nparts = c.nparts, part = c.part;            // nparts == 0: padding CTA of a cluster
    if (nparts > 0 && part == 0 && tid == 0) cl_mbar_init_expect(&s_bar, (uint32_t)nparts * KCAP * kSlotBytes);
    cl_arrive();                  // "this CTA runs, and if it collects candidates its mbarrier is armed" -- waited for below
    if (nparts == 0) return;      // padding: it has arrived, nobody waits for more
    // this CTA's slice of the document: rows [row0, row0 + rows), an even split over the document's CTAs
    const int per = (n + nparts - 1) / nparts;
    const int row0 = part * per;
    const int rows = max(0, min(per, n - row0));
    const float4* __restrict__ Q = reinterpret_cast<const float4*>(p.q) + (size_t)b * d4;
    const float4* __restrict__ E = reinterpret_cast<const float4*>(c.src) + (size_t)row0 * d4;
    float* __restrict__ out = p.sims + c.sims_off + row0;

    // requested now, without holding a register or a warp: what the gather will want after the streaming phase -- the
    // chunk records / bounding boxes of this CTA's rows and three per-document scalars
    if constexpr (MODE == 2) {
        const size_t gc = (size_t)c.sims_off + row0;                       // global chunk number of the slice's first row
        const uint4* g_rec = reinterpret_cast<const uint4*>(G.ds.chunk_rec + gc);
        for (int i = tid; i < 2 * rows; i += kScoreThreads) cp_async16(&s_tile_rec[0][0] + i, g_rec + i);
        if (G.ds.chunk_bbox) {
            const uint4* g_box = reinterpret_cast<const uint4*>(G.ds.chunk_bbox + gc * 4);
            for (int i = tid; i < 2 * rows; i += kScoreThreads) cp_async16(&s_tile_bbox[0][0] + i, g_box + i);
        }
        if (part == 0) {
            if (tid == 128 && G.ds.doc_page_off && G.ds.page_wh) cp_async4(&s_doc[0], G.ds.doc_page_off + b);
            if (tid == 129) cp_async4(&s_doc[1], G.a.prompt_off + b);
            if (tid == 130) cp_async4(&s_doc[2], G.a.prompt_off + b + 1);
        }
    }

    // ---- stream this CTA's rows: a warp owns rows r, r + 1 with r = 2 * warp, + 16, ... ------------------------------
    if constexpr (VPL > 0) {
        // (The question in shared memory instead of 4 * VPL registers per lane -- 40 registers, 6 CTAs per SM -- was
        // measured: the streaming phase went from 5.1 to 9.7 us.  It stays in registers.)
        float4 qv[VPL];
        if (rows > warp * ROWS) {
#pragma unroll
            for (int i = 0; i < VPL; ++i) qv[i] = __ldg(Q + lane + 32 * i);   // in flight together with the rows
        }
        float ss_q = 0.f;
        bool have_q = false;
        for (int r = warp * ROWS; r < rows; r += kScoreWarps * ROWS) {
            float4 ev[ROWS][VPL];
#pragma unroll
            for (int j = 0; j < ROWS; ++j) {
                const bool ok = r + j < rows;
                const float4* src = E + (size_t)(ok ? r + j : 0) * d4 + lane;
#pragma unroll
                for (int i = 0; i < VPL; ++i)
                    ev[j][i] = ok ? ldg_stream(src + 32 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            if (!have_q) {
                float s = 0.f;
#pragma unroll
                for (int i = 0; i < VPL; ++i) {
                    s = fmaf(qv[i].x, qv[i].x, s); s = fmaf(qv[i].y, qv[i].y, s);
                    s = fmaf(qv[i].z, qv[i].z, s); s = fmaf(qv[i].w, qv[i].w, s);
                }
                ss_q = warp_sum(s);
                have_q = true;
            }
            float mine = 0.f;
#pragma unroll
            for (int j = 0; j < ROWS; ++j) {
                float dot = 0.f, ss = 0.f;
#pragma unroll
                for (int i = 0; i < VPL; ++i) {
                    dot = fmaf(ev[j][i].x, qv[i].x, dot); ss = fmaf(ev[j][i].x, ev[j][i].x, ss);
                    dot = fmaf(ev[j][i].y, qv[i].y, dot); ss = fmaf(ev[j][i].y, ev[j][i].y, ss);
                    dot = fmaf(ev[j][i].z, qv[i].z, dot); ss = fmaf(ev[j][i].z, ev[j][i].z, ss);
                    dot = fmaf(ev[j][i].w, qv[i].w, dot); ss = fmaf(ev[j][i].w, ev[j][i].w, ss);
                }
                const float sim = cl_cosine(warp_sum(dot), warp_sum(ss), ss_q);
                if (lane == j) mine = sim;
            }
            if (lane < ROWS && r + lane < rows) {
                out[r + lane] = mine;
                s_sims[r + lane] = mine;
            }
        }
    } else {
        float ss_q = 0.f;
        if (rows > 0) {
            for (int i = lane; i < d4; i += 32) {
                const float4 v = __ldg(Q + i);
                ss_q = fmaf(v.x, v.x, ss_q); ss_q = fmaf(v.y, v.y, ss_q);
                ss_q = fmaf(v.z, v.z, ss_q); ss_q = fmaf(v.w, v.w, ss_q);
            }
            ss_q = warp_sum(ss_q);
        }
        for (int r = warp * ROWS; r < rows; r += kScoreWarps * ROWS) {
            float dot[ROWS], ss[ROWS];
#pragma unroll
            for (int j = 0; j < ROWS; ++j) { dot[j] = 0.f; ss[j] = 0.f; }
#pragma unroll 2
            for (int i = lane; i < d4; i += 32) {
                const float4 qv = __ldg(Q + i);
#pragma unroll
                for (int j = 0; j < ROWS; ++j) {
                    const bool ok = r + j < rows;
                    const float4 e = ok ? ldg_stream(E + (size_t)(r + j) * d4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
                    dot[j] = fmaf(e.x, qv.x, dot[j]); ss[j] = fmaf(e.x, e.x, ss[j]);
                    dot[j] = fmaf(e.y, qv.y, dot[j]); ss[j] = fmaf(e.y, e.y, ss[j]);
                    dot[j] = fmaf(e.z, qv.z, dot[j]); ss[j] = fmaf(e.z, e.z, ss[j]);
                    dot[j] = fmaf(e.w, qv.w, dot[j]); ss[j] = fmaf(e.w, e.w, ss[j]);
                }
            }
            float mine = 0.f;
#pragma unroll
            for (int j = 0; j < ROWS; ++j) {
                const float sim = cl_cosine(warp_sum(dot[j]), warp_sum(ss[j]), ss_q);
                if (lane == j) mine = sim;
            }
            if (lane < ROWS && r + lane < rows) {
                out[r + lane] = mine;
                s_sims[r + lane] = mine;
            }
        }
    }
    if constexpr (MODE == 2) cp_async_wait_all();
    cl_stamp(p, 1);               // rows streamed
    __syncthreads();
    cl_wait();                    // every CTA of the cluster runs; the mbarriers of the collecting CTAs are armed
    cl_stamp(p, 2);

    // ---- level 1: this CTA's k best, pushed into the document's first CTA -------------------------------------------
    const int k = p.k;
    if (warp == 0) {
        unsigned long long key[NK];
#pragma unroll
        for (int j = 0; j < NK; ++j) {
            const int li = lane + 32 * j;                                          // row of this CTA's slice
            key[j] = li < rows ? pack_key(s_sims[li], (uint32_t)(row0 + li)) : 0ull;   // keys carry the row of the DOCUMENT
        }
        const unsigned long long mine = warp_rounds<NK>(key, min(k, rows), lane);  // lane r: this CTA's r-th best (0: none)
        const int mine_local = mine ? (int)key_index(mine) - row0 : 0;             // its row within this CTA's slice
        const float mine_val = mine ? s_sims[mine_local] : 0.f;
        if (lane < KCAP) {
            // every lane pushes the same number of bytes whether it holds a candidate or not: the receiver counts bytes
            const uint32_t first = rank - (uint32_t)part;                          // the document's first CTA
            const uint32_t bar = cl_map(&s_bar, first);
            cl_sta_u64(cl_map(&s_ex.key[part][lane], first), lane < k ? mine : 0ull, bar);
            cl_sta_u32(cl_map(&s_ex.val[part][lane], first), __float_as_uint(mine_val), bar);
            if constexpr (MODE == 2) {
                const uint32_t a_rec = cl_map(&s_ex.rec[part][lane][0], first);
                const uint32_t a_box = cl_map(&s_ex.bbox[part][lane][0], first);
                cl_sta_u4(a_rec, s_tile_rec[mine_local][0], bar); cl_sta_u4(a_rec + 16, s_tile_rec[mine_local][1], bar);
                cl_sta_u4(a_box, s_tile_bbox[mine_local][0], bar); cl_sta_u4(a_box + 16, s_tile_bbox[mine_local][1], bar);
            }
        }
    }
    cl_stamp(p, 3);               // candidates pushed
    if (part != 0) return;        // nothing left to do: the stores complete on their own
    cl_mbar_wait(&s_bar, 0);      // all candidates of this CTA's document have landed
    cl_stamp(p, 4);

    // ---- level 2 (first CTA of the document): merge nparts x KCAP slots ---------------------------------------------
    const int k_min = n < k ? n : k;
    if (warp == 0) {
        constexpr int NS = kClMax * KCAP / 32;
        const unsigned long long* flat = &s_ex.key[0][0];
        const int used = nparts * KCAP;                                            // slots beyond were never written
        unsigned long long key[NS];
#pragma unroll
        for (int j = 0; j < NS; ++j) key[j] = lane + 32 * j < used ? flat[lane + 32 * j] : 0ull;
        unsigned long long mine = 0;
        int mine_slot = 0;
        for (int r = 0; r < k_min; ++r) {
            unsigned long long best = key[0];
#pragma unroll
            for (int j = 1; j < NS; ++j) best = key[j] > best ? key[j] : best;
            const unsigned long long win = warp_max_key(best);
            int slot = -1;
#pragma unroll
            for (int j = 0; j < NS; ++j) {
                if (key[j] == win) { slot = lane + 32 * j; key[j] = 0ull; }
            }
            const unsigned who = __ballot_sync(0xffffffffu, slot >= 0);
            slot = __shfl_sync(0xffffffffu, slot, who ? __ffs(who) - 1 : 0);
            if (lane == r) { mine = win; mine_slot = slot < 0 ? 0 : slot; }
        }
        if (lane < k) {
            const size_t o = (size_t)b * k + lane;
            if (lane < k_min) {
                const uint32_t idx = key_index(mine);
                p.topk_idx[o] = (int32_t)idx;
                p.topk_val[o] = (&s_ex.val[0][0])[mine_slot];
                if constexpr (MODE == 2) {
                    S.hit[lane] = (int)idx;
                    const uint4* rs = &s_ex.rec[0][0][0] + 2 * mine_slot;
                    const uint4* bs = &s_ex.bbox[0][0][0] + 2 * mine_slot;
                    s_rec[lane][0] = rs[0]; s_rec[lane][1] = rs[1];
                    s_bbox[lane][0] = bs[0]; s_bbox[lane][1] = bs[1];
                }
            } else {
                p.topk_idx[o] = -1;
                p.topk_val[o] = -INFINITY;
            }
        }
        if (lane == 0) p.topk_cnt[b] = k_min;
    }
    cl_stamp(p, 5);               // merged
    if constexpr (MODE == 2) {
        __syncthreads();
        gather_document<false>(G.ds, G.a, b, k_min, S, (int64_t)c.sims_off, reinterpret_cast<const rdv_chunk_rec*>(&s_rec[0][0]),
                               G.ds.chunk_bbox ? reinterpret_cast<const double*>(&s_bbox[0][0]) : nullptr, s_doc);
        __syncthreads();
        cl_stamp(p, 6);           // gathered
    }
}

// ---- launch plumbing ----------------------------------------------------------------------------------------------
struct ClusterLaunch {
    int cluster;                 // CTAs per cluster
    int max_rows;
    cudaStream_t stream;
    int* query;                  // non-null: do not launch; query[0] = largest cluster size the device takes for this
                                 // variant, query[1] = clusters of min(that, 16) CTAs that can be resident at once
};

// The kernel is a template ARGUMENT so that the once-per-device flag below exists per kernel instantiation (all of them
// share one signature: a flag keyed by the function-pointer type would opt in the first variant only).
template <void (*kernel)(ClusterParams, GatherParams)>
static cudaError_t launch_clustered(const ClusterLaunch& L, size_t smem, dim3 grid, const ClusterParams& p, const GatherParams& G) {
    static PerDeviceOnce once;                  // one bit per device
    int dev;
    if (once.pending(&dev)) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        once.mark(dev);
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(kScoreThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = L.stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = L.cluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (pdl_mask() & kPdlStream) ? 2 : 1;
    if (L.query) {
        cfg.gridDim = dim3(kClMax * 8);
        cfg.numAttrs = 0;                       // the query answers for "any cluster size"
        cudaError_t e = cudaOccupancyMaxPotentialClusterSize(&L.query[0], kernel, &cfg);
        if (e != cudaSuccess) return e;
        attr[0].val.clusterDim.x = L.query[0] >= kClMax ? kClMax : 8;
        cfg.numAttrs = 1;
        return cudaOccupancyMaxActiveClusters(&L.query[1], kernel, &cfg);
    }
    return cudaLaunchKernelEx(&cfg, kernel, p, G);
}

template <int VPL, int NK, int KCAP, int MINB>
static int launch_cluster_v(const ClusterParams& p, const GatherParams* G, const ClusterLaunch& L) {
    static const GatherParams kNoGather = {};
    const dim3 grid(p.n_ctas);
    cudaError_t e = G ? launch_clustered<retrieve_cluster_kernel<VPL, NK, KCAP, MINB, 2>>(L, sizeof(Exchange<KCAP, 2>), grid, p, *G)
                      : launch_clustered<retrieve_cluster_kernel<VPL, NK, KCAP, MINB, 1>>(L, sizeof(Exchange<KCAP, 1>), grid, p, kNoGather);
    if (e != cudaSuccess) return cuda_fail(e, "retrieve_cluster_kernel");
    return RDV_OK;
}

template <int VPL, int MINB>
static int launch_cluster_d(const ClusterParams& p, const GatherParams* G, const ClusterLaunch& L) {
    const bool short_docs = L.max_rows <= 2 * L.cluster * kClTileRows;   // slices of at most 64 rows: 2 keys per lane
    if (p.k <= 8)
        return short_docs ? launch_cluster_v<VPL, 2, 8, MINB>(p, G, L) : launch_cluster_v<VPL, kClMaxLocalTiles, 8, MINB>(p, G, L);
    if (p.k <= 16)
        return short_docs ? launch_cluster_v<VPL, 2, 16, MINB>(p, G, L) : launch_cluster_v<VPL, kClMaxLocalTiles, 16, MINB>(p, G, L);
    return short_docs ? launch_cluster_v<VPL, 2, 32, MINB>(p, G, L) : launch_cluster_v<VPL, kClMaxLocalTiles, 32, MINB>(p, G, L);
}

static int launch_cluster(const ClusterParams& p, const GatherParams* G, const ClusterLaunch& L) {
    switch (p.d) {
        case 128:  return launch_cluster_d<1, 5>(p, G, L);
        case 256:  return launch_cluster_d<2, 5>(p, G, L);
        case 384:  return launch_cluster_d<3, 5>(p, G, L);
        case 512:  return launch_cluster_d<4, 4>(p, G, L);
        case 768:  return launch_cluster_d<6, 3>(p, G, L);
        case 1024: return launch_cluster_d<8, 3>(p, G, L);
        default:   return launch_cluster_d<0, 4>(p, G, L);
    }
}

// slots (CTAs) a document takes in its cluster: slices of at most slice_rows rows, at most the whole cluster
static inline int doc_slots(int64_t rows, int cluster, int slice_rows) {
    const int64_t t = (rows + slice_rows - 1) / slice_rows;
    return (int)(t < 1 ? 1 : (t > cluster ? cluster : t));
}
static inline bool slice_ok(int s) { return s >= kClTileRows && s <= 2 * kClTileRows && (s & 7) == 0; }

}  // namespace rdv

using namespace rdv;

// What the device takes for the kernel variant a batch would run, asked of the occupancy calculator once per (device,
// variant): the cluster size -- 16 CTAs (non-portable size) where it places 16 of them in one GPC, else the portable 8 --
// and how many such clusters can be resident at once.
static void cluster_capacity(int32_t d, int32_t k, int32_t max_rows, int32_t with_gather, int* size, int* active) {
    static const int forced = [] { const char* v = getenv("RDV_CLUSTER_SIZE"); return v ? atoi(v) : 0; }();
    *size = 8; *active = 0;                              // 0: unknown (no device: host-only callers)
    if (d < 4 || (d & 3) || k < 1 || k > 32 || max_rows < 0) return;
    static std::atomic<int> cache[64][7][3][2][2];       // device, width case, k case, slice depth, gather: 0 unknown
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) { if (forced == 16) *size = 16; return; }
    const int wcase = d == 128 ? 0 : d == 256 ? 1 : d == 384 ? 2 : d == 512 ? 3 : d == 768 ? 4 : d == 1024 ? 5 : 6;
    const int kcase = k <= 8 ? 0 : k <= 16 ? 1 : 2;
    const int deep = max_rows > 2 * kClMax * kClTileRows ? 1 : 0;
    std::atomic<int>& slot = cache[dev][wcase][kcase][deep][with_gather ? 1 : 0];
    int packed = slot.load(std::memory_order_relaxed);   // size | active << 8
    if (packed == 0) {
        ClusterParams p = {};
        p.d = d; p.k = k; p.n_ctas = kClMax;
        static const GatherParams kNoGather = {};
        int q[2] = {0, 0};
        ClusterLaunch L = {kClMax, deep ? 2 * kClMax * kClTileRows + 1 : 0, nullptr, q};
        const int rc = launch_cluster(p, with_gather ? &kNoGather : nullptr, L);
        const int sz = (rc == RDV_OK && q[0] >= kClMax) ? kClMax : 8;
        packed = sz | ((rc == RDV_OK && q[1] > 0 ? q[1] : 0) << 8);
        slot.store(packed, std::memory_order_relaxed);
    }
    *size = packed & 0xff;
    *active = packed >> 8;
    if (forced == 8 || forced == 16) {                   // measurement knob; the residency figure scales with the size
        if (forced != *size) *active = forced == 8 ? *active * 2 : *active / 2;
        *size = forced;
    }
}

// Cluster size and slice height for a batch: the smallest slices (most parallelism per document) for which all clusters
// of the batch are resident at once -- a second wave of clusters starts only when whole clusters retire, and costs the step
// a second latency chain (measured at C2: 44 clusters of 16 against 38 resident: 12.4 us).
extern "C" int rdv_cluster_plan(const int64_t* rows, int32_t B, int32_t d, int32_t k, int32_t with_gather, int32_t* cluster,
                                int32_t* slice_rows, int64_t* n_ctas) {
    RDV_REQUIRE(cluster && slice_rows && n_ctas && (rows || B == 0) && B >= 0, RDV_E_INVALID, "cluster_plan: bad argument");
    int64_t mx = 0;
    for (int32_t b = 0; b < B; ++b) {
        RDV_REQUIRE(rows[b] >= 0, RDV_E_INVALID, "cluster_plan: negative size");
        if (rows[b] > mx) mx = rows[b];
    }
    int size = 8, active = 0;
    cluster_capacity(d, k, (int32_t)(mx > (1 << 30) ? (1 << 30) : mx), with_gather, &size, &active);
    static const int forced_slice = [] { const char* v = getenv("RDV_CLUSTER_SLICE"); return v ? atoi(v) : 0; }();
    *cluster = size;
    *slice_rows = kClTileRows;
    *n_ctas = 0;
    if (B == 0 || mx > kClMaxLocalTiles * size * kClTileRows) return RDV_OK;       // n_ctas == 0: outside the limits
    for (int s = kClTileRows; s <= 2 * kClTileRows; s += 8) {
        if (forced_slice && slice_ok(forced_slice)) s = forced_slice;
        *slice_rows = s;
        *n_ctas = rdv_cluster_table_size(rows, B, size, s);
        if (forced_slice || active <= 0 || *n_ctas <= (int64_t)active * size) break;
    }
    return RDV_OK;
}

extern "C" int32_t rdv_cluster_max_rows(int32_t cluster) { return kClMaxLocalTiles * (cluster == 8 ? 8 : kClMax) * kClTileRows; }
extern "C" int32_t rdv_cluster_max_k(void) { return 32; }

// Measured on B200 at C2 (64 documents of <= 600 chunks, 384-d, one dependent chain, profiles/r2_step_probe.md): score +
// top-k in one launch 10.3-10.7 us, the whole step (+ gather) 14.4-16.3 us, against 6.9 + 5.9 = 13.3 us for the streaming kernel
// followed by the select + gather kernel.  The cluster path loses the step: with 576 CTAs of 3 load rounds in flight instead of
// 684 of 2 its streaming phase alone takes 6.9 us, and the hand-over (2.3 us) and the gather by one CTA per document (3.6 us)
// come on top with nothing left to overlap them.  So the plan answers 1 only for batches of at most 1 MB (the one-call
// host path, C1: one launch less in a call that is all fixed cost) unless RDV_CLUSTER=1 / 0 forces it; the kernels stay
// selectable (tests run them; `cluster=True` in the Python layer).
extern "C" int rdv_retrieve_plan(int64_t total_rows, int32_t max_rows, int32_t B, int32_t d, int32_t k, int32_t* use_cluster) {
    RDV_REQUIRE(use_cluster, RDV_E_INVALID, "retrieve_plan: null output");
    RDV_REQUIRE(total_rows >= 0 && max_rows >= 0 && B >= 0 && d >= 4, RDV_E_INVALID, "retrieve_plan: bad sizes");
    static const int forced = [] { const char* v = getenv("RDV_CLUSTER"); return v ? atoi(v) : -1; }();
    const bool fits = B >= 1 && k >= 1 && k <= rdv_cluster_max_k() && max_rows <= rdv_cluster_max_rows(8) && (d & 3) == 0 &&
                      total_rows < (1ll << 31);
    const bool tiny = (total_rows + B) * (int64_t)d * 4 <= (1ll << 20);
    *use_cluster = fits && (forced >= 0 ? forced != 0 : tiny);
    return RDV_OK;
}

// Best-fit decreasing: documents sorted by the slots they need (stable), each placed into the open cluster with the
// least free space that still takes it; clusters are padded to `cluster` CTAs with barrier-only descriptors.
static int pack_documents(const int64_t* rows, int32_t B, int cluster, int slice_rows, std::vector<int32_t>* doc_cluster,
                          std::vector<int32_t>* doc_first, int64_t* n_clusters) {
    std::vector<int32_t> order(B);
    for (int32_t b = 0; b < B; ++b) order[b] = b;
    std::stable_sort(order.begin(), order.end(), [&](int32_t x, int32_t y) {
        return doc_slots(rows[x], cluster, slice_rows) > doc_slots(rows[y], cluster, slice_rows); });
    std::vector<std::vector<int32_t>> open(cluster + 1);          // open[f]: clusters with f free slots
    std::vector<int32_t> used;                                    // slots used per cluster
    doc_cluster->assign(B, 0);
    doc_first->assign(B, 0);
    for (int32_t b : order) {
        const int s = doc_slots(rows[b], cluster, slice_rows);
        int f = s;
        while (f <= cluster && open[f].empty()) ++f;
        int32_t c;
        if (f > cluster) {
            c = (int32_t)used.size();
            used.push_back(0);
            f = cluster;
        } else {
            c = open[f].back();
            open[f].pop_back();
        }
        (*doc_cluster)[b] = c;
        (*doc_first)[b] = used[c];
        used[c] += s;
        if (f - s > 0) open[f - s].push_back(c);
    }
    *n_clusters = (int64_t)used.size();
    return RDV_OK;
}

extern "C" int64_t rdv_cluster_table_size(const int64_t* rows, int32_t B, int32_t cluster, int32_t slice_rows) {
    if (!rows || B < 0 || (cluster != 8 && cluster != kClMax) || !slice_ok(slice_rows)) return -1;
    for (int32_t b = 0; b < B; ++b)
        if (rows[b] < 0) return -1;
    if (B == 0) return 0;
    std::vector<int32_t> dc, df;
    int64_t n_clusters = 0;
    pack_documents(rows, B, cluster, slice_rows, &dc, &df, &n_clusters);
    return n_clusters * cluster;
}

extern "C" int rdv_build_cluster_table(const void* const* d_docs, const int64_t* rows, int32_t B, int32_t d, int32_t cluster,
                                       int32_t slice_rows, rdv_cta_desc* h_ctas, int64_t n_ctas) {
    RDV_REQUIRE(B >= 0 && d >= 4 && (d & 3) == 0, RDV_E_INVALID, "build_cluster_table: bad B / d");
    RDV_REQUIRE(cluster == 8 || cluster == kClMax, RDV_E_INVALID, "build_cluster_table: cluster size %d is not 8 or %d", cluster, kClMax);
    RDV_REQUIRE(slice_ok(slice_rows), RDV_E_INVALID, "build_cluster_table: slice_rows=%d is not a multiple of 8 in [32, 64]", slice_rows);
    if (B == 0) return RDV_OK;
    RDV_REQUIRE(d_docs && rows && h_ctas, RDV_E_INVALID, "build_cluster_table: null pointer");
    std::vector<int32_t> dc, df;
    int64_t n_clusters = 0;
    for (int32_t b = 0; b < B; ++b) {
        RDV_REQUIRE(rows[b] >= 0 && rows[b] <= rdv_cluster_max_rows(cluster), RDV_E_LIMIT,
                    "build_cluster_table: document %d has %lld rows (clusters of %d take up to %d)", b, (long long)rows[b], cluster,
                    rdv_cluster_max_rows(cluster));
        RDV_REQUIRE(rows[b] == 0 || (d_docs[b] && aligned16(d_docs[b])), RDV_E_ALIGN,
                    "build_cluster_table: document %d is null / not 16-byte aligned", b);
    }
    pack_documents(rows, B, cluster, slice_rows, &dc, &df, &n_clusters);
    RDV_REQUIRE(n_clusters * cluster == n_ctas, RDV_E_INVALID, "build_cluster_table: %lld descriptors provided, %lld needed",
                (long long)n_ctas, (long long)(n_clusters * cluster));
    memset(h_ctas, 0, (size_t)n_ctas * sizeof(rdv_cta_desc));                    // nparts == 0: padding
    int64_t off = 0;
    for (int32_t b = 0; b < B; ++b) {
        RDV_REQUIRE(off + rows[b] < (1ll << 31), RDV_E_LIMIT, "build_cluster_table: more than 2^31 rows");
        const int s = doc_slots(rows[b], cluster, slice_rows);
        for (int part = 0; part < s; ++part) {
            rdv_cta_desc& c = h_ctas[(int64_t)dc[b] * cluster + df[b] + part];
            c.src = d_docs[b];
            c.sims_off = (int32_t)off;
            c.doc_rows = (int32_t)rows[b];
            c.doc = b;
            c.part = (int16_t)part;
            c.nparts = (int16_t)s;
        }
        off += rows[b];
    }
    return RDV_OK;
}

static std::atomic<unsigned long long*> g_trace{nullptr};

// Diagnostics: while d_trace is non-null every CTA of the cluster kernels writes %globaltimer (ns) stamps to
// d_trace[cta * 8 + s]: s = 0 started, 1 rows streamed, 2 cluster complete, 3 candidates pushed, 4 (merging CTA) candidates
// here, 5 merged, 6 gathered.  The buffer must hold 8 x n_ctas entries; pass NULL to switch it off.
extern "C" int rdv_debug_trace(void* d_trace) {
    g_trace.store(static_cast<unsigned long long*>(d_trace), std::memory_order_relaxed);
    return RDV_OK;
}

static int fill_cluster_params(const char* who, ClusterParams* p, const rdv_cta_desc* d_ctas, int64_t n_ctas, int32_t cluster,
                               const float* d_q, int32_t B, int32_t d, int32_t k, int32_t max_rows, float* d_sims,
                               int32_t* d_topk_idx, float* d_topk_val, int32_t* d_topk_cnt) {
    RDV_REQUIRE(B >= 1, RDV_E_INVALID, "%s: B=%d", who, B);
    RDV_REQUIRE(d >= 4 && d <= 8192 && (d & 3) == 0, RDV_E_INVALID, "%s: d=%d must be a multiple of 4 in [4, 8192]", who, d);
    RDV_REQUIRE(k >= 1 && k <= rdv_cluster_max_k(), RDV_E_LIMIT, "%s: k=%d outside [1, %d]", who, k, rdv_cluster_max_k());
    RDV_REQUIRE(cluster == 8 || cluster == kClMax, RDV_E_INVALID, "%s: cluster size %d is not 8 or %d", who, cluster, kClMax);
    RDV_REQUIRE(max_rows >= 0 && max_rows <= rdv_cluster_max_rows(cluster), RDV_E_LIMIT,
                "%s: documents of up to %d rows exceed the cluster kernel's %d (use rdv_score_topk_f32)", who, max_rows,
                rdv_cluster_max_rows(cluster));
    RDV_REQUIRE(n_ctas >= 1 && n_ctas < (1ll << 31) && n_ctas % cluster == 0, RDV_E_INVALID,
                "%s: n_ctas=%lld is not a positive multiple of the cluster size %d", who, (long long)n_ctas, cluster);
    RDV_REQUIRE(d_q && d_ctas && d_topk_idx && d_topk_val && d_topk_cnt, RDV_E_INVALID, "%s: null pointer", who);
    RDV_REQUIRE(aligned16(d_q) && aligned16(d_ctas), RDV_E_ALIGN, "%s: q / descriptors not 16-byte aligned", who);
    memset(p, 0, sizeof(*p));
    p->ctas = d_ctas; p->q = d_q; p->n_ctas = (int32_t)n_ctas; p->d = d; p->k = k;
    p->sims = d_sims; p->topk_idx = d_topk_idx; p->topk_val = d_topk_val; p->topk_cnt = d_topk_cnt;
    p->trace = g_trace.load(std::memory_order_relaxed);
    return RDV_OK;
}

extern "C" int rdv_score_topk_cluster_f32(const rdv_cta_desc* d_ctas, int64_t n_ctas, int32_t cluster, const float* d_q, int32_t B,
                                          int32_t d, int32_t k, int32_t max_rows, float* d_sims, int32_t* d_topk_idx,
                                          float* d_topk_val, int32_t* d_topk_cnt, void* stream) {
    if (B == 0) return RDV_OK;
    ClusterParams p;
    int rc = fill_cluster_params("score_topk_cluster_f32", &p, d_ctas, n_ctas, cluster, d_q, B, d, k, max_rows, d_sims,
                                 d_topk_idx, d_topk_val, d_topk_cnt);
    if (rc) return rc;
    const ClusterLaunch L = {cluster, max_rows, static_cast<cudaStream_t>(stream), nullptr};
    return launch_cluster(p, nullptr, L);
}

extern "C" int rdv_retrieve_vt5_f32(const rdv_cta_desc* d_ctas, int64_t n_ctas, int32_t cluster, const float* d_q, int32_t d,
                                    int32_t max_rows, float* d_sims, const rdv_docstore* ds, const rdv_gather_args* args,
                                    void* stream) {
    RDV_REQUIRE(ds && args, RDV_E_INVALID, "retrieve_vt5_f32: null struct");
    if (ds->B == 0) return RDV_OK;
    int rc = rdv_gather_check_args(ds, args);
    if (rc) return rc;
    RDV_REQUIRE(args->include_surroundings == 0 && !args->emit_order, RDV_E_INVALID,
                "retrieve_vt5_f32: include_surroundings / emit_order need rdv_score_f32 + rdv_gather_vt5_inputs");
    RDV_REQUIRE(args->topk_val, RDV_E_INVALID, "retrieve_vt5_f32: null topk_val");
    ClusterParams p;
    rc = fill_cluster_params("retrieve_vt5_f32", &p, d_ctas, n_ctas, cluster, d_q, ds->B, d, args->k, max_rows, d_sims,
                             args->topk_idx, args->topk_val, args->topk_cnt);
    if (rc) return rc;
    GatherParams G;
    G.ds = *ds;
    G.a = *args;
    G.a.sims = nullptr;
    const ClusterLaunch L = {cluster, max_rows, static_cast<cudaStream_t>(stream), nullptr};
    return launch_cluster(p, &G, L);
}
