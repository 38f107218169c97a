// bf16 tensor-core scoring for the two dense contractions of the path (sm_100a: tcgen05 + TMEM + TMA).
//
//   * corpus mode (BASELINE.json configs[4]): Q questions x N chunk embeddings, cosine + per-question
//     top-k with the top-k folded into the accumulator epilogue -- the (Q x N) score matrix is never
//     written.  Scoring formula: Retriever._get_similarities (reference src/_modules.py:1990-1993) applied
//     to every (question, chunk) pair; the reference itself only ever scores one question per document.
//   * MaxSim mode (fast mode of late_interaction, reference src/utils.py:442-458): operands are the
//     L2-normalised bf16 copies of Q and P; epilogue = running row maximum over a strip's tokens.
//
// One GEMM main loop serves both:  D[128 x 256] (fp32, TMEM) += A[128 x K] * B[256 x K]^T, both K-major.
//   A rows = questions / question tokens  -> TMEM lanes  -> one epilogue THREAD per row, so the per-row
//     reduction over B rows (top-k / max) is register-resident and needs no cross-thread traffic;
//   B rows = corpus chunks / strip tokens -> TMEM columns.
// Two shapes of the same kernel (template NCTA): one CTA per 128 x 256 tile, or a CTA PAIR (cluster of 2,
// tcgen05 cta_group::2) per 256 x 256 tile -- each CTA keeps its own 128 A rows and half of the B rows in
// shared memory, CTA 0 issues the MMAs for both, commits are multicast to both CTAs' barriers.
// Warp roles (320 threads): warp 0 = TMA producer (cp.async.bulk.tensor, 128B swizzle, 4-stage mbarrier
// ring of 48 KB stages), warp 1 = TMEM allocator + single-thread tcgen05.mma issuer (M=128, N=256, K=16,
// kind::f16 bf16 x bf16 -> fp32), warps 2-9 = epilogue (tcgen05.ld 32x32b.x32; two warps per TMEM lane
// quarter, one column half each).  The accumulator is double-buffered in TMEM (2 x 256 columns) so the epilogue of tile i overlaps
// the MMAs of tile i+1.  Persistent: one block per SM loops over work items.
#include "tc_common.cuh"

#include <cuda_bf16.h>
#include <stdlib.h>

namespace rdv {
namespace tc {

constexpr int BM = 128, BN = 256, BK = 64;
constexpr int kABytes = BM * BK * 2;
constexpr int kThreads = 320;                // TMA warp + MMA warp + 8 epilogue warps
// NCTA = 1: one CTA per tile, stage = A (16 KB) + B (32 KB), 4 stages.
// NCTA = 2: a CTA pair (cluster of 2, tcgen05 cta_group::2) shares each B tile: D is 256 x 256, every CTA holds its
//           own 128 A rows and HALF of the B rows (16 KB), so a stage is 32 KB and 6 fit -- a third less
//           shared-memory fill and read per SM and a deeper TMA look-ahead.
template <int NCTA> struct Cfg {
    static constexpr int kBRows = BN / NCTA;                 // B rows this CTA loads per stage
    static constexpr int kBBytes = kBRows * BK * 2;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kStages = NCTA == 1 ? 4 : 6;
};
constexpr int kTopK = 16;                 // register-resident candidates per question and work item
constexpr uint32_t kTmemCols = 512;       // 2 accumulators x 256 fp32 columns

enum Mode { kCorpus = 0, kMaxSim = 1 };

struct Params {
    int mode;
    int n_a;                // A tiles (question blocks / question-token blocks)
    int n_groups;           // corpus: row chunks; maxsim: strips
    int tiles_total;        // corpus: total B tiles;  maxsim: B tiles per strip
    int k_blocks;           // ceil(K / 64)
    int a_rows, b_rows;     // valid A rows (Q / Lq) and B rows (N_local / Lp)
    // corpus epilogue
    const float* inv_norm;  // (N_local) 1 / max(||e||, tiny)
    float* part_val;        // (n_groups, n_a*128, kTopK)
    int32_t* part_idx;
    // maxsim epilogue
    float* partial;         // (n_groups = strips, n_a)
};

// kind::f16 instruction descriptor: D=f32, A=B=bf16, both K-major, M=128*NCTA, N=256
template <int NCTA> struct IdescBf16 {
    static constexpr uint32_t value = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((BM * NCTA) >> 4) << 24);
};

struct Item { int a_tile, group, t0, t1, z; };
// work item -> (A tile, row chunk).  With CTA pairs an item covers NCTA consecutive A tiles, one per CTA rank.
template <int NCTA>
__device__ __forceinline__ Item decode_item(const Params& p, int item, int rank) {
    Item it;
    const int n_ap = (p.n_a + NCTA - 1) / NCTA;
    it.group = item / n_ap;
    it.a_tile = (item - it.group * n_ap) * NCTA + rank;
    if (p.mode == kCorpus) {
        it.t0 = (int)((long long)it.group * p.tiles_total / p.n_groups);
        it.t1 = (int)((long long)(it.group + 1) * p.tiles_total / p.n_groups);
        it.z = 0;
    } else {
        it.t0 = 0; it.t1 = p.tiles_total; it.z = it.group;
    }
    return it;
}

template <int NCTA>
__global__ void __launch_bounds__(kThreads, 1)
tc_score_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const Params p) {
    using C = Cfg<NCTA>;
    constexpr int kStages = C::kStages, kStageBytes = C::kStageBytes;
    constexpr uint32_t kIdesc = IdescBf16<NCTA>::value;
    extern __shared__ __align__(1024) unsigned char smem_unaligned[];
    // 128-byte swizzle atoms repeat every 1024 bytes: the operand ring must start on a 1024-byte boundary
    unsigned char* smem = smem_unaligned + ((1024u - (s32(smem_unaligned) & 1023u)) & 1023u);
    __shared__ __align__(8) uint64_t s_full[kStages], s_empty[kStages], s_tfull[2], s_tempty[2];
    __shared__ uint32_t s_tmem_base;
    __shared__ __align__(16) float s_inv[2][BN];
    __shared__ float s_sum[4];
    // per-epilogue-thread scratch row of 16 scores, slot-major so a warp's accesses are conflict-free
    float* q_val = reinterpret_cast<float*>(smem + (size_t)kStages * kStageBytes);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rank = NCTA == 2 ? (int)cluster_ctarank() : 0;                 // CTA within its pair
    const int n_items = ((p.n_a + NCTA - 1) / NCTA) * p.n_groups;
    const int first_item = blockIdx.x / NCTA, item_step = gridDim.x / NCTA;  // per pair

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(&s_full[s], 1); mbar_init(&s_empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&s_tfull[s], 1); mbar_init(&s_tempty[s], 8 * NCTA); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        if constexpr (NCTA == 1) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&s_tmem_base)), "r"(kTmemCols));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
        } else {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&s_tmem_base)), "r"(kTmemCols));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
        }
    }
    tc_fence_before();
    __syncthreads();
    if constexpr (NCTA == 2) cluster_sync_all();      // the peer's barriers are initialised before anyone signals them
    tc_fence_after();
    const uint32_t tmem_base = s_tmem_base;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int item = first_item; item < n_items; item += item_step) {
                const Item it = decode_item<NCTA>(p, item, rank);
                for (int t = it.t0; t < it.t1; ++t) {
                    for (int kb = 0; kb < p.k_blocks; ++kb) {
                        mbar_wait(&s_empty[stage], phase ^ 1);
                        unsigned char* sa = smem + (size_t)stage * kStageBytes;
                        if constexpr (NCTA == 1) {
                            mbar_expect_tx(&s_full[stage], kStageBytes);
                            tma_load_3d(sa, &map_a, &s_full[stage], kb * BK, it.a_tile * BM, 0);
                            tma_load_3d(sa + kABytes, &map_b, &s_full[stage], kb * BK, t * BN, it.z);
                        } else {
                            // both CTAs load their own A rows and their half of the B rows; all bytes complete on
                            // CTA 0's barrier, which CTA 0 arms for the pair
                            if (rank == 0) mbar_expect_tx(&s_full[stage], NCTA * kStageBytes);
                            tma_load_3d_pair(sa, &map_a, &s_full[stage], kb * BK, it.a_tile * BM, 0);
                            tma_load_3d_pair(sa + kABytes, &map_b, &s_full[stage], kb * BK, t * BN + rank * C::kBRows, it.z);
                        }
                        if (++stage == kStages) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (one thread) =====================
        if (lane == 0 && rank == 0) {                                    // CTA 0 issues for the pair
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            for (int item = first_item; item < n_items; item += item_step) {
                const Item it = decode_item<NCTA>(p, item, rank);
                for (int t = it.t0; t < it.t1; ++t) {
                    mbar_wait(&s_tempty[acc], acc_phase ^ 1);          // epilogue(s) have drained this accumulator
                    tc_fence_after();
                    const uint32_t tmem_d = tmem_base + (uint32_t)acc * BN;
                    for (int kb = 0; kb < p.k_blocks; ++kb) {
                        mbar_wait(&s_full[stage], phase);               // TMA bytes have landed
                        tc_fence_after();
                        const uint32_t a_addr = s32(smem + (size_t)stage * kStageBytes);
                        const uint64_t da = smem_desc(a_addr), db = smem_desc(a_addr + kABytes);
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k) {               // 32 bytes (>>4 = 2) per K=16 step
                            if constexpr (NCTA == 1)
                                tc_mma_bf16(tmem_d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), kIdesc, (kb | k) ? 1u : 0u);
                            else
                                tc_mma_bf16_pair(tmem_d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), kIdesc, (kb | k) ? 1u : 0u);
                        }
                        // frees the smem stage (in both CTAs) when the MMAs retire
                        if constexpr (NCTA == 1) tc_commit(&s_empty[stage]); else tc_commit_pair(&s_empty[stage]);
                        if (++stage == kStages) { stage = 0; phase ^= 1; }
                    }
                    if constexpr (NCTA == 1) tc_commit(&s_tfull[acc]); else tc_commit_pair(&s_tfull[acc]);   // accumulator complete
                    if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                }
            }
        }
    } else {
        // ===================== epilogue: 8 warps =====================
        // Two warps per TMEM lane quarter; each takes one half (128) of the accumulator's columns, so a
        // thread = (one A row, one column half).  Two warps per scheduler hide each other's latencies.
        const int quarter = warp & 3;                                   // TMEM lane quarter this warp may read
        const int half = (warp - 2) >> 2;                               // column half: warps 2-5 -> 0, 6-9 -> 1
        const int row_in_tile = quarter * 32 + lane;
        const int et = (warp - 2) * 32 + lane;                          // 0..255 among epilogue threads
        int acc = 0; uint32_t acc_phase = 0;
        for (int item = first_item; item < n_items; item += item_step) {
            const Item it = decode_item<NCTA>(p, item, rank);
            float vals[kTopK]; int idxs[kTopK];
#pragma unroll
            for (int j = 0; j < kTopK; ++j) { vals[j] = -INFINITY; idxs[j] = -1; }
            float runmax = -INFINITY;
            // inverse norms of the tile's 256 rows, one per epilogue thread; rows past the shard get NaN so they never
            // compare greater.  The load for tile t+1 is issued before tile t is processed: it was 32 % of the
            // epilogue's stall samples when it sat in front of the accumulator wait (profiles/).
            auto load_inv = [&](int t) {
                const int row = t * BN + et;
                return row < p.b_rows ? __ldg(p.inv_norm + row) : __int_as_float(0x7fc00000);
            };
            float inv_next = (p.mode == kCorpus && it.t0 < it.t1) ? load_inv(it.t0) : 0.f;
            for (int t = it.t0; t < it.t1; ++t) {
                if (p.mode == kCorpus) {
                    s_inv[acc][et] = inv_next;
                    if (t + 1 < it.t1) inv_next = load_inv(t + 1);
                }
                mbar_wait(&s_tfull[acc], acc_phase);
                tc_fence_after();
                asm volatile("bar.sync 2, 256;" ::: "memory");          // s_inv visible to all epilogue threads
                const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BN + half * (BN / 2));
#pragma unroll 1
                for (int c0 = 0; c0 < BN / 2; c0 += 32) {
                    uint32_t r[32];
                    tc_ld32(taddr + c0, r);
                    if (p.mode == kCorpus) {
                        // Common case (nothing beats the thread's 16th best): ~3 independent instructions per
                        // score.  Otherwise the 16 scaled scores go to a per-thread shared-memory row and the
                        // few that passed are folded in by ONE rolled insertion loop (the fully unrolled variant
                        // was instruction-fetch bound).
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const int cbase = half * (BN / 2) + c0 + 16 * h;
                            const float thr = vals[kTopK - 1];
                            float sc[16];
                            unsigned mask = 0;
#pragma unroll
                            for (int j4 = 0; j4 < 4; ++j4) {
                                const float4 iv = *reinterpret_cast<const float4*>(&s_inv[acc][cbase + 4 * j4]);
                                sc[4 * j4 + 0] = __uint_as_float(r[16 * h + 4 * j4 + 0]) * iv.x;
                                sc[4 * j4 + 1] = __uint_as_float(r[16 * h + 4 * j4 + 1]) * iv.y;
                                sc[4 * j4 + 2] = __uint_as_float(r[16 * h + 4 * j4 + 2]) * iv.z;
                                sc[4 * j4 + 3] = __uint_as_float(r[16 * h + 4 * j4 + 3]) * iv.w;
                            }
#pragma unroll
                            for (int j = 0; j < 16; ++j) mask |= (sc[j] > thr ? 1u : 0u) << j;
                            if (mask) {
#pragma unroll
                                for (int j = 0; j < 16; ++j) q_val[j * 256 + et] = sc[j];
                                const int nbase = t * BN + cbase;
                                while (mask) {
                                    const int j = __ffs(mask) - 1;
                                    mask &= mask - 1;
                                    const float v = q_val[j * 256 + et];
                                    const int id = nbase + j;
                                    if (v > vals[kTopK - 1]) {      // sorted insert: later (higher id) ties stay below
#pragma unroll
                                        for (int q = kTopK - 1; q >= 1; --q) {
                                            const bool here = v > vals[q], above = v > vals[q - 1];
                                            vals[q] = here ? (above ? vals[q - 1] : v) : vals[q];
                                            idxs[q] = here ? (above ? idxs[q - 1] : id) : idxs[q];
                                        }
                                        const bool top = v > vals[0];
                                        vals[0] = top ? v : vals[0];
                                        idxs[0] = top ? id : idxs[0];
                                    }
                                }
                            }
                        }
                    } else {
                        const int jbase = t * BN + half * (BN / 2) + c0;
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (jbase + j < p.b_rows) runmax = fmaxf(runmax, __uint_as_float(r[j]));
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {                                        // accumulator may be overwritten
                    if constexpr (NCTA == 1) mbar_arrive(&s_tempty[acc]); else mbar_arrive_cta0(&s_tempty[acc]);
                }
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
            const int a_row = it.a_tile * BM + row_in_tile;
            if (p.mode == kCorpus) {
                if (it.a_tile < p.n_a) {                                // an odd tile count leaves the pair's second CTA idle
                    const size_t o = (((size_t)it.group * 2 + half) * ((size_t)p.n_a * BM) + (size_t)a_row) * kTopK;
#pragma unroll
                    for (int j = 0; j < kTopK; ++j) { p.part_val[o + j] = vals[j]; p.part_idx[o + j] = idxs[j]; }
                }
            } else {
                // max over the two column halves of the same row, then the sum over the block's 128 rows
                q_val[et] = runmax;
                asm volatile("bar.sync 2, 256;" ::: "memory");
                float v = 0.f;
                if (half == 0 && a_row < p.a_rows) v = fmaxf(runmax, q_val[et + 128]);
                v = warp_sum(v);
                if (half == 0 && lane == 0) s_sum[warp - 2] = v;
                asm volatile("bar.sync 2, 256;" ::: "memory");
                if (et == 0 && it.a_tile < p.n_a)
                    p.partial[(size_t)it.group * p.n_a + it.a_tile] = (s_sum[0] + s_sum[1]) + (s_sum[2] + s_sum[3]);
                asm volatile("bar.sync 2, 256;" ::: "memory");
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if constexpr (NCTA == 2) cluster_sync_all();      // nobody leaves while its peer may still touch its barriers / TMEM
    if (warp == 1) {
        tc_fence_after();
        if constexpr (NCTA == 1)
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols));
        else
            asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols));
    }
}

// ---- small helper kernels ------------------------------------------------------------------------
// out_bf16[r,:] = x[r,:] (optionally / max(||x[r]||, 1e-12));  inv_norm[r] = 1 / max(||bf16(x[r])||, 1e-30)
__global__ void __launch_bounds__(256) rows_to_bf16_kernel(const float* __restrict__ x, int64_t rows, int d, int normalise,
                                                           __nv_bfloat16* __restrict__ out, float* __restrict__ inv_norm) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= rows) return;
    const float4* src = reinterpret_cast<const float4*>(x) + row * (d >> 2);
    float scale = 1.f;
    if (normalise) {
        float ss = 0.f;
        for (int i = lane; i < (d >> 2); i += 32) {
            const float4 v = src[i];
            ss = fmaf(v.x, v.x, ss); ss = fmaf(v.y, v.y, ss); ss = fmaf(v.z, v.z, ss); ss = fmaf(v.w, v.w, ss);
        }
        scale = __fdiv_rn(1.0f, fmaxf(__fsqrt_rn(warp_sum(ss)), 1e-12f));
    }
    float ssb = 0.f;
    __nv_bfloat162* dst = reinterpret_cast<__nv_bfloat162*>(out) + row * (d >> 1);
    for (int i = lane; i < (d >> 2); i += 32) {
        const float4 v = src[i];
        const __nv_bfloat162 lo = __floats2bfloat162_rn(v.x * scale, v.y * scale);
        const __nv_bfloat162 hi = __floats2bfloat162_rn(v.z * scale, v.w * scale);
        dst[2 * i] = lo; dst[2 * i + 1] = hi;
        const float2 a = __bfloat1622float2(lo), b = __bfloat1622float2(hi);
        ssb = fmaf(a.x, a.x, ssb); ssb = fmaf(a.y, a.y, ssb); ssb = fmaf(b.x, b.x, ssb); ssb = fmaf(b.y, b.y, ssb);
    }
    if (inv_norm) {
        ssb = warp_sum(ssb);
        if (lane == 0) inv_norm[row] = __fdiv_rn(1.0f, fmaxf(__fsqrt_rn(ssb), 1e-30f));
    }
}

// inv_norm[r] = 1 / max(||E[r]||, 1e-30) for a bf16 matrix
__global__ void __launch_bounds__(256) bf16_inv_norm_kernel(const __nv_bfloat16* __restrict__ x, int64_t rows, int d,
                                                            float* __restrict__ inv_norm) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= rows) return;
    const uint4* src = reinterpret_cast<const uint4*>(x) + row * (d >> 3);
    float ss = 0.f;
    for (int i = lane; i < (d >> 3); i += 32) {
        const uint4 v = ldg_stream_u4(src + i);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[e]));
            ss = fmaf(f.x, f.x, ss); ss = fmaf(f.y, f.y, ss);
        }
    }
    ss = warp_sum(ss);
    if (lane == 0) inv_norm[row] = __fdiv_rn(1.0f, fmaxf(__fsqrt_rn(ss), 1e-30f));
}

// out[s] = sum_t partial[s, t]  (fixed order)
__global__ void strip_sum_kernel(const float* __restrict__ partial, int n, int tiles, float* __restrict__ out) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    float acc = 0.f;
    for (int t = 0; t < tiles; ++t) acc += partial[(size_t)s * tiles + t];
    out[s] = acc;
}

// (groups, rows, 16) partial candidates -> (rows, groups*16) candidate lists with global int64 ids, scaled by 1/||q||
__global__ void corpus_candidates_kernel(const float* __restrict__ part_val, const int32_t* __restrict__ part_idx,
                                         const float* __restrict__ inv_q, int groups, int rows_padded, int Q,
                                         int64_t id_offset, float* __restrict__ cand_val, int64_t* __restrict__ cand_idx) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t total = (int64_t)Q * groups * kTopK;
    if (i >= total) return;
    const int j = (int)(i % kTopK);
    const int g = (int)((i / kTopK) % groups);
    const int q = (int)(i / ((int64_t)kTopK * groups));
    const size_t src = ((size_t)g * rows_padded + q) * kTopK + j;
    const int32_t id = part_idx[src];
    cand_val[i] = id >= 0 ? part_val[src] * inv_q[q] : -INFINITY;
    cand_idx[i] = id >= 0 ? id_offset + id : -1;
}

// bf16 tensor (z, rows, K) row-major -> 3-D map with a (64 x box_rows x 1) box, 128-byte swizzle
static int make_map(CUtensorMap* map, const void* base, int64_t K, int64_t rows, int64_t z, int box_rows) {
    return make_map_bytes(map, base, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, K, rows, z, box_rows);
}

template <int NCTA>
static int launch_n(const CUtensorMap& ma, const CUtensorMap& mb, const Params& p, cudaStream_t stream) {
    const size_t smem = (size_t)Cfg<NCTA>::kStages * Cfg<NCTA>::kStageBytes + 1024 + 16 * 256 * 4;
    RDV_ONCE_PER_DEVICE(cudaFuncSetAttribute(tc_score_kernel<NCTA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                        "cudaFuncSetAttribute(tc_score_kernel)");
    int grid = sm_count() / NCTA * NCTA;
    const int items = ((p.n_a + NCTA - 1) / NCTA) * p.n_groups;
    if (grid > items * NCTA) grid = items * NCTA;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = NCTA; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = NCTA > 1 ? 1 : 0;
    cudaError_t e = cudaLaunchKernelEx(&cfg, tc_score_kernel<NCTA>, ma, mb, p);
    if (e != cudaSuccess) return cuda_fail(e, "tc_score_kernel");
    return RDV_OK;
}

// Which shape runs is a measured choice (B200, 1024 questions x 768-d bf16; scripts/probe_c5_pairs.sh, round 2):
//   rows per launch      1.25 M     2.5 M      5 M       10 M
//   one CTA   (q/s)      631 k      323 k      154 k     78 k
//   CTA pairs (q/s)      656 k      335 k      124 k     66 k
// Pairs read a third less from L2 per tile (each CTA loads half of the B tile) and win by 4 % while the launch is short;
// in launches that run long enough for the 1 kW power cap to bite they draw more power per clock and lose 15-20 %.
// So: pairs for launches of up to ~3 M rows x 1024 questions x 768 (4.7 TFLOP; the per-rank share of the 10 M corpus at
// N >= 4), the single-CTA kernel above that; RDV_TC_CTAS=1 / 2 in the environment forces either (pairs only when the A
// tiles pair up: with an odd tile count the pair's second CTA would idle).
static int tc_ctas(const Params& p) {
    const char* v = getenv("RDV_TC_CTAS");
    if (p.n_a & 1) return 1;
    if (v && v[0] == '2') return 2;
    if (v && v[0] == '1') return 1;
    const double flops = 2.0 * (double)p.a_rows * (double)p.b_rows * 64.0 * (double)p.k_blocks;
    return (p.mode == kCorpus && flops <= 4.7e12) ? 2 : 1;
}

static int launch(const CUtensorMap& ma, const CUtensorMap& mb, const Params& p, cudaStream_t stream) {
    return tc_ctas(p) == 2 ? launch_n<2>(ma, mb, p, stream) : launch_n<1>(ma, mb, p, stream);
}

}  // namespace tc
}  // namespace rdv

using namespace rdv;

extern "C" int rdv_rows_to_bf16(const float* d_x, int64_t rows, int32_t d, int32_t normalise, void* d_out_bf16,
                                float* d_inv_norm, void* stream) {
    RDV_REQUIRE(rows >= 0, RDV_E_INVALID, "rows_to_bf16: negative size");
    if (rows == 0) return RDV_OK;
    RDV_REQUIRE(d_x && d_out_bf16, RDV_E_INVALID, "rows_to_bf16: null pointer");
    RDV_REQUIRE(d >= 8 && (d & 7) == 0, RDV_E_INVALID, "rows_to_bf16: d=%d must be a multiple of 8", d);
    RDV_REQUIRE(aligned16(d_x) && aligned16(d_out_bf16), RDV_E_ALIGN, "rows_to_bf16: buffers not 16-byte aligned");
    const int64_t blocks = (rows + 7) / 8;
    RDV_REQUIRE(blocks < (1ll << 31), RDV_E_LIMIT, "rows_to_bf16: too many rows");
    tc::rows_to_bf16_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        d_x, rows, d, normalise, static_cast<__nv_bfloat16*>(d_out_bf16), d_inv_norm);
    RDV_LAUNCH_CHECK("rows_to_bf16_kernel");
    return RDV_OK;
}

extern "C" int rdv_bf16_inv_norm(const void* d_x_bf16, int64_t rows, int32_t d, float* d_inv_norm, void* stream) {
    RDV_REQUIRE(rows >= 0, RDV_E_INVALID, "bf16_inv_norm: negative size");
    if (rows == 0) return RDV_OK;
    RDV_REQUIRE(d_x_bf16 && d_inv_norm, RDV_E_INVALID, "bf16_inv_norm: null pointer");
    RDV_REQUIRE(d >= 8 && (d & 7) == 0, RDV_E_INVALID, "bf16_inv_norm: d=%d must be a multiple of 8", d);
    RDV_REQUIRE(aligned16(d_x_bf16), RDV_E_ALIGN, "bf16_inv_norm: x not 16-byte aligned");
    const int64_t blocks = (rows + 7) / 8;
    RDV_REQUIRE(blocks < (1ll << 31), RDV_E_LIMIT, "bf16_inv_norm: too many rows");
    tc::bf16_inv_norm_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(d_x_bf16), rows, d, d_inv_norm);
    RDV_LAUNCH_CHECK("bf16_inv_norm_kernel");
    return RDV_OK;
}

extern "C" int32_t rdv_corpus_groups(int64_t n_rows, int32_t n_questions) {
    // work items = question blocks x row chunks; aim at ~2 items per SM, at least one B tile per chunk
    const int n_a = (n_questions + tc::BM - 1) / tc::BM;
    const int64_t tiles = (n_rows + tc::BN - 1) / tc::BN;
    int64_t g = (2 * (int64_t)sm_count() + n_a - 1) / (n_a > 0 ? n_a : 1);
    if (g > tiles) g = tiles;
    if (g < 1) g = 1;
    return (int32_t)g;
}

extern "C" int rdv_corpus_score_topk_bf16(const void* d_e_bf16, const float* d_e_inv_norm, int64_t n_rows, int32_t d,
                                          const void* d_q_bf16, const float* d_q_inv_norm, int32_t n_questions,
                                          int32_t k, int64_t id_offset, int32_t groups, float* d_part_val,
                                          int32_t* d_part_idx, float* d_cand_val, int64_t* d_cand_idx, void* stream) {
    RDV_REQUIRE(n_rows >= 0 && n_questions >= 0, RDV_E_INVALID, "corpus_score_topk_bf16: negative size");
    RDV_REQUIRE(d_e_bf16 && d_e_inv_norm && d_q_bf16 && d_q_inv_norm && d_part_val && d_part_idx && d_cand_val &&
                d_cand_idx, RDV_E_INVALID, "corpus_score_topk_bf16: null pointer");
    RDV_REQUIRE(d >= 8 && (d & 7) == 0, RDV_E_INVALID, "corpus_score_topk_bf16: d=%d must be a multiple of 8", d);
    RDV_REQUIRE(k >= 1 && k <= tc::kTopK, RDV_E_LIMIT, "corpus_score_topk_bf16: k=%d outside [1, %d]", k, tc::kTopK);
    RDV_REQUIRE(n_rows < (1ll << 31) - tc::BN, RDV_E_LIMIT, "corpus_score_topk_bf16: shard rows must fit int32");
    RDV_REQUIRE(n_rows >= 1 && n_questions >= 1 && groups >= 1, RDV_E_INVALID, "corpus_score_topk_bf16: empty shard / batch");
    RDV_REQUIRE(aligned16(d_e_bf16) && aligned16(d_q_bf16), RDV_E_ALIGN, "corpus_score_topk_bf16: operands not 16-byte aligned");
    tc::Params p = {};
    p.mode = tc::kCorpus;
    p.n_a = (n_questions + tc::BM - 1) / tc::BM;
    p.tiles_total = (int)((n_rows + tc::BN - 1) / tc::BN);
    p.n_groups = groups < p.tiles_total ? groups : p.tiles_total;
    RDV_REQUIRE(p.n_groups == groups, RDV_E_INVALID, "corpus_score_topk_bf16: groups=%d exceeds the %d row tiles", groups, p.tiles_total);
    p.k_blocks = (d + tc::BK - 1) / tc::BK;
    p.a_rows = n_questions; p.b_rows = (int)n_rows;
    p.inv_norm = d_e_inv_norm; p.part_val = d_part_val; p.part_idx = d_part_idx;
    CUtensorMap ma, mb;
    int rc = tc::make_map(&ma, d_q_bf16, d, n_questions, 1, tc::BM);
    if (rc) return rc;
    rc = tc::make_map(&mb, d_e_bf16, d, n_rows, 1, tc::BN / tc::tc_ctas(p));
    if (rc) return rc;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    rc = tc::launch(ma, mb, p, s);
    if (rc) return rc;
    const int64_t total = (int64_t)n_questions * groups * 2 * tc::kTopK;
    tc::corpus_candidates_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(
        d_part_val, d_part_idx, d_q_inv_norm, groups * 2, p.n_a * tc::BM, n_questions, id_offset, d_cand_val, d_cand_idx);
    RDV_LAUNCH_CHECK("corpus_candidates_kernel");
    return RDV_OK;
}

extern "C" int rdv_maxsim_bf16_tc(const void* d_qn_bf16, const void* d_pn_bf16, int32_t n, int32_t Lq, int32_t Lp,
                                  int32_t d, float* d_partial, float* d_out, void* stream) {
    RDV_REQUIRE(n >= 0 && Lq >= 0, RDV_E_INVALID, "maxsim_bf16_tc: negative size");
    if (n == 0) return RDV_OK;
    RDV_REQUIRE(Lp >= 1 && Lq >= 1, RDV_E_INVALID, "maxsim_bf16_tc: empty operand");
    RDV_REQUIRE(d_qn_bf16 && d_pn_bf16 && d_partial && d_out, RDV_E_INVALID, "maxsim_bf16_tc: null pointer");
    RDV_REQUIRE(d >= 8 && (d & 7) == 0, RDV_E_INVALID, "maxsim_bf16_tc: d=%d must be a multiple of 8", d);
    RDV_REQUIRE(aligned16(d_qn_bf16) && aligned16(d_pn_bf16), RDV_E_ALIGN, "maxsim_bf16_tc: operands not 16-byte aligned");
    tc::Params p = {};
    p.mode = tc::kMaxSim;
    p.n_a = (Lq + tc::BM - 1) / tc::BM;
    p.n_groups = n;
    p.tiles_total = (Lp + tc::BN - 1) / tc::BN;
    p.k_blocks = (d + tc::BK - 1) / tc::BK;
    p.a_rows = Lq; p.b_rows = Lp;
    p.partial = d_partial;
    CUtensorMap ma, mb;
    int rc = tc::make_map(&ma, d_qn_bf16, d, Lq, 1, tc::BM);
    if (rc) return rc;
    rc = tc::make_map(&mb, d_pn_bf16, d, Lp, n, tc::BN / tc::tc_ctas(p));
    if (rc) return rc;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    rc = tc::launch(ma, mb, p, s);
    if (rc) return rc;
    tc::strip_sum_kernel<<<(n + 127) / 128, 128, 0, s>>>(d_partial, n, p.n_a, d_out);
    RDV_LAUNCH_CHECK("strip_sum_kernel");
    return RDV_OK;
}

extern "C" int32_t rdv_tc_tile_m(void) { return tc::BM; }
extern "C" int32_t rdv_tc_candidates_per_group(void) { return 2 * tc::kTopK; }   // two column halves x 16
