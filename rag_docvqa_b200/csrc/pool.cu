// Masked mean pooling (+ optional fused L2 normalisation, optional bf16 copy) -- sm_100a.
//
// Replaces mean_pooling (reference src/_model_utils.py:49-61, called at src/_modules.py:1474):
//     out[i,:] = sum_t embs[i,t,:] * mask[i,t] / max(sum_t mask[i,t], 1e-9)
// The reference materialises embs*mask (a full-size temporary) and reads it again; here every token
// row is read exactly once and tokens whose mask is 0 are not read at all, so the traffic is
// n*Lvalid*d*4 instead of 3*n*L*d*4.  HBM-bound: one block per chunk, thread = one float4 column of
// d, G token-groups per block accumulate in parallel and are folded through shared memory.  When there are few rows of many
// tokens, a row is split over a thread-block cluster instead (mean_pool_split_kernel below).
//
// Deviation, documented: a NaN/Inf sitting at a *masked-out* token position is ignored here, whereas
// the reference's `embs * 0` would propagate it (finite inputs: bit-for-bit the same sum order per
// group; result within fp32 rounding of the reference).
#include "rdv_common.cuh"
#include <cuda_bf16.h>

namespace rdv {

constexpr int kPoolThreads = 256;
constexpr int kPoolMaxL = 4096;
constexpr int kPoolUnroll = 8;

struct PoolParams {
    const float* embs;       // (n, L, d)
    const int64_t* mask;     // (n, L)
    int32_t n, L, d;
    int32_t normalise;       // 1: divide the pooled row by max(||row||, 1e-12)  (F.normalize semantics)
    float* out;              // (n, d) or null
    __nv_bfloat16* out_bf16; // (n, d) or null
    float* out_norm;         // (n,) L2 norm of the pooled row before normalisation, or null
};

__global__ void __launch_bounds__(kPoolThreads, 4) mean_pool_kernel(const PoolParams p) {
    extern __shared__ float4 s_part[];            // [G][d4] partial sums (G > 1 only)
    __shared__ float s_mask[kPoolMaxL];
    __shared__ float s_red[kPoolThreads / 32];
    __shared__ float s_scalar[2];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int d4 = p.d >> 2;
    const int row = blockIdx.x;
    const int64_t* mrow = p.mask + (size_t)row * p.L;

    // mask row -> shared (as float, the reference multiplies by it), count = sum(mask)
    float cnt = 0.f;
    for (int t = tid; t < p.L; t += kPoolThreads) {
        const float m = (float)mrow[t];
        s_mask[t] = m;
        cnt += m;
    }
    cnt = warp_sum(cnt);
    if (lane == 0) s_red[warp] = cnt;
    __syncthreads();
    if (tid == 0) {
        float c = 0.f;
        for (int w = 0; w < kPoolThreads / 32; ++w) c += s_red[w];
        s_scalar[0] = fmaxf(c, 1e-9f);            // clamp(min=1e-9)
    }
    __syncthreads();
    const float denom = s_scalar[0];

    // thread layout: column c = tid % cols, token group g = tid / cols
    const int cols = d4 < kPoolThreads ? d4 : kPoolThreads;
    const int G = kPoolThreads / cols;            // >= 1
    const int g = tid / cols, c0 = tid - g * cols;
    const bool active = g < G;
    const float4* base = reinterpret_cast<const float4*>(p.embs) + (size_t)row * p.L * d4;

    float ss_local = 0.f;
    for (int c = c0; c < d4; c += cols) {         // more than one pass only when d4 > 256
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        if (active) {
            int t = g;
            // kPoolUnroll independent 128-bit loads in flight per thread (masked tokens are never requested)
            for (; t + (kPoolUnroll - 1) * G < p.L; t += kPoolUnroll * G) {
                float m[kPoolUnroll];
                float4 v[kPoolUnroll];
#pragma unroll
                for (int u = 0; u < kPoolUnroll; ++u) {
                    m[u] = s_mask[t + u * G];
                    if (m[u] != 0.f) v[u] = ldg_stream(base + (size_t)(t + u * G) * d4 + c);
                }
#pragma unroll
                for (int u = 0; u < kPoolUnroll; ++u) {
                    if (m[u] != 0.f) {
                        acc.x = fmaf(v[u].x, m[u], acc.x); acc.y = fmaf(v[u].y, m[u], acc.y);
                        acc.z = fmaf(v[u].z, m[u], acc.z); acc.w = fmaf(v[u].w, m[u], acc.w);
                    }
                }
            }
            for (; t < p.L; t += G) {
                const float m = s_mask[t];
                if (m != 0.f) {
                    const float4 v = ldg_stream(base + (size_t)t * d4 + c);
                    acc.x = fmaf(v.x, m, acc.x); acc.y = fmaf(v.y, m, acc.y);
                    acc.z = fmaf(v.z, m, acc.z); acc.w = fmaf(v.w, m, acc.w);
                }
            }
        }
        if (G > 1) {                               // fold the token groups (fixed order: deterministic)
            if (active) s_part[g * cols + c0] = acc;
            __syncthreads();
            if (g == 0) {
                for (int gg = 1; gg < G; ++gg) {
                    const float4 o = s_part[gg * cols + c0];
                    acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w;
                }
            }
        }
        if (g == 0) {
            acc.x = __fdiv_rn(acc.x, denom); acc.y = __fdiv_rn(acc.y, denom);
            acc.z = __fdiv_rn(acc.z, denom); acc.w = __fdiv_rn(acc.w, denom);
            ss_local += acc.x * acc.x + acc.y * acc.y + acc.z * acc.z + acc.w * acc.w;
            if (!p.normalise) {
                if (p.out) reinterpret_cast<float4*>(p.out)[(size_t)row * d4 + c] = acc;
                if (p.out_bf16) {
                    __nv_bfloat162* o = reinterpret_cast<__nv_bfloat162*>(p.out_bf16) + ((size_t)row * d4 + c) * 2;
                    o[0] = __floats2bfloat162_rn(acc.x, acc.y);
                    o[1] = __floats2bfloat162_rn(acc.z, acc.w);
                }
            } else {
                s_part[(size_t)G * cols + c] = acc; // stash the mean row for the second pass
            }
        }
        if (G > 1) __syncthreads();
    }
    if (!p.normalise && !p.out_norm) return;

    // row L2 norm
    ss_local = warp_sum(ss_local);
    __syncthreads();
    if (lane == 0) s_red[warp] = ss_local;
    __syncthreads();
    if (tid == 0) {
        float s = 0.f;
        for (int w = 0; w < kPoolThreads / 32; ++w) s += s_red[w];
        const float nrm = __fsqrt_rn(s);
        s_scalar[1] = nrm;
        if (p.out_norm) p.out_norm[row] = nrm;
    }
    __syncthreads();
    if (!p.normalise) return;
    const float inv_den = fmaxf(s_scalar[1], 1e-12f);
    for (int c = tid; c < d4; c += kPoolThreads) {
        float4 v = s_part[(size_t)G * cols + c];
        v.x = __fdiv_rn(v.x, inv_den); v.y = __fdiv_rn(v.y, inv_den);
        v.z = __fdiv_rn(v.z, inv_den); v.w = __fdiv_rn(v.w, inv_den);
        if (p.out) reinterpret_cast<float4*>(p.out)[(size_t)row * d4 + c] = v;
        if (p.out_bf16) {
            __nv_bfloat162* o = reinterpret_cast<__nv_bfloat162*>(p.out_bf16) + ((size_t)row * d4 + c) * 2;
            o[0] = __floats2bfloat162_rn(v.x, v.y);
            o[1] = __floats2bfloat162_rn(v.z, v.w);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Few rows, many tokens (8 rendered questions x 2048 patch vectors x 768: 50 MB that one block per row would pull through 8
// of the 148 SMs, 284 us measured): one row per thread-block CLUSTER of S CTAs.  CTA r sums the tokens [r * per, (r + 1) * per)
// of the row exactly as the kernel above does, stores its partial row into the LEADER's shared memory (st.shared::cluster),
// and after one cluster barrier the leader adds the S partial rows in rank order (deterministic), divides and normalises.
// ---------------------------------------------------------------------------------------------------------------------
// 1024 threads: with D = 768 that is five token groups per CTA, 960 threads x 8 independent 128-bit loads = 120 KB in
// flight per SM (256 threads: one group, 24 KB in flight, 60 us for the 50 MB of 8 x 2048 x 768 questions; 512 threads
// with clusters of 16: 38 us; 1024: 32-36 us).
constexpr int kSplitThreads = 1024;

__device__ __forceinline__ uint32_t pool_cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void pool_remote_store(const void* local, uint32_t rank, float4 v) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(local);
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank));
    asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(r), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__global__ void __launch_bounds__(kSplitThreads) mean_pool_split_kernel(const PoolParams p, const int S) {
    extern __shared__ float4 s_dyn[];             // [G][cols] group partials | [S][d4] the ranks' partial rows (leader's copy is read)
    __shared__ float s_mask[kPoolMaxL];
    __shared__ float s_red[kSplitThreads / 32];
    __shared__ float s_scalar[2];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int d4 = p.d >> 2;
    const uint32_t rank = pool_cluster_rank();
    const int row = blockIdx.x / S;
    const int64_t* mrow = p.mask + (size_t)row * p.L;

    float cnt = 0.f;                              // the whole mask row in every CTA: the same count, summed in the same order
    for (int t = tid; t < p.L; t += kSplitThreads) {
        const float m = (float)mrow[t];
        s_mask[t] = m;
        cnt += m;
    }
    cnt = warp_sum(cnt);
    if (lane == 0) s_red[warp] = cnt;
    __syncthreads();
    if (tid == 0) {
        float c = 0.f;
        for (int w = 0; w < kSplitThreads / 32; ++w) c += s_red[w];
        s_scalar[0] = fmaxf(c, 1e-9f);
    }
    __syncthreads();
    const float denom = s_scalar[0];

    const int cols = d4 < kSplitThreads ? d4 : kSplitThreads;
    const int G = kSplitThreads / cols;
    const int g = tid / cols, c0 = tid - g * cols;
    const bool active = g < G;
    const int per = (p.L + S - 1) / S;
    const int t_lo = (int)rank * per, t_hi = min(p.L, t_lo + per);
    const float4* base = reinterpret_cast<const float4*>(p.embs) + (size_t)row * p.L * d4;
    float4* s_fold = s_dyn;
    float4* s_rows = s_dyn + (size_t)G * cols;

    for (int c = c0; c < d4; c += cols) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        if (active) {
            // rounds of kPoolUnroll independent loads; the last round is bounds-checked per load instead of falling back to
            // one dependent load at a time (a CTA's slice is ~26 tokens per group: the tail was a third of the work)
            for (int t = t_lo + g; t < t_hi; t += kPoolUnroll * G) {
                float m[kPoolUnroll];
                float4 v[kPoolUnroll];
#pragma unroll
                for (int u = 0; u < kPoolUnroll; ++u) {
                    const int tu = t + u * G;
                    m[u] = tu < t_hi ? s_mask[tu] : 0.f;
                    if (m[u] != 0.f) v[u] = ldg_stream(base + (size_t)tu * d4 + c);
                }
#pragma unroll
                for (int u = 0; u < kPoolUnroll; ++u) {
                    if (m[u] != 0.f) {
                        acc.x = fmaf(v[u].x, m[u], acc.x); acc.y = fmaf(v[u].y, m[u], acc.y);
                        acc.z = fmaf(v[u].z, m[u], acc.z); acc.w = fmaf(v[u].w, m[u], acc.w);
                    }
                }
            }
        }
        if (G > 1) {
            if (active) s_fold[g * cols + c0] = acc;
            __syncthreads();
            if (g == 0) {
                for (int gg = 1; gg < G; ++gg) {
                    const float4 o = s_fold[gg * cols + c0];
                    acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w;
                }
            }
        }
        if (g == 0) pool_remote_store(s_rows + (size_t)rank * d4 + c, 0, acc);      // into the leader's copy of s_rows
        if (G > 1) __syncthreads();
    }
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (rank != 0) return;

    float ss_local = 0.f;
    for (int c = tid; c < d4; c += kSplitThreads) {
        float4 acc = s_rows[c];
        for (int r = 1; r < S; ++r) {
            const float4 o = s_rows[(size_t)r * d4 + c];
            acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w;
        }
        acc.x = __fdiv_rn(acc.x, denom); acc.y = __fdiv_rn(acc.y, denom);
        acc.z = __fdiv_rn(acc.z, denom); acc.w = __fdiv_rn(acc.w, denom);
        ss_local += acc.x * acc.x + acc.y * acc.y + acc.z * acc.z + acc.w * acc.w;
        if (!p.normalise) {
            if (p.out) reinterpret_cast<float4*>(p.out)[(size_t)row * d4 + c] = acc;
            if (p.out_bf16) {
                __nv_bfloat162* o = reinterpret_cast<__nv_bfloat162*>(p.out_bf16) + ((size_t)row * d4 + c) * 2;
                o[0] = __floats2bfloat162_rn(acc.x, acc.y);
                o[1] = __floats2bfloat162_rn(acc.z, acc.w);
            }
        } else {
            s_rows[c] = acc;
        }
    }
    if (!p.normalise && !p.out_norm) return;
    ss_local = warp_sum(ss_local);
    __syncthreads();
    if (lane == 0) s_red[warp] = ss_local;
    __syncthreads();
    if (tid == 0) {
        float s = 0.f;
        for (int w = 0; w < kSplitThreads / 32; ++w) s += s_red[w];
        const float nrm = __fsqrt_rn(s);
        s_scalar[1] = nrm;
        if (p.out_norm) p.out_norm[row] = nrm;
    }
    __syncthreads();
    if (!p.normalise) return;
    const float inv_den = fmaxf(s_scalar[1], 1e-12f);
    for (int c = tid; c < d4; c += kSplitThreads) {
        float4 v = s_rows[c];
        v.x = __fdiv_rn(v.x, inv_den); v.y = __fdiv_rn(v.y, inv_den);
        v.z = __fdiv_rn(v.z, inv_den); v.w = __fdiv_rn(v.w, inv_den);
        if (p.out) reinterpret_cast<float4*>(p.out)[(size_t)row * d4 + c] = v;
        if (p.out_bf16) {
            __nv_bfloat162* o = reinterpret_cast<__nv_bfloat162*>(p.out_bf16) + ((size_t)row * d4 + c) * 2;
            o[0] = __floats2bfloat162_rn(v.x, v.y);
            o[1] = __floats2bfloat162_rn(v.z, v.w);
        }
    }
}

constexpr size_t kPoolSplitSmem = 160 * 1024;

// CTAs per row: a power of two <= 16 that brings the grid to about two CTAs per SM, leaves every CTA >= 64 tokens and keeps
// the S partial rows in the leader's shared memory.  RDV_POOL_SPLIT=0 turns the split off, =S forces it (tests).
static int pool_split(int n, int L, int d, size_t base_smem) {
    int want = 0;
    if (const char* e = getenv("RDV_POOL_SPLIT")) want = atoi(e), want = want <= 0 ? 1 : want;
    int S = 1;
    while (S < 16 && (want ? S * 2 <= want : ((int64_t)n * S * 2 <= 2 * 148 && L / (S * 2) >= 64))) S *= 2;
    while (S > 1 && base_smem + (size_t)S * d * 4 > kPoolSplitSmem) S /= 2;
    return S;
}

// Clusters of S CTAs of this kernel the device holds at once (cudaOccupancyMaxActiveClusters), remembered per device, S and d.
static int split_clusters_resident(int S, int d, const cudaLaunchConfig_t& cfg) {
    struct Entry { int dev, S, d, active; };
    static Entry cache[16];
    static std::atomic<int> used{0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 1 << 30;
    const int n_used = used.load(std::memory_order_acquire);
    for (int i = 0; i < n_used; ++i)
        if (cache[i].dev == dev && cache[i].S == S && cache[i].d == d) return cache[i].active;
    int active = 0;
    if (cudaOccupancyMaxActiveClusters(&active, mean_pool_split_kernel, &cfg) != cudaSuccess || active <= 0) {
        cudaGetLastError();
        return 1 << 30;                                     // unknown: do not shrink the clusters on a guess
    }
    static std::atomic<int> claim{0};
    const int slot = claim.fetch_add(1, std::memory_order_relaxed);
    if (slot < 16) {
        cache[slot] = {dev, S, d, active};
        int expect = slot;                                  // publish in order; a racing reader simply asks the driver again
        used.compare_exchange_strong(expect, slot + 1, std::memory_order_release);
    }
    return active;
}

}  // namespace rdv

extern "C" int rdv_mean_pool_f32(const float* d_embs, const int64_t* d_mask, int32_t n, int32_t L, int32_t d,
                                 int32_t normalise, float* d_out, void* d_out_bf16, float* d_out_norm,
                                 void* stream) {
    using namespace rdv;
    RDV_REQUIRE(n >= 0 && L >= 0, RDV_E_INVALID, "mean_pool_f32: negative size");
    if (n == 0) return RDV_OK;
    RDV_REQUIRE(d_embs && d_mask && (d_out || d_out_bf16), RDV_E_INVALID, "mean_pool_f32: null pointer");
    RDV_REQUIRE(d >= 4 && d <= 8192 && (d & 3) == 0, RDV_E_INVALID,
                "mean_pool_f32: d=%d must be a multiple of 4 in [4, 8192]", d);
    RDV_REQUIRE(L <= kPoolMaxL, RDV_E_LIMIT, "mean_pool_f32: L=%d > %d tokens", L, kPoolMaxL);
    RDV_REQUIRE(aligned16(d_embs) && (!d_out || aligned16(d_out)) && (!d_out_bf16 || aligned16(d_out_bf16)),
                RDV_E_ALIGN, "mean_pool_f32: buffers must be 16-byte aligned");
    PoolParams p;
    p.embs = d_embs; p.mask = d_mask; p.n = n; p.L = L; p.d = d; p.normalise = normalise ? 1 : 0;
    p.out = d_out; p.out_bf16 = static_cast<__nv_bfloat16*>(d_out_bf16); p.out_norm = d_out_norm;
    const int d4 = d >> 2;
    const int cols = d4 < kPoolThreads ? d4 : kPoolThreads;
    const int G = kPoolThreads / cols;
    // [G][cols] group partials + [d4] stash for the normalise pass
    const size_t smem = ((size_t)G * cols + (size_t)d4) * sizeof(float4);
    RDV_ONCE_PER_DEVICE(cudaFuncSetAttribute(mean_pool_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024),
                        "cudaFuncSetAttribute(mean_pool)");
    const int cols_s = d4 < kSplitThreads ? d4 : kSplitThreads;
    const int G_s = kSplitThreads / cols_s;
    int S = pool_split(n, L, d, (size_t)G_s * cols_s * sizeof(float4));
    if (S > 1) {
        RDV_ONCE_PER_DEVICE(cudaFuncSetAttribute(mean_pool_split_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPoolSplitSmem),
                            "cudaFuncSetAttribute(mean_pool_split)");
        RDV_ONCE_PER_DEVICE(cudaFuncSetAttribute(mean_pool_split_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1),
                            "cudaFuncSetAttribute(mean_pool_split, cluster)");
        cudaLaunchConfig_t cfg = {};
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        cfg.blockDim = dim3(kSplitThreads);
        cfg.stream = static_cast<cudaStream_t>(stream);
        cfg.attrs = attr; cfg.numAttrs = 1;
        const bool forced = getenv("RDV_POOL_SPLIT") != nullptr;
        for (;; S /= 2) {
            cfg.gridDim = dim3((unsigned)n * S);
            cfg.dynamicSmemBytes = ((size_t)G_s * cols_s + (size_t)S * d4) * sizeof(float4);
            attr[0].val.clusterDim.x = S; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
            // all clusters resident at once: a B200 holds 7 clusters of 16 full-SM CTAs (ncu launch__cluster_max_active), so
            // the 8th row of a C4p batch waited for a second wave (36 us); 8 clusters of 8 are one wave (32 us).  What is
            // left is round latency: a CTA's slice is 3-4 rounds of 8 loads per thread (~2 us each) after a ~3 us mask pass
            // and before a ~2 us hand-over -- a CTA streams at 24 GB/s, and only 64-128 of the 148 SMs have one
            if (forced || S <= 2 || split_clusters_resident(S, d, cfg) >= n) break;
        }
        cudaError_t e = cudaLaunchKernelEx(&cfg, mean_pool_split_kernel, p, S);
        if (e != cudaSuccess) return cuda_fail(e, "mean_pool_split_kernel");
        return RDV_OK;
    }
    mean_pool_kernel<<<n, kPoolThreads, smem, static_cast<cudaStream_t>(stream)>>>(p);
    RDV_LAUNCH_CHECK("mean_pool_kernel");
    return RDV_OK;
}
