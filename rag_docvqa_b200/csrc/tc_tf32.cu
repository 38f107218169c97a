// fp32-grade MaxSim on the tensor pipe: 3xTF32 split products (sm_100a: tcgen05 kind::tf32 + TMEM + TMA).
//
// late_interaction (reference src/utils.py:442-458) is an fp32 contraction: 2 * n * Lq * Lp * d flops
// (322 GFLOP per question at C4).  On CUDA cores (maxsim.cu) that is FFMA-bound at ~42 TFLOP/s.  The tensor
// pipe has no fp32 mode, but every fp32 number is the exact sum of two tf32-representable parts,
//     x = hi + lo,   hi = tf32(x) (10-bit mantissa),   lo = x - hi (exact; its own tf32 rounding is 2^-21 |x|),
// so   q . p  =  q_hi.p_hi + q_hi.p_lo + q_lo.p_hi + (q_lo.p_lo ~ 2^-22 |q||p|, dropped)
// with every partial product exact.  Three tf32 MMAs per product, at a third of the tf32 tensor rate instead of
// the CUDA-core rate.  Accuracy, measured on B200 against float64: the products are exact, but the tensor
// core's fp32 accumulator rounds toward zero at every accumulation step, which biases every cosine low: by
// ~d * 6e-9 relative with all three products in one accumulator (4.6e-6 at d = 768), by ~d * 2.1e-9 with the
// dominant product in its own accumulator as done here (1.6e-6 at d = 768, 4.4e-6 at d = 2048; an fp32 FFMA
// chain is ~1e-7).  That is inside the 1e-5 parity bar and it is the same relative bias on every strip, so
// rankings are unaffected; the CUDA-core kernel stays available for a strict fp32 result
// (functional.late_interaction(mode="ffma")).
//
// Operands are the L2-NORMALISED rows (F.normalize, src/utils.py:445-446), split once by
// rows_split_tf32_kernel.  Main loop = tc_gemm.cu's (TMA producer warp, one MMA-issuing thread, 8 epilogue
// warps) with 32-float k-blocks: a stage holds A_hi | A_lo | B_hi | B_lo (16 + 16 + 32 + 32 KB), two stages;
// TMEM holds two 128 x 256 accumulators per tile (dominant and cross products, see the MMA issuer).  Epilogue = running row max over the strip's
// tokens, summed over the question tokens in a fixed order (the (Lq x Lp) matrix never exists).
#include "tc_common.cuh"

namespace rdv {
namespace tc3 {

using namespace rdv::tc;

constexpr int BM = 128, BN = 256, BK = 32;          // 32 fp32 = one 128-byte swizzle row
constexpr int kStages = 2;
constexpr int kABytes = BM * BK * 4, kBBytes = BN * BK * 4;
constexpr int kStageBytes = 2 * kABytes + 2 * kBBytes;     // 96 KB
constexpr int kThreads = 320;
constexpr uint32_t kTmemCols = 512;
// kind::tf32 instruction descriptor: D = f32, A = B = tf32, both K-major, M = 128, N = 256
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

struct Params {
    int n_a;          // question-token tiles
    int n_strips;
    int b_tiles;      // strip-token tiles per strip
    int k_blocks;     // ceil(d / 32)
    int a_rows, b_rows;
    float* partial;   // (n_strips, n_a)
};

__global__ void __launch_bounds__(kThreads, 1)
maxsim_tf32x3_kernel(const __grid_constant__ CUtensorMap map_ahi, const __grid_constant__ CUtensorMap map_alo,
                     const __grid_constant__ CUtensorMap map_bhi, const __grid_constant__ CUtensorMap map_blo, const Params p) {
    extern __shared__ __align__(1024) unsigned char smem_unaligned[];
    unsigned char* smem = smem_unaligned + ((1024u - (s32(smem_unaligned) & 1023u)) & 1023u);
    __shared__ __align__(8) uint64_t s_full[kStages], s_empty[kStages], s_tfull[2], s_tempty[2];
    __shared__ uint32_t s_tmem_base;
    __shared__ float s_max[256];
    __shared__ float s_sum[4];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_items = p.n_a * p.n_strips;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(&s_full[s], 1); mbar_init(&s_empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&s_tfull[s], 1); mbar_init(&s_tempty[s], 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&s_tmem_base)), "r"(kTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = s_tmem_base;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
                const int strip = item / p.n_a, a_tile = item - strip * p.n_a;
                for (int t = 0; t < p.b_tiles; ++t) {
                    for (int kb = 0; kb < p.k_blocks; ++kb) {
                        mbar_wait(&s_empty[stage], phase ^ 1);
                        unsigned char* st = smem + (size_t)stage * kStageBytes;
                        mbar_expect_tx(&s_full[stage], kStageBytes);
                        tma_load_3d(st, &map_ahi, &s_full[stage], kb * BK, a_tile * BM, 0);
                        tma_load_3d(st + kABytes, &map_alo, &s_full[stage], kb * BK, a_tile * BM, 0);
                        tma_load_3d(st + 2 * kABytes, &map_bhi, &s_full[stage], kb * BK, t * BN, strip);
                        tma_load_3d(st + 2 * kABytes + kBBytes, &map_blo, &s_full[stage], kb * BK, t * BN, strip);
                        if (++stage == kStages) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (one thread) =====================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            uint32_t acc_phase = 0;
            // Two accumulators per tile, no double buffering: MAIN (columns 0..255) takes only the dominant
            // hi*hi products, CROSS (columns 256..511) the two 2^-11-sized cross products.  The accumulator rounds
            // toward zero at every step, so keeping the small terms out of MAIN cuts its rounding steps from
            // 3d/8 to d/8 (measured bias at d = 768: -4.6e-6 -> -1.6e-6 relative); the epilogue adds the two in fp32.
            // A tile is 3 * d/8 MMAs (~19 us at d = 768), the un-overlapped epilogue ~1 us of it.
            const uint32_t tmem_main = tmem_base, tmem_cross = tmem_base + BN;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
                for (int t = 0; t < p.b_tiles; ++t) {
                    mbar_wait(&s_tempty[0], acc_phase ^ 1);
                    tc_fence_after();
                    for (int kb = 0; kb < p.k_blocks; ++kb) {
                        mbar_wait(&s_full[stage], phase);
                        tc_fence_after();
                        const uint32_t base = s32(smem + (size_t)stage * kStageBytes);
                        const uint64_t ahi = smem_desc(base), alo = smem_desc(base + kABytes);
                        const uint64_t bhi = smem_desc(base + 2 * kABytes), blo = smem_desc(base + 2 * kABytes + kBBytes);
#pragma unroll
                        for (int k = 0; k < BK / 8; ++k) {                 // K = 8 tf32 = 32 bytes (>>4 = 2) per MMA
                            const uint64_t o = (uint64_t)(2 * k);
                            tc_mma_tf32(tmem_cross, alo + o, bhi + o, kIdesc, (kb | k) ? 1u : 0u);
                            tc_mma_tf32(tmem_cross, ahi + o, blo + o, kIdesc, 1u);
                            tc_mma_tf32(tmem_main, ahi + o, bhi + o, kIdesc, (kb | k) ? 1u : 0u);
                        }
                        tc_commit(&s_empty[stage]);
                        if (++stage == kStages) { stage = 0; phase ^= 1; }
                    }
                    tc_commit(&s_tfull[0]);
                    acc_phase ^= 1;
                }
            }
        }
    } else {
        // ===================== epilogue: 8 warps, thread = (question token, column half) =====================
        const int quarter = warp & 3;
        const int half = (warp - 2) >> 2;
        const int row_in_tile = quarter * 32 + lane;
        const int et = (warp - 2) * 32 + lane;
        uint32_t acc_phase = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            const int strip = item / p.n_a, a_tile = item - strip * p.n_a;
            float runmax = -INFINITY;
            for (int t = 0; t < p.b_tiles; ++t) {
                mbar_wait(&s_tfull[0], acc_phase);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(half * (BN / 2));
#pragma unroll 1
                for (int c0 = 0; c0 < BN / 2; c0 += 32) {
                    uint32_t r[32], x[32];
                    tc_ld32(taddr + c0, r);                  // MAIN:  sum q_hi p_hi
                    tc_ld32(taddr + BN + c0, x);             // CROSS: sum q_lo p_hi + q_hi p_lo
                    const int jbase = t * BN + half * (BN / 2) + c0;
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (jbase + j < p.b_rows) runmax = fmaxf(runmax, __uint_as_float(r[j]) + __uint_as_float(x[j]));
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&s_tempty[0]);
                acc_phase ^= 1;
            }
            const int a_row = a_tile * BM + row_in_tile;
            s_max[et] = runmax;
            asm volatile("bar.sync 2, 256;" ::: "memory");
            float v = 0.f;
            if (half == 0 && a_row < p.a_rows) v = fmaxf(runmax, s_max[et + 128]);
            v = warp_sum(v);
            if (half == 0 && lane == 0) s_sum[warp - 2] = v;
            asm volatile("bar.sync 2, 256;" ::: "memory");
            if (et == 0) p.partial[(size_t)strip * p.n_a + a_tile] = (s_sum[0] + s_sum[1]) + (s_sum[2] + s_sum[3]);
            asm volatile("bar.sync 2, 256;" ::: "memory");
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols));
    }
}

// hi[r,:] = tf32(xn[r,:]), lo[r,:] = xn[r,:] - hi[r,:], xn = x / max(||x||, 1e-12) (F.normalize) or x itself
__global__ void __launch_bounds__(256) rows_split_tf32_kernel(const float* __restrict__ x, int64_t rows, int d, int normalise,
                                                              float* __restrict__ hi, float* __restrict__ lo) {
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
    float4* dh = reinterpret_cast<float4*>(hi) + row * (d >> 2);
    float4* dl = reinterpret_cast<float4*>(lo) + row * (d >> 2);
    for (int i = lane; i < (d >> 2); i += 32) {
        const float4 v = src[i];
        const float e[4] = {v.x * scale, v.y * scale, v.z * scale, v.w * scale};
        float h[4], l[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            uint32_t t;
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(e[c]));       // round to nearest, low 13 bits zero
            h[c] = __uint_as_float(t);
            l[c] = e[c] - h[c];                                             // exact
        }
        dh[i] = make_float4(h[0], h[1], h[2], h[3]);
        dl[i] = make_float4(l[0], l[1], l[2], l[3]);
    }
}

__global__ void strip_sum3_kernel(const float* __restrict__ partial, int n, int tiles, float* __restrict__ out) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    float acc = 0.f;
    for (int t = 0; t < tiles; ++t) acc += partial[(size_t)s * tiles + t];
    out[s] = acc;
}

}  // namespace tc3
}  // namespace rdv

using namespace rdv;

extern "C" int rdv_rows_split_tf32(const float* d_x, int64_t rows, int32_t d, int32_t normalise, float* d_hi, float* d_lo,
                                   void* stream) {
    RDV_REQUIRE(rows >= 0, RDV_E_INVALID, "rows_split_tf32: negative size");
    if (rows == 0) return RDV_OK;
    RDV_REQUIRE(d_x && d_hi && d_lo, RDV_E_INVALID, "rows_split_tf32: null pointer");
    RDV_REQUIRE(d >= 4 && (d & 3) == 0, RDV_E_INVALID, "rows_split_tf32: d=%d must be a multiple of 4", d);
    RDV_REQUIRE(aligned16(d_x) && aligned16(d_hi) && aligned16(d_lo), RDV_E_ALIGN, "rows_split_tf32: buffers not 16-byte aligned");
    const int64_t blocks = (rows + 7) / 8;
    RDV_REQUIRE(blocks < (1ll << 31), RDV_E_LIMIT, "rows_split_tf32: too many rows");
    tc3::rows_split_tf32_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(d_x, rows, d, normalise, d_hi, d_lo);
    RDV_LAUNCH_CHECK("rows_split_tf32_kernel");
    return RDV_OK;
}

extern "C" int rdv_maxsim_tf32x3_tc(const float* d_q_hi, const float* d_q_lo, const float* d_p_hi, const float* d_p_lo,
                                    int32_t n, int32_t Lq, int32_t Lp, int32_t d, float* d_partial, float* d_out, void* stream) {
    RDV_REQUIRE(n >= 0 && Lq >= 0, RDV_E_INVALID, "maxsim_tf32x3_tc: negative size");
    if (n == 0) return RDV_OK;
    RDV_REQUIRE(Lp >= 1 && Lq >= 1, RDV_E_INVALID, "maxsim_tf32x3_tc: empty operand");
    RDV_REQUIRE(d_q_hi && d_q_lo && d_p_hi && d_p_lo && d_partial && d_out, RDV_E_INVALID, "maxsim_tf32x3_tc: null pointer");
    RDV_REQUIRE(d >= 4 && (d & 3) == 0, RDV_E_INVALID, "maxsim_tf32x3_tc: d=%d must be a multiple of 4", d);
    RDV_REQUIRE(aligned16(d_q_hi) && aligned16(d_q_lo) && aligned16(d_p_hi) && aligned16(d_p_lo), RDV_E_ALIGN,
                "maxsim_tf32x3_tc: operands not 16-byte aligned");
    tc3::Params p = {};
    p.n_a = (Lq + tc3::BM - 1) / tc3::BM;
    p.n_strips = n;
    p.b_tiles = (Lp + tc3::BN - 1) / tc3::BN;
    p.k_blocks = (d + tc3::BK - 1) / tc3::BK;
    p.a_rows = Lq; p.b_rows = Lp;
    p.partial = d_partial;
    CUtensorMap mah, mal, mbh, mbl;
    int rc = tc::make_map_bytes(&mah, d_q_hi, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d, Lq, 1, tc3::BM);
    if (!rc) rc = tc::make_map_bytes(&mal, d_q_lo, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d, Lq, 1, tc3::BM);
    if (!rc) rc = tc::make_map_bytes(&mbh, d_p_hi, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d, Lp, n, tc3::BN);
    if (!rc) rc = tc::make_map_bytes(&mbl, d_p_lo, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d, Lp, n, tc3::BN);
    if (rc) return rc;
    const size_t smem = (size_t)tc3::kStages * tc3::kStageBytes + 1024;
    RDV_ONCE_PER_DEVICE(cudaFuncSetAttribute(tc3::maxsim_tf32x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                        "cudaFuncSetAttribute(maxsim_tf32x3_kernel)");
    int grid = sm_count();
    const int items = p.n_a * p.n_strips;
    if (grid > items) grid = items;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    tc3::maxsim_tf32x3_kernel<<<grid, tc3::kThreads, smem, s>>>(mah, mal, mbh, mbl, p);
    RDV_LAUNCH_CHECK("maxsim_tf32x3_kernel");
    tc3::strip_sum3_kernel<<<(n + 127) / 128, 128, 0, s>>>(d_partial, n, p.n_a, d_out);
    RDV_LAUNCH_CHECK("strip_sum3_kernel");
    return RDV_OK;
}
