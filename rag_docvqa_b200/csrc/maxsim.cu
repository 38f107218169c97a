// ColBERT late interaction (MaxSim) -- fp32 FFMA parity mode, sm_100a.
//
// Replaces late_interaction (reference src/utils.py:442-458, called per document from
// VisualRetriever._get_similarities, src/_modules.py:2191-2205):
//     Qn = Q / max(||Q||_row, 1e-12);  Pn likewise           (F.normalize)
//     S[n] = Qn @ Pn[n]^T                (Lq x Lp)            (torch.bmm)
//     score[n] = sum_i max_j S[n][i][j]                       (max over the strip's tokens, summed over
//                                                              the question's tokens; no mask)
// The reference materialises S (n x Lq x Lp fp32 = 16 MiB per strip at 2048 x 2048) and an expanded
// copy of Q.  Here S never exists: a block owns a 128-row slice of the question for one strip, walks
// all 128-column tiles of the strip with a register-tiled FFMA GEMM (8x8 outputs per thread), folds
// each tile into a running row maximum (scaled by the strip token's inverse norm), and finally sums
// its 128 row maxima (scaled by the question token's inverse norm).  The per-strip sum over the
// question's row slices is folded in fixed order by the last block to finish (deterministic).
//
// This is the one dense contraction of the path that must stay fp32 for parity, so it is bound by
// the CUDA-core FFMA rate, not HBM (both operands of a document are L2-resident).  The bf16 tensor-core
// variant lives in maxsim_tc.cu.
#include "rdv_common.cuh"

namespace rdv {

constexpr int kMsBM = 128, kMsBN = 128, kMsBK = 16, kMsThreads = 256;
constexpr int kMsLd = kMsBM + 4;   // padded leading dimension of the k-major shared tiles

// one warp per row: inv[r] = 1 / max(||x_r||, 1e-12)
__global__ void __launch_bounds__(256) row_inv_norm_kernel(const float* __restrict__ x, int64_t rows, int d,
                                                           float* __restrict__ inv) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= rows) return;
    const float4* src = reinterpret_cast<const float4*>(x) + row * (d >> 2);
    float ss = 0.f;
    for (int i = lane; i < (d >> 2); i += 32) {
        const float4 v = ldg_stream(src + i);
        ss = fmaf(v.x, v.x, ss); ss = fmaf(v.y, v.y, ss); ss = fmaf(v.z, v.z, ss); ss = fmaf(v.w, v.w, ss);
    }
    ss = warp_sum(ss);
    if (lane == 0) inv[row] = __fdiv_rn(1.0f, fmaxf(__fsqrt_rn(ss), 1e-12f));
}

__device__ __forceinline__ float nan_max(float a, float b) {   // torch.max propagates NaN
    return (a > b || a != a) ? a : b;
}

struct MaxSimParams {
    const float* q;       // (Lq, d)
    const float* p;       // (n, Lp, d)
    const float* inv_q;   // (Lq)
    const float* inv_p;   // (n * Lp)
    int32_t n, Lq, Lp, d;
    float* partial;       // (n, tiles_i) row-slice sums
    int32_t* counter;     // (n) zero on entry, left zero
    float* out;           // (n)
};

__global__ void __launch_bounds__(kMsThreads, 2) maxsim_f32_kernel(const MaxSimParams p) {
    __shared__ __align__(16) float As[2][kMsBK * kMsLd];
    __shared__ __align__(16) float Bs[2][kMsBK * kMsLd];
    __shared__ float s_rowsum[kMsThreads / 32];
    __shared__ int s_last;

    const int tid = threadIdx.x;
    const int strip = blockIdx.y;
    const int i0 = blockIdx.x * kMsBM;
    const int tiles_i = gridDim.x;
    const int d4 = p.d >> 2;
    const int k_tiles = (p.d + kMsBK - 1) / kMsBK;

    // global -> shared loader mapping: 4 adjacent lanes read one row's 64-byte k-slice
    const int ld_row = tid >> 2;          // 0..63 (+64 for the second half)
    const int ld_kq = tid & 3;            // float4 index inside the 16-wide k slice
    // compute mapping: 2x2 blocks of 4x4 (rows ty*4 / ty*4+64, cols tx*4 / tx*4+64)
    const int ty = tid >> 4, tx = tid & 15;

    const float4* Qg = reinterpret_cast<const float4*>(p.q);
    const float4* Pg = reinterpret_cast<const float4*>(p.p) + (size_t)strip * p.Lp * d4;
    const float* inv_p = p.inv_p + (size_t)strip * p.Lp;

    float rowmax[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) rowmax[r] = -INFINITY;

    const int tiles_j = (p.Lp + kMsBN - 1) / kMsBN;
    for (int tj = 0; tj < tiles_j; ++tj) {
        const int j0 = tj * kMsBN;
        float acc[8][8];
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int c = 0; c < 8; ++c) acc[r][c] = 0.f;

        float4 ra[2], rb[2];
        auto load_tile = [&](int kt) {
            const int kf4 = kt * (kMsBK / 4) + ld_kq;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int ri = i0 + ld_row + 64 * h, rj = j0 + ld_row + 64 * h;
                ra[h] = (ri < p.Lq && kf4 < d4) ? Qg[(size_t)ri * d4 + kf4] : make_float4(0.f, 0.f, 0.f, 0.f);
                rb[h] = (rj < p.Lp && kf4 < d4) ? Pg[(size_t)rj * d4 + kf4] : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        };
        auto store_tile = [&](int buf) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int r = ld_row + 64 * h;
                float* a = &As[buf][(ld_kq * 4) * kMsLd + r];
                float* b = &Bs[buf][(ld_kq * 4) * kMsLd + r];
                a[0] = ra[h].x; a[kMsLd] = ra[h].y; a[2 * kMsLd] = ra[h].z; a[3 * kMsLd] = ra[h].w;
                b[0] = rb[h].x; b[kMsLd] = rb[h].y; b[2 * kMsLd] = rb[h].z; b[3 * kMsLd] = rb[h].w;
            }
        };
        load_tile(0);
        __syncthreads();                 // previous j-tile's readers are done with both buffers
        store_tile(0);
        __syncthreads();
        for (int kt = 0; kt < k_tiles; ++kt) {
            const int buf = kt & 1;
            if (kt + 1 < k_tiles) load_tile(kt + 1);
#pragma unroll
            for (int k = 0; k < kMsBK; ++k) {
                const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k * kMsLd + ty * 4]);
                const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k * kMsLd + ty * 4 + 64]);
                const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k * kMsLd + tx * 4]);
                const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k * kMsLd + tx * 4 + 64]);
                const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                for (int r = 0; r < 8; ++r)
#pragma unroll
                    for (int c = 0; c < 8; ++c) acc[r][c] = fmaf(a[r], b[c], acc[r][c]);
            }
            if (kt + 1 < k_tiles) {
                store_tile(buf ^ 1);     // buf^1 was last read in iteration kt-1, fenced by the barrier below
                __syncthreads();
            }
        }
        // fold this tile into the running row maxima (scaled by the strip token's inverse norm)
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const int j = j0 + tx * 4 + (c & 3) + 64 * (c >> 2);
            if (j < p.Lp) {
                const float ip = inv_p[j];
#pragma unroll
                for (int r = 0; r < 8; ++r) rowmax[r] = nan_max(rowmax[r], acc[r][c] * ip);
            }
        }
    }

    // max across the 16 threads (tx) that share each row, then scale by the question token's inverse
    // norm and sum the block's 128 rows in fixed order
    float mine = 0.f;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        float m = rowmax[r];
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) m = nan_max(m, __shfl_xor_sync(0xffffffffu, m, o));
        const int i = i0 + ty * 4 + (r & 3) + 64 * (r >> 2);
        if (tx == 0 && i < p.Lq) mine += m * p.inv_q[i];
    }
    // lanes 0 and 16 of each warp hold partial sums (ty even / odd)
    mine += __shfl_xor_sync(0xffffffffu, mine, 16);
    if ((tid & 31) == 0) s_rowsum[tid >> 5] = mine;
    __syncthreads();
    if (tid == 0) {
        float s = 0.f;
        for (int w = 0; w < kMsThreads / 32; ++w) s += s_rowsum[w];
        p.partial[(size_t)strip * tiles_i + blockIdx.x] = s;
        __threadfence();
        const int prev = atomicAdd(p.counter + strip, 1);
        s_last = (prev == tiles_i - 1);
        __threadfence();
    }
    __syncthreads();
    if (s_last && tid == 0) {
        float s = 0.f;
        for (int t = 0; t < tiles_i; ++t) s += __ldcg(p.partial + (size_t)strip * tiles_i + t);
        p.out[strip] = s;
        p.counter[strip] = 0;
    }
}

}  // namespace rdv

extern "C" int rdv_row_inv_norm_f32(const float* d_x, int64_t rows, int32_t d, float* d_inv, void* stream) {
    using namespace rdv;
    RDV_REQUIRE(rows >= 0, RDV_E_INVALID, "row_inv_norm_f32: negative size");
    if (rows == 0) return RDV_OK;
    RDV_REQUIRE(d_x && d_inv, RDV_E_INVALID, "row_inv_norm_f32: null pointer");
    RDV_REQUIRE(d >= 4 && (d & 3) == 0, RDV_E_INVALID, "row_inv_norm_f32: d=%d must be a multiple of 4", d);
    RDV_REQUIRE(aligned16(d_x), RDV_E_ALIGN, "row_inv_norm_f32: x not 16-byte aligned");
    const int64_t blocks = (rows + 7) / 8;
    RDV_REQUIRE(blocks < (1ll << 31), RDV_E_LIMIT, "row_inv_norm_f32: too many rows");
    row_inv_norm_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(d_x, rows, d, d_inv);
    RDV_LAUNCH_CHECK("row_inv_norm_kernel");
    return RDV_OK;
}

extern "C" int rdv_maxsim_f32(const float* d_q, const float* d_p, const float* d_inv_q, const float* d_inv_p,
                              int32_t n, int32_t Lq, int32_t Lp, int32_t d, float* d_partial,
                              int32_t* d_counter, float* d_out, void* stream) {
    using namespace rdv;
    RDV_REQUIRE(n >= 0 && Lq >= 0, RDV_E_INVALID, "maxsim_f32: negative size");
    if (n == 0) return RDV_OK;
    RDV_REQUIRE(Lp >= 1, RDV_E_INVALID, "maxsim_f32: a strip needs at least one token (max over an empty set)");
    RDV_REQUIRE(d_q && d_p && d_inv_q && d_inv_p && d_partial && d_counter && d_out, RDV_E_INVALID,
                "maxsim_f32: null pointer");
    RDV_REQUIRE(d >= 4 && (d & 3) == 0, RDV_E_INVALID, "maxsim_f32: d=%d must be a multiple of 4", d);
    RDV_REQUIRE(aligned16(d_q) && aligned16(d_p), RDV_E_ALIGN, "maxsim_f32: q/p not 16-byte aligned");
    RDV_REQUIRE(n <= 65535, RDV_E_LIMIT, "maxsim_f32: n=%d strips > 65535 per launch", n);
    MaxSimParams p;
    p.q = d_q; p.p = d_p; p.inv_q = d_inv_q; p.inv_p = d_inv_p; p.n = n; p.Lq = Lq; p.Lp = Lp; p.d = d;
    p.partial = d_partial; p.counter = d_counter; p.out = d_out;
    const int tiles_i = Lq > 0 ? (Lq + kMsBM - 1) / kMsBM : 1;
    dim3 grid(tiles_i, n);
    maxsim_f32_kernel<<<grid, kMsThreads, 0, static_cast<cudaStream_t>(stream)>>>(p);
    RDV_LAUNCH_CHECK("maxsim_f32_kernel");
    return RDV_OK;
}

extern "C" int32_t rdv_maxsim_tiles_i(int32_t Lq) { return Lq > 0 ? (Lq + rdv::kMsBM - 1) / rdv::kMsBM : 1; }
