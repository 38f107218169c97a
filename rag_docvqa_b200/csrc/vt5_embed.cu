// Generator-input embeddings on the device, straight from the gather's input_ids / boxes / layout labels -- sm_100a.
//
// Replaces the last stage of VT5.prepare_inputs_for_vqa (reference src/VT5.py:194-204):
//     input_embeds = shared(ids) + SpatialEmbeddings(boxes) [+ layout_embedding(labels) * layout_embedding_scale]
// with SpatialEmbeddings.forward (src/_modules.py:70-86, inference: dropout is the identity)
//     s = x_emb[l] + y_emb[u] + x_emb[r] + y_emb[b];   spatial = Linear(LayerNorm(s))
// The reference runs it as ~15 torch kernels that each materialise a (B, L, D) temporary, including a (B*L, D) x (D, D)
// GEMM.  Both LayerNorm's centring and the Linear are LINEAR in s, and s is a sum of four table rows, so the GEMM moves
// into the tables (built once per model by rdv_vt5_embed_tables_build, fp64 accumulation):
//     x~_i = x_emb[i] - mean(x_emb[i])                      s - mean(s) = x~_l + y~_u + x~_r + y~_b     (exactly)
//     XW[i] = (x~_i * gamma) W^T,  YW likewise              c = beta W^T + bias
//     var(s) = |x~_l + y~_u + x~_r + y~_b|^2 / D            = ten entries of the Gram tables x~x~^T, x~y~^T, y~y~^T
//     spatial = (XW[l] + YW[u] + XW[r] + YW[b]) * rsqrt(var + eps) + c
// One pass: per token four projected rows (L2-resident tables, 3 MB each at D = 768), ten scalars, the token's own
// embedding row, one row written.  HBM-bound on the output + the token-table rows, L2-bound on the coordinate rows; no
// tensor cores, by construction.  A thread owns one float4 column of the rows and keeps the spatial value of its column in
// a register while the box repeats (all tokens of a word, the whole prompt and the whole padding share one box).
#include "rdv_common.cuh"

namespace rdv {

constexpr int kEmbChunk = 32;          // tokens whose scalars one producer warp prepares at a time (lane = token)
constexpr int kEmbDepth = 4;           // tokens a consumer thread keeps in flight (register ring)
constexpr int kEmbMaxCols = 256;       // consumer threads = D / 4 <= 256
constexpr int kEmbLayoutSmem = 48 * 1024;

struct EmbedParams {
    rdv_vt5_embed_tables t;
    const int64_t* ids;      // (B, L) rows `ld` apart; null: spatial embedding only
    const int64_t* boxes;    // (B, L, 4) rows `ld * 4` apart
    const int64_t* labels;   // (B, L) rows `ld` apart or null
    int32_t B, L;
    int64_t ld;
    float* out;              // (B, out_ld, D): row b of the tokens starts b * out_ld rows in (out_ld = L: contiguous)
    int64_t out_ld;
    int32_t* bad;            // |= 1 box coordinate, |= 2 token id, |= 4 layout label outside its table
    int32_t layout_in_smem;  // the layout table fits kEmbLayoutSmem: staged once per block
    int32_t* work;           // {next chunk, blocks done}: zero before the launch, zero again after it
};

struct __align__(16) TokMeta {   // what the consumers need about one token: two 16-byte shared-memory loads
    int l, u, r, b;
    int id;
    int lab_same;                // label << 1 | "same box as the previous token of the chunk"
    int out_tok;                 // row of the output this token is written to: b * out_ld + t
    float rstd;
};

__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ uint32_t emb_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void emb_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(emb_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void emb_mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(emb_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void emb_mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(emb_smem_u32(bar)), "r"(parity) : "memory");
    } while (!done);
}

// Persistent.  A block = D / 4 consumer threads, one float4 COLUMN each, + one producer warp; chunks of 32 tokens are
// handed out by an atomic counter (a chunk of padding costs a fifth of a chunk of words: a static split left SMs idle).
//   producer warp: claims the next chunk and describes it, up to kEmbSlots chunks ahead.  Lane = token: ids, box and label,
//     validation, the ten Gram entries of the variance (ten independent loads per lane: one round for 32 tokens),
//     1 / sqrt(var + eps), "same box as the previous token" -> a shared-memory slot, published through an mbarrier.
//   consumer thread: a register ring of kEmbDepth tokens in flight (the four coordinate rows' and the token row's float4 of
//     ITS column: 5 independent 128-bit loads per token, coalesced across the block), adds, scales, writes its column of
//     the output row; the ring runs on into the next chunk.  A repeated box (all tokens of a word, the prompt, the padding)
//     costs one load: the spatial value of the column stays in a register.  A warp releases a slot when it is through
//     with it: consumer warps never wait for each other.
// History, 64 x 512 tokens at D = 768: a warp per token with the rows in registers (four rounds of dependent loads per
// token) 122 us; per-warp shared-memory rings filled by 1-D bulk copies (TMA) or by cp.async 67-77 us -- six warps per SM
// spent ~460 instructions per token on addressing and issue, and the rings sat half empty; a thread per column with a
// static split of the tokens and a __syncthreads per chunk 59 us (26 % of the stall samples at the barrier); this kernel
// with the thread's ring in shared memory instead of registers (cp.async per thread, wait_group): 81 us; coordinate rows
// in registers + token rows in a cp.async ring + three blocks per SM (80 registers, spills): 103 us.
// scripts/probe_embed_parts.py takes the 59 us apart: writing the output alone 17 us, the kernel with no row to load 28 us,
// with every load a cache hit 44 us, coordinate rows from L2 +7 us, token rows from L2 / HBM +7 us.
constexpr int kEmbSlots = 4;

template <bool LAYOUT, bool WIDE>      // WIDE: D > 896 (more than 224 consumer threads): one block per SM
__global__ void __launch_bounds__(WIDE ? kEmbMaxCols + 32 : 256, WIDE ? 1 : 2) vt5_embed_kernel(const EmbedParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];              // the layout table, when it is staged
    __shared__ TokMeta s_meta[kEmbSlots][kEmbChunk];
    __shared__ int s_chunk[kEmbSlots];                                      // the chunk a slot describes, -1: no more chunks
    __shared__ __align__(8) uint64_t s_full[kEmbSlots], s_empty[kEmbSlots];
    const rdv_vt5_embed_tables& T = p.t;
    const int d4 = T.D >> 2, np = T.n_pos;
    const int n_cons = (int)blockDim.x - 32;                                // consumer threads (>= d4, multiple of 32)
    const int tid = threadIdx.x;
    const int64_t n = (int64_t)p.B * p.L;
    const int n_chunks = (int)((n + kEmbChunk - 1) / kEmbChunk);

    if (tid == 0) {
        for (int i = 0; i < kEmbSlots; ++i) {
            emb_mbar_init(&s_full[i], 1);
            emb_mbar_init(&s_empty[i], (uint32_t)(n_cons >> 5));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (LAYOUT && p.layout_in_smem) {
        float4* dst = reinterpret_cast<float4*>(smem_raw);
        const float4* src = reinterpret_cast<const float4*>(T.layout);
        for (int i = tid; i < T.n_labels * d4; i += blockDim.x) dst[i] = __ldg(src + i);
    }
    __syncthreads();

    if (tid >= n_cons) {
        // ------------------------------------------------ producer warp ------------------------------------------------
        const int lane = tid - n_cons;
        int bad = 0;
        for (int k = 0;; ++k) {
            const int slot = k % kEmbSlots;
            if (k >= kEmbSlots) emb_mbar_wait(&s_empty[slot], (uint32_t)(k / kEmbSlots - 1) & 1u);
            int chunk = 0;
            if (lane == 0) chunk = atomicAdd(p.work, 1);
            chunk = __shfl_sync(0xffffffffu, chunk, 0);
            if (chunk >= n_chunks) {
                if (lane == 0) {
                    s_chunk[slot] = -1;
                    emb_mbar_arrive(&s_full[slot]);
                }
                break;
            }
            TokMeta m = {};
            const int64_t tok = (int64_t)chunk * kEmbChunk + lane;
            if (tok < n) {
                const int64_t src = (tok / p.L) * p.ld + (tok % p.L);
                const longlong2* bx = reinterpret_cast<const longlong2*>(p.boxes + src * 4);
                const longlong2 b0 = bx[0], b1 = bx[1];
                long long id = p.ids ? p.ids[src] : 0, lab = p.labels ? p.labels[src] : 0;
                long long c4[4] = {b0.x, b0.y, b1.x, b1.y};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    if ((unsigned long long)c4[i] >= (unsigned long long)np) {  // nn.Embedding would raise: flag, stay in bounds
                        bad |= 1;
                        c4[i] = min(max(c4[i], 0ll), (long long)np - 1);
                    }
                }
                if (p.ids && (unsigned long long)id >= (unsigned long long)T.V) { bad |= 2; id = 0; }
                if (p.labels && (unsigned long long)lab >= (unsigned long long)T.n_labels) { bad |= 4; lab = 0; }
                m.l = (int)c4[0]; m.u = (int)c4[1]; m.r = (int)c4[2]; m.b = (int)c4[3]; m.id = (int)id;
                m.lab_same = (int)lab << 1;
                m.out_tok = (int)((tok / p.L) * p.out_ld + (tok % p.L));
                // |x~_l + y~_u + x~_r + y~_b|^2: four squares and six cross terms, summed in fp64
                const size_t lo = (size_t)m.l * np, uo = (size_t)m.u * np, ro = (size_t)m.r * np, bo = (size_t)m.b * np;
                const float g0 = __ldg(T.gxx + lo + m.l), g1 = __ldg(T.gyy + uo + m.u), g2 = __ldg(T.gxx + ro + m.r);
                const float g3 = __ldg(T.gyy + bo + m.b), g4 = __ldg(T.gxy + lo + m.u), g5 = __ldg(T.gxx + lo + m.r);
                const float g6 = __ldg(T.gxy + lo + m.b), g7 = __ldg(T.gxy + ro + m.u), g8 = __ldg(T.gyy + uo + m.b);
                const float g9 = __ldg(T.gxy + ro + m.b);
                const double sq = ((double)g0 + (double)g1) + ((double)g2 + (double)g3);
                const double cross = (((double)g4 + (double)g5) + ((double)g6 + (double)g7)) + ((double)g8 + (double)g9);
                m.rstd = (float)(1.0 / sqrt(fmax(sq + 2.0 * cross, 0.0) / (double)T.D + (double)T.eps));
            }
            const int pl = __shfl_up_sync(0xffffffffu, m.l, 1), pu = __shfl_up_sync(0xffffffffu, m.u, 1);
            const int pr = __shfl_up_sync(0xffffffffu, m.r, 1), pb = __shfl_up_sync(0xffffffffu, m.b, 1);
            m.lab_same |= (lane > 0 && m.l == pl && m.u == pu && m.r == pr && m.b == pb) ? 1 : 0;   // a chunk's first token always loads
            s_meta[slot][lane] = m;
            if (lane == 0) s_chunk[slot] = chunk;
            __syncwarp();
            if (lane == 0) emb_mbar_arrive(&s_full[slot]);  // release: the slot's contents are visible to whoever sees the flip
        }
        if (bad && p.bad) atomicOr(p.bad, bad);
        if (lane == 0) {
            // every block passes here only after the counter has run out: the last one puts the workspace back to zero
            __threadfence();
            if (atomicAdd(p.work + 1, 1) == (int)gridDim.x - 1) {
                atomicExch(p.work, 0);
                atomicExch(p.work + 1, 0);
            }
        }
        return;
    }

    // ---------------------------------------------------- consumers ----------------------------------------------------
    const int c = tid;                                       // this thread's float4 column
    const bool active = c < d4;
    const int lane = tid & 31;
    const float4* xw = reinterpret_cast<const float4*>(T.xw) + c;
    const float4* yw = reinterpret_cast<const float4*>(T.yw) + c;
    const float4* sem = reinterpret_cast<const float4*>(T.shared) + c;
    const float4* lay_g = reinterpret_cast<const float4*>(T.layout) + c;
    const float4* lay_s = reinterpret_cast<const float4*>(smem_raw) + c;
    const float4 cc = active ? __ldg(reinterpret_cast<const float4*>(T.c) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float scale = T.layout_scale;
    const bool has_ids = p.ids != nullptr;
    float4* out = reinterpret_cast<float4*>(p.out) + c;

    float4 buf[kEmbDepth][5];
    float4 sp = cc;
    auto load = [&](int slot, const TokMeta& m) {            // the five loads of a token
        const int4 lurb = *reinterpret_cast<const int4*>(&m.l);
        const int4 rest = *reinterpret_cast<const int4*>(&m.id);            // id, label / same, output row, rstd
        if (!(rest.y & 1)) {
            buf[slot][0] = ldg_stream(xw + (size_t)lurb.x * d4); buf[slot][1] = ldg_stream(yw + (size_t)lurb.y * d4);
            buf[slot][2] = ldg_stream(xw + (size_t)lurb.z * d4); buf[slot][3] = ldg_stream(yw + (size_t)lurb.w * d4);
        }
        if (has_ids) buf[slot][4] = ldg_stream(sem + (size_t)rest.x * d4);
    };
    auto consume = [&](int slot, const TokMeta& m) {
        const int4 rest = *reinterpret_cast<const int4*>(&m.id);
        float4* dst = out + (size_t)rest.z * d4;
        if (!(rest.y & 1)) {
            const float rstd = __int_as_float(rest.w);
            const float4 a0 = buf[slot][0], a1 = buf[slot][1], a2 = buf[slot][2], a3 = buf[slot][3];
            sp.x = fmaf(((a0.x + a1.x) + a2.x) + a3.x, rstd, cc.x);
            sp.y = fmaf(((a0.y + a1.y) + a2.y) + a3.y, rstd, cc.y);
            sp.z = fmaf(((a0.z + a1.z) + a2.z) + a3.z, rstd, cc.z);
            sp.w = fmaf(((a0.w + a1.w) + a2.w) + a3.w, rstd, cc.w);
        }
        float4 o = sp;
        if (has_ids) {                                       // semantic + spatial (src/VT5.py:202)
            const float4 e = buf[slot][4];
            o.x = __fadd_rn(e.x, o.x); o.y = __fadd_rn(e.y, o.y); o.z = __fadd_rn(e.z, o.z); o.w = __fadd_rn(e.w, o.w);
        }
        if (LAYOUT) {                                        // + layout * scale, the product rounded first (src/VT5.py:204)
            const int lab = rest.y >> 1;
            const float4 e = p.layout_in_smem ? lay_s[(size_t)lab * d4] : __ldg(lay_g + (size_t)lab * d4);
            o.x = __fadd_rn(o.x, __fmul_rn(e.x, scale)); o.y = __fadd_rn(o.y, __fmul_rn(e.y, scale));
            o.z = __fadd_rn(o.z, __fmul_rn(e.z, scale)); o.w = __fadd_rn(o.w, __fmul_rn(e.w, scale));
        }
        __stcs(dst, o);                                      // written once, read by the next model stage
    };
    auto tokens_of = [&](int chunk) { return chunk < 0 ? 0 : (int)min((int64_t)kEmbChunk, n - (int64_t)chunk * kEmbChunk); };

    emb_mbar_wait(&s_full[0], 0);
    int chunk = s_chunk[0];
    int ntok = tokens_of(chunk);
    if (active) {
#pragma unroll
        for (int i = 0; i < kEmbDepth; ++i)
            if (i < ntok) load(i, s_meta[0][i]);
    }
    for (int k = 0; chunk >= 0; ++k) {
        const int slot = k % kEmbSlots, nslot = (k + 1) % kEmbSlots;
        const TokMeta* cur = s_meta[slot];
        const TokMeta* nxt = s_meta[nslot];
        int nchunk = -1, nntok = 0;
#pragma unroll 1
        for (int q0 = 0; q0 < kEmbChunk; q0 += kEmbDepth) {
            if (q0 == kEmbChunk - kEmbDepth) {               // the ring is about to run on into the next chunk
                emb_mbar_wait(&s_full[nslot], (uint32_t)((k + 1) / kEmbSlots) & 1u);
                nchunk = s_chunk[nslot];
                nntok = tokens_of(nchunk);
            }
            if (active) {
#pragma unroll
                for (int i = 0; i < kEmbDepth; ++i) {        // one revolution of the ring: slots are literals
                    const int q = q0 + i;
                    if (q < ntok) consume(i, cur[q]);
                    const int kk = q + kEmbDepth;
                    if (kk < kEmbChunk) {
                        if (kk < ntok) load(i, cur[kk]);
                    } else if (kk - kEmbChunk < nntok) {
                        load(i, nxt[kk - kEmbChunk]);
                    }
                }
            }
        }
        __syncwarp();
        if (lane == 0) emb_mbar_arrive(&s_empty[slot]);      // this warp is through with the slot
        chunk = nchunk;
        ntok = nntok;
    }
}

// out[i][j] = sum_k (A[i][k] - ma[i]) * g[k] * (Bm[j][k] - mb[j]) (+ add[j]), fp64 accumulation, fp32 result.
// ma / g / mb / add may be null (0 / 1 / 0 / 0).  Runs once per model: clarity over speed.
constexpr int kTile = 16;
__global__ void __launch_bounds__(kTile * kTile) abt_f64_kernel(const float* __restrict__ A, const double* __restrict__ ma,
                                                                const float* __restrict__ g, const float* __restrict__ Bm,
                                                                const double* __restrict__ mb, const float* __restrict__ add,
                                                                int M, int N, int K, float* __restrict__ out) {
    __shared__ double sa[kTile][kTile + 1], sb[kTile][kTile + 1];
    const int tx = threadIdx.x % kTile, ty = threadIdx.x / kTile;
    const int i = blockIdx.y * kTile + ty, j = blockIdx.x * kTile + tx;
    const int ia = blockIdx.y * kTile + ty, jb = blockIdx.x * kTile + ty;     // rows this thread stages
    double acc = 0.0;
    for (int k0 = 0; k0 < K; k0 += kTile) {
        const int k = k0 + tx;
        double a = 0.0, b = 0.0;
        if (k < K) {
            if (ia < M) a = ((double)A[(size_t)ia * K + k] - (ma ? ma[ia] : 0.0)) * (g ? (double)g[k] : 1.0);
            if (jb < N) b = (double)Bm[(size_t)jb * K + k] - (mb ? mb[jb] : 0.0);
        }
        sa[ty][tx] = a;
        sb[ty][tx] = b;
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < kTile; ++kk) acc = fma(sa[ty][kk], sb[tx][kk], acc);
        __syncthreads();
    }
    if (i < M && j < N) out[(size_t)i * N + j] = (float)(acc + (add ? (double)add[j] : 0.0));
}

__global__ void __launch_bounds__(256) row_mean_f64_kernel(const float* __restrict__ x, int rows, int D, double* __restrict__ mean) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rows) return;
    double s = 0.0;
    for (int k = lane; k < D; k += 32) s += (double)x[(size_t)row * D + k];
    s = warp_sum_f64(s);
    if (lane == 0) mean[row] = s / (double)D;
}

static int launch_abt(const float* A, const double* ma, const float* g, const float* Bm, const double* mb, const float* add,
                      int M, int N, int K, float* out, cudaStream_t s) {
    dim3 grid((N + kTile - 1) / kTile, (M + kTile - 1) / kTile);
    abt_f64_kernel<<<grid, kTile * kTile, 0, s>>>(A, ma, g, Bm, mb, add, M, N, K, out);
    RDV_LAUNCH_CHECK("abt_f64_kernel");
    return RDV_OK;
}

}  // namespace rdv

extern "C" int rdv_vt5_embed_tables_build(const float* d_x_emb, const float* d_y_emb, int32_t n_pos, int32_t D,
                                          const float* d_ln_weight, const float* d_ln_bias, const float* d_lin_weight,
                                          const float* d_lin_bias, double* d_ws_means, float* d_xw, float* d_yw,
                                          float* d_gxx, float* d_gxy, float* d_gyy, float* d_c, void* stream) {
    using namespace rdv;
    RDV_REQUIRE(n_pos >= 1 && D >= 4 && D <= 1024 && (D & 3) == 0, RDV_E_INVALID,
                "vt5_embed_tables_build: n_pos=%d, D=%d (D must be a multiple of 4 in [4, 1024])", n_pos, D);
    RDV_REQUIRE(d_x_emb && d_y_emb && d_ln_weight && d_ln_bias && d_lin_weight && d_ws_means && d_xw && d_yw && d_gxx &&
                d_gxy && d_gyy && d_c, RDV_E_INVALID, "vt5_embed_tables_build: null pointer");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    double* mx = d_ws_means;
    double* my = d_ws_means + n_pos;
    row_mean_f64_kernel<<<(n_pos + 7) / 8, 256, 0, s>>>(d_x_emb, n_pos, D, mx);
    RDV_LAUNCH_CHECK("row_mean_f64_kernel");
    row_mean_f64_kernel<<<(n_pos + 7) / 8, 256, 0, s>>>(d_y_emb, n_pos, D, my);
    RDV_LAUNCH_CHECK("row_mean_f64_kernel");
    int rc;
    if ((rc = launch_abt(d_x_emb, mx, d_ln_weight, d_lin_weight, nullptr, nullptr, n_pos, D, D, d_xw, s))) return rc;
    if ((rc = launch_abt(d_y_emb, my, d_ln_weight, d_lin_weight, nullptr, nullptr, n_pos, D, D, d_yw, s))) return rc;
    if ((rc = launch_abt(d_x_emb, mx, nullptr, d_x_emb, mx, nullptr, n_pos, n_pos, D, d_gxx, s))) return rc;
    if ((rc = launch_abt(d_x_emb, mx, nullptr, d_y_emb, my, nullptr, n_pos, n_pos, D, d_gxy, s))) return rc;
    if ((rc = launch_abt(d_y_emb, my, nullptr, d_y_emb, my, nullptr, n_pos, n_pos, D, d_gyy, s))) return rc;
    return launch_abt(d_ln_bias, nullptr, nullptr, d_lin_weight, nullptr, d_lin_bias, 1, D, D, d_c, s);
}

extern "C" int rdv_vt5_input_embeds_f32(const rdv_vt5_embed_tables* t, const int64_t* d_ids, const int64_t* d_boxes,
                                        const int64_t* d_labels, int32_t B, int32_t L, int64_t ld, float* d_out,
                                        int64_t out_ld, int32_t* d_bad, int32_t* d_work, void* stream) {
    using namespace rdv;
    RDV_REQUIRE(t, RDV_E_INVALID, "vt5_input_embeds_f32: null tables");
    RDV_REQUIRE(B >= 0 && L >= 0 && ld >= L && out_ld >= L, RDV_E_INVALID, "vt5_input_embeds_f32: B=%d, L=%d, ld=%lld, out_ld=%lld", B, L,
                (long long)ld, (long long)out_ld);
    RDV_REQUIRE((int64_t)B * out_ld < (1ll << 31), RDV_E_LIMIT, "vt5_input_embeds_f32: %lld output rows", (long long)B * out_ld);
    RDV_REQUIRE(t->D >= 4 && t->D <= 1024 && (t->D & 3) == 0 && t->n_pos >= 1, RDV_E_INVALID,
                "vt5_input_embeds_f32: D=%d must be a multiple of 4 in [4, 1024]", t->D);
    if ((int64_t)B * L == 0) return RDV_OK;
    RDV_REQUIRE(d_work, RDV_E_INVALID, "vt5_input_embeds_f32: null workspace");
    RDV_REQUIRE(d_boxes && d_out && t->xw && t->yw && t->gxx && t->gxy && t->gyy && t->c, RDV_E_INVALID,
                "vt5_input_embeds_f32: null pointer");
    RDV_REQUIRE(!d_ids || (t->shared && t->V >= 1), RDV_E_INVALID, "vt5_input_embeds_f32: input ids without a token table");
    RDV_REQUIRE(!d_labels || (t->layout && t->n_labels >= 1), RDV_E_INVALID, "vt5_input_embeds_f32: labels without a layout table");
    RDV_REQUIRE(aligned16(d_boxes) && aligned16(d_out) && aligned16(t->xw) && aligned16(t->yw) && aligned16(t->c) &&
                (!d_ids || aligned16(t->shared)) && (!d_labels || aligned16(t->layout)), RDV_E_ALIGN,
                "vt5_input_embeds_f32: boxes, out and the tables must be 16-byte aligned");
    EmbedParams p;
    p.t = *t; p.ids = d_ids; p.boxes = d_boxes; p.labels = d_labels; p.B = B; p.L = L; p.ld = ld; p.out = d_out; p.out_ld = out_ld; p.bad = d_bad; p.work = d_work;
    const int d4 = t->D / 4;
    const int threads = (d4 + 31) / 32 * 32 + 32;           // one float4 column per consumer thread + the producer warp
    const size_t layout_bytes = d_labels ? (size_t)t->n_labels * t->D * 4 : 0;
    p.layout_in_smem = d_labels && layout_bytes <= (size_t)kEmbLayoutSmem;
    const size_t smem = p.layout_in_smem ? layout_bytes : 0;
    const int64_t n = (int64_t)B * L;
    const int64_t chunks = (n + kEmbChunk - 1) / kEmbChunk;
    RDV_REQUIRE(chunks < (1ll << 31), RDV_E_LIMIT, "vt5_input_embeds_f32: %lld tokens", (long long)n);
    int64_t blocks = (threads > 256 ? 1 : 2) * (int64_t)sm_count();      // every resident slot of the device, once
    if (blocks > chunks) blocks = chunks;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const bool wide = threads > 256;
    if (d_labels && wide) {
        RDV_ONCE_PER_DEVICE(cudaFuncSetAttribute(vt5_embed_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kEmbLayoutSmem),
                            "cudaFuncSetAttribute(vt5_embed)");
        vt5_embed_kernel<true, true><<<(unsigned)blocks, threads, smem, s>>>(p);
    } else if (d_labels) {
        RDV_ONCE_PER_DEVICE(cudaFuncSetAttribute(vt5_embed_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kEmbLayoutSmem),
                            "cudaFuncSetAttribute(vt5_embed)");
        vt5_embed_kernel<true, false><<<(unsigned)blocks, threads, smem, s>>>(p);
    } else if (wide) {
        vt5_embed_kernel<false, true><<<(unsigned)blocks, threads, 0, s>>>(p);
    } else {
        vt5_embed_kernel<false, false><<<(unsigned)blocks, threads, 0, s>>>(p);
    }
    RDV_LAUNCH_CHECK("vt5_embed_kernel");
    return RDV_OK;
}
