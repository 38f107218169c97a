// Stand-alone probe (not part of the product): what does a ~32 MB streaming pass cost on B200, and which
// launch shape of the cosine-score kernel gets closest to it?   nvcc -O3 -gencode arch=compute_100a,code=sm_100a
//   K0  pure read (float4 loads, sum)                      -- the floor for a given grid shape
//   K1  block-per-tile score kernel (the product's r1 shape), tile_rows / ROWS / blocks-per-SM sweep
//   K2  warp-unit persistent score kernel: every warp owns a contiguous run of ROWS-row units, next unit's
//       loads issued before the current unit is reduced
// Each config: 200 back-to-back launches rotating over 9 distinct buffers (> 2x L2), CUDA events.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

struct Tile { const void* src; long long sims_off; int rows, doc, doc_rows, reserved; };

__device__ __forceinline__ float4 ldg_stream(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float cosine(float dot, float ss_e, float ss_q) {
    return __fdiv_rn(dot, __fadd_rn(__fmul_rn(__fsqrt_rn(ss_e), __fsqrt_rn(ss_q)), 1e-8f));
}
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---- K0: pure read -----------------------------------------------------------------------------------
template <int U>
__global__ void __launch_bounds__(256) k0_read(const float4* __restrict__ src, size_t n4, float* out, int pdl) {
    if (pdl) { pdl_launch(); pdl_wait(); }
    const size_t stride = (size_t)gridDim.x * 256;
    size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    float acc = 0.f;
    for (; i + (U - 1) * stride < n4; i += U * stride) {
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = ldg_stream(src + i + u * stride);
#pragma unroll
        for (int u = 0; u < U; ++u) acc += v[u].x + v[u].y + v[u].z + v[u].w;
    }
    for (; i < n4; i += stride) { float4 v = ldg_stream(src + i); acc += v.x + v.y + v.z + v.w; }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0 && acc == 123.456f) out[0] = acc;
}

// ---- K1: block per tile ----------------------------------------------------------------------------------
template <int VPL, int ROWS, int MINB>
__global__ void __launch_bounds__(256, MINB) k1_tile(const Tile* __restrict__ tiles, const float* __restrict__ q, int d,
                                                     float* __restrict__ sims, int pdl) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int d4 = d >> 2;
    if (pdl) { pdl_launch(); pdl_wait(); }
    const Tile t = tiles[blockIdx.x];
    const float4* __restrict__ E = reinterpret_cast<const float4*>(t.src);
    const float4* __restrict__ Q = reinterpret_cast<const float4*>(q) + (size_t)t.doc * d4;
    float* __restrict__ out = sims + t.sims_off;
    float4 qv[VPL];
#pragma unroll
    for (int i = 0; i < VPL; ++i) qv[i] = __ldg(Q + lane + 32 * i);
    float ss_q = 0.f;
    bool have_q = false;
    for (int r = warp * ROWS; r < t.rows; r += 8 * ROWS) {
        float4 ev[ROWS][VPL];
#pragma unroll
        for (int j = 0; j < ROWS; ++j) {
            const bool ok = r + j < t.rows;
            const float4* src = E + (size_t)(ok ? r + j : 0) * d4 + lane;
#pragma unroll
            for (int i = 0; i < VPL; ++i) ev[j][i] = ok ? ldg_stream(src + 32 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (!have_q) {
            float s = 0.f;
#pragma unroll
            for (int i = 0; i < VPL; ++i) { s = fmaf(qv[i].x, qv[i].x, s); s = fmaf(qv[i].y, qv[i].y, s); s = fmaf(qv[i].z, qv[i].z, s); s = fmaf(qv[i].w, qv[i].w, s); }
            ss_q = warp_sum(s); have_q = true;
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
            const float sim = cosine(warp_sum(dot), warp_sum(ss), ss_q);
            if (lane == j) mine = sim;
        }
        if (lane < ROWS && r + lane < t.rows) out[r + lane] = mine;
    }
}

// ---- K2: warp-unit persistent --------------------------------------------------------------------------
// unit u = (tile u / UPT, rows [sub*ROWS, sub*ROWS+ROWS) of that tile), UPT = tile_rows / ROWS.
// warp gw owns units [gw*U/GW, (gw+1)*U/GW): contiguous in memory.
template <int VPL, int ROWS>
struct Unit {
    float4 ev[ROWS][VPL];
    float4 qv[VPL];
    float* out;
    int nrows;
};

template <int VPL, int ROWS>
__device__ __forceinline__ void unit_load(Unit<VPL, ROWS>& u, const Tile* __restrict__ tiles, const float* __restrict__ q, int d4,
                                          float* __restrict__ sims, int unit, int upt, int lane) {
    const int ti = unit / upt, sub = unit - ti * upt;
    const Tile t = tiles[ti];
    const int r0 = sub * ROWS;
    u.nrows = min(ROWS, t.rows - r0);      // may be <= 0 for the ragged last tile of a doc
    u.out = sims + t.sims_off + r0;
    const float4* __restrict__ E = reinterpret_cast<const float4*>(t.src) + (size_t)r0 * d4 + lane;
    const float4* __restrict__ Q = reinterpret_cast<const float4*>(q) + (size_t)t.doc * d4 + lane;
#pragma unroll
    for (int j = 0; j < ROWS; ++j) {
        const bool ok = j < u.nrows;
#pragma unroll
        for (int i = 0; i < VPL; ++i) u.ev[j][i] = ok ? ldg_stream(E + (size_t)j * d4 + 32 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int i = 0; i < VPL; ++i) u.qv[i] = __ldg(Q + 32 * i);
}

template <int VPL, int ROWS>
__device__ __forceinline__ void unit_reduce(const Unit<VPL, ROWS>& u, int lane) {
    if (u.nrows <= 0) return;
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) { s = fmaf(u.qv[i].x, u.qv[i].x, s); s = fmaf(u.qv[i].y, u.qv[i].y, s); s = fmaf(u.qv[i].z, u.qv[i].z, s); s = fmaf(u.qv[i].w, u.qv[i].w, s); }
    const float ss_q = warp_sum(s);
    float mine = 0.f;
#pragma unroll
    for (int j = 0; j < ROWS; ++j) {
        float dot = 0.f, ss = 0.f;
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            dot = fmaf(u.ev[j][i].x, u.qv[i].x, dot); ss = fmaf(u.ev[j][i].x, u.ev[j][i].x, ss);
            dot = fmaf(u.ev[j][i].y, u.qv[i].y, dot); ss = fmaf(u.ev[j][i].y, u.ev[j][i].y, ss);
            dot = fmaf(u.ev[j][i].z, u.qv[i].z, dot); ss = fmaf(u.ev[j][i].z, u.ev[j][i].z, ss);
            dot = fmaf(u.ev[j][i].w, u.qv[i].w, dot); ss = fmaf(u.ev[j][i].w, u.ev[j][i].w, ss);
        }
        const float sim = cosine(warp_sum(dot), warp_sum(ss), ss_q);
        if (lane == j) mine = sim;
    }
    if (lane < u.nrows) u.out[lane] = mine;
}

template <int VPL, int ROWS, int MINB>
__global__ void __launch_bounds__(256, MINB) k2_units(const Tile* __restrict__ tiles, int total_tiles, int tile_rows,
                                                      const float* __restrict__ q, int d, float* __restrict__ sims, int pdl) {
    const int lane = threadIdx.x & 31;
    const int gw = blockIdx.x * 8 + (threadIdx.x >> 5), GW = gridDim.x * 8;
    const int upt = tile_rows / ROWS;
    const long long U = (long long)total_tiles * upt;
    const int u0 = (int)(U * gw / GW), u1 = (int)(U * (gw + 1) / GW);
    if (pdl) { pdl_launch(); pdl_wait(); }
    if (u0 >= u1) return;
    const int d4 = d >> 2;
    Unit<VPL, ROWS> a, b;
    unit_load<VPL, ROWS>(a, tiles, q, d4, sims, u0, upt, lane);
    int u = u0;
    while (true) {
        if (u + 1 < u1) unit_load<VPL, ROWS>(b, tiles, q, d4, sims, u + 1, upt, lane);
        unit_reduce<VPL, ROWS>(a, lane);
        if (++u >= u1) break;
        if (u + 1 < u1) unit_load<VPL, ROWS>(a, tiles, q, d4, sims, u + 1, upt, lane);
        unit_reduce<VPL, ROWS>(b, lane);
        if (++u >= u1) break;
    }
}

// ---- host ----------------------------------------------------------------------------------------------
struct Batch {
    float* emb; float* q; float* sims; Tile* tiles; int total_tiles; long long rows;
};

static std::vector<int> doc_sizes(int B, int max_pages, int cpp, unsigned seed) {
    std::vector<int> s(B);
    for (int b = 0; b < B; ++b) { seed = seed * 1664525u + 1013904223u; s[b] = cpp * (1 + (int)((seed >> 8) % (unsigned)max_pages)); }
    return s;
}

static Batch make_batch(const std::vector<int>& sizes, int d, int tile_rows) {
    Batch bt{};
    long long N = 0;
    for (int s : sizes) N += s;
    bt.rows = N;
    CK(cudaMalloc(&bt.emb, (size_t)N * d * 4));
    CK(cudaMemset(bt.emb, 0x3c, (size_t)N * d * 4));
    CK(cudaMalloc(&bt.q, sizes.size() * d * 4));
    CK(cudaMemset(bt.q, 0x3c, sizes.size() * d * 4));
    CK(cudaMalloc(&bt.sims, (size_t)N * 4));
    std::vector<Tile> tiles;
    long long off = 0;
    for (size_t b = 0; b < sizes.size(); ++b) {
        for (int r = 0; r < sizes[b]; r += tile_rows) {
            Tile t{};
            t.src = bt.emb + (size_t)(off + r) * d; t.sims_off = off + r;
            t.rows = sizes[b] - r < tile_rows ? sizes[b] - r : tile_rows; t.doc = (int)b; t.doc_rows = sizes[b];
            tiles.push_back(t);
        }
        off += sizes[b];
    }
    bt.total_tiles = (int)tiles.size();
    CK(cudaMalloc(&bt.tiles, tiles.size() * sizeof(Tile)));
    CK(cudaMemcpy(bt.tiles, tiles.data(), tiles.size() * sizeof(Tile), cudaMemcpyHostToDevice));
    return bt;
}

template <class KernelT, class... Args>
static void launch(KernelT kern, dim3 grid, dim3 block, cudaStream_t s, int pdl, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = 0; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
    CK(cudaLaunchKernelEx(&cfg, kern, args...));
}

template <class F>
static float time_us(F fn, int iters, cudaStream_t s) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 10; ++i) fn(i);
    CK(cudaStreamSynchronize(s));
    CK(cudaEventRecord(e0, s));
    for (int i = 0; i < iters; ++i) fn(i);
    CK(cudaEventRecord(e1, s));
    CK(cudaStreamSynchronize(s));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    return ms * 1e3f / iters;
}

int main(int argc, char** argv) {
    const int R = 9, d = 384, B = 64;
    const int max_pages = argc > 1 ? atoi(argv[1]) : 20;
    cudaStream_t s;
    CK(cudaStreamCreate(&s));
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    std::vector<std::vector<int>> sizes;
    for (int r = 0; r < R; ++r) sizes.push_back(doc_sizes(B, max_pages, 30, 1234 + r));
    long long N0 = 0;
    for (int x : sizes[0]) N0 += x;
    const double bytes0 = (double)N0 * d * 4;
    printf("SMs %d, batch 0: %lld rows, %.1f MB\n", sms, N0, bytes0 / 1e6);
    float* sink;
    CK(cudaMalloc(&sink, 4));

    for (int tr : {16, 32, 64, 128}) {
        std::vector<Batch> bs;
        for (int r = 0; r < R; ++r) bs.push_back(make_batch(sizes[r], d, tr));
        double mb = 0;
        for (auto& b : bs) mb += (double)b.rows * d * 4 / R / 1e6;
        if (tr == 16) {
            for (int pdl = 0; pdl < 2; ++pdl)
                for (int per_sm : {2, 4, 8, 16}) {
                    float us4 = time_us([&](int i) { const Batch& b = bs[i % R]; launch(k0_read<4>, dim3(sms * per_sm), dim3(256), s, pdl, (const float4*)b.emb, (size_t)b.rows * d / 4, sink, pdl); }, 200, s);
                    float us8 = time_us([&](int i) { const Batch& b = bs[i % R]; launch(k0_read<8>, dim3(sms * per_sm), dim3(256), s, pdl, (const float4*)b.emb, (size_t)b.rows * d / 4, sink, pdl); }, 200, s);
                    printf("K0 read   pdl=%d blocks/SM=%2d  U=4: %6.2f us (%5.0f GB/s)   U=8: %6.2f us (%5.0f GB/s)\n", pdl, per_sm, us4, mb * 1e3 / us4, us8, mb * 1e3 / us8);
                }
        }
#define RUN_K1(ROWS, MINB)                                                                                          \
        for (int pdl = 0; pdl < 2; ++pdl) {                                                                         \
            float us = time_us([&](int i) { const Batch& b = bs[i % R];                                             \
                launch(k1_tile<3, ROWS, MINB>, dim3(b.total_tiles), dim3(256), s, pdl, (const Tile*)b.tiles, (const float*)b.q, d, b.sims, pdl); }, 200, s); \
            printf("K1 tile   tile_rows=%3d ROWS=%d MINB=%d pdl=%d tiles=%4d: %6.2f us (%5.0f GB/s)\n", tr, ROWS, MINB, pdl, bs[0].total_tiles, us, mb * 1e3 / us); \
        }
        RUN_K1(4, 1) RUN_K1(4, 3) RUN_K1(4, 4) RUN_K1(2, 4) RUN_K1(2, 5) RUN_K1(2, 6)
#define RUN_K2(ROWS, MINB)                                                                                          \
        for (int pdl = 0; pdl < 2; ++pdl) {                                                                         \
            float us = time_us([&](int i) { const Batch& b = bs[i % R];                                             \
                launch(k2_units<3, ROWS, MINB>, dim3(sms * MINB), dim3(256), s, pdl, (const Tile*)b.tiles, b.total_tiles, tr, (const float*)b.q, d, b.sims, pdl); }, 200, s); \
            printf("K2 units  tile_rows=%3d ROWS=%d blocks/SM=%d pdl=%d: %6.2f us (%5.0f GB/s)\n", tr, ROWS, MINB, pdl, us, mb * 1e3 / us); \
        }
        RUN_K2(1, 2) RUN_K2(1, 3) RUN_K2(1, 4) RUN_K2(1, 6) RUN_K2(2, 2) RUN_K2(2, 3) RUN_K2(2, 4) RUN_K2(4, 2)
        CK(cudaDeviceSynchronize());
        for (auto& b : bs) { cudaFree(b.emb); cudaFree(b.q); cudaFree(b.sims); cudaFree(b.tiles); }
    }
    printf("done\n");
    return 0;
}
