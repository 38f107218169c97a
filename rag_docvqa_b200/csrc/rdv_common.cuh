// Shared device/host helpers for librdv (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include <atomic>

#include "../../include/rdv.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "librdv is written for sm_100a (B200) only"
#endif

namespace rdv {

// ---- host-side error plumbing ---------------------------------------------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t err, const char* what);
int sm_count();

#define RDV_REQUIRE(cond, code, ...)        \
    do {                                    \
        if (!(cond)) {                      \
            ::rdv::set_error(__VA_ARGS__);  \
            return (code);                  \
        }                                   \
    } while (0)

#define RDV_LAUNCH_CHECK(what)                                   \
    do {                                                         \
        cudaError_t e__ = cudaGetLastError();                    \
        if (e__ != cudaSuccess) return ::rdv::cuda_fail(e__, what); \
    } while (0)

// Function attributes (opt-in dynamic shared memory, carve-out preference) belong to ONE device's context, so a
// "set it once" guard has to be per device: a process that launches on cuda:0 and then on cuda:1 must opt in on
// both.  One bit per device ordinal; ordinals >= 64 simply set the attribute every time (the call is cheap).
struct PerDeviceOnce {
    std::atomic<unsigned long long> done{0ull};
    bool pending(int* dev) {
        *dev = 0;
        if (cudaGetDevice(dev) != cudaSuccess) return true;
        return *dev >= 64 || !((done.load(std::memory_order_acquire) >> *dev) & 1ull);
    }
    void mark(int dev) {
        if (dev >= 0 && dev < 64) done.fetch_or(1ull << dev, std::memory_order_release);
    }
};

#define RDV_ONCE_PER_DEVICE(expr, what)                                 \
    do {                                                                \
        static ::rdv::PerDeviceOnce once__;                             \
        int dev__;                                                      \
        if (once__.pending(&dev__)) {                                   \
            cudaError_t e__ = (expr);                                   \
            if (e__ != cudaSuccess) return ::rdv::cuda_fail(e__, what); \
            once__.mark(dev__);                                         \
        }                                                               \
    } while (0)

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ---- programmatic dependent launch (PDL) --------------------------------------------------------
// The per-document batches the reference runs are tens of MB: a kernel is ~5 us of HBM time, so the ~2 us
// drain + launch gap between two dependent kernels is a third of the step (measured: scripts/probe_stream.cu,
// 34 MB pure read: 8.1 us per launch back to back, 6.3 us with PDL).  Every kernel launched through
// launch_pdl() (a) calls pdl_launch_dependents() first, so the next kernel in the stream can be scheduled
// while this one still executes, and (b) calls pdl_wait() BEFORE its first global-memory access --
// griddepcontrol.wait returns once the preceding grid has completed and its writes are visible, so stream
// order is preserved exactly.  Which launches carry the attribute is a measured choice (RDV_PDL bit mask in the
// environment; scripts/probe_step.py on B200, C2 step score -> select+gather): plain launches 17.1 us with
// none, 15.1 us with the streaming kernels only (default), 15.7 us with both; inside a CUDA graph 13.2 us with
// or without the streaming kernels, 15.1 us when the select/gather kernel is launched early as well.
// RDV_CARVEOUT: preferred shared-memory carve-out (percent) of the kernels launched through launch_pdl, -1 = driver
// default (the default).  Measured on B200 (C2 step, with / without PDL on the select+gather launch): matching carve-outs
// do not make the early-launched gather pay off inside a graph (13.4 us per step either way, 5.0-5.2 us with 8 lanes);
// with plain stream launches RDV_CARVEOUT=0 RDV_PDL=3 gives 14.3 us instead of 16.3 us, but costs the 8-lane mode 15 %.
int carveout_pct();
int pdl_mask();                       // RDV_PDL bit 0: streaming kernels (default), bit 1: selection / gather kernels
constexpr int kPdlStream = 1, kPdlSelect = 2;

__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <class... KArgs>
static inline void prefer_carveout(void (*kernel)(KArgs...)) {
    const int pct = carveout_pct();
    if (pct >= 0) cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
}

template <class... KArgs, class... Args>
static inline cudaError_t launch_pdl(int cls, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                     Args&&... args) {
    prefer_carveout(kernel);                  // a measurement knob (RDV_CARVEOUT), off by default: set on every launch,
                                              // because a function-local flag here would exist once per kernel SIGNATURE
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (pdl_mask() & cls) ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---- ordering keys ----------------------------------------------------------------------------
// fp32 -> u32 whose unsigned order equals torch.topk's order: NaN (either sign) greatest, -0 == +0.
__device__ __forceinline__ uint32_t order_key(float v) {
    uint32_t u = __float_as_uint(v);
    if (v != v) return 0xFFFFFFFFu;
    if (u == 0x80000000u) u = 0u;
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
// (score, index) -> u64: greater key == better hit; equal scores prefer the LOWER index.
__device__ __forceinline__ unsigned long long pack_key(float v, uint32_t idx) {
    return (static_cast<unsigned long long>(order_key(v)) << 32) | static_cast<unsigned long long>(~idx);
}
__device__ __forceinline__ uint32_t key_index(unsigned long long key) { return ~static_cast<uint32_t>(key); }

// ---- warp / block reductions ------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---- streaming loads: read-once data should not pollute L1 -------------------------------------
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ uint4 ldg_stream_u4(const uint4* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

}  // namespace rdv
