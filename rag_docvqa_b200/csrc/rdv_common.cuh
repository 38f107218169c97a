// Shared device/host helpers for librdv (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

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

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

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
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ unsigned long long warp_max_u64(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long other = __shfl_xor_sync(0xffffffffu, v, o);
        v = other > v ? other : v;
    }
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
