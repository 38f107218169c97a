// librdv: error plumbing and device queries shared by every entry point.
#include "rdv_common.cuh"

#include <stdlib.h>
#include <string.h>

namespace rdv {

static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t err, const char* what) {
    set_error("%s: %s (%s)", what, cudaGetErrorString(err), cudaGetErrorName(err));
    return RDV_E_CUDA;
}

int pdl_mask() {
    static const int mask = [] {
        const char* v = getenv("RDV_PDL");
        return v ? atoi(v) : kPdlStream;
    }();
    return mask;
}

int sm_count() {
    static thread_local int cached_dev = -1, cached_sms = 148;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != cached_dev) {
        int sms = 0;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && sms > 0) {
            cached_sms = sms;
            cached_dev = dev;
        }
    }
    return cached_sms;
}

}  // namespace rdv

extern "C" int rdv_abi_version(void) { return 8; }

extern "C" const char* rdv_last_error(void) { return rdv::g_error; }

extern "C" int rdv_device_info(int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return rdv::cuda_fail(e, "cudaGetDevice");
    int sms = 0, major = 0, minor = 0;
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
    if (e != cudaSuccess) return rdv::cuda_fail(e, "cudaDeviceGetAttribute");
    if (sm_count) *sm_count = sms;
    if (cc_major) *cc_major = major;
    if (cc_minor) *cc_minor = minor;
    return RDV_OK;
}
