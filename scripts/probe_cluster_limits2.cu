// Cluster size 16 vs registers per thread (256 threads per CTA).
#include <cuda_runtime.h>
#include <stdio.h>

template <int LIVE, int MINB>
__global__ void __launch_bounds__(256, MINB) kern(float* out, const float* in, int n) {
    float v[LIVE];
#pragma unroll
    for (int i = 0; i < LIVE; ++i) v[i] = in[threadIdx.x + 256 * i];
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    for (int it = 0; it < n; ++it) {
#pragma unroll
        for (int i = 0; i < LIVE; ++i) v[i] = v[i] * v[(i + 1) % LIVE] + 1.0f;
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < LIVE; ++i) s += v[i];
    out[blockIdx.x * 256 + threadIdx.x] = s;
}

template <int LIVE, int MINB>
static void test(float* out, const float* in) {
    auto k = kern<LIVE, MINB>;
    cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(128); cfg.blockDim = dim3(256);
    int potential = -1;
    cudaError_t e1 = cudaOccupancyMaxPotentialClusterSize(&potential, k, &cfg);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 16; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    int active = -1;
    cudaError_t e2 = cudaOccupancyMaxActiveClusters(&active, k, &cfg);
    cudaError_t e3 = cudaLaunchKernelEx(&cfg, k, out, in, 1);
    cudaError_t e4 = cudaDeviceSynchronize();
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, k);
    printf("live %2d minb %d regs %3d: potential %2d (%s)  active clusters of 16: %3d (%s)  launch16: %s / %s\n", LIVE, MINB,
           fa.numRegs, potential, cudaGetErrorName(e1), active, cudaGetErrorName(e2), cudaGetErrorName(e3), cudaGetErrorName(e4));
    cudaGetLastError();
}

int main() {
    float *out, *in; cudaMalloc(&out, 1 << 22); cudaMalloc(&in, 1 << 22); cudaMemset(in, 0, 1 << 22);
    test<4, 5>(out, in);
    test<12, 5>(out, in);
    test<16, 5>(out, in);
    test<20, 5>(out, in);
    test<24, 5>(out, in);
    test<28, 5>(out, in);
    test<32, 5>(out, in);
    test<40, 5>(out, in);
    test<40, 3>(out, in);
    test<64, 3>(out, in);
    test<64, 1>(out, in);
    test<100, 1>(out, in);
    return 0;
}
