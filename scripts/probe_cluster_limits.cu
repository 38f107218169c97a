// What limits the cluster size a kernel can be launched with on B200?  Kernels of 256 threads with different static
// shared memory and launch bounds: cudaOccupancyMaxPotentialClusterSize, cudaOccupancyMaxActiveClusters and a launch.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/probe_cluster_limits scripts/probe_cluster_limits.cu
#include <cuda_runtime.h>
#include <stdio.h>

template <int SMEM, int MINB>
__global__ void __launch_bounds__(256, MINB) kern(float* out) {
    __shared__ float s[SMEM / 4 > 0 ? SMEM / 4 : 1];
    s[threadIdx.x % (SMEM / 4 > 0 ? SMEM / 4 : 1)] = (float)threadIdx.x;
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (s[0] == 12345.678f) out[blockIdx.x] = s[1];
}

template <int SMEM, int MINB>
static void test(float* out, int dyn) {
    auto k = kern<SMEM, MINB>;
    cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (dyn) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(128); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = dyn;
    int potential = -1;
    cudaError_t e1 = cudaOccupancyMaxPotentialClusterSize(&potential, k, &cfg);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 16; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    int active = -1;
    cudaError_t e2 = cudaOccupancyMaxActiveClusters(&active, k, &cfg);
    cudaError_t e3 = cudaLaunchKernelEx(&cfg, k, out);
    cudaError_t e4 = cudaDeviceSynchronize();
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, k);
    printf("static smem %6d dyn %6d minb %d regs %3d: potential %2d (%s)  active clusters of 16: %3d (%s)  launch16: %s / %s\n", SMEM, dyn, MINB,
           fa.numRegs, potential, cudaGetErrorName(e1), active, cudaGetErrorName(e2), cudaGetErrorName(e3), cudaGetErrorName(e4));
    cudaGetLastError();
}

int main() {
    float* out; cudaMalloc(&out, 1 << 20);
    test<0, 5>(out, 0);
    test<1024, 5>(out, 0);
    test<2048, 5>(out, 0);
    test<4096, 5>(out, 0);
    test<8192, 5>(out, 0);
    test<16384, 5>(out, 0);
    test<32768, 5>(out, 0);
    test<16384, 3>(out, 0);
    test<16384, 1>(out, 0);
    test<0, 5>(out, 2560);
    test<0, 5>(out, 16384);
    test<0, 1>(out, 65536);
    test<14000, 5>(out, 2560);
    return 0;
}
