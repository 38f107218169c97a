// How fast does the hardware launch thread-block clusters?  An (almost) empty kernel, grid of 512 / 1024 CTAs x 256
// threads, cluster sizes 1 .. 16, 300 back-to-back launches timed with CUDA events.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/probe_cluster_launch scripts/probe_cluster_launch.cu
#include <cuda_runtime.h>
#include <stdio.h>

__global__ void __launch_bounds__(256, 5) tiny(float* out, int work) {
    float acc = 0.f;
    for (int i = 0; i < work; ++i) acc = acc * 1.0001f + (float)threadIdx.x;
    if (acc == 12345.678f) out[blockIdx.x] = acc;
}

__global__ void __launch_bounds__(256, 5) tiny_sync(float* out, int work) {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    float acc = 0.f;
    for (int i = 0; i < work; ++i) acc = acc * 1.0001f + (float)threadIdx.x;
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (acc == 12345.678f) out[blockIdx.x] = acc;
}

static float run(void (*k)(float*, int), int grid, int cluster, int work, float* out) {
    cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(256);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = cluster > 0 ? 1 : 0;
    for (int i = 0; i < 20; ++i) cudaLaunchKernelEx(&cfg, k, out, work);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    const int n = 300;
    for (int i = 0; i < n; ++i) cudaLaunchKernelEx(&cfg, k, out, work);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    cudaError_t err = cudaGetLastError();
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    if (err != cudaSuccess) { printf("  (%s) ", cudaGetErrorString(err)); return -1.f; }
    return ms / n * 1e3f;
}

int main() {
    float* out; cudaMalloc(&out, 1 << 20);
    for (int work : {0, 2000}) {
        for (int grid : {512, 1024}) {
            printf("work %d, grid %d x 256 threads:\n", work, grid);
            for (int cluster : {0, 1, 2, 4, 8, 16}) {
                float a = run(tiny, grid, cluster, work, out);
                float b = cluster >= 1 ? run(tiny_sync, grid, cluster, work, out) : -1.f;
                printf("  cluster %2d: %7.2f us per launch   with cluster barrier: %7.2f us\n", cluster, a, b);
            }
        }
    }
    return 0;
}
