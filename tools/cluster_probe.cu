// Scratch: how many clusters of 512-thread, ~90 KB-smem, 128-register CTAs fit on this GPU at once?
// nvcc -arch=sm_100a -O3 -o tools/cluster_probe tools/cluster_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(512, 1) k(double *p) {
    extern __shared__ double s[];
    double a[56];
#pragma unroll
    for (int i = 0; i < 56; ++i) a[i] = p[threadIdx.x + i * 512];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int i = 0; i < 56; ++i) a[i] = a[i] * a[(i + 1) % 56] + s[threadIdx.x];
    double t = 0;
#pragma unroll
    for (int i = 0; i < 56; ++i) t += a[i];
    p[threadIdx.x] = t;
}
int main() {
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 92160);
    cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, k);
    printf("regs %d\n", fa.numRegs);
    for (int cs : {1, 2, 4, 8, 16}) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(cs * 64); cfg.blockDim = dim3(512); cfg.dynamicSmemBytes = 92160;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        int n = -1;
        cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k, &cfg);
        printf("cluster size %2d: %3d clusters = %3d CTAs  (%s)\n", cs, n, n * cs, cudaGetErrorString(e));
    }
    return 0;
}
