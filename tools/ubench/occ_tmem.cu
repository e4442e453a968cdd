// Does a kernel that allocates TMEM get a lower occupancy from the runtime's calculator?  (B200 probe)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(256, 2) k_plain(float* out) {
    extern __shared__ float sm[];
    sm[threadIdx.x] = threadIdx.x;
    __syncthreads();
    out[threadIdx.x] = sm[255 - threadIdx.x];
}
__global__ void __launch_bounds__(256, 2) k_tmem(float* out, int cols) {
    extern __shared__ float sm[];
    __shared__ unsigned slot;
    if (threadIdx.x < 32) {
        unsigned a = (unsigned)__cvta_generic_to_shared(&slot);
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(a), "r"(128u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    __syncthreads();
    sm[threadIdx.x] = slot;
    __syncthreads();
    out[threadIdx.x] = sm[255 - threadIdx.x];
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(slot), "r"(128u) : "memory");
}
int main() {
    for (int kb : {0, 16, 48, 64, 66, 90, 100}) {
        int a = 0, b = 0;
        cudaFuncSetAttribute(k_plain, cudaFuncAttributeMaxDynamicSharedMemorySize, kb * 1024);
        cudaFuncSetAttribute(k_tmem, cudaFuncAttributeMaxDynamicSharedMemorySize, kb * 1024);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, k_plain, 256, kb * 1024);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, k_tmem, 256, kb * 1024);
        printf("dyn smem %3d KB: plain %d CTAs/SM, tmem kernel %d CTAs/SM\n", kb, a, b);
    }
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, k_tmem);
    printf("k_tmem regs %d static smem %zu\n", fa.numRegs, fa.sharedSizeBytes);
    // actually run 2 CTAs per SM of the TMEM kernel to see that co-residency works
    float* out; cudaMalloc(&out, 1024);
    k_tmem<<<296, 256, 64 * 1024>>>(out, 128);
    printf("run: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
