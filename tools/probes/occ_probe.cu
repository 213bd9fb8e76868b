// Probe: what limits cudaOccupancyMaxActiveBlocksPerMultiprocessor to 1 for the tcgen05 kernels?
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
extern __shared__ uint8_t dyn[];
__global__ void __launch_bounds__(192, 1) k_plain(float* o) { dyn[threadIdx.x] = 1; __syncthreads(); o[threadIdx.x] = dyn[0]; }
__global__ void __launch_bounds__(192, 1) k_alloc_rt(float* o, int cols) {
    __shared__ uint32_t slot;
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    __syncthreads();
    o[threadIdx.x] = (float)slot + dyn[0];
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(cols) : "memory");
}
__global__ void __launch_bounds__(192, 1) k_alloc_64(float* o) {
    __shared__ uint32_t slot;
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    __syncthreads();
    o[threadIdx.x] = (float)slot + dyn[0];
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(slot) : "memory");
}
__global__ void __launch_bounds__(192, 1) k_cluster(float* o) {
    asm volatile("barrier.cluster.arrive.release;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire;" ::: "memory");
    o[threadIdx.x] = dyn[0];
}
__global__ void __launch_bounds__(192) k_alloc_rt_nolb(float* o, int cols) {
    __shared__ uint32_t slot;
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    __syncthreads();
    o[threadIdx.x] = (float)slot + dyn[0];
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(cols) : "memory");
}
template <typename K> void show(const char* name, K k) {
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200000);
    int o[4]; size_t sz[4] = {1024, 49152, 99584, 200000};
    for (int i = 0; i < 4; ++i) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o[i], k, 192, sz[i]);
    printf("%-16s occupancy @1K/48K/97K/195K = %d %d %d %d\n", name, o[0], o[1], o[2], o[3]);
}
int main() {
    show("plain", k_plain); show("tcgen05.alloc rt", k_alloc_rt); show("tcgen05.alloc 64", k_alloc_64); show("cluster barrier", k_cluster);
    show("alloc rt no lb", k_alloc_rt_nolb);
    return 0;
}
