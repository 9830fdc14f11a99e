// How fast can all SMs stream L2-resident data into shared memory with cp.async.bulk?
// (sizing the operand traffic a wider INT8 MMA tiling may ask for)
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(2); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
template <int STAGES>
__global__ void __launch_bounds__(64) stream_kernel(const uint8_t* __restrict__ src, size_t region, int chunk, int iters) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full[STAGES];
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; s++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full[s])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    size_t off = ((size_t)blockIdx.x * 7919 * chunk) % (region - chunk);
    off &= ~(size_t)1023;
    // keep STAGES copies in flight
    for (int it = 0; it < iters + STAGES; it++) {
      const int s = it % STAGES;
      if (it >= STAGES) {
        const uint32_t ph = ((it / STAGES) - 1) & 1;
        long long g = 0;
        while (!try_wait(&full[s], ph)) if (++g > 100000000LL) __trap();
      }
      if (it < iters) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full[s])), "r"((uint32_t)chunk) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(smem + (size_t)s * chunk)), "l"(src + off), "r"((uint32_t)chunk), "r"(smem_u32(&full[s])) : "memory");
        off += (size_t)chunk * 148;
        if (off + chunk > region) off = (off + chunk) % (region - chunk) & ~(size_t)1023;
      }
    }
  }
}
int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  const int n_sm = p.multiProcessorCount;
  for (size_t region_mb : {16, 48, 96, 600}) {
    const size_t region = region_mb << 20;
    uint8_t* d; CK(cudaMalloc(&d, region)); CK(cudaMemset(d, 1, region));
    for (int chunk : {14336, 43008}) {
      constexpr int ST = 4;
      const size_t smem = (size_t)ST * chunk;
      CK(cudaFuncSetAttribute(stream_kernel<ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      const int iters = 4000;
      cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
      stream_kernel<ST><<<n_sm, 64, smem>>>(d, region, chunk, 200);
      CK(cudaDeviceSynchronize());
      CK(cudaEventRecord(e0));
      stream_kernel<ST><<<n_sm, 64, smem>>>(d, region, chunk, iters);
      CK(cudaEventRecord(e1));
      CK(cudaDeviceSynchronize());
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
      printf("region %4zu MB  chunk %5d B x %d stages: %.2f TB/s into shared memory (%.1f B/clk/SM at 1.9 GHz)\n", region_mb, chunk, ST,
             (double)n_sm * iters * chunk / (ms * 1e-3) * 1e-12, (double)iters * chunk / (ms * 1e-3 * 1.9e9));
    }
    cudaFree(d);
  }
  return 0;
}
