// Probe: tcgen05.mma kind::i8 (INT8 x INT8 -> INT32 in TMEM) on sm_100a.
//  (1) correctness of one 128 x N x K tile against the host, operands in the canonical
//      K-major no-swizzle shared-memory layout ([k16][row][16 B]);
//  (2) issue rate: every SM loops over MMAs on resident operands.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o i8_probe i8_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(2); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout = 0) {
  uint64_t d = (uint64_t)layout << 61;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version 1 (sm_100)
  return d;
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
  return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}
__device__ __forceinline__ void mma_i8_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}
__device__ __forceinline__ void tmem_cp_128x256b(uint32_t tmem_dst, uint64_t sdesc) {
  asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(tmem_dst), "l"(sdesc) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_bounded(uint64_t* bar, uint32_t parity) {
  for (long long it = 0; it < 200000000LL; it++)
    if (mbar_try_wait(bar, parity)) return;
  printf("mbarrier timeout (block %d)\n", blockIdx.x);
  __trap();
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
#define TMEM_LD16(taddr, r)                                                                        \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];" \
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),        \
                 "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) \
               : "r"(taddr))

// A: [K/16][128][16] bytes, B: [K/16][N][16] bytes in global (already in smem image order)
// TS variant (SWZ = 0 layout only): per k-chunk the A operand is first copied to TMEM columns
// [448, 456) with tcgen05.cp, then `pairs` MMAs read it from there.
template <int N>
__global__ void __launch_bounds__(128) ts_kernel(const int8_t* __restrict__ A, const int8_t* __restrict__ B,
                                                 int K, int32_t* __restrict__ D, int reps, int pairs) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  uint8_t* sA = smem;
  uint8_t* sB = smem + (size_t)K * 128;
  for (int e = tid; e < K * 128 / 16; e += 128) ((uint4*)sA)[e] = ((const uint4*)A)[e];
  for (int e = tid; e < K * N / 16; e += 128) ((uint4*)sB)[e] = ((const uint4*)B)[e];
  if (tid == 0) mbar_init(&bar, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base;
  constexpr uint32_t idesc = make_idesc(128, N);
  uint32_t phase = 0;
  for (int r = 0; r < reps; r++) {
    if (tid == 0) {
      const uint64_t da0 = make_desc(smem_u32(sA), 2048, 128);
      const uint64_t db0 = make_desc(smem_u32(sB), N * 16, 128);
#pragma unroll
      for (int kc = 0; kc < 4; kc++) {
        // 7 slices' worth of copies per chunk in the real kernel: issue 7 here too when timing
        const int ncp = pairs > 1 ? 7 : 1;
        for (int c = 0; c < ncp; c++)
          tmem_cp_128x256b(tmem + 448 + (uint32_t)(c == 0 ? 0 : 8 * (c % 7)), da0 + (uint64_t)((kc * 2 * 2048) >> 4));
        if (pairs == 28) {
#pragma unroll
          for (int p = 0; p < 28; p++)
            mma_i8_ts(tmem + (uint32_t)((p % 7) * N), tmem + 448, db0 + (uint64_t)((kc * 2 * N * 16) >> 4), idesc, 1u);
        } else {
          mma_i8_ts(tmem, tmem + 448, db0 + (uint64_t)((kc * 2 * N * 16) >> 4), idesc, kc > 0 ? 1u : 0u);
        }
      }
      umma_commit(&bar);
    }
    mbar_wait_bounded(&bar, phase);
    phase ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  if (D != nullptr && blockIdx.x == 0) {
    for (int c0 = 0; c0 < N; c0 += 16) {
      uint32_t v[16];
      const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
      TMEM_LD16(taddr, v);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int c = 0; c < 16; c++) D[(size_t)tid * N + c0 + c] = (int32_t)v[c];
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

template <int N, int SWZ>
__global__ void __launch_bounds__(128) tile_kernel(const int8_t* __restrict__ A, const int8_t* __restrict__ B,
                                                   int K, int32_t* __restrict__ D, int reps, int pairs) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  uint8_t* sA = smem;                       // K/16 * 2048
  uint8_t* sB = smem + (size_t)K * 128;     // K/16 * N*16
  for (int e = tid; e < K * 128 / 16; e += 128) ((uint4*)sA)[e] = ((const uint4*)A)[e];
  for (int e = tid; e < K * N / 16; e += 128) ((uint4*)sB)[e] = ((const uint4*)B)[e];
  if (tid == 0) mbar_init(&bar, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base;
  constexpr uint32_t idesc = make_idesc(128, N);
  uint32_t phase = 0;
  for (int r = 0; r < reps; r++) {
    if (tid == 0) {
      // descriptors: one base per operand, k-steps and slices are immediates added to the low word
      const uint32_t stepA = SWZ == 0 ? 2 * 2048 : (SWZ == 32 ? 128 * 32 : 32);
      const uint32_t stepB = SWZ == 0 ? 2 * N * 16 : (SWZ == 32 ? N * 32 : 32);
      const uint64_t da0 = SWZ == 0 ? make_desc(smem_u32(sA), 2048, 128)
                                    : (SWZ == 32 ? make_desc(smem_u32(sA), 16, 256, 6) : make_desc(smem_u32(sA), 16, 1024, 2));
      const uint64_t db0 = SWZ == 0 ? make_desc(smem_u32(sB), N * 16, 128)
                                    : (SWZ == 32 ? make_desc(smem_u32(sB), 16, 256, 6) : make_desc(smem_u32(sB), 16, 1024, 2));
      if (pairs == 28) {
#pragma unroll
        for (int p = 0; p < 28; p++) {
#pragma unroll
          for (int kc = 0; kc < 4; kc++)
            mma_i8(tmem + (uint32_t)((p * N) % 512), da0 + (uint64_t)((kc * stepA) >> 4),
                   db0 + (uint64_t)((kc * stepB) >> 4), idesc, kc > 0 ? 1u : 0u);
        }
      } else {
        for (int kc = 0; kc < K / 32; kc++)
          mma_i8(tmem, da0 + (uint64_t)((kc * stepA) >> 4), db0 + (uint64_t)((kc * stepB) >> 4), idesc,
                 kc > 0 ? 1u : 0u);
      }
      umma_commit(&bar);
    }
    mbar_wait_bounded(&bar, phase);
    phase ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  if (D != nullptr && blockIdx.x == 0) {
    // thread t <-> TMEM lane t (row), N columns
    for (int c0 = 0; c0 < N; c0 += 16) {
      uint32_t v[16];
      const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
      TMEM_LD16(taddr, v);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int c = 0; c < 16; c++) D[(size_t)tid * N + c0 + c] = (int32_t)v[c];
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

template <int N, int SWZ>
static void run(int n_sm) {
  const int K = 128;     // 4 chunks of 32
  std::vector<int8_t> hA((size_t)128 * K), hB((size_t)N * K), iA(hA.size()), iB(hB.size());
  srand(7);
  for (auto& x : hA) x = (int8_t)(rand() % 256 - 128);
  for (auto& x : hB) x = (int8_t)(rand() % 256 - 128);
  // smem image: [k16][row][16]
  auto pos = [&](int rows, int r, int k) -> size_t {
    if (SWZ == 0) return ((size_t)(k / 16) * rows + r) * 16 + k % 16;
    if (SWZ == 32) return ((size_t)(k / 32) * rows + r) * 32 + ((((k % 32) / 16) ^ ((r / 4) & 1)) * 16) + k % 16;
    return ((size_t)(k / 128) * rows + r) * 128 + ((((k % 128) / 16) ^ (r & 7)) * 16) + k % 16;
  };
  for (int r = 0; r < 128; r++) for (int k = 0; k < K; k++) iA[pos(128, r, k)] = hA[(size_t)r * K + k];
  for (int r = 0; r < N; r++) for (int k = 0; k < K; k++) iB[pos(N, r, k)] = hB[(size_t)r * K + k];
  int8_t *dA, *dB; int32_t* dD;
  CK(cudaMalloc(&dA, iA.size())); CK(cudaMalloc(&dB, iB.size())); CK(cudaMalloc(&dD, (size_t)128 * N * 4));
  CK(cudaMemcpy(dA, iA.data(), iA.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, iB.data(), iB.size(), cudaMemcpyHostToDevice));
  const size_t smem = (size_t)K * 128 + (size_t)K * N;
  CK(cudaFuncSetAttribute(tile_kernel<N, SWZ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  tile_kernel<N, SWZ><<<1, 128, smem>>>(dA, dB, K, dD, 1, 1);
  CK(cudaDeviceSynchronize());
  std::vector<int32_t> hD((size_t)128 * N);
  CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
  long long bad = 0;
  for (int m = 0; m < 128; m++) for (int n = 0; n < N; n++) {
    int32_t ref = 0;
    for (int k = 0; k < K; k++) ref += (int32_t)hA[(size_t)m * K + k] * (int32_t)hB[(size_t)n * K + k];
    if (ref != hD[(size_t)m * N + n]) { if (bad < 5) printf("  mismatch m=%d n=%d got %d want %d\n", m, n, hD[(size_t)m * N + n], ref); bad++; }
  }
  printf("SWZ=%d N=%d correctness: %lld mismatches of %d\n", SWZ, N, bad, 128 * N);
  // rate: all SMs, reps x pairs x (K/32) MMAs of 128 x N x 32
  const int reps = 200, pairs = 28;
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  tile_kernel<N, SWZ><<<n_sm, 128, smem>>>(dA, dB, K, nullptr, 5, pairs);
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  tile_kernel<N, SWZ><<<n_sm, 128, smem>>>(dA, dB, K, nullptr, reps, pairs);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  double macs = (double)n_sm * reps * pairs * (K / 32) * 128.0 * N * 32.0;
  printf("SWZ=%d N=%d rate: %.3f ms, %.1f TOPS (2 ops per MAC), %.1f clk per MMA at 1.9 GHz\n", SWZ, N, ms, 2 * macs / ms * 1e-9,
         ms * 1e-3 * 1.9e9 / (reps * pairs * (K / 32)));
  cudaFree(dA); cudaFree(dB); cudaFree(dD);
}

// two issuing threads (warp 0 and warp 1), disjoint accumulator columns
template <int N>
__global__ void __launch_bounds__(128) dual_kernel(const int8_t* __restrict__ A, const int8_t* __restrict__ B,
                                                   int K, int reps, int nissue) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  uint8_t* sA = smem;
  uint8_t* sB = smem + (size_t)K * 128;
  for (int e = tid; e < K * 128 / 16; e += 128) ((uint4*)sA)[e] = ((const uint4*)A)[e];
  for (int e = tid; e < K * N / 16; e += 128) ((uint4*)sB)[e] = ((const uint4*)B)[e];
  if (tid == 0) mbar_init(&bar, nissue);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base;
  constexpr uint32_t idesc = make_idesc(128, N);
  uint32_t phase = 0;
  for (int r = 0; r < reps; r++) {
    if ((tid & 31) == 0 && warp < nissue) {
      const uint64_t da0 = make_desc(smem_u32(sA), 2048, 128);
      const uint64_t db0 = make_desc(smem_u32(sB), N * 16, 128);
      const int per = 28 / nissue;
      if (warp == 0) {
#pragma unroll
        for (int p = 0; p < 28; p++) {
          if (p < per) {
#pragma unroll
            for (int kc = 0; kc < 4; kc++)
              mma_i8(tmem + (uint32_t)((p % 4) * N), da0 + (uint64_t)((kc * 2 * 2048) >> 4),
                     db0 + (uint64_t)((kc * 2 * N * 16) >> 4), idesc, kc > 0 ? 1u : 0u);
          }
        }
      } else {
#pragma unroll
        for (int p = 0; p < 14; p++) {
#pragma unroll
          for (int kc = 0; kc < 4; kc++)
            mma_i8(tmem + (uint32_t)((4 + p % 3) * N), da0 + (uint64_t)((kc * 2 * 2048) >> 4),
                   db0 + (uint64_t)((kc * 2 * N * 16) >> 4), idesc, kc > 0 ? 1u : 0u);
        }
      }
      umma_commit(&bar);
    }
    mbar_wait_bounded(&bar, phase);
    phase ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}
template <int N>
static void run_dual(int n_sm) {
  const int K = 128;
  int8_t *dA, *dB;
  CK(cudaMalloc(&dA, 128 * K)); CK(cudaMalloc(&dB, N * K));
  CK(cudaMemset(dA, 1, 128 * K)); CK(cudaMemset(dB, 1, N * K));
  const size_t smem = (size_t)K * 128 + (size_t)K * N;
  CK(cudaFuncSetAttribute(dual_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  for (int nissue = 1; nissue <= 2; nissue++) {
    const int reps = 200;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    dual_kernel<N><<<n_sm, 128, smem>>>(dA, dB, K, 5, nissue);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    dual_kernel<N><<<n_sm, 128, smem>>>(dA, dB, K, reps, nissue);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("DUAL N=%d issuers=%d: %.3f ms, %.1f clk per MMA at 1.9 GHz\n", N, nissue, ms,
           ms * 1e-3 * 1.9e9 / (reps * 28 * 4));
  }
  cudaFree(dA); cudaFree(dB);
}

template <int N>
static void run_ts(int n_sm) {
  const int K = 128;
  std::vector<int8_t> hA((size_t)128 * K), hB((size_t)N * K), iA(hA.size()), iB(hB.size());
  srand(11);
  for (auto& x : hA) x = (int8_t)(rand() % 256 - 128);
  for (auto& x : hB) x = (int8_t)(rand() % 256 - 128);
  for (int r = 0; r < 128; r++) for (int k = 0; k < K; k++) iA[((size_t)(k / 16) * 128 + r) * 16 + k % 16] = hA[(size_t)r * K + k];
  for (int r = 0; r < N; r++) for (int k = 0; k < K; k++) iB[((size_t)(k / 16) * N + r) * 16 + k % 16] = hB[(size_t)r * K + k];
  int8_t *dA, *dB; int32_t* dD;
  CK(cudaMalloc(&dA, iA.size())); CK(cudaMalloc(&dB, iB.size())); CK(cudaMalloc(&dD, (size_t)128 * N * 4));
  CK(cudaMemcpy(dA, iA.data(), iA.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, iB.data(), iB.size(), cudaMemcpyHostToDevice));
  const size_t smem = (size_t)K * 128 + (size_t)K * N;
  CK(cudaFuncSetAttribute(ts_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ts_kernel<N><<<1, 128, smem>>>(dA, dB, K, dD, 1, 1);
  CK(cudaDeviceSynchronize());
  std::vector<int32_t> hD((size_t)128 * N);
  CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
  long long bad = 0;
  for (int m = 0; m < 128; m++) for (int n = 0; n < N; n++) {
    int32_t ref = 0;
    for (int k = 0; k < K; k++) ref += (int32_t)hA[(size_t)m * K + k] * (int32_t)hB[(size_t)n * K + k];
    if (ref != hD[(size_t)m * N + n]) { if (bad < 5) printf("  TS mismatch m=%d n=%d got %d want %d\n", m, n, hD[(size_t)m * N + n], ref); bad++; }
  }
  printf("TS N=%d correctness: %lld mismatches of %d\n", N, bad, 128 * N);
  const int reps = 200, pairs = 28;
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  ts_kernel<N><<<n_sm, 128, smem>>>(dA, dB, K, nullptr, 5, pairs);
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  ts_kernel<N><<<n_sm, 128, smem>>>(dA, dB, K, nullptr, reps, pairs);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  double macs = (double)n_sm * reps * pairs * (K / 32) * 128.0 * N * 32.0;
  printf("TS N=%d rate (7 cp + 28 mma per chunk): %.3f ms, %.1f TOPS, %.1f clk per chunk at 1.9 GHz\n", N, ms,
         2 * macs / ms * 1e-9, ms * 1e-3 * 1.9e9 / (reps * (K / 32)));
  cudaFree(dA); cudaFree(dB); cudaFree(dD);
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
  run<64, 0>(p.multiProcessorCount);
  run<128, 32>(p.multiProcessorCount);
  run<256, 128>(p.multiProcessorCount);
  run_ts<64>(p.multiProcessorCount);
  run_dual<64>(p.multiProcessorCount);
  return 0;
}
