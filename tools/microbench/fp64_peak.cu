// FP64 peak microbenchmark for B200 (sm_100a): DFMA vs DMMA.8x8x4 issue rates, plus a
// fragment-layout self check for mma.sync.m8n8k4.f64.  Used to pick the roofline denominator
// for the variance contraction (DESIGN.md "Roofline").  Not part of the product path.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int NACC>
__global__ void __launch_bounds__(1024) dmma_peak(double* out, int iters, double a0, double b0) {
  double c[NACC][2];
#pragma unroll
  for (int i = 0; i < NACC; i++) { c[i][0] = 0.0; c[i][1] = 0.0; }
  double a = a0 + threadIdx.x * 1e-9, b = b0;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) dmma884(c[i][0], c[i][1], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void __launch_bounds__(1024) dfma_peak(double* out, int iters, double a0, double b0) {
  double c[NACC];
#pragma unroll
  for (int i = 0; i < NACC; i++) c[i] = i;
  double a = a0 + threadIdx.x * 1e-9, b = b0;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) c[i] = fma(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// mixed: DMMA and DFMA interleaved, to see whether they share one pipe
template <int NACC>
__global__ void __launch_bounds__(1024) mixed_peak(double* out, int iters, double a0, double b0) {
  double c[NACC][2]; double f[NACC];
#pragma unroll
  for (int i = 0; i < NACC; i++) { c[i][0] = 0.0; c[i][1] = 0.0; f[i] = i; }
  double a = a0 + threadIdx.x * 1e-9, b = b0;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) { dmma884(c[i][0], c[i][1], a, b); f[i] = fma(f[i], a, b); }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1] + f[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void exp_peak(double* out, int iters, double x0) {
  double x = x0 - threadIdx.x * 1e-3, s = 0;
  for (int it = 0; it < iters; it++) { s += exp(x); x -= 1e-6; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void layout_check(const double* A, const double* B, double* C) {
  // A is 8x4 row-major, B is 4x8 (k x n) row-major, C is 8x8 row-major
  int l = threadIdx.x, g = l >> 2, t = l & 3;
  double a = A[g * 4 + t];      // A[row=g][k=t]
  double b = B[t * 8 + g];      // B[k=t][n=g]
  double c0 = 0, c1 = 0;
  dmma884(c0, c1, a, b);
  C[g * 8 + 2 * t] = c0; C[g * 8 + 2 * t + 1] = c1;
}

template <typename F>
float time_ms(F launch, int reps) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  launch(); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; r++) {
    CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  printf("device %s sm_%d%d SMs=%d smem/blk optin=%zu L2=%d MB\n", p.name, p.major, p.minor, p.multiProcessorCount, p.sharedMemPerBlockOptin, p.l2CacheSize >> 20);
  int nsm = p.multiProcessorCount;
  double* out; CK(cudaMalloc(&out, sizeof(double) * nsm * 8 * 1024));
  // layout check
  {
    double hA[32], hB[32], hC[64], ref[64];
    for (int i = 0; i < 32; i++) { hA[i] = 1 + i * 0.5; hB[i] = 2 - i * 0.25; }
    for (int i = 0; i < 8; i++) for (int j = 0; j < 8; j++) { double s = 0; for (int k = 0; k < 4; k++) s = fma(hA[i * 4 + k], hB[k * 8 + j], s); ref[i * 8 + j] = s; }
    double *dA, *dB, *dC; CK(cudaMalloc(&dA, 256)); CK(cudaMalloc(&dB, 256)); CK(cudaMalloc(&dC, 512));
    CK(cudaMemcpy(dA, hA, 256, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, hB, 256, cudaMemcpyHostToDevice));
    layout_check<<<1, 32>>>(dA, dB, dC); CK(cudaMemcpy(hC, dC, 512, cudaMemcpyDeviceToHost));
    double maxd = 0; for (int i = 0; i < 64; i++) { double d = fabs(hC[i] - ref[i]); if (d > maxd) maxd = d; }
    printf("layout_check m8n8k4: max|diff|=%g (%s)\n", maxd, maxd < 1e-12 ? "OK" : "MISMATCH");
  }
  const int iters = 2048;
  int cfgs[][2] = {{1, 128}, {1, 256}, {1, 512}, {2, 256}, {1, 1024}, {2, 512}, {4, 256}};
  for (auto& cf : cfgs) {
    int bps = cf[0], thr = cf[1];
    int grid = nsm * bps;
    {
      float ms = time_ms([&] { dmma_peak<8><<<grid, thr>>>(out, iters, 1.0000001, 0.5); }, 5);
      double flops = 2.0 * 256 * 8 * (double)iters * (thr / 32) * grid;
      printf("DMMA.8x8x4 x8acc  blocks/SM=%d threads=%4d : %8.3f ms  %7.2f TFLOP/s\n", bps, thr, ms, flops / ms * 1e-9);
    }
    {
      float ms = time_ms([&] { dmma_peak<16><<<grid, thr>>>(out, iters, 1.0000001, 0.5); }, 5);
      double flops = 2.0 * 256 * 16 * (double)iters * (thr / 32) * grid;
      printf("DMMA.8x8x4 x16acc blocks/SM=%d threads=%4d : %8.3f ms  %7.2f TFLOP/s\n", bps, thr, ms, flops / ms * 1e-9);
    }
    {
      float ms = time_ms([&] { dfma_peak<16><<<grid, thr>>>(out, iters, 1.0000001, 0.5); }, 5);
      double flops = 2.0 * 16 * (double)iters * thr * grid;
      printf("DFMA x16chains    blocks/SM=%d threads=%4d : %8.3f ms  %7.2f TFLOP/s\n", bps, thr, ms, flops / ms * 1e-9);
    }
    {
      float ms = time_ms([&] { mixed_peak<8><<<grid, thr>>>(out, iters, 1.0000001, 0.5); }, 5);
      double flops = 2.0 * (256 * 8 * (thr / 32) + 8.0 * thr) * (double)iters * grid;
      printf("DMMA+DFMA mixed   blocks/SM=%d threads=%4d : %8.3f ms  %7.2f TFLOP/s\n", bps, thr, ms, flops / ms * 1e-9);
    }
  }
  {
    int grid = nsm * 4, thr = 512, it = 4096;
    float ms = time_ms([&] { exp_peak<<<grid, thr>>>(out, it, -0.5); }, 5);
    printf("exp(double): %.3f ms -> %.2f Gexp/s\n", ms, (double)it * thr * grid / ms * 1e-6);
  }
  // H2D / D2H pinned bandwidth
  {
    size_t bytes = 1ull << 30; void *h, *d; CK(cudaMallocHost(&h, bytes)); CK(cudaMalloc(&d, bytes));
    float ms = time_ms([&] { cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, 0); }, 3);
    printf("H2D pinned 1 GiB: %.2f ms -> %.1f GB/s\n", ms, bytes / ms * 1e-6);
    ms = time_ms([&] { cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, 0); }, 3);
    printf("D2H pinned 1 GiB: %.2f ms -> %.1f GB/s\n", ms, bytes / ms * 1e-6);
  }
  return 0;
}
