// Shared device/host helpers for the gpry_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/gpry_b200.h"

namespace gpry {

// ---------------------------------------------------------------------------------------
// Error plumbing: every C-ABI entry catches GpryError and stores the text for
// gpry_last_error().
// ---------------------------------------------------------------------------------------
struct GpryError {
  int code;
  std::string msg;
};

void set_last_error(const std::string& msg);

#define GPRY_CUDA(expr)                                                                   \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      char _buf[512];                                                                     \
      snprintf(_buf, sizeof(_buf), "CUDA error '%s' at %s:%d (%s)", cudaGetErrorString(_e), \
               __FILE__, __LINE__, #expr);                                                \
      throw ::gpry::GpryError{GPRY_ERR_CUDA, _buf};                                       \
    }                                                                                     \
  } while (0)

#define GPRY_CHECK_ARG(cond, text)                                         \
  do {                                                                     \
    if (!(cond)) throw ::gpry::GpryError{GPRY_ERR_ARG, std::string(text)}; \
  } while (0)

// ---------------------------------------------------------------------------------------
// Tiling constants shared by the build kernel (producer of K* tiles), the state upload
// (producer of V tiles) and the contraction kernel (consumer of both).
//
// A "tile" is 128 rows x 16 k-columns of FP64 stored as [4 k-panels][128 rows][4 k] =
// 2048 doubles = 16 KB, contiguous in global memory so that ONE cp.async.bulk (TMA, SASS
// UBLKCP) moves it into shared memory, where the 8x4 / 4x8 fragments of
// mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4) are 256 contiguous bytes: lane l reads double l.
// ---------------------------------------------------------------------------------------
constexpr int TILE_ROWS = 128;                          // rows (V) or candidates (K*) per tile
constexpr int TILE_K = 16;                              // k columns per tile
constexpr int TILE_DOUBLES = TILE_ROWS * TILE_K;        // 2048
constexpr int TILE_BYTES = TILE_DOUBLES * 8;            // 16384
constexpr int KT_PER_BLOCK = TILE_ROWS / TILE_K;        // k-tiles per 128-row block = 8

// ---------------------------------------------------------------------------------------
// INT8 split of the FP64 variance contraction (ozaki.cu): operands as OZ_NS balanced base-256
// digits; tcgen05.mma kind::i8 tiles of 128 candidates x OZ_ROWS rows of V x OZ_KC k.
// ---------------------------------------------------------------------------------------
constexpr int OZ_NS = 7;                                // int8 slices per operand
constexpr int OZ_ROWS = 64;                             // rows of V per MMA (N dimension)
constexpr int OZ_KC = 32;                               // k per MMA (32 int8)
constexpr int OZ_A_BYTES = TILE_ROWS * OZ_KC;           // one K* slice of one chunk: 4096
constexpr int OZ_B_BYTES = OZ_ROWS * OZ_KC;             // one V slice of one chunk: 2048
constexpr int OZ_STAGE_BYTES = OZ_NS * (OZ_A_BYTES + OZ_B_BYTES);   // 43008
constexpr int OZ_STAGES = 5;

__host__ __device__ inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

// offset (in tiles) of V tile (jb, kt), kt < (jb+1)*8, lower block triangle packed by rows
__host__ __device__ inline int64_t vtile_index(int jb, int kt) {
  return (int64_t)KT_PER_BLOCK * jb * (jb + 1) / 2 + kt;
}
__host__ __device__ inline int64_t vtile_count(int n_row_blocks) {
  return (int64_t)KT_PER_BLOCK * n_row_blocks * (n_row_blocks + 1) / 2;
}
// position of element (row r in 0..127, k in 0..15) inside a tile
__host__ __device__ inline int tile_elem(int r, int k) {
  return ((k >> 2) * TILE_ROWS + r) * 4 + (k & 3);
}

#ifdef __CUDACC__
// ---------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

// TMA 1-D bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP).
// dst/src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void tma_bulk_g2s(void* smem_dst, const void* gmem_src,
                                             uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// FP64 tensor-core MMA, D(8x8) += A(8x4, row) * B(4x8, col).  Lane l holds
// A[l/4][l%4], B[k=l%4][n=l/4], C[l/4][2*(l%4) + {0,1}].  SASS: DMMA.8x8x4.
__device__ __forceinline__ void dmma_8x8x4(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// shared-memory counter increment with acquire-release semantics at CTA scope
__device__ __forceinline__ int atom_add_acq_rel_shared(int* p, int v) {
  int old;
  asm volatile("atom.acq_rel.cta.shared::cta.add.s32 %0, [%1], %2;"
               : "=r"(old)
               : "r"(smem_u32(p)), "r"(v)
               : "memory");
  return old;
}

#endif  // __CUDACC__

}  // namespace gpry
