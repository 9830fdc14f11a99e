// Variance contraction on the INT8 tensor cores (tcgen05.mma kind::i8, accumulators in TMEM).
//
// ssq_i = sum_j (sum_k V_jk k*_ik)^2 is ~97 % of the flops of predict(return_std) and runs at
// the FP64 tensor peak in predict.cu (DMMA, 0.99 of cuBLAS DGEMM).  B200's INT8 tensor pipe is
// two orders of magnitude faster than its FP64 pipe, so the product is re-expressed exactly in
// integers (the "Ozaki scheme"):
//   k*_ik / c      in [0, 1]  -> round(. 2^54) = sum_p a_p 256^(6-p),  a_p int8 (balanced digits)
//   V_jk / 2^e_j   in (-1, 1) -> round(. 2^54) = sum_q b_q 256^(6-q),  e_j: exponent of max_k |V_jk|
//   sum_k V_jk k*_ik = c 2^e_j 2^-12 sum_g 256^-g S_g,   S_g = sum_{p+q=g} sum_k a_p b_q
// Every S_g is an exact int32 sum (|S_g| <= 7 N 2^14 < 2^31 for N <= 16384); groups g <= 6 are
// kept (28 digit products), which leaves a relative error of ~2^-50 per row, the size of the
// rounding error of an FP64 dot product of this length.  The digits of K* are produced by
// kstar_build (predict.cu, WMODE 2), the digits of V once per upload (slice_v_kernel).
//
// oz_contract_kernel: one CTA per (row split, candidate tile); warp 4 streams the operand
// chunks with TMA bulk copies into a 5-stage ring, one thread of warp 5 issues the 28 MMAs of a
// chunk (128 candidates x 64 rows x 32 k each, 7 accumulators of 64 columns in TMEM), warps 0-3
// read the accumulators back after the last chunk of a row block (thread = candidate), rebuild
// the FP64 row products and accumulate their squares.
#include "state.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <vector>

namespace gpry {

namespace {

__device__ __forceinline__ uint64_t oz_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  // K-major, no swizzle: 8 rows x 16 B core matrices (128 contiguous bytes); SBO between 8-row
  // groups, LBO between the two 16-byte k halves; descriptor version 1 (sm_100)
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
// instruction descriptor: D = S32, A = B = signed 8 bit, both K-major, N >> 3, M >> 4
constexpr uint32_t OZ_IDESC = (2u << 4) | (1u << 7) | (1u << 10) |
                              ((uint32_t)(OZ_ROWS >> 3) << 17) | ((uint32_t)(TILE_ROWS >> 4) << 24);

__device__ __forceinline__ void oz_mma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(OZ_IDESC), "r"(accumulate), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void oz_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
               ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void oz_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// one lane of a converged warp (elect.sync): the compiler then treats the region as run by a
// single, known thread and issues the uniform-datapath tcgen05 instructions straight-line,
// without an ELECT + branch loop around each of them
__device__ __forceinline__ bool oz_elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
// a wait that cannot hang the GPU: traps after ~seconds
__device__ __forceinline__ void oz_wait(uint64_t* bar, uint32_t parity) {
  for (long long it = 0; it < 400000000LL; it++)
    if (mbar_try_wait(bar, parity)) return;
  __trap();
}
// exact int32 -> double without the (quarter-rate) I2F.F64: the double with high word 0x43300000
// and low word L is 2^52 + L; L = v + 2^31 (mod 2^32) for signed v
__device__ __forceinline__ double oz_i2d(uint32_t v) {
  return __hiloint2double(0x43300000, (int)(v ^ 0x80000000u)) - 4503601774854144.0;   // 2^52 + 2^31
}
#define OZ_TMEM_LD16(taddr, r)                                                                   \
  asm volatile(                                                                                  \
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "                                                  \
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"                          \
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),      \
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), \
        "=r"(r[14]), "=r"(r[15])                                                                 \
      : "r"(taddr))

constexpr int OZ_THREADS = 192;   // warps 0-3 epilogue, 4 TMA producer, 5 MMA issuer
constexpr size_t OZ_SMEM = (size_t)OZ_STAGES * OZ_STAGE_BYTES + 256;

__global__ void __launch_bounds__(OZ_THREADS, 1)
oz_contract_kernel(const uint8_t* __restrict__ Ksl, const uint8_t* __restrict__ Vs, int nKC,
                   const double* __restrict__ row_scale, const int* __restrict__ rb_list,
                   const int* __restrict__ rb_count, int max_rb, double* __restrict__ ssqp,
                   int chunk_cands) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)OZ_STAGES * OZ_STAGE_BYTES);
  uint64_t* empty = full + OZ_STAGES;
  uint64_t* tmem_full = empty + OZ_STAGES;
  uint64_t* tmem_empty = tmem_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int split = blockIdx.x, tile = blockIdx.y;
  const int n_rb = rb_count[split];
  const int* my_rb = rb_list + (size_t)split * max_rb;

  if (tid == 0) {
    for (int s = 0; s < OZ_STAGES; s++) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(tmem_full, 1);
    mbar_init(tmem_empty, 128);
    mbar_fence_init();
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;

  if (warp == 4) {
    if (lane == 0) {   // ---- producer: one 28 KB + one 14 KB bulk copy per chunk
      const uint8_t* Abase = Ksl + (size_t)tile * nKC * (size_t)(OZ_NS * OZ_A_BYTES);
      int it = 0;
      for (int r = 0; r < n_rb; r++) {
        const int rb = my_rb[r];
        const uint8_t* Bbase = Vs + (size_t)rb * nKC * (size_t)(OZ_NS * OZ_B_BYTES);
        const int nch = 2 * (rb + 1);      // rows 64 rb .. 64 rb + 63 are zero beyond k = 64 (rb + 1)
        for (int kc = 0; kc < nch; kc++, it++) {
          const int s = it % OZ_STAGES;
          oz_wait(&empty[s], ((it / OZ_STAGES) & 1) ^ 1);
          uint8_t* dst = smem + (size_t)s * OZ_STAGE_BYTES;
          mbar_expect_tx(&full[s], OZ_STAGE_BYTES);
          tma_bulk_g2s(dst, Abase + (size_t)kc * (OZ_NS * OZ_A_BYTES), OZ_NS * OZ_A_BYTES, &full[s]);
          tma_bulk_g2s(dst + OZ_NS * OZ_A_BYTES, Bbase + (size_t)kc * (OZ_NS * OZ_B_BYTES),
                       OZ_NS * OZ_B_BYTES, &full[s]);
        }
      }
    }
  } else if (warp == 5) {
    if (lane == 0) {   // ---- MMA issuer
      int it = 0;
      for (int r = 0; r < n_rb; r++) {
        const int nch = 2 * (my_rb[r] + 1);
        oz_wait(tmem_empty, (r & 1) ^ 1);          // epilogue has drained the accumulators
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int kc = 0; kc < nch; kc++, it++) {
          const int s = it % OZ_STAGES;
          oz_wait(&full[s], (it / OZ_STAGES) & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t sa = smem_u32(smem + (size_t)s * OZ_STAGE_BYTES);
          const uint64_t da0 = oz_desc(sa, OZ_A_BYTES / 2, 128);
          const uint64_t db0 = oz_desc(sa + OZ_NS * OZ_A_BYTES, OZ_B_BYTES / 2, 128);
          const uint32_t first = kc > 0 ? 1u : 0u;
#pragma unroll
          for (int g = 0; g < OZ_NS; g++) {
#pragma unroll
            for (int p = 0; p <= g; p++)
              oz_mma(tmem + (uint32_t)(g * OZ_ROWS), da0 + (uint64_t)((p * OZ_A_BYTES) >> 4),
                     db0 + (uint64_t)(((g - p) * OZ_B_BYTES) >> 4), p > 0 ? 1u : first);
          }
          oz_commit(&empty[s]);        // frees the ring slot once these MMAs have read it
        }
        oz_commit(tmem_full);          // all products of this row block are in TMEM
      }
    }
  } else {
    // ---- epilogue: thread = candidate (TMEM lane), 64 columns = rows of V, 7 digit groups
    double ssq = 0.0;
    for (int r = 0; r < n_rb; r++) {
      const int rb = my_rb[r];
      oz_wait(tmem_full, r & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
      for (int cc = 0; cc < OZ_ROWS; cc += 16) {
        uint32_t v[OZ_NS][16];
#pragma unroll
        for (int g = 0; g < OZ_NS; g++) OZ_TMEM_LD16(lane_base + (uint32_t)(g * OZ_ROWS + cc), v[g]);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int c = 0; c < 16; c++) {
          double acc = oz_i2d(v[OZ_NS - 1][c]);
#pragma unroll
          for (int g = OZ_NS - 2; g >= 0; g--) acc = fma(acc, 0.00390625, oz_i2d(v[g][c]));
          const double w = acc * __ldg(row_scale + rb * OZ_ROWS + cc + c);
          ssq = fma(w, w, ssq);
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      oz_arrive(tmem_empty);
    }
    ssqp[(size_t)split * chunk_cands + tile * TILE_ROWS + tid] = ssq;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u)
                 : "memory");
}

// ---------------------------------------------------------------------------------------
// Two-pass variant with 128 x 128 x 32 MMAs (1.5x the issue rate of the 128 x 64 shape).
// Only 4 accumulators of 128 columns fit in TMEM, so the digit groups of a 128-row block of V are
// produced in two passes over its k range:
//   pass 1: groups 0..3 (10 products, digits 0..3 of both operands)  -> hi = sum_{g<=3} 256^-g S_g,
//           exact in FP64, parked in an L2-resident scratch line of this SM (128 x 128 doubles)
//   pass 2: groups 4..6 (18 products, all digits)                   -> lo = sum_{g=4..6} 256^(4-g) S_g
//   row product = scale_j (hi + 2^-32 lo)
// ---------------------------------------------------------------------------------------
constexpr int OZ2_ROWS = 128;
#ifndef OZ2_EPI_COLS
#define OZ2_EPI_COLS OZ2_ROWS   // (timing experiments only: fewer columns = shorter epilogue, wrong results)
#endif
constexpr int OZ2_B_BYTES = OZ2_ROWS * OZ_KC;                        // 4096 per digit and chunk
constexpr int OZ2_STAGE_BYTES = OZ_NS * (OZ_A_BYTES + OZ2_B_BYTES);  // 57344
constexpr int OZ2_STAGES = 4;                                        // 224 KB
constexpr size_t OZ2_SMEM = (size_t)OZ2_STAGES * OZ2_STAGE_BYTES + 256;
constexpr int OZ2_PARK_SLOTS = 256;                                  // indexed by %smid

// idesc of a 128 x n x 32 MMA (n a multiple of 16)
__device__ __forceinline__ constexpr uint32_t oz2_idesc(int n) {
  return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(TILE_ROWS >> 4) << 24);
}
__device__ __forceinline__ void oz2_mma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}
// Structural zeros of the lower-triangular V inside a 128-row block rb (k chunks of 32):
//  * chunk kc = 4 rb + q (q = 0..3, the diagonal 128 x 128 block): rows below 32 q are zero in
//    this chunk, so the MMA covers rows 32 q .. only (N = rows_blk - 32 q, operand and
//    accumulator start shifted by 32 q rows / columns);
//  * the last row block holds rows_blk = round_up(N_train - 128 rb, 16) <= 128 real rows: the
//    MMAs stop there and its diagonal chunks with 32 q >= rows_blk do not exist.
__device__ __forceinline__ int oz2_rows_blk(int rb, int n_train) {
  const int left = n_train - rb * OZ2_ROWS;
  return left >= OZ2_ROWS ? OZ2_ROWS : ((left + 15) & ~15);
}
__device__ __forceinline__ int oz2_nch(int rb, int rows_blk) {
  return 4 * rb + ((rows_blk + 31) >> 5);
}

// The MMAs of one k chunk of pass PASS (0: digit groups 0..3, 1: groups 4..6).  db0 describes V
// digit 0 at the chunk's first row (k halves nd * 2048 bytes apart); digit q is 2048 bytes on.
// PAIR: one N = 256 MMA per two consecutive V digits (idesc = the 256-column one), with the
// accumulator-enable of the pair taken from its first group (both groups are in the same state:
// the first product into every accumulator is the p = 0 one).
template <int PASS, bool PAIR>
__device__ __forceinline__ void oz2_issue_chunk(uint32_t td, uint64_t da0, uint64_t db0,
                                                uint32_t idesc, uint32_t first) {
  constexpr int G0 = PASS == 0 ? 0 : 4, G1 = PASS == 0 ? 3 : 6;
  constexpr uint32_t HALF = OZ2_B_BYTES / 2;
  const uint32_t idesc1 = oz2_idesc(OZ2_ROWS);       // (PAIR only: the odd digit left over)
#pragma unroll
  for (int p = 0; p <= G1; p++) {
    const int qlo = G0 - p > 0 ? G0 - p : 0, qhi = G1 - p;
    const uint64_t da = da0 + (uint64_t)((p * OZ_A_BYTES) >> 4);
    const uint32_t en = p == 0 ? first : 1u;
#pragma unroll
    for (int q = qlo; q <= qhi; q++) {
      const uint32_t tcol = td + (uint32_t)((p + q - G0) * OZ2_ROWS);
      const uint64_t db = db0 + (uint64_t)((q * HALF) >> 4);
      if (PAIR) {
        if (((q - qlo) & 1) == 0) oz2_mma(tcol, da, db, q + 1 <= qhi ? idesc : idesc1, en);
      } else {
        oz2_mma(tcol, da, db, idesc, en);
      }
    }
  }
}

__global__ void __launch_bounds__(OZ_THREADS, 1)
oz2_contract_kernel(const uint8_t* __restrict__ Ksl, const uint8_t* __restrict__ Vs, int nKC,
                    const double* __restrict__ row_scale, const int* __restrict__ rb_list,
                    const int* __restrict__ rb_count, int max_rb, double* __restrict__ park,
                    double* __restrict__ ssqp, int chunk_cands, int n_splits, int n_tiles,
                    int n_train, int dbg) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)OZ2_STAGES * OZ2_STAGE_BYTES);
  uint64_t* empty = full + OZ2_STAGES;
  uint64_t* tmem_full = empty + OZ2_STAGES;
  uint64_t* tmem_empty = tmem_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // persistent: work item w = (tile, split), split fastest, so that the row splits of one
  // candidate tile run at the same time on neighbouring SMs and share its K* digits in L2
  const int n_items = n_splits * n_tiles;

  if (tid == 0) {
    for (int s = 0; s < OZ2_STAGES; s++) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(tmem_full, 1);
    mbar_init(tmem_empty, 128);
    mbar_fence_init();
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;

  if (warp == 4) {
    if (oz_elect_one()) {   // ---- producer
      int it = 0;
      for (int w = blockIdx.x; w < n_items; w += gridDim.x) {
      const int split = w % n_splits, tile = w / n_splits;
      const int n_rb = rb_count[split];
      const int* my_rb = rb_list + (size_t)split * max_rb;
      const uint8_t* Abase = Ksl + (size_t)tile * nKC * (size_t)(OZ_NS * OZ_A_BYTES);
      for (int r = 0; r < n_rb; r++) {
        const int rb = my_rb[r];
        const uint8_t* Bbase = Vs + (size_t)rb * nKC * (size_t)(OZ_NS * OZ2_B_BYTES);
        const int nch = oz2_nch(rb, oz2_rows_blk(rb, n_train));   // no k beyond the block's last row
        for (int pass = 0; pass < 2; pass++) {
          const uint32_t nd = pass == 0 ? 4u : (uint32_t)OZ_NS;      // digits of each operand needed
          for (int kc = 0; kc < nch; kc++, it++) {
            const int s = it % OZ2_STAGES;
            oz_wait(&empty[s], ((it / OZ2_STAGES) & 1) ^ 1);
            uint8_t* dst = smem + (size_t)s * OZ2_STAGE_BYTES;
            if (dbg & 1) {                 // (timing experiments: no operand traffic, wrong results)
              oz_arrive(&full[s]);
              continue;
            }
            if (dbg & 8) {                 // (timing experiments: no V traffic)
              mbar_expect_tx(&full[s], nd * OZ_A_BYTES);
              tma_bulk_g2s(dst, Abase + (size_t)kc * (OZ_NS * OZ_A_BYTES), nd * OZ_A_BYTES, &full[s]);
              continue;
            }
            mbar_expect_tx(&full[s], nd * (OZ_A_BYTES + OZ2_B_BYTES));
            tma_bulk_g2s(dst, Abase + (size_t)kc * (OZ_NS * OZ_A_BYTES), nd * OZ_A_BYTES, &full[s]);
            // V digits: global [k16][digit][row][16 B]; the stage keeps the nd digits of each k
            // half together (pass 1: two pieces, pass 2: the whole chunk in one)
            const uint8_t* Bsrc = Bbase + (size_t)kc * (OZ_NS * OZ2_B_BYTES);
            uint8_t* Bdst = dst + OZ_NS * OZ_A_BYTES;
            if (pass == 0) {
              tma_bulk_g2s(Bdst, Bsrc, nd * (OZ2_B_BYTES / 2), &full[s]);
              tma_bulk_g2s(Bdst + nd * (OZ2_B_BYTES / 2), Bsrc + OZ_NS * (OZ2_B_BYTES / 2),
                           nd * (OZ2_B_BYTES / 2), &full[s]);
            } else {
              tma_bulk_g2s(Bdst, Bsrc, OZ_NS * OZ2_B_BYTES, &full[s]);
            }
          }
        }
      }
      }
    }
  } else if (warp == 5) {
    if (oz_elect_one()) {   // ---- MMA issuer
      int it = 0, t = 0;
      for (int w = blockIdx.x; w < n_items; w += gridDim.x) {
      const int split = w % n_splits;
      const int n_rb = rb_count[split];
      const int* my_rb = rb_list + (size_t)split * max_rb;
      for (int r = 0; r < n_rb; r++) {
        const int rb = my_rb[r];
        const int rows_blk = oz2_rows_blk(rb, n_train);
        const int nch = oz2_nch(rb, rows_blk);
        for (int pass = 0; pass < 2; pass++, t++) {
          oz_wait(tmem_empty, (t & 1) ^ 1);        // the previous epilogue has drained TMEM
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          for (int kc = 0; kc < nch; kc++, it++) {
            const int s = it % OZ2_STAGES;
            oz_wait(&full[s], (it / OZ2_STAGES) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t sa = smem_u32(smem + (size_t)s * OZ2_STAGE_BYTES);
            // rows [row0, rows_blk) of the block are non-zero in this chunk
            const int row0 = kc > 4 * rb ? 32 * (kc - 4 * rb) : 0;
            const int nrows = (dbg & 2) ? 64 : rows_blk - row0;
            const uint32_t nd = pass == 0 ? 4u : (uint32_t)OZ_NS;
            const uint32_t half = OZ2_B_BYTES / 2;                  // one digit, one k half
            const uint64_t da0 = oz_desc(sa, OZ_A_BYTES / 2, 128);
            const uint64_t db0 = oz_desc(sa + OZ_NS * OZ_A_BYTES + row0 * 16, nd * half, 128);
            const uint32_t td = tmem + (uint32_t)row0;
            const uint32_t first = kc > 0 ? 1u : 0u;
            // Full 128-row chunks: ONE MMA of N = 256 covers two consecutive V digits q, q + 1
            // (adjacent in shared memory) against the same K* digit p, i.e. the adjacent
            // accumulators of groups p + q and p + q + 1: 6 + 12 instructions per chunk instead of
            // 10 + 18, and the 4 KB K* operand is read once per pair.  Narrow chunks (diagonal,
            // ragged last block): one digit at a time, N = nrows.  (Four fully unrolled variants:
            // every descriptor is a base plus an immediate, see DESIGN.md "MMA issue loop".)
            const bool pair = (nrows == OZ2_ROWS) && !(dbg & 32);
            if (pass == 0) {
              if (pair) oz2_issue_chunk<0, true>(td, da0, db0, oz2_idesc(2 * OZ2_ROWS), first);
              else oz2_issue_chunk<0, false>(td, da0, db0, oz2_idesc(nrows), first);
            } else {
              if (pair) oz2_issue_chunk<1, true>(td, da0, db0, oz2_idesc(2 * OZ2_ROWS), first);
              else oz2_issue_chunk<1, false>(td, da0, db0, oz2_idesc(nrows), first);
            }
            oz_commit(&empty[s]);
          }
          oz_commit(tmem_full);
        }
      }
      }
    }
  } else {
    // ---- epilogue: thread = candidate (TMEM lane); 128 columns = rows of V
    uint32_t smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    double* my_park = park + ((size_t)(smid % OZ2_PARK_SLOTS) * OZ2_ROWS) * TILE_ROWS + tid;
    const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
    int t = 0;
    for (int w = blockIdx.x; w < n_items; w += gridDim.x) {
    const int split = w % n_splits, tile = w / n_splits;
    const int n_rb = rb_count[split];
    const int* my_rb = rb_list + (size_t)split * max_rb;
    double ssq = 0.0;
    for (int r = 0; r < n_rb; r++) {
      const int rb = my_rb[r];
      const int ncols = (dbg & 4) ? 16 : min(OZ2_EPI_COLS, oz2_rows_blk(rb, n_train));   // columns the MMAs wrote
      // pass 1: hi = S0 + S1/256 + S2/256^2 + S3/256^3 (exact), parked per (column, candidate)
      oz_wait(tmem_full, t & 1);
      t++;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
      for (int cc = 0; cc < ncols; cc += 16) {
        uint32_t v[4][16];
#pragma unroll
        for (int g = 0; g < 4; g++) OZ_TMEM_LD16(lane_base + (uint32_t)(g * OZ2_ROWS + cc), v[g]);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int c = 0; c < 16; c++) {
          double hi = oz_i2d(v[3][c]);
#pragma unroll
          for (int g = 2; g >= 0; g--) hi = fma(hi, 0.00390625, oz_i2d(v[g][c]));
          my_park[(size_t)(cc + c) * TILE_ROWS] = hi;
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      oz_arrive(tmem_empty);
      // pass 2: lo = S4 + S5/256 + S6/256^2; row product = scale (hi + 2^-32 lo)
      oz_wait(tmem_full, t & 1);
      t++;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
      for (int cc = 0; cc < ncols; cc += 16) {
        uint32_t v[3][16];
#pragma unroll
        for (int g = 0; g < 3; g++) OZ_TMEM_LD16(lane_base + (uint32_t)(g * OZ2_ROWS + cc), v[g]);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int c = 0; c < 16; c++) {
          double lo = oz_i2d(v[2][c]);
          lo = fma(lo, 0.00390625, oz_i2d(v[1][c]));
          lo = fma(lo, 0.00390625, oz_i2d(v[0][c]));
          const double acc = fma(lo, 2.3283064365386963e-10, my_park[(size_t)(cc + c) * TILE_ROWS]);
          const double w = acc * __ldg(row_scale + rb * OZ2_ROWS + cc + c);
          ssq = fma(w, w, ssq);
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      oz_arrive(tmem_empty);
    }
    ssqp[(size_t)split * chunk_cands + tile * TILE_ROWS + tid] = ssq;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u)
                 : "memory");
}

// per row of V (row-major, padded to Np): 2^e with |V_jk| / 2^e < 1
// row_stat[j] = |V_j|_2^2 + 4^e_j (j + 1): what the a-priori error estimate of the split needs
__global__ void __launch_bounds__(256)
oz_row_exponent_kernel(const double* __restrict__ Vrm, int Np, double* __restrict__ row_pow2,
                       double* __restrict__ row_stat) {
  const int j = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (j >= Np) return;
  double mx = 0.0, ss = 0.0;
  for (int k = lane; k <= j; k += 32) {
    const double v = Vrm[(size_t)j * Np + k];
    mx = fmax(mx, fabs(v));
    ss = fma(v, v, ss);
  }
  for (int o = 16; o > 0; o >>= 1) {
    mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    ss += __shfl_xor_sync(0xffffffffu, ss, o);
  }
  if (lane == 0) {
    int e = 0;
    if (mx > 0.0) frexp(mx, &e);       // mx = m 2^e, m in [0.5, 1)
    const double p2 = ldexp(1.0, e);
    row_pow2[j] = p2;
    row_stat[j] = mx > 0.0 ? ss + p2 * p2 * (double)(j + 1) : 0.0;
  }
}
// digits of V: thread per (row j, 16 consecutive k); layout [row block][k chunk of 32][slice]
// [k16 (2)][row in block][16 B], row blocks of 64 (one-pass kernel) or 128 (two-pass kernel)
__global__ void __launch_bounds__(256)
oz_slice_v_kernel(const double* __restrict__ Vrm, int Np, const double* __restrict__ row_pow2,
                  int rows_per_block, uint8_t* __restrict__ Vs) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int n16 = Np / 16;
  if (e >= (int64_t)Np * n16) return;
  const int j = (int)(e / n16), k0 = (int)(e % n16) * 16;
  const double sc = 18014398509481984.0 / row_pow2[j];     // 2^54 / 2^e_j (exact)
  uint32_t packs[OZ_NS][4];
#pragma unroll
  for (int p = 0; p < OZ_NS; p++)
#pragma unroll
    for (int w = 0; w < 4; w++) packs[p][w] = 0u;
#pragma unroll
  for (int b = 0; b < 16; b++) {
    const int k = k0 + b;
    const double val = k <= j ? Vrm[(size_t)j * Np + k] : 0.0;
    long long t = __double2ll_rn(val * sc);
#pragma unroll
    for (int p = OZ_NS - 1; p >= 1; p--) {
      const int dg = (int)(signed char)(t & 0xFF);
      t = (t - dg) >> 8;
      packs[p][b >> 2] |= (uint32_t)(dg & 0xFF) << (8 * (b & 3));
    }
    packs[0][b >> 2] |= (uint32_t)((int)t & 0xFF) << (8 * (b & 3));
  }
  const int nKC = Np / OZ_KC;
  const int b_bytes = rows_per_block * OZ_KC;      // one digit of one chunk
  uint8_t* chunk = Vs + ((size_t)(j / rows_per_block) * nKC + (k0 >> 5)) * (size_t)(OZ_NS * b_bytes);
  if (rows_per_block == 128) {
    // two-pass kernel: [k16 (2)][digit][row (128)][16 B] -- the rows of consecutive digits are
    // contiguous, so ONE MMA of N = 256 can take two digits (two digit groups) at a time
    uint8_t* base = chunk + ((k0 >> 4) & 1) * (OZ_NS * (b_bytes / 2)) + (j % rows_per_block) * 16;
#pragma unroll
    for (int p = 0; p < OZ_NS; p++)
      *reinterpret_cast<uint4*>(base + (size_t)p * (b_bytes / 2)) =
          make_uint4(packs[p][0], packs[p][1], packs[p][2], packs[p][3]);
    return;
  }
  uint8_t* base = chunk + ((k0 >> 4) & 1) * (b_bytes / 2) + (j % rows_per_block) * 16;
#pragma unroll
  for (int p = 0; p < OZ_NS; p++)
    *reinterpret_cast<uint4*>(base + (size_t)p * b_bytes) =
        make_uint4(packs[p][0], packs[p][1], packs[p][2], packs[p][3]);
}
// row_scale[j] = c 2^e_j 2^-12
__global__ void oz_row_scale_kernel(const double* __restrict__ row_pow2, int Np, double c,
                                    double* __restrict__ row_scale) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < Np) row_scale[j] = c * row_pow2[j] * 0.000244140625;
}

// ---------------------------------------------------------------------------------------
// Roofline denominator: issue rate of tcgen05.mma kind::i8 at its best shape (128 x 256 x 32)
// on operands resident in shared memory, every SM busy.  No memory traffic, no epilogue.
// ---------------------------------------------------------------------------------------
constexpr uint32_t OZ_IDESC_PEAK = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(256 >> 3) << 17) |
                                   ((uint32_t)(128 >> 4) << 24);
__global__ void __launch_bounds__(128, 1) oz_peak_kernel(int reps) {
  extern __shared__ __align__(1024) uint8_t smem[];     // A: 4 x 4096, B: 4 x 8192 (values irrelevant)
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int e = tid; e < (4 * 4096 + 4 * 8192) / 16; e += 128)
    reinterpret_cast<uint4*>(smem)[e] = make_uint4(0x01010101u, 0x01010101u, 0x01010101u, 0x01010101u);
  if (tid == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  uint32_t phase = 0;
  for (int r = 0; r < reps; r++) {
    if (warp == 0 && oz_elect_one()) {
      const uint64_t da0 = oz_desc(smem_u32(smem), 2048, 128);
      const uint64_t db0 = oz_desc(smem_u32(smem) + 4 * 4096, 4096, 128);
#pragma unroll
      for (int p = 0; p < 16; p++) {
#pragma unroll
        for (int kc = 0; kc < 4; kc++)
          asm volatile(
              "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
              "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
              ::"r"(tmem + (uint32_t)((p & 1) * 256)), "l"(da0 + (uint64_t)((kc * 4096) >> 4)),
              "l"(db0 + (uint64_t)((kc * 8192) >> 4)), "r"(OZ_IDESC_PEAK), "r"(kc > 0 ? 1u : 0u), "r"(0u)
              : "memory");
      }
      oz_commit(&bar);
    }
    oz_wait(&bar, phase);
    phase ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u)
                 : "memory");
}

}  // namespace

double ozaki_int8_peak_tops(gpry_state* st) {
  GPRY_CUDA(cudaSetDevice(st->device));
  const size_t smem = 4 * 4096 + 4 * 8192;
  GPRY_CUDA(cudaFuncSetAttribute(oz_peak_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)smem));
  cudaEvent_t e0, e1;
  GPRY_CUDA(cudaEventCreate(&e0));
  GPRY_CUDA(cudaEventCreate(&e1));
  const int reps = 400;
  oz_peak_kernel<<<st->n_sm, 128, smem>>>(20);
  float best = 1e30f;
  for (int t = 0; t < 3; t++) {
    GPRY_CUDA(cudaEventRecord(e0));
    oz_peak_kernel<<<st->n_sm, 128, smem>>>(reps);
    GPRY_CUDA(cudaEventRecord(e1));
    GPRY_CUDA(cudaEventSynchronize(e1));
    float ms;
    GPRY_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    best = std::min(best, ms);
  }
  GPRY_CUDA(cudaGetLastError());
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  const double ops = 2.0 * (double)st->n_sm * reps * 16 * 4 * 128.0 * 256.0 * 32.0;
  return ops / (best * 1e-3) * 1e-12;
}

// The same kernel launched back to back for `seconds`: the rate the board sustains under its
// power cap (what a kernel timed inside a long step can reach), measured over the last 3/4.
double ozaki_int8_peak_sustained_tops(gpry_state* st, double seconds) {
  GPRY_CUDA(cudaSetDevice(st->device));
  const size_t smem = 4 * 4096 + 4 * 8192;
  GPRY_CUDA(cudaFuncSetAttribute(oz_peak_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)smem));
  const int reps = 2000;                                   // ~10 ms per launch
  const int n = std::max(8, (int)(seconds / 0.0095));
  cudaEvent_t e0, e1;
  GPRY_CUDA(cudaEventCreate(&e0));
  GPRY_CUDA(cudaEventCreate(&e1));
  int timed = 0;
  for (int i = 0; i < n; i++) {
    if (i == n / 4) GPRY_CUDA(cudaEventRecord(e0));
    oz_peak_kernel<<<st->n_sm, 128, smem>>>(reps);
    if (i >= n / 4) timed++;
  }
  GPRY_CUDA(cudaEventRecord(e1));
  GPRY_CUDA(cudaEventSynchronize(e1));
  GPRY_CUDA(cudaGetLastError());
  float ms;
  GPRY_CUDA(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  const double ops = 2.0 * (double)st->n_sm * reps * 16 * 4 * 128.0 * 256.0 * 32.0 * timed;
  return ops / (ms * 1e-3) * 1e-12;
}

bool ozaki_supported(const gpry_state* st) {
  if (!(st->has_V && st->Npad >= 512 && st->Npad <= 16384)) return false;
  // the guard (ozaki_prepare + ozaki_validate) has ruled the split out for this model
  return !(st->oz_guard && st->oz_valid && st->oz_checked && !st->oz_ok);
}

// A-priori error model of the split.  Operands are rounded to 55-bit fixed
// point: k*_k / c to a multiple of 2^-54 and V_jk to a multiple of 2^(e_j - 54); digit groups
// g >= 7 of the product are dropped.  For the row product w_j = sum_k V_jk k*_k over n_j = j + 1
// terms:
//   worst case   |dw_j| <= c 2^e_j n_j (1 + 6.02) 2^-54              (all errors aligned)
//   statistical  sd(dw_j) ~= 2.2 c 2^-55 sqrt((|V_j|_2^2 + 4^e_j n_j) / 3)
//                (independent uniform roundings of both operands, rho = rms(k* / c) <= 1;
//                 the dropped groups add about as much again: factor 2.2)
// and for the variance c - sum_j w_j^2 (sum_j w_j^2 <= c):  d var <= 2 sqrt(c) max_j |dw_j|.
static void ozaki_error_model(gpry_state* st, const std::vector<double>& row_stat,
                              const std::vector<double>& row_pow2) {
  double worst = 0.0, stat = 0.0;
  for (int j = 0; j < st->N; j++) {
    worst = std::max(worst, row_pow2[j] * (double)(j + 1));
    stat = std::max(stat, row_stat[j]);
  }
  const double c = st->c, eps = ldexp(1.0, -54);
  st->oz_bound_worst = 2.0 * sqrt(c) * c * worst * 7.02 * eps;
  st->oz_est_sigma = 2.0 * sqrt(c) * 2.2 * c * 0.5 * eps * sqrt(stat / 3.0);
}

// digits of V and the balanced assignment of row blocks to row splits (once per upload and
// variant: contract_mode 1 = two-pass 128-row blocks, 2 = one-pass 64-row blocks)
void ozaki_prepare(gpry_state* st, cudaStream_t s) {
  const int rows = st->contract_mode == 2 ? OZ_ROWS : OZ2_ROWS;
  if (st->oz_valid && st->oz_rows == rows) return;
  const int Np = st->Npad, nRB = Np / rows, nKC = Np / OZ_KC;
  st->oz_Vs.reserve((size_t)nRB * nKC * OZ_NS * rows * OZ_KC);
  st->oz_scale.reserve(3 * (size_t)Np);
  double* row_pow2 = st->oz_scale.p + Np;
  double* row_stat = st->oz_scale.p + 2 * (size_t)Np;
  oz_row_exponent_kernel<<<(Np + 7) / 8, 256, 0, s>>>(st->Vrm.p, Np, row_pow2, row_stat);
  GPRY_CUDA(cudaGetLastError());
  const int64_t tot = (int64_t)Np * (Np / 16);
  oz_slice_v_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(st->Vrm.p, Np, row_pow2, rows,
                                                                  st->oz_Vs.p);
  GPRY_CUDA(cudaGetLastError());
  oz_row_scale_kernel<<<(Np + 255) / 256, 256, 0, s>>>(row_pow2, Np, st->c, st->oz_scale.p);
  GPRY_CUDA(cudaGetLastError());
  // Row blocks beyond the last training row are all zero: skip them.  Row splits: enough of
  // them that the K* digits of the candidate tiles resident at one time (n_sm / splits tiles)
  // stay in L2 (~40 MB), then the split count in that neighbourhood with the best balance.
  const int used_rb = (st->N + rows - 1) / rows;
  const double tile_mb = (double)nKC * OZ_NS * OZ_A_BYTES / 1048576.0;
  int want = (int)std::ceil(st->n_sm * tile_mb / 40.0);
  want = std::max(1, std::min(want, used_rb));
  int splits = want;
  double best_cost = 1e300;
  std::vector<std::vector<int>> lists;
  for (int cand = want; cand <= std::min(used_rb, want + 3); cand++) {
    std::vector<std::vector<int>> l(cand);
    std::vector<long long> load(cand, 0);
    for (int rb = used_rb - 1; rb >= 0; rb--) {     // longest first, to the least loaded split
      int b = 0;
      for (int q = 1; q < cand; q++)
        if (load[q] < load[b]) b = q;
      l[b].push_back(rb);
      load[b] += rb + 1;
    }
    const long long mx = *std::max_element(load.begin(), load.end());
    const double cost = (double)mx * cand;          // makespan x CTAs ~ total SM time
    if (cost < best_cost * 0.999) {
      best_cost = cost;
      splits = cand;
      lists = l;
    }
  }
  int max_rb = 0;
  for (auto& l : lists) max_rb = std::max(max_rb, (int)l.size());
  std::vector<int> h((size_t)splits * max_rb + splits, 0);
  for (int q = 0; q < splits; q++) {
    for (size_t i = 0; i < lists[q].size(); i++) h[(size_t)q * max_rb + i] = lists[q][i];
    h[(size_t)splits * max_rb + q] = (int)lists[q].size();
  }
  st->oz_rb.reserve(h.size());
  GPRY_CUDA(cudaMemcpyAsync(st->oz_rb.p, h.data(), h.size() * sizeof(int), cudaMemcpyHostToDevice, s));
  if (rows == OZ2_ROWS) st->oz_park.reserve((size_t)OZ2_PARK_SLOTS * OZ2_ROWS * TILE_ROWS);
  std::vector<double> h_stat(2 * (size_t)Np);
  GPRY_CUDA(cudaMemcpyAsync(h_stat.data(), row_pow2, 2 * (size_t)Np * 8, cudaMemcpyDeviceToHost, s));
  GPRY_CUDA(cudaStreamSynchronize(s));
  ozaki_error_model(st, std::vector<double>(h_stat.begin() + Np, h_stat.end()),
                    std::vector<double>(h_stat.begin(), h_stat.begin() + Np));
  st->oz_splits = splits;
  st->oz_max_rb = max_rb;
  st->oz_rows = rows;
  GPRY_CUDA(cudaFuncSetAttribute(oz_contract_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)OZ_SMEM));
  GPRY_CUDA(cudaFuncSetAttribute(oz2_contract_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)OZ2_SMEM));
  st->oz_valid = true;
  st->oz_checked = false;      // ozaki_validate (predict.cu) decides before the first use
  st->oz_ok = true;
  st->oz_probe_err = -1.0;
}

size_t ozaki_kslices_bytes(const gpry_state* st, int tiles) {
  return (size_t)tiles * (st->Npad / OZ_KC) * OZ_NS * OZ_A_BYTES;
}

static int oz_dbg() {      // GPRY_B200_OZ_DBG: timing experiments only (results are wrong)
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("GPRY_B200_OZ_DBG");
    v = e ? atoi(e) : 0;
  }
  return v;
}

// ssqp[split][chunk_cands] for `tiles` candidate tiles whose digits are in Ksl
void ozaki_contract(gpry_state* st, const uint8_t* Ksl, int tiles, int chunk_cands,
                    cudaStream_t s) {
  dim3 grid(st->oz_splits, tiles);
  const int* counts = st->oz_rb.p + (size_t)st->oz_splits * st->oz_max_rb;
  if (st->oz_rows == OZ2_ROWS)
    oz2_contract_kernel<<<std::min(st->n_sm, st->oz_splits * tiles), OZ_THREADS, OZ2_SMEM, s>>>(
        Ksl, st->oz_Vs.p, st->Npad / OZ_KC, st->oz_scale.p, st->oz_rb.p, counts, st->oz_max_rb,
        st->oz_park.p, st->ssqp.p, chunk_cands, st->oz_splits, tiles, st->N, oz_dbg());
  else
    oz_contract_kernel<<<grid, OZ_THREADS, OZ_SMEM, s>>>(
        Ksl, st->oz_Vs.p, st->Npad / OZ_KC, st->oz_scale.p, st->oz_rb.p, counts, st->oz_max_rb,
        st->ssqp.p, chunk_cands);
  GPRY_CUDA(cudaGetLastError());
}

}  // namespace gpry
