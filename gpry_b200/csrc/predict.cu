// Posterior mean / std / LogExp over a candidate pool -- the candidate-side hot path.
//
// Reference arithmetic (GPry 3.0.0): gpr.py:1176-1227 (predict), 1325-1347 (predict_std),
// acquisition_functions.py:1068-1074 (LogExp.f); kernel values sklearn:kernels.py:971, 1278,
// 1569-1570 (RBF), 1720-1729 (Matern).
//
// Three kernels per chunk of candidates (K* is only ever materialised one chunk at a time,
// in a tile layout private to this library):
//
//   kstar_build   k*[i, j] = c g(|u_i - t_j|) for a tile of 128 candidates x a slice of
//                 training points; training points (pre-divided by the length scales) are
//                 staged into shared memory by TMA bulk copy and broadcast to the warp,
//                 candidate coordinates live in registers.  Emits K* tiles in the
//                 [k-panel][candidate][4] layout the contraction wants plus the partial
//                 means sum_j k*_ij alpha_j.                       (FP64 ALU / exp bound)
//   var_contract  ssq[i] = sum_j (sum_k V[j,k] k*[i,k])^2 with V = L^-1 lower triangular:
//                 block-lower-triangular GEMM on FP64 tensor cores (mma.sync m8n8k4 ->
//                 DMMA.8x8x4), operands streamed by TMA bulk copies through a 6-stage
//                 mbarrier ring, squared-row-sum epilogue kept in registers.  (DMMA bound)
//   finish        var = c - ssq, clamp, sqrt, de-normalise, clip, LogExp: fused into the
//                 tile epilogue of var_contract (a separate small kernel only when the row
//                 blocks of V are split over several CTAs for small pools, or mean only).
#include <math.h>

#include <algorithm>
#include <cstdlib>

#include "state.cuh"

namespace gpry {

// ---------------------------------------------------------------------------------------
// stationary kernel g(r2) * c
// ---------------------------------------------------------------------------------------
template <int KIND>
__device__ __forceinline__ double kernel_value(double r2, double c) {
  if (KIND == GPRY_KERNEL_RBF) {
    return c * exp(-0.5 * r2);                        // sklearn:kernels.py:1570
  } else if (KIND == GPRY_KERNEL_MATERN15) {
    double K = sqrt(r2) * 1.7320508075688772;         // dists * sqrt(3)   :1725
    return c * ((1.0 + K) * exp(-K));                 // :1726
  } else {
    double K = sqrt(r2) * 2.23606797749979;           // dists * sqrt(5)   :1728
    return c * ((1.0 + K + K * K / 3.0) * exp(-K));   // :1729
  }
}

// ---------------------------------------------------------------------------------------
// INT8 digits of one kernel value (ozaki.cu): k*/c in [0, 1] as the 55-bit fixed-point number
// t = hi 2^24 + lo in OZ_NS = 7 balanced base-256 digits, digit p into byte jj of packs[p][q4].
// Both halves come out of the mantissa of (x + 1.5 2^52) -- no 64-bit integer arithmetic, no
// F2I: hi = rint(x 2^30) exactly, lo = rint of the exact remainder times 2^24.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void oz_push_digits(double kv, double slice_scale /* 2^30 / c */,
                                               uint32_t (&packs)[OZ_NS][4], int q4, int jj) {
  const double xs = kv * slice_scale;
  const double m1 = xs + 6755399441055744.0;
  const int hi = __double2loint(m1);
  const double rem = xs - (m1 - 6755399441055744.0);    // exact, |rem| <= 0.5
  const int lo = __double2loint(fma(rem, 16777216.0, 6755399441055744.0));
  // Balanced digits without a digit-by-digit carry chain: with B = 0x80 in each of the six low
  // byte positions, the bytes of W = t + B are the digits plus 128, so (byte ^ 0x80) read as a
  // signed byte IS the balanced digit (their weighted sum is t + B - B = t, digits in
  // [-128, 127], and that representation is unique); bits 48.. of W are the top digit as is.
  uint32_t wl, wh;
  asm("{\n\t.reg .u32 t0, t1;\n\t"
      "add.cc.u32 t0, %2, 0x80808080;\n\t"      // sign-extended lo + B
      "addc.u32 t1, %3, 0x8080;\n\t"
      "add.cc.u32 %0, t0, %4;\n\t"              // + hi 2^24
      "addc.u32 %1, t1, %5;\n\t}"
      : "=r"(wl), "=r"(wh)
      : "r"(lo), "r"(lo >> 31), "r"((uint32_t)hi << 24), "r"(hi >> 8));
  wl ^= 0x80808080u;
  wh ^= 0x00008080u;
  // byte jj of packs[p][q4] <- digit p  (digit 6 = byte 0 of wl ... digit 0 = byte 2 of wh)
  const uint32_t keep = 0x3210u & ~(0xFu << (4 * jj));
#define OZ_SEL(b) (keep | ((4u + (b)) << (4 * jj)))
  packs[6][q4] = __byte_perm(packs[6][q4], wl, OZ_SEL(0u));
  packs[5][q4] = __byte_perm(packs[5][q4], wl, OZ_SEL(1u));
  packs[4][q4] = __byte_perm(packs[4][q4], wl, OZ_SEL(2u));
  packs[3][q4] = __byte_perm(packs[3][q4], wl, OZ_SEL(3u));
  packs[2][q4] = __byte_perm(packs[2][q4], wh, OZ_SEL(0u));
  packs[1][q4] = __byte_perm(packs[1][q4], wh, OZ_SEL(1u));
  packs[0][q4] = __byte_perm(packs[0][q4], wh, OZ_SEL(2u));
#undef OZ_SEL
}
// [tile][k-chunk of 32][digit][k16 (2)][candidate (128)][16 B]: the shared-memory image of the
// K-major operand of tcgen05.mma kind::i8
__device__ __forceinline__ void oz_store_digits(void* Kout, int tile, int nKT, int k0, int tid,
                                                const uint32_t (&packs)[OZ_NS][4]) {
  uint8_t* base = reinterpret_cast<uint8_t*>(Kout) +
                  ((size_t)tile * (nKT >> 1) + (k0 >> 5)) * (size_t)(OZ_NS * OZ_A_BYTES) +
                  ((k0 >> 4) & 1) * (OZ_A_BYTES / 2) + tid * 16;
#pragma unroll
  for (int p = 0; p < OZ_NS; p++)
    *reinterpret_cast<uint4*>(base + p * OZ_A_BYTES) =
        make_uint4(packs[p][0], packs[p][1], packs[p][2], packs[p][3]);
}

// ---------------------------------------------------------------------------------------
// kstar_build
//   grid  = (tiles in chunk, JS)      block = 128 threads = the 128 candidates of a tile
//   each block: NJ = Npad / JS training points (a multiple of 16)
// ---------------------------------------------------------------------------------------
template <int DP, int KIND, int WMODE>
__global__ void __launch_bounds__(128)
kstar_build_kernel(const double* __restrict__ X, int64_t M, int d, int64_t cand0,
                   const double* __restrict__ T, const double* __restrict__ alpha, int NJ,
                   int nKT, double c, XformParams prm, void* __restrict__ Kout,
                   double* __restrict__ meanp, int chunk_cands, double slice_scale) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* Ts = reinterpret_cast<double*>(smem_raw);          // [NJ][DP]
  double* As = Ts + (size_t)NJ * DP;                         // [NJ]
  double* Xs = As + NJ;                                      // [128][d]
  uint64_t* bar = reinterpret_cast<uint64_t*>(Xs + 128 * d + (d & 1));

  const int tid = threadIdx.x;
  const int tile = blockIdx.x;
  const int j0 = blockIdx.y * NJ;

  if (tid == 0) {
    mbar_init(bar, 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (tid == 0) {
    const uint32_t bytesT = (uint32_t)NJ * DP * 8, bytesA = (uint32_t)NJ * 8;
    mbar_expect_tx(bar, bytesT + bytesA);
    tma_bulk_g2s(Ts, T + (size_t)j0 * DP, bytesT, bar);
    tma_bulk_g2s(As, alpha + j0, bytesA, bar);
  }
  // coalesced load of the candidate tile (contiguous 128 * d doubles)
  const int64_t cbase = cand0 + (int64_t)tile * TILE_ROWS;
  const int64_t avail = (M - cbase) * d;   // doubles available from X + cbase*d
  const double* Xg = X + cbase * d;
  for (int e = tid; e < TILE_ROWS * d; e += 128) Xs[e] = (e < avail) ? __ldg(Xg + e) : 0.0;
  __syncthreads();

  // ((x - min) / width) / ell : preprocessing.py:380 then sklearn:kernels.py:1569 (X / l)
  double u[DP];
#pragma unroll
  for (int k = 0; k < DP; k++) {
    if (k < d) {
      double x = Xs[tid * d + k];
      u[k] = ((x - prm.x_min[k]) / prm.x_width[k]) / prm.ell[k];
    } else {
      u[k] = 0.0;
    }
  }
  mbar_wait(bar, 0);

  double mp = 0.0;
  double* kout = reinterpret_cast<double*>(Kout) + ((size_t)tile * nKT + (j0 >> 4)) * TILE_DOUBLES +
                 tid * 4;
  for (int j16 = 0; j16 < NJ; j16 += 16) {
    uint32_t packs[OZ_NS][4];     // WMODE 2: 16 int8 digits per slice for k = j0 + j16 .. + 15
    if (WMODE == 2) {
#pragma unroll
      for (int p = 0; p < OZ_NS; p++)
#pragma unroll
        for (int w = 0; w < 4; w++) packs[p][w] = 0u;
    }
#pragma unroll
    for (int q4 = 0; q4 < 4; q4++) {
      const int j4 = j16 + q4 * 4;
      double kv[4];
#pragma unroll
      for (int jj = 0; jj < 4; jj++) {
        const double2* tp = reinterpret_cast<const double2*>(Ts + (size_t)(j4 + jj) * DP);
        double r2 = 0.0;      // summed in index order, as cdist does (sklearn:kernels.py:1569)
#pragma unroll
        for (int k2 = 0; k2 < DP / 2; k2++) {
          double2 t = tp[k2];
          double d0 = u[2 * k2] - t.x;
          r2 = fma(d0, d0, r2);
          double d1 = u[2 * k2 + 1] - t.y;
          r2 = fma(d1, d1, r2);
        }
        kv[jj] = kernel_value<KIND>(r2, c);
        mp = fma(kv[jj], As[j4 + jj], mp);
        if (WMODE == 2) oz_push_digits(kv[jj], slice_scale, packs, q4, jj);
      }
      if (WMODE == 1) {
        // tile (j4 / 16), panel (j4 / 4) % 4, row tid
        double* p = kout + (size_t)(j4 >> 4) * TILE_DOUBLES + ((j4 >> 2) & 3) * (TILE_ROWS * 4);
        reinterpret_cast<double2*>(p)[0] = make_double2(kv[0], kv[1]);
        reinterpret_cast<double2*>(p)[1] = make_double2(kv[2], kv[3]);
      }
    }
    if (WMODE == 2) oz_store_digits(Kout, tile, nKT, j0 + j16, tid, packs);
  }
  meanp[(size_t)blockIdx.y * chunk_cands + tile * TILE_ROWS + tid] = mp;
}

// generic-d variant (d > 32): candidate coordinates stay in shared memory
template <int KIND, int WMODE>
__global__ void __launch_bounds__(128)
kstar_build_generic_kernel(const double* __restrict__ X, int64_t M, int d, int DP, int64_t cand0,
                           const double* __restrict__ T, const double* __restrict__ alpha,
                           int NJ, int nKT, double c, const double* __restrict__ prm,
                           void* __restrict__ Kout, double* __restrict__ meanp,
                           int chunk_cands, double slice_scale) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* Ts = reinterpret_cast<double*>(smem_raw);   // [NJ][DP]
  double* As = Ts + (size_t)NJ * DP;                  // [NJ]
  double* Us = As + NJ;                               // [DP][129]  (transposed, padded)
  const int tid = threadIdx.x, tile = blockIdx.x, j0 = blockIdx.y * NJ;
  for (int e = tid; e < NJ * DP; e += 128) Ts[e] = T[(size_t)j0 * DP + e];
  for (int e = tid; e < NJ; e += 128) As[e] = alpha[j0 + e];
  const int64_t cbase = cand0 + (int64_t)tile * TILE_ROWS;
  const int64_t cand = cbase + tid;
  for (int k = 0; k < DP; k++) {
    double v = 0.0;
    if (k < d && cand < M) {
      double x = X[cand * d + k];
      v = ((x - prm[k]) / prm[MAX_DIM + k]) / prm[2 * MAX_DIM + k];
    }
    Us[k * 129 + tid] = v;
  }
  __syncthreads();
  double mp = 0.0;
  double* kout = reinterpret_cast<double*>(Kout) + ((size_t)tile * nKT + (j0 >> 4)) * TILE_DOUBLES +
                 tid * 4;
  for (int j16 = 0; j16 < NJ; j16 += 16) {
    uint32_t packs[OZ_NS][4];
    if (WMODE == 2) {
#pragma unroll
      for (int p = 0; p < OZ_NS; p++)
#pragma unroll
        for (int w = 0; w < 4; w++) packs[p][w] = 0u;
    }
#pragma unroll
    for (int q4 = 0; q4 < 4; q4++) {
      const int j4 = j16 + q4 * 4;
      double kv[4];
#pragma unroll
      for (int jj = 0; jj < 4; jj++) {
        const double* tp = Ts + (size_t)(j4 + jj) * DP;
        double r2 = 0.0;
        for (int k = 0; k < DP; k++) {
          double d0 = Us[k * 129 + tid] - tp[k];
          r2 = fma(d0, d0, r2);
        }
        kv[jj] = kernel_value<KIND>(r2, c);
        mp = fma(kv[jj], As[j4 + jj], mp);
        if (WMODE == 2) oz_push_digits(kv[jj], slice_scale, packs, q4, jj);
      }
      if (WMODE == 1) {
        double* p = kout + (size_t)(j4 >> 4) * TILE_DOUBLES + ((j4 >> 2) & 3) * (TILE_ROWS * 4);
        reinterpret_cast<double2*>(p)[0] = make_double2(kv[0], kv[1]);
        reinterpret_cast<double2*>(p)[1] = make_double2(kv[2], kv[3]);
      }
    }
    if (WMODE == 2) oz_store_digits(Kout, tile, nKT, j0 + j16, tid, packs);
  }
  meanp[(size_t)blockIdx.y * chunk_cands + tile * TILE_ROWS + tid] = mp;
}

// ---------------------------------------------------------------------------------------
// finishing arithmetic (gpr.py:1180-1195, 1207-1227; LogExp.f acquisition_functions.py:
// 1071-1074), shared by the fused epilogue of var_contract and by finish_kernel
// ---------------------------------------------------------------------------------------
struct FinishParams {
  const double* meanp;   // [JS][chunk_cands] partial means
  int JS;
  int chunk_cands;
  int n_valid;           // candidates of this chunk that exist (the last tile may be ragged)
  double c, y_mean, y_std, clip_hi;
  int want_acq;
  double two_zeta, sigma_n2, y_max;
  double* o_mean;        // outputs, already offset to the chunk; any may be NULL
  double* o_std;
  double* o_acq;
};

// mean, std and acquisition of candidate i of the chunk (in registers)
__device__ __forceinline__ void finish_values(const FinishParams& f, int i, bool have_var,
                                              double ssq, double& mean, double& sd, double& acq) {
  double m_ = 0.0;
  for (int s = 0; s < f.JS; s++) m_ += f.meanp[(size_t)s * f.chunk_cands + i];
  mean = m_ * f.y_std + f.y_mean;                 // preprocessing.py:620
  mean = fmin(mean, f.clip_hi);                   // gpr.py:1187-1195 (np.clip upper)
  if (isnan(m_)) mean = m_;
  sd = 0.0;
  acq = 0.0;
  if (have_var) {
    if (isnan(m_)) ssq = m_;                      // a NaN in k* reaches every output (INT8 digits
                                                  // of a NaN are meaningless, the FP64 sum is not)
    double var = f.c - ssq;                       // gpr.py:1207-1208
    if (var < 0.0) var = 0.0;                     // :1214-1219
    sd = sqrt(var) * f.y_std;                     // :1220, preprocessing.py:630
    if (f.want_acq) {
      double v = sd * sd - f.sigma_n2;            // std**2 - noise_level**2
      v = v > 0.0 ? v : 0.0;                      // np.clip(., 0, None)
      acq = f.two_zeta * (mean - f.y_max) + log(sqrt(v));
    }
  }
}

__device__ __forceinline__ void finish_one(const FinishParams& f, int i, bool have_var,
                                           double ssq) {
  double mean, sd, acq;
  finish_values(f, i, have_var, ssq, mean, sd, acq);
  if (f.o_mean) f.o_mean[i] = mean;
  if (have_var) {
    if (f.o_std) f.o_std[i] = sd;
    if (f.want_acq && f.o_acq) f.o_acq[i] = acq;
  }
}

// ---------------------------------------------------------------------------------------
// var_contract
// ---------------------------------------------------------------------------------------
constexpr int VC_STAGES = 6;
constexpr int VC_WARPS = 8;
constexpr int VC_THREADS = VC_WARPS * 32;

struct VcSmem {
  double A[VC_STAGES][TILE_DOUBLES];   // V tiles      [4 panels][128 rows][4]
  double B[VC_STAGES][TILE_DOUBLES];   // K* tiles     [4 panels][128 cands][4]
  double red[2][TILE_ROWS];
  uint64_t full[VC_STAGES];
  int released[VC_STAGES];            // warps done with the stage's current contents
};

// which row-split owns row block jb (snake order: balanced triangular work)
__device__ __forceinline__ int row_block_owner(int jb, int S) {
  int pos = jb % (2 * S);
  return pos < S ? pos : 2 * S - 1 - pos;
}
__device__ __forceinline__ int next_owned_block(int jb, int nJ, int S, int split) {
  while (jb < nJ && row_block_owner(jb, S) != split) jb++;
  return jb;
}

// The (tile, row block, k-tile) sequence a CTA walks; every warp also tracks the position
// VC_STAGES steps ahead of the MMA loop (what the slot it is consuming is refilled with).
struct VcSeq {
  int t, jb, kt, kt_end;
  __device__ __forceinline__ void start(int n_tiles, int nJ, int S, int split, int kt_total) {
    t = blockIdx.x;
    jb = next_owned_block(0, nJ, S, split);
    kt = 0;
    kt_end = min((jb + 1) * KT_PER_BLOCK, kt_total);
    if (jb >= nJ) t = n_tiles;
  }
  __device__ __forceinline__ bool valid(int n_tiles) const { return t < n_tiles; }
  __device__ __forceinline__ void advance(int nJ, int S, int split, int kt_total) {
    if (++kt < kt_end) return;
    kt = 0;
    jb = next_owned_block(jb + 1, nJ, S, split);
    if (jb >= nJ) {
      t += gridDim.x;
      jb = next_owned_block(0, nJ, S, split);
    }
    kt_end = min((jb + 1) * KT_PER_BLOCK, kt_total);
  }
};

// One k-tile (16 columns = 4 panels) of a warp's accumulator tile.  The warp owns the 8-row
// fragments f = 2 mi + rw (mi = 0..7) of the 128-row block -- interleaved between the two
// warps that share an SM sub-partition so that both see the same triangular structure -- and
// 32 candidates.  Fragment f meets only columns k <= row: inside the diagonal block
// (qd = k-tile index relative to the block's first column tile, >= 0) panel p is skipped for
// f < (4 qd + p) / 2; fragments at or beyond mi_max hold only padded rows (>= N).  All
// predicates are warp-uniform; the DMMA pipe (one issue per 16 cycles per sub-partition)
// leaves ample issue slots for them.
#define VC_FRAG(mi)                                                                  \
  {                                                                                  \
    dmma_8x8x4(acc[mi][0][0], acc[mi][0][1], a[mi], b[0]);                           \
    dmma_8x8x4(acc[mi][1][0], acc[mi][1][1], a[mi], b[1]);                           \
    dmma_8x8x4(acc[mi][2][0], acc[mi][2][1], a[mi], b[2]);                           \
    dmma_8x8x4(acc[mi][3][0], acc[mi][3][1], a[mi], b[3]);                           \
  }

// One k-tile for the fragments LO <= mi < HI of the warp (compile-time range: straight-line
// code, no predicates, no branches).  <0, 8> is the interior tile; LO > 0 skips the fragments
// that lie entirely above the diagonal inside a diagonal block (k-tile qd of the block only
// meets fragments f >= 2 qd, i.e. mi >= qd for both row parities); HI < 8 drops fragments made of padded rows only.
template <int LO, int HI>
__device__ __forceinline__ void vc_stage(double (&acc)[8][4][2], const double* __restrict__ As,
                                         const double* __restrict__ Bs) {
#pragma unroll
  for (int p = 0; p < 4; p++) {
    double a[8], b[4];
#pragma unroll
    for (int mi = LO; mi < HI; mi++) a[mi] = As[p * (TILE_ROWS * 4) + mi * 64];
#pragma unroll
    for (int ni = 0; ni < 4; ni++) b[ni] = Bs[p * (TILE_ROWS * 4) + ni * 32];
#pragma unroll
    for (int mi = LO; mi < HI; mi++) VC_FRAG(mi)
  }
}
#undef VC_FRAG

// grid = (min(tiles, #SM), row_splits); 256 threads = 8 MMA warps (warp tile: 8 interleaved
// 8-row fragments x 32 candidates = 32 DMMA accumulator fragments); the TMA refill of a slot
// is issued by the last warp that releases it.
__global__ void __launch_bounds__(VC_THREADS, 1)
var_contract_kernel(const double* __restrict__ Vt, const double* __restrict__ Ks, int n_tiles,
                    int N, int nJ, int nKT, int row_splits, double* __restrict__ ssqp,
                    int chunk_cands, int fused, FinishParams fin) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  VcSmem& sm = *reinterpret_cast<VcSmem*>(smem_raw);
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int split = blockIdx.y;
  const int kt_total = (N + TILE_K - 1) / TILE_K;   // k-tiles that contain real columns

  if (tid == 0) {
    for (int s = 0; s < VC_STAGES; s++) {
      mbar_init(&sm.full[s], 1);
      sm.released[s] = 0;
    }
    mbar_fence_init();
  }
  __syncthreads();

  // ---- TMA issue.  Slot s is refilled by whichever warp releases it last (shared-memory
  // counter), so no warp ever waits for another one except through the data itself. Every
  // warp tracks the step that will next be loaded into the slot it is consuming: the
  // consumer position + VC_STAGES. ----
  auto issue = [&](const VcSeq& q, int slot) {
    mbar_expect_tx(&sm.full[slot], 2 * TILE_BYTES);
    tma_bulk_g2s(sm.A[slot], Vt + (vtile_index(q.jb, 0) + q.kt) * TILE_DOUBLES, TILE_BYTES,
                 &sm.full[slot]);
    tma_bulk_g2s(sm.B[slot], Ks + ((size_t)q.t * nKT + q.kt) * TILE_DOUBLES, TILE_BYTES,
                 &sm.full[slot]);
  };
  VcSeq nseq;
  nseq.start(n_tiles, nJ, row_splits, split, kt_total);
  for (int i = 0; i < VC_STAGES; i++) {
    if (!nseq.valid(n_tiles)) break;
    if (tid == 0) issue(nseq, i);
    nseq.advance(nJ, row_splits, split, kt_total);
  }

  const int rw = warp >> 2;        // row parity: this warp owns the 8-row fragments 2*mi + rw
  const int cw = warp & 3;         // candidate quarter (cands cw*32 .. +31)
  int stage = 0;
  uint32_t phase = 0;

  for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    double ssq[4][2];
#pragma unroll
    for (int ni = 0; ni < 4; ni++) ssq[ni][0] = ssq[ni][1] = 0.0;

    for (int jb = next_owned_block(0, nJ, row_splits, split); jb < nJ;
         jb = next_owned_block(jb + 1, nJ, row_splits, split)) {
      const int kt_end = min((jb + 1) * KT_PER_BLOCK, kt_total);
      // fragments of this warp: f = 2 mi + rw; those with f * 8 >= (N - first row) are padding
      const int f_max = (N - jb * TILE_ROWS + 7) / 8;          // fragments with real rows
      const int mi_max = max(0, min(8, (f_max - rw + 1) / 2));
      const int kt_diag0 = jb * KT_PER_BLOCK;                  // first k-tile of the diagonal block

      double acc[8][4][2];
#pragma unroll
      for (int mi = 0; mi < 8; mi++)
#pragma unroll
        for (int ni = 0; ni < 4; ni++) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;

      for (int kt = 0; kt < kt_end; kt++) {
        mbar_wait(&sm.full[stage], phase);
        if (mi_max > 0) {
          const double* As = sm.A[stage] + rw * 32 + lane;
          const double* Bs = sm.B[stage] + (cw * 32) * 4 + lane;
          const int qd = kt - kt_diag0;
          if (mi_max >= 8) {
            switch (qd < 0 ? 0 : qd) {
              case 0: vc_stage<0, 8>(acc, As, Bs); break;
              case 1: vc_stage<1, 8>(acc, As, Bs); break;
              case 2: vc_stage<2, 8>(acc, As, Bs); break;
              case 3: vc_stage<3, 8>(acc, As, Bs); break;
              case 4: vc_stage<4, 8>(acc, As, Bs); break;
              case 5: vc_stage<5, 8>(acc, As, Bs); break;
              case 6: vc_stage<6, 8>(acc, As, Bs); break;
              default: vc_stage<7, 8>(acc, As, Bs);
            }
          } else {   // last row block: only the fragments that hold real rows
            switch (mi_max) {
              case 1: vc_stage<0, 1>(acc, As, Bs); break;
              case 2: vc_stage<0, 2>(acc, As, Bs); break;
              case 3: vc_stage<0, 3>(acc, As, Bs); break;
              case 4: vc_stage<0, 4>(acc, As, Bs); break;
              case 5: vc_stage<0, 5>(acc, As, Bs); break;
              case 6: vc_stage<0, 6>(acc, As, Bs); break;
              default: vc_stage<0, 7>(acc, As, Bs);
            }
          }
        }
        __syncwarp();
        if (lane == 0) {
          // release (this warp's reads of the slot are done) / acquire (the last warp out sees
          // every other warp's release) on the slot counter; the last warp out refills
          if (atom_add_acq_rel_shared(&sm.released[stage], 1) == VC_WARPS - 1) {
            sm.released[stage] = 0;
            if (nseq.valid(n_tiles)) issue(nseq, stage);
          }
        }
        if (nseq.valid(n_tiles)) nseq.advance(nJ, row_splits, split, kt_total);
        if (++stage == VC_STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
      // squared-row-sum epilogue (gpr.py:1208 einsum("ji,ji->i", M, M)), kept per lane
#pragma unroll
      for (int mi = 0; mi < 8; mi++)
#pragma unroll
        for (int ni = 0; ni < 4; ni++) {
          ssq[ni][0] = fma(acc[mi][ni][0], acc[mi][ni][0], ssq[ni][0]);
          ssq[ni][1] = fma(acc[mi][ni][1], acc[mi][ni][1], ssq[ni][1]);
        }
    }

    // reduce over the 8 row groups of the warp (lanes with equal lane%4), then over rw
#pragma unroll
    for (int ni = 0; ni < 4; ni++)
#pragma unroll
      for (int e = 0; e < 2; e++) {
        double v = ssq[ni][e];
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        v += __shfl_xor_sync(0xffffffffu, v, 8);
        v += __shfl_xor_sync(0xffffffffu, v, 16);
        if (lane < 4) sm.red[rw][cw * 32 + ni * 8 + 2 * lane + e] = v;
      }
    __syncthreads();
    if (tid < TILE_ROWS) {
      const double tot = sm.red[0][tid] + sm.red[1][tid];
      const int i = t * TILE_ROWS + tid;
      if (fused) {   // one CTA saw every row block: finish mean / std / LogExp right here
        if (i < fin.n_valid) finish_one(fin, i, true, tot);
      } else {
        ssqp[(size_t)split * chunk_cands + i] = tot;
      }
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------
// finish: gpr.py:1180-1195, 1207-1227 and LogExp.f acquisition_functions.py:1071-1074
// ---------------------------------------------------------------------------------------
__global__ void finish_kernel(FinishParams f, const double* __restrict__ ssqp, int row_splits) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= f.n_valid) return;
  double ssq = 0.0;
  if (ssqp)
    for (int s = 0; s < row_splits; s++) ssq += ssqp[(size_t)s * f.chunk_cands + i];
  finish_one(f, i, ssqp != nullptr, ssq);
}

// ---------------------------------------------------------------------------------------
// finish + selection (gpry_predict_logexp_topk): the finishing arithmetic above, the two device
// masks (classifier: gpr.py:1145, 1172; trust region: :1201; either makes LogExp -inf,
// acquisition_functions.py:983-992), the skip list of already proposed rows
// (gp_acquisition.py:1037-1047), and a threshold filter against the K'-th best acquisition
// value seen so far: warp ballot + one atomic per warp reserve the slots, only the records that
// can still be ranked are written (see topk.cu).
// ---------------------------------------------------------------------------------------
struct SelectArgs {
  double* acq;
  int64_t* idx;
  double* mean;
  double* sd;
  unsigned long long* ctl;   // [0] count, [1] threshold key, [2] overflow
  int cap;
  int64_t gbase;             // global index of candidate 0 of this chunk
  int64_t lbase;             // row number in the pool of candidate 0 of this chunk
  const int64_t* excl;       // sorted row numbers to skip
  int n_excl;
  const double* X;           // candidate rows of this chunk (trust region) or NULL
  int d;
  const double* trust_lohi;
  double mask_value;
  const double* clf_dec;     // classifier decisions of this chunk's rows or NULL
};

__device__ __forceinline__ unsigned long long select_key(double x) {
  if (x != x) return 0ull;   // NaN ranks below everything (as topk.cu sortable_key)
  unsigned long long b = (unsigned long long)__double_as_longlong(x);
  return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}

__global__ void __launch_bounds__(256)
finish_select_kernel(FinishParams f, const double* __restrict__ ssqp, int row_splits,
                     SelectArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  bool pass = false;
  double mean = 0.0, sd = 0.0, acq = 0.0;
  if (i < f.n_valid) {
    double ssq = 0.0;
    for (int s = 0; s < row_splits; s++) ssq += ssqp[(size_t)s * f.chunk_cands + i];
    finish_values(f, i, true, ssq, mean, sd, acq);
    if (a.clf_dec && !(a.clf_dec[i] > 0.0)) {
      mean = a.mask_value;
      sd = 0.0;
      acq = -INFINITY;
    }
    if (a.X) {
      bool inside = true;
      for (int k = 0; k < a.d; k++) {
        const double x = a.X[(size_t)i * a.d + k];
        inside = inside && (x >= a.trust_lohi[k]) && (x <= a.trust_lohi[MAX_DIM + k]);
      }
      if (!inside) {
        mean = a.mask_value;
        acq = -INFINITY;
      }
    }
    pass = select_key(acq) >= a.ctl[1];
    if (pass && a.n_excl > 0) {      // binary search of the row number in the skip list
      const int64_t row = a.lbase + i;
      int lo = 0, hi = a.n_excl;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (a.excl[mid] < row) lo = mid + 1; else hi = mid;
      }
      if (lo < a.n_excl && a.excl[lo] == row) pass = false;
    }
  }
  const unsigned ballot = __ballot_sync(0xffffffffu, pass);
  if (ballot == 0u) return;
  const int lane = threadIdx.x & 31;
  unsigned long long base = 0;
  if (lane == 0) base = atomicAdd(a.ctl, (unsigned long long)__popc(ballot));
  base = __shfl_sync(0xffffffffu, base, 0);
  if (pass) {
    const unsigned long long slot = base + __popc(ballot & ((1u << lane) - 1u));
    if (slot < (unsigned long long)a.cap) {
      a.acq[slot] = acq;
      a.idx[slot] = a.gbase + i;
      a.mean[slot] = mean;
      a.sd[slot] = sd;
    } else {
      a.ctl[2] = 1ull;
    }
  }
}

// ---------------------------------------------------------------------------------------
// mean gradient at one point: kernels.py:257-278 (RBF), 363-393 (nu=1.5), 395-432 (nu=2.5),
// Product rule :687-699, gpr.py:1237-1242.   grid = d blocks, block k reduces dimension k.
// ---------------------------------------------------------------------------------------
template <int KIND>
__global__ void __launch_bounds__(256)
mean_grad_kernel(const double* __restrict__ Xt, const double* __restrict__ alpha, int N, int d,
                 const double* __restrict__ x_t /* transformed point */,
                 const double* __restrict__ ell, double c, double y_std,
                 double* __restrict__ out) {
  __shared__ double red[256];
  const int k = blockIdx.x;
  double s = 0.0;
  for (int j = threadIdx.x; j < N; j += 256) {
    double r2 = 0.0, dk = 0.0;
    for (int q = 0; q < d; q++) {
      double diff = (x_t[q] - Xt[(size_t)j * d + q]) / ell[q];
      r2 = fma(diff, diff, r2);
      if (q == k) dk = diff;
    }
    double g;
    if (KIND == GPRY_KERNEL_RBF) {
      g = (-exp(-0.5 * r2) * dk) / ell[k];
    } else if (KIND == GPRY_KERNEL_MATERN15) {
      double dist = sqrt(r2);
      double s3d = 1.7320508075688772 * dist;
      double by = dist != 0.0 ? 1.7320508075688772 / dist : 0.0;
      double f_grad = (dk / ell[k]) * by;
      g = exp(-s3d) * f_grad * (1.0 - (1.0 + s3d));
    } else {
      double dist = sqrt(r2);
      double s5d = 2.23606797749979 * dist;
      double f = (5.0 / 3.0) * r2 + s5d + 1.0;
      double inv = dist != 0.0 ? 2.23606797749979 * (1.0 / dist) : 0.0;
      double dl = dk / ell[k];
      double f1g = inv * dl, f2g = (10.0 / 3.0) * dl;
      double gg = exp(-s5d);
      g = f * (-gg * f1g) + gg * (f1g + f2g);
    }
    s = fma(c * g, alpha[j], s);
  }
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[k] = red[0] * y_std;
}

// ---------------------------------------------------------------------------------------
// model upload helpers
// ---------------------------------------------------------------------------------------
// T[j][k] = X_train_[j][k] / ell[k], zero padded to [Npad][DP]
__global__ void scale_train_kernel(const double* __restrict__ Xt, int N, int d, int Npad, int DP,
                                   const double* __restrict__ ell, double* __restrict__ T) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (int64_t)Npad * DP) return;
  int j = (int)(e / DP), k = (int)(e % DP);
  T[e] = (j < N && k < d) ? Xt[(size_t)j * d + k] / ell[k] : 0.0;
}

// pack row-major V (lower triangular N x N, leading dimension ld; or its transpose) into
// the tile layout.  One thread per packed element.
__global__ void pack_v_kernel(const double* __restrict__ V, int N, int ld, int transposed, int nJ,
                              double* __restrict__ Vt) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t total = vtile_count(nJ) * TILE_DOUBLES;
  if (e >= total) return;
  int64_t tile = e / TILE_DOUBLES;
  int within = (int)(e % TILE_DOUBLES);
  // invert vtile_index: find jb with 4*jb*(jb+1) <= tile
  int jb = (int)((sqrt(1.0 + (double)tile) - 1.0) * 0.5);
  while (vtile_index(jb + 1, 0) <= tile) jb++;
  while (vtile_index(jb, 0) > tile) jb--;
  int kt = (int)(tile - vtile_index(jb, 0));
  int p = within / (TILE_ROWS * 4), rem = within % (TILE_ROWS * 4);
  int r = rem >> 2, kk = rem & 3;
  int row = jb * TILE_ROWS + r, col = kt * TILE_K + p * 4 + kk;
  double v = 0.0;
  if (row < N && col <= row)
    v = transposed ? V[(size_t)col * ld + row] : V[(size_t)row * ld + col];
  Vt[e] = v;
}

// zero-padded row-major copy of the lower triangle of V (or of its transpose)
__global__ void pad_v_rowmajor_kernel(const double* __restrict__ V, int N, int ld, int transposed,
                                      int Np, double* __restrict__ out) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (int64_t)Np * Np) return;
  int row = (int)(e / Np), col = (int)(e % Np);
  double v = 0.0;
  if (row < N && col <= row)
    v = transposed ? V[(size_t)col * ld + row] : V[(size_t)row * ld + col];
  out[e] = v;
}

__global__ void pad_copy_kernel(const double* __restrict__ src, int n, int npad,
                                double* __restrict__ dst) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < npad) dst[i] = i < n ? src[i] : 0.0;
}

void upload_model(gpry_state* st, int kind, int N, int d, const double* X_train_t,
                  const double* alpha_, const double* V_host, const double* V_dev_rowmajor,
                  const double* VT_dev_rowmajor, int ldV, const double* alpha_dev, double c,
                  const double* ell, const double* x_min, const double* x_width, double y_mean,
                  double y_std, double clip_hi) {
  GPRY_CHECK_ARG(kind >= 0 && kind <= 2, "unknown kernel kind");
  GPRY_CHECK_ARG(N >= 1 && d >= 1 && d <= MAX_DIM, "need N >= 1 and 1 <= d <= 128");
  GPRY_CUDA(cudaSetDevice(st->device));
  st->loaded = false;
  st->kind = kind;
  st->N = N;
  st->d = d;
  st->DP = round_up(d, 4);
  st->Npad = round_up(N, TILE_ROWS);
  st->nJ = st->Npad / TILE_ROWS;
  st->nKT = st->Npad / TILE_K;
  st->c = c;
  st->y_mean = y_mean;
  st->y_std = y_std;
  st->clip_hi = clip_hi;
  std::vector<double> prm(3 * MAX_DIM, 1.0);
  for (int k = 0; k < d; k++) {
    prm[k] = x_min ? x_min[k] : 0.0;
    prm[MAX_DIM + k] = x_width ? x_width[k] : 1.0;
    prm[2 * MAX_DIM + k] = ell[k];
    if (k < MAX_DIM_REG) {
      st->prm.x_min[k] = prm[k];
      st->prm.x_width[k] = prm[MAX_DIM + k];
      st->prm.ell[k] = ell[k];
    }
  }
  for (int k = d; k < MAX_DIM_REG; k++) {
    st->prm.x_min[k] = 0.0;
    st->prm.x_width[k] = 1.0;
    st->prm.ell[k] = 1.0;
  }
  cudaStream_t s = 0;
  st->prm_dev.reserve(3 * MAX_DIM);
  GPRY_CUDA(cudaMemcpyAsync(st->prm_dev.p, prm.data(), 3 * MAX_DIM * 8, cudaMemcpyHostToDevice, s));
  st->Xt.reserve((size_t)N * d);
  GPRY_CUDA(cudaMemcpyAsync(st->Xt.p, X_train_t, (size_t)N * d * 8, cudaMemcpyHostToDevice, s));
  st->T.reserve((size_t)st->Npad * st->DP);
  {
    int64_t total = (int64_t)st->Npad * st->DP;
    scale_train_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(
        st->Xt.p, N, d, st->Npad, st->DP, st->prm_dev.p + 2 * MAX_DIM, st->T.p);
    GPRY_CUDA(cudaGetLastError());
  }
  st->alpha.reserve(st->Npad);
  if (alpha_dev) {
    pad_copy_kernel<<<(st->Npad + 255) / 256, 256, 0, s>>>(alpha_dev, N, st->Npad, st->alpha.p);
    GPRY_CUDA(cudaGetLastError());
  } else {
    GPRY_CUDA(cudaMemsetAsync(st->alpha.p, 0, (size_t)st->Npad * 8, s));
    GPRY_CUDA(cudaMemcpyAsync(st->alpha.p, alpha_, (size_t)N * 8, cudaMemcpyHostToDevice, s));
  }
  const double* Vsrc = V_dev_rowmajor;
  int transposed = 0;
  if (ldV <= 0) ldV = N;
  if (V_host) {
    st->tmp.reserve((size_t)N * N);
    GPRY_CUDA(cudaMemcpyAsync(st->tmp.p, V_host, (size_t)N * N * 8, cudaMemcpyHostToDevice, s));
    Vsrc = st->tmp.p;
    ldV = N;
  } else if (VT_dev_rowmajor) {
    Vsrc = VT_dev_rowmajor;
    transposed = 1;
  }
  st->has_V = Vsrc != nullptr;     // mean-only states (the classifier's decision function)
  if (st->has_V) {
    int64_t total = vtile_count(st->nJ) * TILE_DOUBLES;
    st->Vt.reserve((size_t)total);
    pack_v_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(Vsrc, N, ldV, transposed, st->nJ,
                                                                 st->Vt.p);
    GPRY_CUDA(cudaGetLastError());
    int64_t tot = (int64_t)st->Npad * st->Npad;
    st->Vrm.reserve((size_t)tot);
    pad_v_rowmajor_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(Vsrc, N, ldV, transposed,
                                                                       st->Npad, st->Vrm.p);
    GPRY_CUDA(cudaGetLastError());
  }
  GPRY_CUDA(cudaStreamSynchronize(s));
  st->vtrm_valid = false;
  st->oz_valid = false;
  st->loaded = true;
}

// ---------------------------------------------------------------------------------------
// launch helpers
// ---------------------------------------------------------------------------------------
// Shared-memory floor of kstar_build blocks (bytes): pads the request so that FEWER blocks fit on
// an SM.  A chunk is 2 n_sm tiles x JS slices = 16 n_sm blocks at JS = 8: with 4 resident blocks
// per SM that is exactly 4 waves, with the 5 the kernel's own footprint allows it is 3.2.
// GPRY_B200_BUILD_SMEM_KB overrides (experiments).
static size_t build_smem_floor() {
  static long v = -1;
  if (v < 0) {
    const char* e = getenv("GPRY_B200_BUILD_SMEM_KB");
    v = e ? atol(e) * 1024 : 0;
  }
  return (size_t)v;
}
template <int KIND, int WMODE>
static void launch_build(gpry_state* st, const double* dX, int64_t M, int64_t cand0, int tiles,
                         int JS, int chunk_cands, cudaStream_t s) {
  const int NJ = st->Npad / JS;
  const int d = st->d, DP = st->DP;
  dim3 grid(tiles, JS), block(128);
  void* kout = WMODE == 2 ? (void*)st->oz_Ksl_cur : (void*)st->Ks.p;
  const double slice_scale = 1073741824.0 / st->c;           // 2^30 / c (see kstar_build, WMODE 2)
  if (d <= MAX_DIM_REG) {
    size_t smem = ((size_t)NJ * DP + NJ + 128 * d + (d & 1)) * 8 + 16;
    smem = std::max(smem, build_smem_floor());
#define GPRY_LAUNCH_BUILD(DPV)                                                                \
  case DPV: {                                                                                 \
    auto kern = kstar_build_kernel<DPV, KIND, WMODE>;                                         \
    if (smem > 48 * 1024)                                                                     \
      GPRY_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,       \
                                     (int)smem));                                             \
    kern<<<grid, block, smem, s>>>(dX, M, d, cand0, st->T.p, st->alpha.p, NJ, st->nKT, st->c, \
                                   st->prm, kout, st->meanp_cur, chunk_cands, slice_scale);   \
  } break;
    switch (DP) {
      GPRY_LAUNCH_BUILD(4)
      GPRY_LAUNCH_BUILD(8)
      GPRY_LAUNCH_BUILD(12)
      GPRY_LAUNCH_BUILD(16)
      GPRY_LAUNCH_BUILD(20)
      GPRY_LAUNCH_BUILD(24)
      GPRY_LAUNCH_BUILD(28)
      GPRY_LAUNCH_BUILD(32)
      default:
        throw GpryError{GPRY_ERR_ARG, "internal: bad DP"};
    }
#undef GPRY_LAUNCH_BUILD
  } else {
    size_t smem = ((size_t)NJ * DP + NJ + (size_t)DP * 129) * 8;
    auto kern = kstar_build_generic_kernel<KIND, WMODE>;
    GPRY_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, block, smem, s>>>(dX, M, d, DP, cand0, st->T.p, st->alpha.p, NJ, st->nKT, st->c,
                                   st->prm_dev.p, kout, st->meanp_cur, chunk_cands, slice_scale);
  }
  GPRY_CUDA(cudaGetLastError());
}

template <int WMODE>
static void launch_build_kind(gpry_state* st, const double* dX, int64_t M, int64_t cand0,
                              int tiles, int JS, int chunk_cands, cudaStream_t s) {
  switch (st->kind) {
    case GPRY_KERNEL_RBF:
      launch_build<GPRY_KERNEL_RBF, WMODE>(st, dX, M, cand0, tiles, JS, chunk_cands, s);
      break;
    case GPRY_KERNEL_MATERN15:
      launch_build<GPRY_KERNEL_MATERN15, WMODE>(st, dX, M, cand0, tiles, JS, chunk_cands, s);
      break;
    default:
      launch_build<GPRY_KERNEL_MATERN25, WMODE>(st, dX, M, cand0, tiles, JS, chunk_cands, s);
  }
}

// Largest j-split count JS (power of two) such that the per-block training slice
// NJ = Npad / JS stays a multiple of 16, fits in shared memory and gives >= ~2 blocks/SM.
static int choose_jsplit(const gpry_state* st, int tiles) {
  if (const char* e = getenv("GPRY_B200_BUILD_JS")) {      // experiments
    const int js = atoi(e);
    if (js >= 1 && st->Npad % (js * 16) == 0) return js;
  }
  int JS = 1;
  if (st->d > MAX_DIM_REG) {   // generic path: [NJ][DP] slice + [DP][129] candidates in smem
    int maxJSg = st->Npad / 16;
    while (JS < maxJSg && ((size_t)(st->Npad / JS) * (st->DP + 1) * 8 > 64 * 1024)) JS *= 2;
    return JS;
  }
  const int maxJS = st->Npad / 128;   // NJ >= 128
  // shared-memory bound on NJ (keep a block under ~48 KB so several are resident per SM)
  while (JS < maxJS && ((size_t)(st->Npad / JS) * (st->DP + 1) * 8 > 40 * 1024)) JS *= 2;
  while (JS * 2 <= maxJS && (int64_t)tiles * JS < 4 * st->n_sm && (st->Npad / (JS * 2)) % 16 == 0)
    JS *= 2;
  while (JS > 1 && (st->Npad % JS != 0 || (st->Npad / JS) % 16 != 0)) JS /= 2;
  return JS;
}

// ---------------------------------------------------------------------------------------
// Latency path for a handful of candidates (M <= SMALL_M_MAX: single-point calls of samplers and
// optimisers, predict_std on a few pool rows): no 128-wide tiles, V is streamed once.
//   small_kstar   ks[m][j] = k(x_m, X_j)                      (thread per training point)
//   small_vk      u[m][j]  = sum_k V[j][k] ks[m][k]           (warp per row of V, M accumulators)
//   small_finish  mean_m = sum_j ks[m][j] alpha_j, ssq_m = sum_j u[m][j]^2, then finish_one
// ---------------------------------------------------------------------------------------
constexpr int SMALL_M = 8;        // candidates per pass over V
constexpr int SMALL_M_MAX = 64;   // largest batch served by this path

template <int KIND>
__global__ void small_kstar_kernel(const double* __restrict__ X, int M, int d,
                                   const double* __restrict__ T, int N, int Np, int DP,
                                   const double* __restrict__ prm, double c,
                                   double* __restrict__ ks) {
  __shared__ double us[SMALL_M * MAX_DIM];
  for (int e = threadIdx.x; e < M * DP; e += blockDim.x) {
    int m = e / DP, k = e % DP;
    double v = 0.0;
    if (k < d) {
      double x = X[(size_t)m * d + k];
      v = ((x - prm[k]) / prm[MAX_DIM + k]) / prm[2 * MAX_DIM + k];
    }
    us[m * MAX_DIM + k] = v;
  }
  __syncthreads();
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= Np) return;
  for (int m = 0; m < M; m++) {
    double v = 0.0;
    if (j < N) {
      double r2 = 0.0;
      for (int k = 0; k < DP; k++) {
        double df = us[m * MAX_DIM + k] - T[(size_t)j * DP + k];
        r2 = fma(df, df, r2);
      }
      v = kernel_value<KIND>(r2, c);
    }
    ks[(size_t)m * Np + j] = v;
  }
}

__global__ void __launch_bounds__(256)
small_vk_kernel(const double* __restrict__ V, int Np, int N, const double* __restrict__ ks, int M,
                double* __restrict__ u) {
  const int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (j >= Np) return;
  double acc[SMALL_M];
#pragma unroll
  for (int m = 0; m < SMALL_M; m++) acc[m] = 0.0;
  if (j < N) {
    for (int k = lane; k <= j; k += 32) {
      const double v = V[(size_t)j * Np + k];
#pragma unroll
      for (int m = 0; m < SMALL_M; m++)
        if (m < M) acc[m] = fma(v, ks[(size_t)m * Np + k], acc[m]);
    }
  }
#pragma unroll
  for (int m = 0; m < SMALL_M; m++) {
    double v = acc[m];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0 && m < M) u[(size_t)m * Np + j] = v;
  }
}

// one block per candidate
__global__ void __launch_bounds__(256)
small_finish_kernel(const double* __restrict__ ks, const double* __restrict__ u,
                    const double* __restrict__ alpha, int N, int Np, int have_var,
                    FinishParams f, double* __restrict__ meanp_out) {
  __shared__ double r0[256], r1[256];
  const int m = blockIdx.x;
  double a = 0.0, b = 0.0;
  for (int j = threadIdx.x; j < N; j += 256) {
    a = fma(ks[(size_t)m * Np + j], alpha[j], a);
    if (have_var) {
      double uu = u[(size_t)m * Np + j];
      b = fma(uu, uu, b);
    }
  }
  r0[threadIdx.x] = a;
  r1[threadIdx.x] = b;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      r0[threadIdx.x] += r0[threadIdx.x + o];
      r1[threadIdx.x] += r1[threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    meanp_out[m] = r0[0];          // f.meanp points here with JS = 1
    finish_one(f, m, have_var != 0, r1[0]);
  }
}

static void predict_small(gpry_state* st, const double* dX, int M, bool want_var, bool want_acq,
                          double zeta, double sigma_n, double y_max, double* d_mean,
                          double* d_std, double* d_acq, cudaStream_t s) {
  const int N = st->N, Np = st->Npad, DP = st->DP, d = st->d;
  st->pc_U.reserve((size_t)(2 * SMALL_M) * Np + SMALL_M);
  double* ks = st->pc_U.p;
  double* u = ks + (size_t)SMALL_M * Np;
  double* mp = u + (size_t)SMALL_M * Np;
  TimedScope ts(st, s, T_BUILD, 3);
  const int nblk = (Np + 127) / 128;
  switch (st->kind) {
    case GPRY_KERNEL_RBF:
      small_kstar_kernel<GPRY_KERNEL_RBF><<<nblk, 128, 0, s>>>(dX, M, d, st->T.p, N, Np, DP,
                                                              st->prm_dev.p, st->c, ks);
      break;
    case GPRY_KERNEL_MATERN15:
      small_kstar_kernel<GPRY_KERNEL_MATERN15><<<nblk, 128, 0, s>>>(dX, M, d, st->T.p, N, Np, DP,
                                                                   st->prm_dev.p, st->c, ks);
      break;
    default:
      small_kstar_kernel<GPRY_KERNEL_MATERN25><<<nblk, 128, 0, s>>>(dX, M, d, st->T.p, N, Np, DP,
                                                                   st->prm_dev.p, st->c, ks);
  }
  GPRY_CUDA(cudaGetLastError());
  if (want_var) {
    small_vk_kernel<<<(Np * 32 + 255) / 256, 256, 0, s>>>(st->Vrm.p, Np, N, ks, M, u);
    GPRY_CUDA(cudaGetLastError());
  }
  FinishParams fin;
  fin.meanp = mp; fin.JS = 1; fin.chunk_cands = SMALL_M; fin.n_valid = M;
  fin.c = st->c; fin.y_mean = st->y_mean; fin.y_std = st->y_std; fin.clip_hi = st->clip_hi;
  fin.want_acq = want_acq ? 1 : 0;
  fin.two_zeta = 2.0 * zeta; fin.sigma_n2 = sigma_n * sigma_n; fin.y_max = y_max;
  fin.o_mean = d_mean; fin.o_std = d_std; fin.o_acq = d_acq;
  small_finish_kernel<<<M, 256, 0, s>>>(ks, u, st->alpha.p, N, Np, want_var ? 1 : 0, fin, mp);
  GPRY_CUDA(cudaGetLastError());
}

// mean[i] = value for candidates outside [lo, hi] (tools.py:263-287 is_in_bounds, gpr.py:1201)
__global__ void trust_mask_kernel(const double* __restrict__ X, int64_t M, int d,
                                  const double* __restrict__ lohi, double value,
                                  double* __restrict__ mean, double* __restrict__ acq) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M) return;
  bool inside = true;
  for (int k = 0; k < d; k++) {
    double x = X[i * d + k];
    inside = inside && (x >= lohi[k]) && (x <= lohi[MAX_DIM + k]);
  }
  if (!inside) {
    if (mean) mean[i] = value;
    if (acq) acq[i] = -INFINITY;   // LogExp of a non-finite mean (acquisition_functions.py:983-992)
  }
}

// rows the classifier calls infinite (decision <= 0): mean = value, std = 0, acq = -inf
// (gpr.py:1145, 1172, 1229-1231; acquisition_functions.py:983-992)
__global__ void classifier_mask_kernel(const double* __restrict__ dec, int64_t M, double value,
                                       double* __restrict__ mean, double* __restrict__ sd,
                                       double* __restrict__ acq) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M || dec[i] > 0.0) return;
  if (mean) mean[i] = value;
  if (sd) sd[i] = 0.0;
  if (acq) acq[i] = -INFINITY;
}

// decision values of the classifier for the candidates dX (device) -> st->clf_dec, then mask
void apply_classifier(gpry_state* st, const double* dX, int64_t M, double* d_mean, double* d_std,
                      double* d_acq, cudaStream_t s) {
  if (!st->clf_on || M <= 0) return;
  st->clf_dec.reserve((size_t)M);
  predict_pipeline(st->clf, dX, M, true, false, false, 0, 0, 0, st->clf_dec.p, nullptr, nullptr, s);
  TimedScope ts(st, s, T_FINISH);
  classifier_mask_kernel<<<(unsigned)((M + 255) / 256), 256, 0, s>>>(st->clf_dec.p, M,
                                                                    st->trust_value, d_mean,
                                                                    d_std, d_acq);
  GPRY_CUDA(cudaGetLastError());
}

void apply_trust_region(gpry_state* st, const double* dX, int64_t M, double* d_mean,
                        double* d_acq, cudaStream_t s) {
  if (!st->trust_on || (!d_mean && !d_acq) || M <= 0) return;
  TimedScope ts(st, s, T_FINISH);
  trust_mask_kernel<<<(unsigned)((M + 255) / 256), 256, 0, s>>>(dX, M, st->d, st->trust.p,
                                                               st->trust_value, d_mean, d_acq);
  GPRY_CUDA(cudaGetLastError());
}

void predict_pipeline(gpry_state* st, const double* dX, int64_t M, bool want_mean, bool want_var,
                      bool want_acq, double zeta, double sigma_n, double y_max, double* d_mean,
                      double* d_std, double* d_acq, cudaStream_t s) {
  if (!st->loaded) throw GpryError{GPRY_ERR_STATE, "no model uploaded into this state"};
  if (M <= 0) return;
  if (want_var && !st->has_V)
    throw GpryError{GPRY_ERR_STATE, "this state was uploaded without V: mean only"};
  GPRY_CUDA(cudaSetDevice(st->device));
  if (M <= SMALL_M_MAX && !st->sel.on) {   // latency path, SMALL_M candidates per pass over V
    for (int m0 = 0; m0 < (int)M; m0 += SMALL_M)
      predict_small(st, dX + (size_t)m0 * st->d, std::min(SMALL_M, (int)M - m0), want_var, want_acq,
                    zeta, sigma_n, y_max, d_mean ? d_mean + m0 : nullptr,
                    d_std ? d_std + m0 : nullptr, d_acq ? d_acq + m0 : nullptr, s);
    return;
  }
  const int64_t total_tiles = (M + TILE_ROWS - 1) / TILE_ROWS;
  // chunking: at most 2 waves of tiles per chunk (bounds the K* scratch: 2 MB/tile at N=2048)
  const int max_chunk_tiles = 2 * st->n_sm;
  const int chunk_tiles = (int)std::min<int64_t>(total_tiles, max_chunk_tiles);
  const int chunk_cands = chunk_tiles * TILE_ROWS;
  // INT8 split of the contraction (ozaki.cu) when selected and the model qualifies
  bool ozaki = want_var && st->contract_mode != 0 && ozaki_supported(st);
  if (ozaki) {
    ozaki_prepare(st, s);
    ozaki_validate(st, s);      // error model + probe against the FP64 kernel, once per model
    ozaki = ozaki_supported(st);
  }
  // row splits: for small pools spread the row blocks of V over more CTAs
  int row_splits = 1;
  if (ozaki) {
    row_splits = st->oz_splits;
  } else if (want_var) {
    while (row_splits * 2 <= st->nJ && (int64_t)chunk_tiles * row_splits * 2 <= st->n_sm)
      row_splits *= 2;
  }
  const int JS = choose_jsplit(st, chunk_tiles);
  if (ozaki)
    st->oz_Ksl.reserve(ozaki_kslices_bytes(st, chunk_tiles));
  else if (want_var)
    st->Ks.reserve((size_t)chunk_tiles * st->nKT * TILE_DOUBLES);
  st->meanp.reserve((size_t)JS * chunk_cands);
  if (want_var) st->ssqp.reserve((size_t)row_splits * chunk_cands);
  if (want_var && !ozaki)
    GPRY_CUDA(cudaFuncSetAttribute(var_contract_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)sizeof(VcSmem)));
  // (Running the K* build of chunk c+1 on a second stream under the INT8 contraction of chunk c
  // was measured: the two kernels do co-reside, but the build's shared-memory reads slow the
  // MMA operand fetch by as much as the overlap hides -- +1 % in total; not kept.)
  for (int64_t t0 = 0; t0 < total_tiles; t0 += chunk_tiles) {
    const int tiles = (int)std::min<int64_t>(chunk_tiles, total_tiles - t0);
    const int64_t cand0 = t0 * TILE_ROWS;
    const int n = (int)std::min<int64_t>((int64_t)tiles * TILE_ROWS, M - cand0);
    st->meanp_cur = st->meanp.p;
    st->oz_Ksl_cur = st->oz_Ksl.p;
    {
      TimedScope ts(st, s, T_BUILD);
      if (!want_var)
        launch_build_kind<0>(st, dX, M, cand0, tiles, JS, chunk_cands, s);
      else if (ozaki)
        launch_build_kind<2>(st, dX, M, cand0, tiles, JS, chunk_cands, s);
      else
        launch_build_kind<1>(st, dX, M, cand0, tiles, JS, chunk_cands, s);
    }
    FinishParams fin;
    fin.meanp = st->meanp_cur;
    fin.JS = JS;
    fin.chunk_cands = chunk_cands;
    fin.n_valid = n;
    fin.c = st->c; fin.y_mean = st->y_mean; fin.y_std = st->y_std; fin.clip_hi = st->clip_hi;
    fin.want_acq = want_acq ? 1 : 0;
    fin.two_zeta = 2.0 * zeta; fin.sigma_n2 = sigma_n * sigma_n; fin.y_max = y_max;
    fin.o_mean = d_mean ? d_mean + cand0 : nullptr;
    fin.o_std = d_std ? d_std + cand0 : nullptr;
    fin.o_acq = d_acq ? d_acq + cand0 : nullptr;
    const bool selecting = st->sel.on;
    const bool fused = want_var && row_splits == 1 && !ozaki && !selecting;
    if (ozaki) {
      TimedScope ts(st, s, T_CONTRACT);
      st->n_contract_launches += 1;
      ozaki_contract(st, st->oz_Ksl_cur, tiles, chunk_cands, s);
    } else if (want_var) {
      TimedScope ts(st, s, T_CONTRACT);
      st->n_contract_launches += 1;
      dim3 grid(std::min(tiles, st->n_sm), row_splits);
      var_contract_kernel<<<grid, VC_THREADS, sizeof(VcSmem), s>>>(
          st->Vt.p, st->Ks.p, tiles, st->N, st->nJ, st->nKT, row_splits, st->ssqp.p, chunk_cands,
          fused ? 1 : 0, fin);
      GPRY_CUDA(cudaGetLastError());
    }
    if (selecting) {
      SelectRun& r = st->sel;
      SelectArgs a;
      a.acq = st->sel_acq[r.cur].p; a.idx = st->sel_idx[r.cur].p;
      a.mean = st->sel_mean[r.cur].p; a.sd = st->sel_std[r.cur].p;
      a.ctl = st->sel_ctl.p; a.cap = r.cap;
      a.gbase = r.gbase + cand0; a.lbase = r.lbase + cand0;
      a.excl = st->excl.p; a.n_excl = st->n_excl;
      a.X = st->trust_on ? dX + (size_t)cand0 * st->d : nullptr;
      a.d = st->d; a.trust_lohi = st->trust.p; a.mask_value = st->trust_value;
      a.clf_dec = r.clf_dec ? r.clf_dec + cand0 : nullptr;
      {
        TimedScope ts(st, s, T_FINISH);
        finish_select_kernel<<<(n + 255) / 256, 256, 0, s>>>(fin, st->ssqp.p, row_splits, a);
        GPRY_CUDA(cudaGetLastError());
      }
      if (++r.pending >= r.max_pending || r.first) select_compact(st, s);
    } else if (!fused) {
      TimedScope ts(st, s, T_FINISH);
      finish_kernel<<<(n + 255) / 256, 256, 0, s>>>(fin, want_var ? st->ssqp.p : nullptr,
                                                    row_splits);
      GPRY_CUDA(cudaGetLastError());
    }
  }
}

// Guard of the INT8 split, once per uploaded model (and digit layout): (1) the a-priori
// statistical estimate of ozaki.cu must be below the tolerance, (2) on 512 probe candidates --
// 256 training points (smallest variances, largest cancellation in c - sum w^2, taken evenly
// from all row blocks of V) and 256 uniform draws from the training set's bounding box -- the
// INT8 and the FP64 (DMMA) contraction must agree to tolerance / 16.  Otherwise this model
// takes the FP64 kernel (st->oz_ok = false).
void ozaki_validate(gpry_state* st, cudaStream_t s) {
  if (st->oz_checked) return;
  st->oz_checked = true;         // also stops the recursion: the probe runs predict_pipeline
  st->oz_ok = true;
  st->oz_probe_err = -1.0;
  if (!st->oz_guard) return;
  if (!(st->oz_est_sigma <= OZ_TOLERANCE)) {
    st->oz_ok = false;
    return;
  }
  const int P = 512, d = st->d, N = st->N;
  std::vector<double> Xt((size_t)N * d), prm(3 * MAX_DIM), probe((size_t)P * d);
  GPRY_CUDA(cudaMemcpyAsync(Xt.data(), st->Xt.p, (size_t)N * d * 8, cudaMemcpyDeviceToHost, s));
  GPRY_CUDA(cudaMemcpyAsync(prm.data(), st->prm_dev.p, 3 * MAX_DIM * 8, cudaMemcpyDeviceToHost, s));
  GPRY_CUDA(cudaStreamSynchronize(s));
  std::vector<double> lo(d, 1e300), hi(d, -1e300);
  for (int j = 0; j < N; j++)
    for (int k = 0; k < d; k++) {
      lo[k] = std::min(lo[k], Xt[(size_t)j * d + k]);
      hi[k] = std::max(hi[k], Xt[(size_t)j * d + k]);
    }
  uint64_t lcg = 0x9E3779B97F4A7C15ull;
  for (int i = 0; i < P; i++)
    for (int k = 0; k < d; k++) {
      double xt;
      if (i < P / 2) {
        xt = Xt[(size_t)((int64_t)i * N / (P / 2)) * d + k];
      } else {
        lcg = lcg * 6364136223846793005ull + 1442695040888963407ull;
        xt = lo[k] + (hi[k] - lo[k]) * ((double)(lcg >> 11) * (1.0 / 9007199254740992.0));
      }
      probe[(size_t)i * d + k] = xt * prm[MAX_DIM + k] + prm[k];     // un-transform
    }
  st->oz_probe.reserve((size_t)P * d + 2 * (size_t)P);
  double* dP = st->oz_probe.p;
  double* sd8 = dP + (size_t)P * d;
  double* sd64 = sd8 + P;
  GPRY_CUDA(cudaMemcpyAsync(dP, probe.data(), (size_t)P * d * 8, cudaMemcpyHostToDevice, s));
  const bool sel_on = st->sel.on;
  const bool prof = st->profiling;
  const int mode = st->contract_mode;
  const double nl = st->n_launches, ncl = st->n_contract_launches;
  st->sel.on = false;
  st->profiling = false;
  try {
    predict_pipeline(st, dP, P, false, true, false, 0, 0, 0, nullptr, sd8, nullptr, s);
    st->contract_mode = GPRY_CONTRACT_FP64;
    predict_pipeline(st, dP, P, false, true, false, 0, 0, 0, nullptr, sd64, nullptr, s);
  } catch (...) {
    st->contract_mode = mode; st->sel.on = sel_on; st->profiling = prof;
    throw;
  }
  st->contract_mode = mode;
  st->sel.on = sel_on;
  st->profiling = prof;
  st->n_launches = nl;
  st->n_contract_launches = ncl;
  std::vector<double> h(2 * (size_t)P);
  GPRY_CUDA(cudaMemcpyAsync(h.data(), sd8, 2 * (size_t)P * 8, cudaMemcpyDeviceToHost, s));
  GPRY_CUDA(cudaStreamSynchronize(s));
  double err = 0.0;
  for (int i = 0; i < P; i++) {
    const double a = h[i] / st->y_std, b = h[P + i] / st->y_std;
    if (!(a == a) || !(b == b)) continue;
    err = std::max(err, fabs(a * a - b * b) / std::max(b * b, 1.0));
  }
  st->oz_probe_err = err;
  st->oz_ok = err * 16.0 <= OZ_TOLERANCE;
}

void mean_grad_device(gpry_state* st, const double* x_host, double* out_host) {
  if (!st->loaded) throw GpryError{GPRY_ERR_STATE, "no model uploaded into this state"};
  GPRY_CUDA(cudaSetDevice(st->device));
  const int d = st->d;
  std::vector<double> xt(d);
  // Normalize_bounds.transform, preprocessing.py:380 (same expression as the reference)
  for (int k = 0; k < d; k++) {
    double mn, w;
    if (d <= MAX_DIM_REG) {
      mn = st->prm.x_min[k];
      w = st->prm.x_width[k];
    } else {
      std::vector<double> prm(3 * MAX_DIM);
      GPRY_CUDA(cudaMemcpy(prm.data(), st->prm_dev.p, 3 * MAX_DIM * 8, cudaMemcpyDeviceToHost));
      mn = prm[k];
      w = prm[MAX_DIM + k];
    }
    xt[k] = (x_host[k] - mn) / w;
  }
  st->small.reserve(2 * MAX_DIM);
  cudaStream_t s = 0;
  GPRY_CUDA(cudaMemcpyAsync(st->small.p, xt.data(), d * 8, cudaMemcpyHostToDevice, s));
  const double* ell = st->prm_dev.p + 2 * MAX_DIM;
  double* out = st->small.p + MAX_DIM;
  switch (st->kind) {
    case GPRY_KERNEL_RBF:
      mean_grad_kernel<GPRY_KERNEL_RBF><<<d, 256, 0, s>>>(st->Xt.p, st->alpha.p, st->N, d,
                                                         st->small.p, ell, st->c, st->y_std, out);
      break;
    case GPRY_KERNEL_MATERN15:
      mean_grad_kernel<GPRY_KERNEL_MATERN15><<<d, 256, 0, s>>>(
          st->Xt.p, st->alpha.p, st->N, d, st->small.p, ell, st->c, st->y_std, out);
      break;
    default:
      mean_grad_kernel<GPRY_KERNEL_MATERN25><<<d, 256, 0, s>>>(
          st->Xt.p, st->alpha.p, st->N, d, st->small.p, ell, st->c, st->y_std, out);
  }
  GPRY_CUDA(cudaGetLastError());
  GPRY_CUDA(cudaMemcpyAsync(out_host, out, d * 8, cudaMemcpyDeviceToHost, s));
  GPRY_CUDA(cudaStreamSynchronize(s));
}

// ---------------------------------------------------------------------------------------
// Posterior covariance among a small set of candidates (Kriging-believer conditioning of the
// ranked pool, gp_acquisition.py:1464-1494, 1522-1555, 1598-1670): instead of re-factorising
// the model augmented with the pool points, the conditioned variance follows from
//   Sigma = k(Xa, Xa) - (V K*a^T)^T (V K*a^T)        (normalised units, no noise term)
// as  var(a | P) = Sigma_aa - Sigma_aP (Sigma_PP + noise I)^-1 Sigma_Pa  (host, |P| <= size).
// ---------------------------------------------------------------------------------------
// scaled coordinates U[a][k] = ((x - min) / width) / ell, zero padded to [rows_pad][DP]
__global__ void scale_candidates_kernel(const double* __restrict__ X, int Ka, int d, int rows_pad,
                                        int DP, const double* __restrict__ prm,
                                        double* __restrict__ U) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= rows_pad * DP) return;
  int a = e / DP, k = e % DP;
  double v = 0.0;
  if (a < Ka && k < d) {
    double x = X[(size_t)a * d + k];
    v = ((x - prm[k]) / prm[MAX_DIM + k]) / prm[2 * MAX_DIM + k];
  }
  U[e] = v;
}

// out[a][j] = c g(|A_a - B_j|) for a < nA, j < nB, else 0; out is [rows_pad][ld] row major.
// If sub != nullptr: out = value - sub[a][j] (same layout).  grid (ld/128, rows_pad/8).
template <int KIND>
__global__ void __launch_bounds__(128)
kcross_kernel(const double* __restrict__ A, int nA, const double* __restrict__ B, int nB, int DP,
              double c, int ld, const double* __restrict__ sub, double* __restrict__ out) {
  extern __shared__ double sh[];
  double* Bs = sh;                     // [128][DP+1]
  double* As = sh + 128 * (DP + 1);    // [8][DP]
  const int tid = threadIdx.x;
  const int j0 = blockIdx.x * 128, a0 = blockIdx.y * 8;
  for (int e = tid; e < 128 * DP; e += 128) {
    int r = e / DP, k = e % DP;
    Bs[r * (DP + 1) + k] = B[(size_t)(j0 + r) * DP + k];
  }
  for (int e = tid; e < 8 * DP; e += 128) As[e] = A[(size_t)a0 * DP + e];
  __syncthreads();
  const int j = j0 + tid;
  for (int q = 0; q < 8; q++) {
    const int a = a0 + q;
    double v = 0.0;
    if (a < nA && j < nB) {
      double r2 = 0.0;
      for (int k = 0; k < DP; k++) {
        double df = As[q * DP + k] - Bs[tid * (DP + 1) + k];
        r2 = fma(df, df, r2);
      }
      v = kernel_value<KIND>(r2, c);
      if (sub) v -= sub[(size_t)a * ld + j];
    }
    out[(size_t)a * ld + j] = v;
  }
}

static void launch_kcross(int kind, const double* A, int nA, int rowsA_pad, const double* B, int nB,
                          int DP, double c, int ld, const double* sub, double* out,
                          cudaStream_t s) {
  dim3 grid(ld / 128, rowsA_pad / 8);
  size_t smem = ((size_t)128 * (DP + 1) + 8 * DP) * 8;
  switch (kind) {
    case GPRY_KERNEL_RBF:
      if (smem > 48 * 1024)
        GPRY_CUDA(cudaFuncSetAttribute(kcross_kernel<GPRY_KERNEL_RBF>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      kcross_kernel<GPRY_KERNEL_RBF><<<grid, 128, smem, s>>>(A, nA, B, nB, DP, c, ld, sub, out);
      break;
    case GPRY_KERNEL_MATERN15:
      if (smem > 48 * 1024)
        GPRY_CUDA(cudaFuncSetAttribute(kcross_kernel<GPRY_KERNEL_MATERN15>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      kcross_kernel<GPRY_KERNEL_MATERN15><<<grid, 128, smem, s>>>(A, nA, B, nB, DP, c, ld, sub,
                                                                 out);
      break;
    default:
      if (smem > 48 * 1024)
        GPRY_CUDA(cudaFuncSetAttribute(kcross_kernel<GPRY_KERNEL_MATERN25>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      kcross_kernel<GPRY_KERNEL_MATERN25><<<grid, 128, smem, s>>>(A, nA, B, nB, DP, c, ld, sub,
                                                                 out);
  }
  GPRY_CUDA(cudaGetLastError());
}

// d_out: [Ka][Ka] row major (device)
void posterior_cov_device(gpry_state* st, const double* dX, int Ka, double* d_out,
                          cudaStream_t s) {
  if (!st->loaded) throw GpryError{GPRY_ERR_STATE, "no model uploaded into this state"};
  if (!st->has_V) throw GpryError{GPRY_ERR_STATE, "this state was uploaded without V"};
  GPRY_CHECK_ARG(Ka >= 1 && Ka <= 8192, "posterior covariance: 1 <= Ka <= 8192");
  GPRY_CUDA(cudaSetDevice(st->device));
  const int Kap = round_up(Ka, TILE_ROWS), Np = st->Npad, DP = st->DP, d = st->d;
  st->pc_U.reserve((size_t)Kap * DP);
  st->pc_Ks.reserve((size_t)Kap * std::max(Np, Kap));
  st->pc_UT.reserve((size_t)Kap * Np);
  st->pc_G.reserve((size_t)Kap * Kap);
  TimedScope ts(st, s, T_FINISH, 5);
  scale_candidates_kernel<<<(Kap * DP + 255) / 256, 256, 0, s>>>(dX, Ka, d, Kap, DP, st->prm_dev.p,
                                                                 st->pc_U.p);
  GPRY_CUDA(cudaGetLastError());
  // K*a (row major, zero padded)
  launch_kcross(st->kind, st->pc_U.p, Ka, Kap, st->T.p, st->N, DP, st->c, Np, nullptr, st->pc_Ks.p,
                s);
  // UT = K*a V^T   (UT[a][j] = sum_k K*[a][k] V[j][k])
  gemm_nt(st->pc_Ks.p, Np, st->Vrm.p, Np, st->pc_UT.p, Np, Kap, Np, Np, 1.0, 0, 0, 0, s);
  // G = UT UT^T
  gemm_nt(st->pc_UT.p, Np, st->pc_UT.p, Np, st->pc_G.p, Kap, Kap, Kap, Np, 1.0, 0, 0, 0, s);
  // Sigma = k(Xa, Xa) - G  -> pc_Ks reused as [Kap][Kap]
  launch_kcross(st->kind, st->pc_U.p, Ka, Kap, st->pc_U.p, Ka, DP, st->c, Kap, st->pc_G.p,
                st->pc_Ks.p, s);
  GPRY_CUDA(cudaMemcpy2DAsync(d_out, (size_t)Ka * 8, st->pc_Ks.p, (size_t)Kap * 8, (size_t)Ka * 8,
                              Ka, cudaMemcpyDeviceToDevice, s));
}

// k_theta(X, Y) for arbitrary host inputs (already transformed): Product.__call__
// sklearn:kernels.py:971.  Host in, host out; API completeness, not a hot path.
void kernel_cross_device(gpry_state* st, int kind, int d, const double* theta, const double* hX,
                         int M, const double* hY, int N, double* h_out) {
  GPRY_CHECK_ARG(kind >= 0 && kind <= 2 && d >= 1 && d <= MAX_DIM && M >= 1 && N >= 1,
                 "bad kernel_cross arguments");
  GPRY_CUDA(cudaSetDevice(st->device));
  cudaStream_t s = 0;
  const int DP = round_up(d, 4), Mp = round_up(M, 8), Npd = round_up(N, 128);
  std::vector<double> A((size_t)Mp * DP, 0.0), B((size_t)Npd * DP, 0.0);
  const double c = exp(theta[0]);
  for (int i = 0; i < M; i++)
    for (int k = 0; k < d; k++) A[(size_t)i * DP + k] = hX[(size_t)i * d + k] / exp(theta[1 + k]);
  for (int j = 0; j < N; j++)
    for (int k = 0; k < d; k++) B[(size_t)j * DP + k] = hY[(size_t)j * d + k] / exp(theta[1 + k]);
  st->pc_U.reserve(A.size());
  st->pc_UT.reserve(B.size());
  st->pc_Ks.reserve((size_t)Mp * Npd);
  GPRY_CUDA(cudaMemcpyAsync(st->pc_U.p, A.data(), A.size() * 8, cudaMemcpyHostToDevice, s));
  GPRY_CUDA(cudaMemcpyAsync(st->pc_UT.p, B.data(), B.size() * 8, cudaMemcpyHostToDevice, s));
  launch_kcross(kind, st->pc_U.p, M, Mp, st->pc_UT.p, N, DP, c, Npd, nullptr, st->pc_Ks.p, s);
  GPRY_CUDA(cudaMemcpy2DAsync(h_out, (size_t)N * 8, st->pc_Ks.p, (size_t)Npd * 8, (size_t)N * 8, M,
                              cudaMemcpyDeviceToHost, s));
  GPRY_CUDA(cudaStreamSynchronize(s));
}

// d k(x, X_train) / d x_ for one point: (N x d) row major (Kernel.gradient_x, kernels.py:
// 257-278, 363-432, 687-699).  Thread per (j, k).
template <int KIND>
__global__ void gradx_kernel(const double* __restrict__ Xt, int N, int d,
                             const double* __restrict__ x_t, const double* __restrict__ ell,
                             double c, double* __restrict__ out) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= N * d) return;
  const int j = e / d, k = e % d;
  double r2 = 0.0, dk = 0.0;
  for (int q = 0; q < d; q++) {
    double diff = (x_t[q] - Xt[(size_t)j * d + q]) / ell[q];
    r2 = fma(diff, diff, r2);
    if (q == k) dk = diff;
  }
  double g;
  if (KIND == GPRY_KERNEL_RBF) {
    g = (-exp(-0.5 * r2) * dk) / ell[k];
  } else if (KIND == GPRY_KERNEL_MATERN15) {
    double dist = sqrt(r2);
    double s3d = 1.7320508075688772 * dist;
    double by = dist != 0.0 ? 1.7320508075688772 / dist : 0.0;
    g = exp(-s3d) * ((dk / ell[k]) * by) * (1.0 - (1.0 + s3d));
  } else {
    double dist = sqrt(r2);
    double s5d = 2.23606797749979 * dist;
    double f = (5.0 / 3.0) * r2 + s5d + 1.0;
    double inv = dist != 0.0 ? 2.23606797749979 * (1.0 / dist) : 0.0;
    double dl = dk / ell[k];
    double f1g = inv * dl, f2g = (10.0 / 3.0) * dl;
    double gg = exp(-s5d);
    g = f * (-gg * f1g) + gg * (f1g + f2g);
  }
  out[e] = c * g;
}

void kernel_gradx_device(gpry_state* st, const double* x_host, double* out_host) {
  if (!st->loaded) throw GpryError{GPRY_ERR_STATE, "no model uploaded into this state"};
  GPRY_CUDA(cudaSetDevice(st->device));
  const int d = st->d, N = st->N;
  cudaStream_t s = 0;
  st->small.reserve(2 * MAX_DIM);
  st->tmp.reserve((size_t)N * d);
  GPRY_CUDA(cudaMemcpyAsync(st->small.p, x_host, d * 8, cudaMemcpyHostToDevice, s));
  const double* ell = st->prm_dev.p + 2 * MAX_DIM;
  const int nblk = (N * d + 255) / 256;
  switch (st->kind) {
    case GPRY_KERNEL_RBF:
      gradx_kernel<GPRY_KERNEL_RBF><<<nblk, 256, 0, s>>>(st->Xt.p, N, d, st->small.p, ell, st->c,
                                                        st->tmp.p);
      break;
    case GPRY_KERNEL_MATERN15:
      gradx_kernel<GPRY_KERNEL_MATERN15><<<nblk, 256, 0, s>>>(st->Xt.p, N, d, st->small.p, ell,
                                                             st->c, st->tmp.p);
      break;
    default:
      gradx_kernel<GPRY_KERNEL_MATERN25><<<nblk, 256, 0, s>>>(st->Xt.p, N, d, st->small.p, ell,
                                                             st->c, st->tmp.p);
  }
  GPRY_CUDA(cudaGetLastError());
  GPRY_CUDA(cudaMemcpyAsync(out_host, st->tmp.p, (size_t)N * d * 8, cudaMemcpyDeviceToHost, s));
  GPRY_CUDA(cudaStreamSynchronize(s));
}

// ---------------------------------------------------------------------------------------
// std gradient at one point (gpr.py:1247-1261):
//   grad_std = -(k*^T V^T V dk*/dx) / sqrt(var_)  then  * y_std * y_std  (the reference applies
//   inverse_transform_scale twice).  u = V k*, z = V^T u, grad_k = sum_j z_j dk*_j/dx_k.
// ---------------------------------------------------------------------------------------
template <int KIND>
__global__ void kstar_single_kernel(const double* __restrict__ T, int N, int Np, int DP,
                                    const double* __restrict__ u_scaled, double c,
                                    double* __restrict__ ks) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= Np) return;
  double v = 0.0;
  if (j < N) {
    double r2 = 0.0;
    for (int k = 0; k < DP; k++) {
      double df = u_scaled[k] - T[(size_t)j * DP + k];
      r2 = fma(df, df, r2);
    }
    v = kernel_value<KIND>(r2, c);
  }
  ks[j] = v;
}
// u = V k*  (V row major lower, warp per row)
__global__ void gemv_lower_kernel(const double* __restrict__ V, int Np, int N,
                                  const double* __restrict__ x, double* __restrict__ y) {
  int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (j >= Np) return;
  double s = 0.0;
  if (j < N)
    for (int k = lane; k <= j; k += 32) s = fma(V[(size_t)j * Np + k], x[k], s);
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) y[j] = s;
}
// z = V^T u  (CTA per 32 columns, 8 row groups, fixed-order sums)
__global__ void __launch_bounds__(256)
gemv_lower_t_kernel(const double* __restrict__ V, int Np, int N, const double* __restrict__ u,
                    double* __restrict__ z) {
  __shared__ double red[8][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int k = blockIdx.x * 32 + cx;
  double s = 0.0;
  if (k < N)
    for (int j = blockIdx.x * 32 + ry; j < N; j += 8)
      if (j >= k) s = fma(V[(size_t)j * Np + k], u[j], s);
  red[ry][cx] = s;
  __syncthreads();
  if (ry == 0) {
    double v = 0.0;
    for (int q = 0; q < 8; q++) v += red[q][cx];
    if (k < Np) z[k] = v;
  }
}
// scale[0] = -y_std^2 / sqrt(max(c - |u|^2, 0))   (0 if the variance vanishes)
__global__ void std_grad_scale_kernel(const double* __restrict__ u, int N, double c, double y_std,
                                      double* __restrict__ out) {
  __shared__ double r[256];
  double s = 0.0;
  for (int i = threadIdx.x; i < N; i += 256) s = fma(u[i], u[i], s);
  r[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) r[threadIdx.x] += r[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    double var = c - r[0];
    out[0] = var > 0.0 ? -(y_std * y_std) / sqrt(var) : 0.0;
    out[1] = var > 0.0 ? sqrt(var) * y_std : 0.0;
  }
}
__global__ void scale_vec_kernel(double* __restrict__ v, int n, const double* __restrict__ scale) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] *= scale[0];
}

void std_grad_device(gpry_state* st, const double* x_host, double* out_grad, double* out_std) {
  if (!st->loaded) throw GpryError{GPRY_ERR_STATE, "no model uploaded into this state"};
  if (!st->has_V) throw GpryError{GPRY_ERR_STATE, "this state was uploaded without V"};
  GPRY_CUDA(cudaSetDevice(st->device));
  const int d = st->d, N = st->N, Np = st->Npad, DP = st->DP;
  cudaStream_t s = 0;
  // host: transformed point x_ and scaled point x_ / ell (same expressions as the device path)
  std::vector<double> prm(3 * MAX_DIM);
  GPRY_CUDA(cudaMemcpy(prm.data(), st->prm_dev.p, 3 * MAX_DIM * 8, cudaMemcpyDeviceToHost));
  std::vector<double> buf(MAX_DIM + DP, 0.0);
  for (int k = 0; k < d; k++) {
    double xt = (x_host[k] - prm[k]) / prm[MAX_DIM + k];
    buf[k] = xt;                                   // transformed (gradient kernel)
    buf[MAX_DIM + k] = xt / prm[2 * MAX_DIM + k];  // scaled (kernel values)
  }
  st->small.reserve(4 * MAX_DIM + 8);
  st->pc_U.reserve(3 * (size_t)Np);
  double* d_xt = st->small.p;                 // [MAX_DIM]
  double* d_us = st->small.p + MAX_DIM;       // [DP]
  double* d_out = st->small.p + 2 * MAX_DIM;  // [MAX_DIM]
  double* d_scale = st->small.p + 3 * MAX_DIM;
  double *ks = st->pc_U.p, *u = ks + Np, *z = u + Np;
  GPRY_CUDA(cudaMemcpyAsync(d_xt, buf.data(), (MAX_DIM + DP) * 8, cudaMemcpyHostToDevice, s));
  const int nblk = (Np + 255) / 256;
  switch (st->kind) {
    case GPRY_KERNEL_RBF:
      kstar_single_kernel<GPRY_KERNEL_RBF><<<nblk, 256, 0, s>>>(st->T.p, N, Np, DP, d_us, st->c, ks);
      break;
    case GPRY_KERNEL_MATERN15:
      kstar_single_kernel<GPRY_KERNEL_MATERN15><<<nblk, 256, 0, s>>>(st->T.p, N, Np, DP, d_us, st->c, ks);
      break;
    default:
      kstar_single_kernel<GPRY_KERNEL_MATERN25><<<nblk, 256, 0, s>>>(st->T.p, N, Np, DP, d_us, st->c, ks);
  }
  GPRY_CUDA(cudaGetLastError());
  gemv_lower_kernel<<<(Np * 32 + 255) / 256, 256, 0, s>>>(st->Vrm.p, Np, N, ks, u);
  GPRY_CUDA(cudaGetLastError());
  gemv_lower_t_kernel<<<Np / 32, 256, 0, s>>>(st->Vrm.p, Np, N, u, z);
  GPRY_CUDA(cudaGetLastError());
  std_grad_scale_kernel<<<1, 256, 0, s>>>(u, N, st->c, st->y_std, d_scale);
  GPRY_CUDA(cudaGetLastError());
  const double* ell = st->prm_dev.p + 2 * MAX_DIM;
  switch (st->kind) {   // sum_j z_j dk*_j/dx_k   (y_std factor 1: the scale carries y_std^2)
    case GPRY_KERNEL_RBF:
      mean_grad_kernel<GPRY_KERNEL_RBF><<<d, 256, 0, s>>>(st->Xt.p, z, N, d, d_xt, ell, st->c, 1.0, d_out);
      break;
    case GPRY_KERNEL_MATERN15:
      mean_grad_kernel<GPRY_KERNEL_MATERN15><<<d, 256, 0, s>>>(st->Xt.p, z, N, d, d_xt, ell, st->c, 1.0, d_out);
      break;
    default:
      mean_grad_kernel<GPRY_KERNEL_MATERN25><<<d, 256, 0, s>>>(st->Xt.p, z, N, d, d_xt, ell, st->c, 1.0, d_out);
  }
  GPRY_CUDA(cudaGetLastError());
  scale_vec_kernel<<<1, MAX_DIM, 0, s>>>(d_out, d, d_scale);
  GPRY_CUDA(cudaGetLastError());
  GPRY_CUDA(cudaMemcpyAsync(out_grad, d_out, d * 8, cudaMemcpyDeviceToHost, s));
  if (out_std) GPRY_CUDA(cudaMemcpyAsync(out_std, d_scale + 1, 8, cudaMemcpyDeviceToHost, s));
  GPRY_CUDA(cudaStreamSynchronize(s));
}

// ---------------------------------------------------------------------------------------
// Batched gradients (the acquisition optimiser's many starts in lock-step).
// Per candidate m the same quantities as gpry_mean_grad / gpry_std_grad (gpr.py:1236-1261):
//   grad_mean[m] = y_std * sum_j alpha_j dk*_mj/dx_
//   grad_std[m]  = -(y_std^2 / sqrt(var_m)) * sum_j z_mj dk*_mj/dx_,  z_m = V^T (V k*_m)
// K* (M x N) comes from kcross_kernel, W = K* V^T and Z = W V from the DMMA GEMM of train.cu
// (Z needs V^T row major: transposed lazily once per upload), the contraction with dk*/dx_ is
// one CTA per (dimension, candidate); dk*/dx_ (M x N x d) is never materialised.
// ---------------------------------------------------------------------------------------
__global__ void transpose_sq_kernel(const double* __restrict__ in, int n, double* __restrict__ out) {
  __shared__ double t[32][33];
  const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += 8) t[r][threadIdx.x] = in[(size_t)(by + r) * n + bx + threadIdx.x];
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += 8) out[(size_t)(bx + r) * n + by + threadIdx.x] = t[threadIdx.x][r];
}

// scale[2m] = var > 0 ? -(y_std^2)/sqrt(var) : 0,  scale[2m+1] = var (normalised), var = c - |W_m|^2
__global__ void __launch_bounds__(256)
grad_scale_batch_kernel(const double* __restrict__ W, int ldw, int N, double c, double y_std,
                        double* __restrict__ scale) {
  __shared__ double r[256];
  const double* w = W + (size_t)blockIdx.x * ldw;
  double s = 0.0;
  for (int i = threadIdx.x; i < N; i += 256) s = fma(w[i], w[i], s);
  r[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) r[threadIdx.x] += r[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    double var = c - r[0];
    scale[2 * blockIdx.x] = var > 0.0 ? -(y_std * y_std) / sqrt(var) : 0.0;
    scale[2 * blockIdx.x + 1] = var;
  }
}

// grid (d, M): block (k, m) reduces dimension k of candidate m over the training points
template <int KIND>
__global__ void __launch_bounds__(256)
grad_batch_kernel(const double* __restrict__ Xt, const double* __restrict__ alpha,
                  const double* __restrict__ Z, int ldz, int N, int d,
                  const double* __restrict__ X, const double* __restrict__ prm, double c,
                  double y_std, const double* __restrict__ scale, double* __restrict__ gmean,
                  double* __restrict__ gstd) {
  __shared__ double xt[MAX_DIM];
  __shared__ double inv_unused;
  __shared__ double red[2][256];
  (void)inv_unused;
  const int k = blockIdx.x, m = blockIdx.y;
  const double* ell = prm + 2 * MAX_DIM;
  for (int q = threadIdx.x; q < d; q += 256)
    xt[q] = (X[(size_t)m * d + q] - prm[q]) / prm[MAX_DIM + q];
  __syncthreads();
  const double* z = Z ? Z + (size_t)m * ldz : nullptr;
  double sm = 0.0, ss = 0.0;
  for (int j = threadIdx.x; j < N; j += 256) {
    double r2 = 0.0, dk = 0.0;
    for (int q = 0; q < d; q++) {
      double diff = (xt[q] - Xt[(size_t)j * d + q]) / ell[q];
      r2 = fma(diff, diff, r2);
      if (q == k) dk = diff;
    }
    double g;
    if (KIND == GPRY_KERNEL_RBF) {
      g = (-exp(-0.5 * r2) * dk) / ell[k];
    } else if (KIND == GPRY_KERNEL_MATERN15) {
      double dist = sqrt(r2);
      double s3d = 1.7320508075688772 * dist;
      double by = dist != 0.0 ? 1.7320508075688772 / dist : 0.0;
      double f_grad = (dk / ell[k]) * by;
      g = exp(-s3d) * f_grad * (1.0 - (1.0 + s3d));
    } else {
      double dist = sqrt(r2);
      double s5d = 2.23606797749979 * dist;
      double f = (5.0 / 3.0) * r2 + s5d + 1.0;
      double inv = dist != 0.0 ? 2.23606797749979 * (1.0 / dist) : 0.0;
      double dl = dk / ell[k];
      double f1g = inv * dl, f2g = (10.0 / 3.0) * dl;
      double gg = exp(-s5d);
      g = f * (-gg * f1g) + gg * (f1g + f2g);
    }
    const double cg = c * g;
    sm = fma(cg, alpha[j], sm);
    if (z) ss = fma(cg, z[j], ss);
  }
  red[0][threadIdx.x] = sm;
  red[1][threadIdx.x] = ss;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      red[0][threadIdx.x] += red[0][threadIdx.x + o];
      red[1][threadIdx.x] += red[1][threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    if (gmean) gmean[(size_t)m * d + k] = red[0][0] * y_std;
    if (gstd) gstd[(size_t)m * d + k] = red[1][0] * scale[2 * m];
  }
}

void predict_grad_device(gpry_state* st, const double* hX, int M, double* h_mean, double* h_std,
                         double* h_gmean, double* h_gstd) {
  if (!st->loaded) throw GpryError{GPRY_ERR_STATE, "no model uploaded into this state"};
  GPRY_CHECK_ARG(M >= 1 && M <= 8192, "batched gradients: 1 <= M <= 8192");
  if (h_gstd && !st->has_V) throw GpryError{GPRY_ERR_STATE, "this state was uploaded without V"};
  GPRY_CHECK_ARG(hX != nullptr, "X is NULL");
  GPRY_CUDA(cudaSetDevice(st->device));
  const int d = st->d, N = st->N, Np = st->Npad, DP = st->DP;
  const int Mp = round_up(M, TILE_ROWS);
  cudaStream_t s = 0;
  st->Xdev.reserve((size_t)M * d);
  GPRY_CUDA(cudaMemcpyAsync(st->Xdev.p, hX, (size_t)M * d * 8, cudaMemcpyHostToDevice, s));
  double *dm = nullptr, *ds = nullptr;
  if (h_mean) { st->o_mean.reserve(M); dm = st->o_mean.p; }
  if (h_std) { st->o_std.reserve(M); ds = st->o_std.p; }
  if (dm || ds)
    predict_pipeline(st, st->Xdev.p, M, dm != nullptr, ds != nullptr, false, 0, 0, 0, dm, ds, nullptr, s);
  st->gr_out.reserve((size_t)2 * M * d + 2 * (size_t)Mp);
  double* d_gm = st->gr_out.p;
  double* d_gs = d_gm + (size_t)M * d;
  double* d_scale = d_gs + (size_t)M * d;
  const double* Z = nullptr;
  if (h_gstd) {
    if (!st->vtrm_valid) {
      st->VTrm.reserve((size_t)Np * Np);
      transpose_sq_kernel<<<dim3(Np / 32, Np / 32), dim3(32, 8), 0, s>>>(st->Vrm.p, Np, st->VTrm.p);
      GPRY_CUDA(cudaGetLastError());
      st->vtrm_valid = true;
    }
    st->pc_U.reserve((size_t)Mp * DP);
    st->pc_Ks.reserve((size_t)Mp * Np);
    st->pc_UT.reserve((size_t)Mp * Np);
    scale_candidates_kernel<<<(Mp * DP + 255) / 256, 256, 0, s>>>(st->Xdev.p, M, d, Mp, DP,
                                                                  st->prm_dev.p, st->pc_U.p);
    GPRY_CUDA(cudaGetLastError());
    launch_kcross(st->kind, st->pc_U.p, M, Mp, st->T.p, N, DP, st->c, Np, nullptr, st->pc_Ks.p, s);
    // W = K* V^T  (W[m][j] = sum_k K*[m][k] V[j][k])
    gemm_nt(st->pc_Ks.p, Np, st->Vrm.p, Np, st->pc_UT.p, Np, Mp, Np, Np, 1.0, 0, 0, 0, s);
    grad_scale_batch_kernel<<<M, 256, 0, s>>>(st->pc_UT.p, Np, N, st->c, st->y_std, d_scale);
    GPRY_CUDA(cudaGetLastError());
    // Z = W V  (Z[m][i] = sum_j W[m][j] V[j][i]) -> reuses the K* buffer
    gemm_nt(st->pc_UT.p, Np, st->VTrm.p, Np, st->pc_Ks.p, Np, Mp, Np, Np, 1.0, 0, 0, 0, s);
    Z = st->pc_Ks.p;
  }
  if (h_gmean || h_gstd) {
    dim3 grid(d, M);
    double* gm = h_gmean ? d_gm : nullptr;
    double* gs = h_gstd ? d_gs : nullptr;
    switch (st->kind) {
      case GPRY_KERNEL_RBF:
        grad_batch_kernel<GPRY_KERNEL_RBF><<<grid, 256, 0, s>>>(
            st->Xt.p, st->alpha.p, Z, Np, N, d, st->Xdev.p, st->prm_dev.p, st->c, st->y_std, d_scale, gm, gs);
        break;
      case GPRY_KERNEL_MATERN15:
        grad_batch_kernel<GPRY_KERNEL_MATERN15><<<grid, 256, 0, s>>>(
            st->Xt.p, st->alpha.p, Z, Np, N, d, st->Xdev.p, st->prm_dev.p, st->c, st->y_std, d_scale, gm, gs);
        break;
      default:
        grad_batch_kernel<GPRY_KERNEL_MATERN25><<<grid, 256, 0, s>>>(
            st->Xt.p, st->alpha.p, Z, Np, N, d, st->Xdev.p, st->prm_dev.p, st->c, st->y_std, d_scale, gm, gs);
    }
    GPRY_CUDA(cudaGetLastError());
  }
  if (h_mean) GPRY_CUDA(cudaMemcpyAsync(h_mean, dm, (size_t)M * 8, cudaMemcpyDeviceToHost, s));
  if (h_std) GPRY_CUDA(cudaMemcpyAsync(h_std, ds, (size_t)M * 8, cudaMemcpyDeviceToHost, s));
  if (h_gmean)
    GPRY_CUDA(cudaMemcpyAsync(h_gmean, d_gm, (size_t)M * d * 8, cudaMemcpyDeviceToHost, s));
  if (h_gstd)
    GPRY_CUDA(cudaMemcpyAsync(h_gstd, d_gs, (size_t)M * d * 8, cudaMemcpyDeviceToHost, s));
  GPRY_CUDA(cudaStreamSynchronize(s));
}

}  // namespace gpry
