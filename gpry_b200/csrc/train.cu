// Training-side hot path: kernel matrix, blocked FP64 Cholesky, L^-1, alpha, log marginal
// likelihood and its gradient -- for a BATCH of hyper-parameter vectors at once.
//
// Reference arithmetic: gpr.py:1015-1017, 1453-1465 (_update_model / _kernel_inverse);
// sklearn:_gpr.py:584-651 (LML + gradient); kernel values and theta-gradients
// sklearn:kernels.py:1561-1584 (RBF), 1716-1771 (Matern), 964-969 (Product), 1283-1292
// (Constant).
//
// Matrices are padded to Np = round_up(N, 128) with an identity block in the padding (unit
// diagonal, zero coupling): every tile is full and the padded rows leave L, L^-1, alpha and
// log det untouched.  All kernels take the batch index from the grid (one theta per z / x).
//
//   kmat_kernel        K = c g(r) + diag(noise2)                            (lower tiles)
//   potf2_inv_kernel   128 x 128 diagonal block: L_jj, W_jj = L_jj^-1 (one CTA per theta:
//                      32 x 32 sub-blocks factored in registers by one warp, all 32^3 block
//                      products on DMMA); also writes the diagonal block of V^T
//   gemm_nt_kernel     C (+)= alpha A B^T on FP64 tensor cores (DMMA.8x8x4), cp.async
//                      4-stage pipeline, up to two row segments sharing the B operand.
//                      LEFT-LOOKING blocked algorithm: per block column j
//                        (1) A[j:, j] -= L[j:, :j] L[j, :j]^T      fused with
//                            TT      =  V^T[:j, :j..] L[j, :j]^T   (same B operand, long K)
//                        (2) potf2_inv on the diagonal block
//                        (3) L[j+1:, j] = A[j+1:, j] W_jj^T        fused with
//                            V^T[:j, j] = -TT W_jj^T               (same B operand W_jj)
//                      then K^-1 = V^T V (block-triangular k range).  Every output tile is
//                      written once; no trailing-matrix read-modify-write sweeps.
//   lml_grad_kernel    fused trace contraction 1/2 sum_ij (a_i a_j - K^-1_ij) dK_ij/dtheta:
//                      kernel values and per-dimension distances are recomputed per pair,
//                      dK/dtheta (N x N x (1+d)) is never materialised
#include <math.h>

#include <algorithm>
#include <vector>

#include <cstdlib>

#include "state.cuh"

namespace gpry {

constexpr int NB = 128;   // block size of the factorization (= GEMM tile)

// ---------------------------------------------------------------------------------------
// kernel matrix
// ---------------------------------------------------------------------------------------
template <int KIND>
__device__ __forceinline__ double stationary_value(double r2) {
  if (KIND == GPRY_KERNEL_RBF) return exp(-0.5 * r2);
  if (KIND == GPRY_KERNEL_MATERN15) {
    double K = sqrt(r2) * 1.7320508075688772;
    return (1.0 + K) * exp(-K);
  }
  double K = sqrt(r2) * 2.23606797749979;
  return (1.0 + K + K * K / 3.0) * exp(-K);
}


// T = X_ / ell (row major N x d per theta).  32 x 32 tile per CTA (lower tiles incl. diagonal).
template <int KIND>
__global__ void __launch_bounds__(256)
kmat_kernel(const double* __restrict__ Tall, int N, int d, int Np, const double* __restrict__ cs,
            const double* __restrict__ noise2, double* __restrict__ Kall) {
  const int bi = blockIdx.y, bj = blockIdx.x, th = blockIdx.z;
  if (bj > bi) return;
  const double* T = Tall + (size_t)th * N * d;
  double* K = Kall + (size_t)th * Np * Np;
  const double c = cs[th];
  extern __shared__ double sh[];
  double* Ti = sh;                    // [32][d+1]
  double* Tj = sh + 32 * (d + 1);     // [32][d+1]
  const int tid = threadIdx.x;
  for (int e = tid; e < 32 * d; e += 256) {
    int r = e / d, k = e % d;
    int gi = bi * 32 + r, gj = bj * 32 + r;
    Ti[r * (d + 1) + k] = gi < N ? T[(size_t)gi * d + k] : 0.0;
    Tj[r * (d + 1) + k] = gj < N ? T[(size_t)gj * d + k] : 0.0;
  }
  __syncthreads();
  const int tx = tid & 31, ty = tid >> 5;   // column tx, rows ty, ty+8, ...
  for (int rr = ty; rr < 32; rr += 8) {
    const int gi = bi * 32 + rr, gj = bj * 32 + tx;
    double v;
    if (gi >= N || gj >= N) {
      v = (gi == gj) ? 1.0 : 0.0;           // identity padding
    } else if (gi == gj) {
      v = c * 1.0 + noise2[gi];             // np.fill_diagonal(K, 1); K1*K2; += alpha
    } else {
      double r2 = 0.0;
      for (int k = 0; k < d; k++) {
        double df = Ti[rr * (d + 1) + k] - Tj[tx * (d + 1) + k];
        r2 = fma(df, df, r2);
      }
      v = c * stationary_value<KIND>(r2);
    }
    K[(size_t)gi * Np + gj] = v;
  }
}

// ---------------------------------------------------------------------------------------
// diagonal block: 128 x 128 Cholesky + triangular inverse in ONE CTA (8 warps) per theta.
//   A: the diagonal block (leading dimension ld), overwritten by L_jj (upper triangle zeroed)
//   W: dense 128 x 128 = L_jj^-1 (upper zeroed);  VTd: diagonal block of V^T = W^T
//   info: first failing global pivot index + 1 (0 = ok)
// The block is split in 4 x 4 sub-blocks of 32 x 32 held in shared memory ([32][36] padded,
// conflict-free for the DMMA fragment loads).  A 32 x 32 diagonal sub-block is factored and
// inverted by one warp entirely in registers (lane = row / column, operands exchanged with
// warp shuffles); every 32^3 product of the blocked algorithm runs on the FP64 tensor cores,
// one warp per product.
// ---------------------------------------------------------------------------------------
constexpr int SB = 32;          // sub-block size
constexpr int SBLD = 36;        // padded leading dimension (= 4 mod 16 doubles)
constexpr int SB_DOUBLES = SB * SBLD;
__host__ __device__ inline int sbidx(int i, int j) { return i * (i + 1) / 2 + j; }   // i >= j

// acc += A (32x32, [r][k]) * op(B);  NT: B stored [n][k];  NN: B stored [k][n]
template <bool NN>
__device__ __forceinline__ void mm32(const double* __restrict__ A, const double* __restrict__ B,
                                     double (&acc)[4][4][2], int lane) {
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int k4 = 0; k4 < 8; k4++) {
    double a[4], b[4];
#pragma unroll
    for (int mi = 0; mi < 4; mi++) a[mi] = A[(mi * 8 + g) * SBLD + k4 * 4 + t];
#pragma unroll
    for (int ni = 0; ni < 4; ni++)
      b[ni] = NN ? B[(k4 * 4 + t) * SBLD + ni * 8 + g] : B[(ni * 8 + g) * SBLD + k4 * 4 + t];
#pragma unroll
    for (int mi = 0; mi < 4; mi++)
#pragma unroll
      for (int ni = 0; ni < 4; ni++) dmma_8x8x4(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
  }
}
__device__ __forceinline__ void zero_acc(double (&acc)[4][4][2]) {
#pragma unroll
  for (int mi = 0; mi < 4; mi++)
#pragma unroll
    for (int ni = 0; ni < 4; ni++) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
}
// C = beta * C + alpha * acc   (C: [32][36] in shared memory)
__device__ __forceinline__ void store_acc(double* __restrict__ C, const double (&acc)[4][4][2],
                                          double alpha, double beta, int lane) {
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int mi = 0; mi < 4; mi++)
#pragma unroll
    for (int ni = 0; ni < 4; ni++) {
      double* c = C + (mi * 8 + g) * SBLD + ni * 8 + 2 * t;
      if (beta != 0.0) {
        c[0] = fma(alpha, acc[mi][ni][0], beta * c[0]);
        c[1] = fma(alpha, acc[mi][ni][1], beta * c[1]);
      } else {
        c[0] = alpha * acc[mi][ni][0];
        c[1] = alpha * acc[mi][ni][1];
      }
    }
}

// One warp: Cholesky of the 32 x 32 block D (lane = row, the row lives in registers,
// right-looking; the pivot column is broadcast through a 32-double shared buffer) and its
// inverse by forward substitution (lane = column; L is re-read from shared memory as
// broadcast loads).  1 / L_kk comes from rsqrt, so the loop has no sqrt / division chain.
// Writes L (upper zeroed) back into D and the inverse into Winv.  scratch: 96 doubles.
__device__ __forceinline__ void factor_inv_32(double* __restrict__ D, double* __restrict__ Winv,
                                              double* __restrict__ scratch, int lane,
                                              int pivot_offset, int* __restrict__ info) {
  double* col = scratch;          // [2][32] pivot column, double buffered
  double* invd = scratch + 64;    // [32]    1 / L_kk
  double a[SB];
#pragma unroll
  for (int k = 0; k < SB; k++) a[k] = D[lane * SBLD + k];
#pragma unroll
  for (int k = 0; k < SB; k++) {
    double dkk = __shfl_sync(0xffffffffu, a[k], k);
    if (!(dkk > 0.0)) {
      if (lane == 0 && *info == 0) *info = pivot_offset + k + 1;
      dkk = 1.0;
    }
    // 1/sqrt and sqrt to < 1 ulp (one Newton step each), then a / sqrt(d) as a correctly
    // rounded quotient by residual correction: no sqrt / division latency chain
    double rs = rsqrt(dkk);
    rs = fma(rs * 0.5, fma(-dkk * rs, rs, 1.0), rs);
    double sq = dkk * rs;
    sq = fma(fma(-sq, sq, dkk), 0.5 * rs, sq);
    double q = a[k] * rs;
    q = fma(fma(-q, sq, a[k]), rs, q);
    const double lk = (lane == k) ? sq : q;
    a[k] = lk;
    double* cb = col + (k & 1) * 32;
    cb[lane] = lk;
    if (lane == k) invd[k] = rs;
    __syncwarp();
#pragma unroll
    for (int j = k + 1; j < SB; j++) {
      const double ljk = cb[j];
      if (lane >= j) a[j] = fma(-lk, ljk, a[j]);
    }
  }
#pragma unroll
  for (int k = 0; k < SB; k++) D[lane * SBLD + k] = (k <= lane) ? a[k] : 0.0;
  __syncwarp();
  // inverse: lane c holds column c of W;  w_i = (delta_ic - sum_{m<i} L_im w_m) / L_ii
  double w[SB];
#pragma unroll
  for (int i = 0; i < SB; i++) {
    double s0 = (i == lane) ? 1.0 : 0.0, s1 = 0.0;
#pragma unroll
    for (int m = 0; m + 1 < i; m += 2) {
      s0 = fma(-D[i * SBLD + m], w[m], s0);
      s1 = fma(-D[i * SBLD + m + 1], w[m + 1], s1);
    }
    if (i & 1) s0 = fma(-D[i * SBLD + i - 1], w[i - 1], s0);
    const double si = s0 + s1, ri = invd[i];
    double wq = si * ri;
    wq = fma(fma(-wq, D[i * SBLD + i], si), ri, wq);     // si / L_ii by residual correction
    w[i] = (i >= lane) ? wq : 0.0;
  }
#pragma unroll
  for (int k = 0; k < SB; k++) Winv[k * SBLD + lane] = w[k];
}

__device__ __forceinline__ void potf2_inv_block(double* __restrict__ A, int ld,
                                                double* __restrict__ W, double* __restrict__ VTd,
                                                int row_offset, int* __restrict__ info,
                                                double* __restrict__ sh) {
  double* Lb = sh;                       // 10 lower sub-blocks of the block / of L
  double* Wb = sh + 10 * SB_DOUBLES;     // 10 lower sub-blocks of the inverse
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int e = tid; e < 128 * 128; e += 256) {
    int r = e >> 7, cidx = e & 127;
    int bi = r >> 5, bj = cidx >> 5;
    if (bj <= bi) Lb[sbidx(bi, bj) * SB_DOUBLES + (r & 31) * SBLD + (cidx & 31)] = A[(size_t)r * ld + cidx];
  }
  __syncthreads();
  double acc[4][4][2];
  for (int kb = 0; kb < 4; kb++) {
    if (warp == 0)
      factor_inv_32(Lb + sbidx(kb, kb) * SB_DOUBLES, Wb + sbidx(kb, kb) * SB_DOUBLES,
                    sh + 20 * SB_DOUBLES, lane, row_offset + kb * SB, info);
    __syncthreads();
    // panel: L[ib][kb] = A[ib][kb] * W_kk^T
    if (warp < 3 - kb) {
      const int ib = kb + 1 + warp;
      double* blk = Lb + sbidx(ib, kb) * SB_DOUBLES;
      zero_acc(acc);
      mm32<false>(blk, Wb + sbidx(kb, kb) * SB_DOUBLES, acc, lane);
      __syncwarp();
      store_acc(blk, acc, 1.0, 0.0, lane);
    }
    __syncthreads();
    // trailing: A[ib][jb] -= L[ib][kb] * L[jb][kb]^T,  kb < jb <= ib
    {
      int p = 0;
      for (int ib = kb + 1; ib < 4; ib++)
        for (int jb = kb + 1; jb <= ib; jb++, p++)
          if (p == warp) {
            zero_acc(acc);
            mm32<false>(Lb + sbidx(ib, kb) * SB_DOUBLES, Lb + sbidx(jb, kb) * SB_DOUBLES, acc, lane);
            store_acc(Lb + sbidx(ib, jb) * SB_DOUBLES, acc, -1.0, 1.0, lane);
          }
    }
    __syncthreads();
  }
  // off-diagonal blocks of the inverse: W[i][j] = -W[i][i] * sum_{k=j}^{i-1} L[i][k] W[k][j]
  for (int dist = 1; dist < 4; dist++) {
    if (warp < 4 - dist) {
      const int j = warp, i = j + dist;
      double* out = Wb + sbidx(i, j) * SB_DOUBLES;
      zero_acc(acc);
      for (int k = j; k < i; k++)
        mm32<true>(Lb + sbidx(i, k) * SB_DOUBLES, Wb + sbidx(k, j) * SB_DOUBLES, acc, lane);
      store_acc(out, acc, 1.0, 0.0, lane);
      __syncwarp();
      zero_acc(acc);
      mm32<true>(Wb + sbidx(i, i) * SB_DOUBLES, out, acc, lane);
      __syncwarp();
      store_acc(out, acc, -1.0, 0.0, lane);
    }
    __syncthreads();
  }
  for (int e = tid; e < 128 * 128; e += 256) {
    int r = e >> 7, cidx = e & 127;
    int bi = r >> 5, bj = cidx >> 5;
    double lv = 0.0, wv = 0.0;
    if (bj <= bi) {
      int o = sbidx(bi, bj) * SB_DOUBLES + (r & 31) * SBLD + (cidx & 31);
      lv = Lb[o];
      wv = Wb[o];
    }
    A[(size_t)r * ld + cidx] = lv;
    W[e] = wv;
    // V^T diagonal block: VTd[a][b] = W[b][a]
    double wt = 0.0;
    if (bi <= bj) wt = Wb[sbidx(bj, bi) * SB_DOUBLES + (cidx & 31) * SBLD + (r & 31)];
    VTd[(size_t)r * ld + cidx] = wt;
  }
}

constexpr size_t POTF2_SMEM_DOUBLES = 20 * (size_t)SB_DOUBLES + 96;

__global__ void __launch_bounds__(256)
potf2_inv_kernel(double* __restrict__ Aall, size_t sA, int ld, double* __restrict__ Wall, size_t sW,
                 double* __restrict__ VTall, size_t sVT, int row_offset,
                 int* __restrict__ info_all) {
  extern __shared__ double sh[];
  potf2_inv_block(Aall + blockIdx.x * sA, ld, Wall + blockIdx.x * sW, VTall + blockIdx.x * sVT,
                  row_offset, info_all + blockIdx.x, sh);
}

// ---------------------------------------------------------------------------------------
// FP64 tensor-core GEMM, C (+)= alpha * A * B^T   (A: M x K, B: N x K, both k-contiguous),
// batched over blockIdx.z, with up to two row segments that share the B operand.
// ---------------------------------------------------------------------------------------
constexpr int G_STAGES = 4;
constexpr int G_THREADS = 256;

struct GemmSeg {
  const double* A;
  double* C;
  int lda, ldc;
  int m_tiles;          // 128-row tiles in this segment
  double alpha;
  int accumulate;       // 0: C = alpha A B^T ; 1: C += alpha A B^T
  int klo_row;          // 1: k starts at the tile's first row (A upper triangular)
  size_t sA, sC;        // batch strides (doubles)
};
struct GemmArgs {
  GemmSeg seg[2];
  const double* B;
  int ldb;
  size_t sB;
  int n_tiles;          // 128-column tiles of C (rows of B)
  int K;                // multiple of 16
  int lower_only;       // only tiles with tile_row >= tile_col (segment 0)
  // fused diagonal-block factorisation: the CTA that owns row tile 0 of segment 0 (the
  // diagonal block of the panel update, dispatched first) goes on to factor and invert it
  // while the other CTAs are still working on the rest of the panel
  int fuse_potf2;
  double* pW;           // [batch] 128 x 128 inverse blocks (stride sW)
  double* pVT;          // [batch] diagonal block of V^T (stride sVT, leading dimension ldc of seg 0)
  size_t sW, sVT;
  int row_offset;
  int* info;
  // split-K: the k range of every tile is cut into `splits` contiguous parts computed by
  // different CTAs (gridDim.x = batch * splits); each writes its partial accumulators to
  // `scratch`, the LAST one to arrive at the tile's counter adds all parts in part order (so
  // the sum does not depend on the arrival order) and runs the epilogue.  Fills the GPU when a
  // launch has fewer tiles than SMs (single factorisations) or leaves a ragged last wave.
  int splits;
  int batch;
  double* scratch;      // [tile slot][split][128 x 128]
  int* counters;        // [tile slot], zero before the launch, reset by the last arriver
};

struct GemmSmem {
  double A[G_STAGES][TILE_DOUBLES];
  double B[G_STAGES][TILE_DOUBLES];
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}


__global__ void __launch_bounds__(G_THREADS, 1) gemm_nt_kernel(GemmArgs g) {
  // grid = (batch, row tiles, column tiles): the batch index varies fastest in the dispatch
  // order, so the tiles of all thetas are issued longest-K-first (row tile 0 of every theta,
  // then row tile 1, ...), which balances the triangular k ranges across the SMs.
  int mt = blockIdx.y;
  const int nt = blockIdx.z;
  const int si = (mt >= g.seg[0].m_tiles) ? 1 : 0;
  if (si) mt -= g.seg[0].m_tiles;
  const GemmSeg& sg = g.seg[si];
  if (g.lower_only && mt < nt) return;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  GemmSmem& sm = *reinterpret_cast<GemmSmem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = mt * TILE_ROWS, n0 = nt * TILE_ROWS;
  int k_lo = sg.klo_row ? m0 : 0;
  int nk = (g.K - k_lo) / TILE_K;
  const size_t bz = blockIdx.x % g.batch;
  const int split = blockIdx.x / g.batch;
  if (g.splits > 1) {       // this CTA's part of the k range
    const int k_a = (int)((long long)nk * split / g.splits);
    const int k_b = (int)((long long)nk * (split + 1) / g.splits);
    k_lo += k_a * TILE_K;
    nk = k_b - k_a;
  }

  const double* Ag = sg.A + bz * sg.sA + (size_t)m0 * sg.lda + k_lo;
  const double* Bg = g.B + bz * g.sB + (size_t)n0 * g.ldb + k_lo;
  const int lda = sg.lda, ldb = g.ldb;
  auto load_stage = [&](int s, int kt) {
    const int k0 = kt * TILE_K;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      int q = tid + G_THREADS * i;        // 0..1023: 16-byte chunk id
      int row = q >> 3, kc = (q & 7) * 2;
      cp_async16(&sm.A[s][tile_elem(row, kc)], Ag + (size_t)row * lda + k0 + kc);
      cp_async16(&sm.B[s][tile_elem(row, kc)], Bg + (size_t)row * ldb + k0 + kc);
    }
  };
  for (int s = 0; s < G_STAGES - 1; s++) {
    if (s < nk) load_stage(s, s);
    cp_async_commit();
  }
  const int rw = warp >> 2, cw = warp & 3;
  double acc[8][4][2];
#pragma unroll
  for (int mi = 0; mi < 8; mi++)
#pragma unroll
    for (int ni = 0; ni < 4; ni++) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;

  for (int kt = 0; kt < nk; kt++) {
    cp_async_wait<G_STAGES - 2>();
    __syncthreads();
    {
      int nxt = kt + G_STAGES - 1;
      if (nxt < nk) load_stage(nxt % G_STAGES, nxt);
      cp_async_commit();
    }
    const int s = kt % G_STAGES;
    const double* As = sm.A[s] + (rw * 64) * 4 + lane;
    const double* Bs = sm.B[s] + (cw * 32) * 4 + lane;
#pragma unroll
    for (int p = 0; p < 4; p++) {
      double a[8], b[4];
#pragma unroll
      for (int mi = 0; mi < 8; mi++) a[mi] = As[p * (TILE_ROWS * 4) + mi * 32];
#pragma unroll
      for (int ni = 0; ni < 4; ni++) b[ni] = Bs[p * (TILE_ROWS * 4) + ni * 32];
#pragma unroll
      for (int mi = 0; mi < 8; mi++)
#pragma unroll
        for (int ni = 0; ni < 4; ni++) dmma_8x8x4(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
    }
  }
  cp_async_wait<0>();
  __syncthreads();   // in-place segments: every warp's operand reads are done before C is written
  if (g.splits > 1) {
    // slot of this tile; partial accumulators as [split][register][thread] (coalesced)
    const size_t slot = ((size_t)blockIdx.z * gridDim.y + blockIdx.y) * g.batch + bz;
    double* part = g.scratch + (slot * g.splits + split) * (size_t)(TILE_ROWS * TILE_ROWS);
#pragma unroll
    for (int mi = 0; mi < 8; mi++)
#pragma unroll
      for (int ni = 0; ni < 4; ni++) {
        __stcg(part + (size_t)((mi * 4 + ni) * 2 + 0) * G_THREADS + tid, acc[mi][ni][0]);
        __stcg(part + (size_t)((mi * 4 + ni) * 2 + 1) * G_THREADS + tid, acc[mi][ni][1]);
      }
    __threadfence();
    __syncthreads();
    __shared__ int s_last;
    if (tid == 0) {
      const int old = atomicAdd(g.counters + slot, 1);
      s_last = old == g.splits - 1;
      if (s_last) g.counters[slot] = 0;       // ready for the next launch
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const double* all = g.scratch + slot * g.splits * (size_t)(TILE_ROWS * TILE_ROWS);
#pragma unroll
    for (int mi = 0; mi < 8; mi++)
#pragma unroll
      for (int ni = 0; ni < 4; ni++) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
    for (int sp = 0; sp < g.splits; sp++) {      // fixed order: parts 0, 1, ...
      const double* q = all + (size_t)sp * (TILE_ROWS * TILE_ROWS);
#pragma unroll
      for (int mi = 0; mi < 8; mi++)
#pragma unroll
        for (int ni = 0; ni < 4; ni++) {
          acc[mi][ni][0] += __ldcg(q + (size_t)((mi * 4 + ni) * 2 + 0) * G_THREADS + tid);
          acc[mi][ni][1] += __ldcg(q + (size_t)((mi * 4 + ni) * 2 + 1) * G_THREADS + tid);
        }
    }
  }
  // epilogue
  const int g8 = lane >> 2, t4 = lane & 3;
  double* Cg = sg.C + bz * sg.sC;
  const double alpha = sg.alpha;
  const int ldc = sg.ldc;
#pragma unroll
  for (int mi = 0; mi < 8; mi++)
#pragma unroll
    for (int ni = 0; ni < 4; ni++) {
      int row = m0 + rw * 64 + mi * 8 + g8;
      int col = n0 + cw * 32 + ni * 8 + 2 * t4;
      double2* cp = reinterpret_cast<double2*>(Cg + (size_t)row * ldc + col);
      double2 v = make_double2(alpha * acc[mi][ni][0], alpha * acc[mi][ni][1]);
      if (sg.accumulate) {
        double2 o = *cp;
        v.x += o.x;
        v.y += o.y;
      }
      *cp = v;
    }
  if (g.fuse_potf2 && si == 0 && mt == 0) {
    __threadfence_block();
    __syncthreads();                     // the whole updated block is in global memory
    potf2_inv_block(Cg, ldc, g.pW + bz * g.sW, g.pVT + bz * g.sVT, g.row_offset, g.info + bz,
                    reinterpret_cast<double*>(smem_raw));
  }
}

// scratch / counters of the split-K path and the number of SMs (set by gemm_prepare)
static double* g_split_scratch = nullptr;
static size_t g_split_scratch_doubles = 0;
static int* g_split_counters = nullptr;
static size_t g_split_counters_n = 0;
static int g_n_sm = 148;
static int g_splitk_enabled = -1;

// Parts per tile.  Measured (profiles/r02_train_ab.txt): splitting pays when a launch leaves SMs
// idle -- fewer tiles than SMs (single factorisations: 32 tiles at N = 4000) or a ragged second
// wave -- and costs ~3 % per part (partial accumulators written and re-read through L2), so a
// split must win 6 % below one wave of tiles and 15 % above.  1 = whole tiles.
static int choose_splits(int tiles, int max_nk) {
  if (g_splitk_enabled < 0) {
    const char* e = getenv("GPRY_B200_SPLITK");
    g_splitk_enabled = (e && atoi(e) == 0) ? 0 : 1;
  }
  if (!g_splitk_enabled || tiles <= 0) return 1;
  const double need = tiles < g_n_sm ? 0.94 : 0.85;
  double best_cost = 1e30;
  int best = 1;
  for (int sp = 1; sp <= 8; sp++) {
    if (sp > 1 && max_nk / sp < 4) break;   // (a short tile may end up with empty parts: fine)
    const int ctas = tiles * sp;
    const double rounds = (double)((ctas + g_n_sm - 1) / g_n_sm) / sp;   // in whole-tile times
    const double cost = rounds * (sp > 1 ? 1.0 + 0.03 * sp : 1.0);
    if (cost < best_cost * need) {
      best_cost = cost;
      best = sp;
    }
  }
  return best;
}

static void launch_gemm(const GemmArgs& g_in, int batch, cudaStream_t s) {
  GemmArgs g = g_in;
  const int mt = g.seg[0].m_tiles + g.seg[1].m_tiles;
  if (mt <= 0 || g.n_tiles <= 0 || g.K <= 0 || batch <= 0) return;
  // active tiles and the shortest k range among them
  int tiles = 0, max_nk = 0;
  for (int nt = 0; nt < g.n_tiles; nt++)
    for (int m = 0; m < mt; m++) {
      const int si = m >= g.seg[0].m_tiles ? 1 : 0;
      const int mm = si ? m - g.seg[0].m_tiles : m;
      if (g.lower_only && mm < nt) continue;
      const int k_lo = g.seg[si].klo_row ? mm * TILE_ROWS : 0;
      tiles++;
      max_nk = std::max(max_nk, (g.K - k_lo) / TILE_K);
    }
  g.batch = batch;
  g.splits = choose_splits(tiles * batch, max_nk);
  if (g.splits > 1) {
    const size_t slots = (size_t)g.n_tiles * mt * batch;
    size_t need = slots * g.splits * (size_t)(TILE_ROWS * TILE_ROWS);
    if (need > g_split_scratch_doubles) {
      need = std::max(need + need / 2, (size_t)32 << 20);     // grow rarely (256 MB at least)
      if (g_split_scratch) cudaFree(g_split_scratch);
      g_split_scratch = nullptr;
      g_split_scratch_doubles = 0;
      if (cudaMalloc((void**)&g_split_scratch, need * 8) != cudaSuccess) {
        cudaGetLastError();
        g.splits = 1;                    // no room: whole tiles
      } else {
        g_split_scratch_doubles = need;
      }
    }
    if (g.splits > 1 && slots > g_split_counters_n) {
      if (g_split_counters) cudaFree(g_split_counters);
      GPRY_CUDA(cudaMalloc((void**)&g_split_counters, slots * sizeof(int)));
      GPRY_CUDA(cudaMemsetAsync(g_split_counters, 0, slots * sizeof(int), s));
      g_split_counters_n = slots;
    }
    g.scratch = g_split_scratch;
    g.counters = g_split_counters;
  }
  dim3 grid(batch * g.splits, mt, g.n_tiles);
  const size_t smem = g.fuse_potf2 ? std::max(sizeof(GemmSmem), POTF2_SMEM_DOUBLES * 8)
                                   : sizeof(GemmSmem);
  gemm_nt_kernel<<<grid, G_THREADS, smem, s>>>(g);
  GPRY_CUDA(cudaGetLastError());
}
static void gemm_prepare() {
  GPRY_CUDA(cudaFuncSetAttribute(gemm_nt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)std::max(sizeof(GemmSmem), POTF2_SMEM_DOUBLES * 8)));
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) == cudaSuccess &&
      cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
    g_n_sm = n;
}

void gemm_nt(const double* A, int lda, const double* B, int ldb, double* C, int ldc, int M, int N,
             int K, double alpha, int accumulate, int lower_only, int klo_row, cudaStream_t s) {
  GemmArgs g{};
  g.seg[0] = GemmSeg{A, C, lda, ldc, M / TILE_ROWS, alpha, accumulate, klo_row, 0, 0};
  g.seg[1].m_tiles = 0;
  g.B = B; g.ldb = ldb; g.sB = 0;
  g.n_tiles = N / TILE_ROWS; g.K = K; g.lower_only = lower_only;
  gemm_prepare();
  launch_gemm(g, 1, s);
}

// ---------------------------------------------------------------------------------------
// small helpers (batch index = blockIdx.y unless noted)
// ---------------------------------------------------------------------------------------
// Tall[th][e] = X[e] / ell[th][e % d]
__global__ void scale_rows_kernel(const double* __restrict__ X, int N, int d,
                                  const double* __restrict__ ells, double* __restrict__ Tall) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  const int th = blockIdx.y;
  if (e < N * d) Tall[(size_t)th * N * d + e] = X[e] / ells[(size_t)th * MAX_DIM + e % d];
}
// t = V y = VT^T y : CTA per 32 columns of VT, 8 row groups, rows j <= i; fixed-order sums
__global__ void __launch_bounds__(256)
gemv_vt_t_kernel(const double* __restrict__ VTall, int Np, int N, const double* __restrict__ y,
                 double* __restrict__ tall) {
  const double* VT = VTall + (size_t)blockIdx.y * Np * Np;
  double* t = tall + (size_t)blockIdx.y * Np;
  __shared__ double red[8][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + cx;
  double s = 0.0;
  if (i < N) {
    const int jmax = min(N - 1, blockIdx.x * 32 + 31);
    for (int j = ry; j <= jmax; j += 8)
      if (j <= i) s = fma(VT[(size_t)j * Np + i], y[j], s);
  }
  red[ry][cx] = s;
  __syncthreads();
  if (ry == 0) {
    double v = 0.0;
    for (int q = 0; q < 8; q++) v += red[q][cx];
    if (i < Np) t[i] = v;
  }
}
// alpha = V^T t = VT t : warp per row j, columns i >= j
__global__ void gemv_vt_kernel(const double* __restrict__ VTall, int Np, int N,
                               const double* __restrict__ tall, double* __restrict__ alpha_all) {
  const double* VT = VTall + (size_t)blockIdx.y * Np * Np;
  const double* t = tall + (size_t)blockIdx.y * Np;
  double* alpha = alpha_all + (size_t)blockIdx.y * Np;
  int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (j >= Np) return;
  double s = 0.0;
  if (j < N)
    for (int i = j + lane; i < N; i += 32) s = fma(VT[(size_t)j * Np + i], t[i], s);
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) alpha[j] = s;
}
// out[th][0] = sum_i log L_ii, out[th][1] = y . alpha      (one block per theta, fixed order)
__global__ void lml_scalars_kernel(const double* __restrict__ Lall, int Np, int N,
                                   const double* __restrict__ y,
                                   const double* __restrict__ alpha_all,
                                   double* __restrict__ out_all) {
  const double* L = Lall + (size_t)blockIdx.x * Np * Np;
  const double* alpha = alpha_all + (size_t)blockIdx.x * Np;
  double* out = out_all + (size_t)blockIdx.x * 2;
  __shared__ double r0[256], r1[256];
  double a = 0.0, b = 0.0;
  for (int i = threadIdx.x; i < N; i += 256) {
    a += log(L[(size_t)i * Np + i]);
    b = fma(y[i], alpha[i], b);
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
    out[0] = r0[0];
    out[1] = r1[0];
  }
}
// dense N x N row-major copies for the host: lower triangle of L; V = VT^T
__global__ void extract_kernel(const double* __restrict__ src, int Np, int N, int transpose,
                               double* __restrict__ dst) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (int64_t)N * N) return;
  int r = (int)(e / N), cidx = (int)(e % N);
  double v = 0.0;
  if (cidx <= r) v = transpose ? src[(size_t)cidx * Np + r] : src[(size_t)r * Np + cidx];
  dst[e] = v;
}

// ---------------------------------------------------------------------------------------
// fused LML-gradient trace contraction
//   grad_p = 1/2 sum_ij (alpha_i alpha_j - Kinv_ij) dK_ij / dtheta_p   (sklearn:_gpr.py:647)
// 32 x 32 tile of pairs per CTA over the lower triangle; off-diagonal tiles count twice.
// partial[th][tile][p] written per CTA, reduced in fixed order by reduce_partials_kernel.
// ---------------------------------------------------------------------------------------
template <int KIND>
__global__ void __launch_bounds__(256)
lml_grad_kernel(const double* __restrict__ X, const double* __restrict__ Tall, int N, int d,
                int Np, const double* __restrict__ cs, const double* __restrict__ ells,
                const double* __restrict__ alpha_all, const double* __restrict__ Kinv_all,
                double* __restrict__ partial_all, int P) {
  const int bi = blockIdx.y, bj = blockIdx.x, th = blockIdx.z;
  const int nb = gridDim.x;
  const double* T = Tall + (size_t)th * N * d;
  const double* ell = ells + (size_t)th * MAX_DIM;
  const double* alpha = alpha_all + (size_t)th * Np;
  const double* Kinv = Kinv_all + (size_t)th * Np * Np;
  const double c = cs[th];
  double* out = partial_all + (((size_t)th * nb + bi) * nb + bj) * P;
  extern __shared__ double sh[];
  double* Xi = sh;                      // [32][d+1]
  double* Xj = Xi + 32 * (d + 1);       // [32][d+1]
  double* Ti = Xj + 32 * (d + 1);       // [32][d+1]  X / ell
  double* Tj = Ti + 32 * (d + 1);       // [32][d+1]
  double* l2 = Tj + 32 * (d + 1);       // [d]  1 / ell^2
  double* red = l2 + d;                 // [8 warps][P]
  const int tid = threadIdx.x;
  if (bj > bi) {
    for (int p = tid; p < P; p += 256) out[p] = 0.0;
    return;
  }
  for (int e = tid; e < 32 * d; e += 256) {
    int r = e / d, k = e % d;
    int gi = bi * 32 + r, gj = bj * 32 + r;
    Xi[r * (d + 1) + k] = gi < N ? X[(size_t)gi * d + k] : 0.0;
    Xj[r * (d + 1) + k] = gj < N ? X[(size_t)gj * d + k] : 0.0;
    Ti[r * (d + 1) + k] = gi < N ? T[(size_t)gi * d + k] : 0.0;
    Tj[r * (d + 1) + k] = gj < N ? T[(size_t)gj * d + k] : 0.0;
  }
  for (int k = tid; k < d; k += 256) l2[k] = 1.0 / (ell[k] * ell[k]);   // 1 / length_scale**2
  __syncthreads();
  const int tx = tid & 31, ty = tid >> 5, lane = tx, warp = ty;
  // each thread: column tx, rows ty + 8 q.  Pass 1 computes w * (kernel factor) per pair and
  // the constant-kernel term; the per-dimension sums are accumulated dimension by dimension
  // so that P can be arbitrary (no P registers per thread).
  double wk[4];        // w_ij * factor multiplying D_k for the 4 pairs of this thread
  double g0 = 0.0;     // d/dlog c term
  const double mult = (bi == bj) ? 1.0 : 2.0;
#pragma unroll
  for (int q = 0; q < 4; q++) {
    const int rr = ty + 8 * q;
    const int gi = bi * 32 + rr, gj = bj * 32 + tx;
    wk[q] = 0.0;
    if (gi < N && gj < N) {
      double w = alpha[gi] * alpha[gj] - Kinv[(size_t)max(gi, gj) * Np + min(gi, gj)];
      if (gi == gj) {
        g0 = fma(w, c, g0);                       // K2 = 1 on the diagonal, D = 0
      } else {
        double sumD = 0.0, r2 = 0.0;
        for (int k = 0; k < d; k++) {
          double df = Xi[rr * (d + 1) + k] - Xj[tx * (d + 1) + k];
          double D = df * df * l2[k];             // (xi - xj)**2 / length_scale**2
          sumD += D;
          double a = Ti[rr * (d + 1) + k] - Tj[tx * (d + 1) + k];
          r2 = fma(a, a, r2);                     // pdist(X / l): the value entering K itself
        }
        double k2, fac;
        if (KIND == GPRY_KERNEL_RBF) {
          k2 = exp(-0.5 * r2);
          fac = k2;                               // K_gradient = D * K2        (:1581-1584)
        } else if (KIND == GPRY_KERNEL_MATERN15) {
          k2 = stationary_value<KIND>(r2);
          fac = 3.0 * exp(-sqrt(3.0 * sumD));     // 3 D exp(-sqrt(3 sum D))    (:1768)
        } else {
          k2 = stationary_value<KIND>(r2);
          double tmp = sqrt(5.0 * sumD);
          fac = 5.0 / 3.0 * (tmp + 1.0) * exp(-tmp);   // (:1770-1771)
        }
        g0 = fma(mult * w, c * k2, g0);           // dK/dlog c = K1_gradient * K2 = c K2
        wk[q] = mult * w * c * fac;               // K2_gradient * K1
      }
    }
  }
  auto block_sum_to = [&](double v, int p) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[warp * P + p] = v;
  };
  block_sum_to(g0, 0);
  for (int k = 0; k < d; k++) {
    double s = 0.0;
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const int rr = ty + 8 * q;
      double df = Xi[rr * (d + 1) + k] - Xj[tx * (d + 1) + k];
      s = fma(wk[q], df * df * l2[k], s);
    }
    block_sum_to(s, 1 + k);
  }
  __syncthreads();
  for (int p = tid; p < P; p += 256) {
    double s = 0.0;
    for (int w8 = 0; w8 < 8; w8++) s += red[w8 * P + p];
    out[p] = 0.5 * s;
  }
}

// grid (P, B): out[th][p] = sum over tiles, fixed-order tree
__global__ void reduce_partials_kernel(const double* __restrict__ partial_all, int n, int P,
                                       double* __restrict__ out_all) {
  __shared__ double r[256];
  const int p = blockIdx.x, th = blockIdx.y;
  const double* partial = partial_all + (size_t)th * n * P;
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) s += partial[(size_t)i * P + p];
  r[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) r[threadIdx.x] += r[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out_all[(size_t)th * (MAX_DIM + 8) + p] = r[0];
}

// ---------------------------------------------------------------------------------------
// host orchestration
// ---------------------------------------------------------------------------------------
struct TrainBuffers {
  int N, d, Np, nb, B;        // B = thetas per sub-batch
  double *K, *VT, *W;         // [B][Np][Np] each (W: K^-1, only for the gradient)
  double *TT;                 // [B][Np][128]
  double *Winv;               // [B][nb][128][128]
  double *X, *noise2, *y;     // shared problem
  double *T, *ells, *cs, *alpha, *t, *scal, *partial, *grad;
  int* info;
};

// f_prob : [y Np][noise2 Np][X_ N*d]
// f_misc : [alpha B*Np][t B*Np][T B*nXd][ells B*MAX_DIM][cs B][scal B*2][grad B*(MAX_DIM+8)]
//          [info B ints][partial B*nb32^2*P]
static TrainBuffers carve(gpry_state* st, int N, int d, bool need_grad, int B) {
  TrainBuffers b;
  b.N = N; b.d = d; b.B = B;
  b.Np = round_up(N, NB);
  b.nb = b.Np / NB;
  const size_t Np = b.Np, NN = Np * Np;
  auto al = [](size_t x) { return (x + 31) / 32 * 32; };   // 256-byte aligned segments
  st->f_K.reserve(NN * B);
  st->f_VT.reserve(NN * B);
  if (need_grad) st->f_W.reserve(NN * B);
  st->f_TT.reserve(Np * NB * B);
  st->f_Winv.reserve((size_t)b.nb * NB * NB * B);
  const size_t nXd = al(Np * d);   // room for rows appended later (factor_append_device)
  st->f_prob.reserve(2 * Np + nXd);
  const int nb32 = (N + 31) / 32, P = d + 1;
  size_t n = 2 * Np * B + nXd * B + al((size_t)MAX_DIM * B) + al(B) + al(2 * (size_t)B) +
             al((size_t)(MAX_DIM + 8) * B) + al(B) + (need_grad ? (size_t)nb32 * nb32 * P * B : 0) + 64;
  st->f_misc.reserve(n);
  b.y = st->f_prob.p;
  b.noise2 = b.y + Np;
  b.X = b.noise2 + Np;
  double* p = st->f_misc.p;
  b.alpha = p; p += Np * B;
  b.t = p; p += Np * B;
  b.T = p; p += nXd * B;
  b.ells = p; p += al((size_t)MAX_DIM * B);
  b.cs = p; p += al(B);
  b.scal = p; p += al(2 * (size_t)B);
  b.grad = p; p += al((size_t)(MAX_DIM + 8) * B);
  b.info = reinterpret_cast<int*>(p); p += al(B);
  b.partial = p;
  b.K = st->f_K.p;
  b.VT = st->f_VT.p;
  b.W = need_grad ? st->f_W.p : nullptr;
  b.TT = st->f_TT.p;
  b.Winv = st->f_Winv.p;
  return b;
}

static void upload_problem(const TrainBuffers& b, const double* X, const double* noise2,
                           const double* y, cudaStream_t s) {
  const size_t Np = b.Np;
  GPRY_CUDA(cudaMemsetAsync(b.y, 0, 2 * Np * 8, s));
  GPRY_CUDA(cudaMemcpyAsync(b.X, X, (size_t)b.N * b.d * 8, cudaMemcpyHostToDevice, s));
  GPRY_CUDA(cudaMemcpyAsync(b.noise2, noise2, (size_t)b.N * 8, cudaMemcpyHostToDevice, s));
  GPRY_CUDA(cudaMemcpyAsync(b.y, y, (size_t)b.N * 8, cudaMemcpyHostToDevice, s));
}

template <int KIND>
static void launch_kmat(const TrainBuffers& b, int nth, cudaStream_t s) {
  const int nb32 = b.Np / 32;
  size_t smem = 2 * 32 * (size_t)(b.d + 1) * 8;
  if (smem > 48 * 1024)
    GPRY_CUDA(cudaFuncSetAttribute(kmat_kernel<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)smem));
  kmat_kernel<KIND><<<dim3(nb32, nb32, nth), 256, smem, s>>>(b.T, b.N, b.d, b.Np, b.cs, b.noise2,
                                                            b.K);
  GPRY_CUDA(cudaGetLastError());
}

template <int KIND>
static void launch_grad(const TrainBuffers& b, int nth, int nb32, cudaStream_t s) {
  const int P = b.d + 1;
  size_t smem = (4 * 32 * (size_t)(b.d + 1) + b.d + 8 * (size_t)P) * 8;
  if (smem > 48 * 1024)
    GPRY_CUDA(cudaFuncSetAttribute(lml_grad_kernel<KIND>,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  lml_grad_kernel<KIND><<<dim3(nb32, nb32, nth), 256, smem, s>>>(
      b.X, b.T, b.N, b.d, b.Np, b.cs, b.ells, b.alpha, b.W, b.partial, P);
  GPRY_CUDA(cudaGetLastError());
}

// For nth thetas (host array, stride P): K -> L (in K), VT = L^-T, alpha, scal = {sum log diag
// L, y.alpha}; optionally W = K^-1 and grad.  Everything stays on the device.
static void factorize_batch(gpry_state* st, TrainBuffers& b, int kind, const double* thetas,
                            int nth, bool need_grad, cudaStream_t s) {
  const int N = b.N, d = b.d, Np = b.Np, nb = b.nb, P = d + 1;
  const size_t NN = (size_t)Np * Np;
  // exp(theta) -> device (pageable copy: staged before this call returns)
  {
    std::vector<double> h((size_t)MAX_DIM * nth + nth, 1.0);
    for (int i = 0; i < nth; i++) {
      for (int k = 0; k < d; k++) h[(size_t)i * MAX_DIM + k] = exp(thetas[(size_t)i * P + 1 + k]);
      h[(size_t)MAX_DIM * nth + i] = exp(thetas[(size_t)i * P]);
    }
    GPRY_CUDA(cudaMemcpyAsync(b.ells, h.data(), (size_t)MAX_DIM * nth * 8, cudaMemcpyHostToDevice, s));
    GPRY_CUDA(cudaMemcpyAsync(b.cs, h.data() + (size_t)MAX_DIM * nth, (size_t)nth * 8,
                              cudaMemcpyHostToDevice, s));
  }
  GPRY_CUDA(cudaMemsetAsync(b.info, 0, (size_t)nth * sizeof(int), s));
  scale_rows_kernel<<<dim3((N * d + 255) / 256, nth), 256, 0, s>>>(b.X, N, d, b.ells, b.T);
  GPRY_CUDA(cudaGetLastError());
  switch (kind) {
    case GPRY_KERNEL_RBF: launch_kmat<GPRY_KERNEL_RBF>(b, nth, s); break;
    case GPRY_KERNEL_MATERN15: launch_kmat<GPRY_KERNEL_MATERN15>(b, nth, s); break;
    default: launch_kmat<GPRY_KERNEL_MATERN25>(b, nth, s);
  }
  gemm_prepare();
  const size_t potf2_smem = POTF2_SMEM_DOUBLES * 8;
  GPRY_CUDA(cudaFuncSetAttribute(potf2_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)potf2_smem));
  const size_t sWinv = (size_t)nb * NB * NB, sTT = (size_t)Np * NB;
  for (int j = 0; j < nb; j++) {
    const size_t jB = (size_t)j * NB;
    if (j > 0) {
      // (1) block column j of A minus the contribution of the finished columns, and the
      //     V^T row-sweep product, sharing the B operand L[j, :j];  (2) the CTA of the diagonal
      //     tile then factors and inverts it (potf2_inv_block) inside the same launch
      GemmArgs g{};
      g.seg[0] = GemmSeg{b.K + jB * Np, b.K + jB * Np + jB, Np, Np, nb - j, -1.0, 1, 0, NN, NN};
      g.seg[1] = GemmSeg{b.VT, b.TT, Np, NB, j, 1.0, 0, 1, NN, sTT};
      g.B = b.K + jB * Np; g.ldb = Np; g.sB = NN;
      g.n_tiles = 1; g.K = (int)jB; g.lower_only = 0;
      g.fuse_potf2 = 1;
      g.pW = b.Winv + jB * NB; g.sW = sWinv;
      g.pVT = b.VT + jB * (Np + 1); g.sVT = NN;
      g.row_offset = (int)jB; g.info = b.info;
      launch_gemm(g, nth, s);
    } else {
      // (2) first diagonal block: nothing to subtract
      potf2_inv_kernel<<<nth, 256, potf2_smem, s>>>(b.K, NN, Np, b.Winv, sWinv, b.VT, NN, 0,
                                                    b.info);
      GPRY_CUDA(cudaGetLastError());
    }
    // (3) panel solve and the new block column of V^T, sharing the B operand W_jj
    if (nb - j - 1 > 0 || j > 0) {
      GemmArgs g{};
      double* panel = b.K + (jB + NB) * Np + jB;
      g.seg[0] = GemmSeg{panel, panel, Np, Np, nb - j - 1, 1.0, 0, 0, NN, NN};
      g.seg[1] = GemmSeg{b.TT, b.VT + jB, NB, Np, j, -1.0, 0, 0, sTT, NN};
      g.B = b.Winv + jB * NB; g.ldb = NB; g.sB = sWinv;
      g.n_tiles = 1; g.K = NB; g.lower_only = 0;
      launch_gemm(g, nth, s);
    }
  }
  // alpha = V^T (V y)
  gemv_vt_t_kernel<<<dim3(Np / 32, nth), 256, 0, s>>>(b.VT, Np, N, b.y, b.t);
  GPRY_CUDA(cudaGetLastError());
  gemv_vt_kernel<<<dim3((Np * 32 + 255) / 256, nth), 256, 0, s>>>(b.VT, Np, N, b.t, b.alpha);
  GPRY_CUDA(cudaGetLastError());
  lml_scalars_kernel<<<nth, 256, 0, s>>>(b.K, Np, N, b.y, b.alpha, b.scal);
  GPRY_CUDA(cudaGetLastError());
  if (need_grad) {
    // K^-1 = V^T V = VT VT^T (lower tiles; VT upper triangular: k >= tile row)
    GemmArgs g{};
    g.seg[0] = GemmSeg{b.VT, b.W, Np, Np, nb, 1.0, 0, 1, NN, NN};
    g.seg[1].m_tiles = 0;
    g.B = b.VT; g.ldb = Np; g.sB = NN;
    g.n_tiles = nb; g.K = Np; g.lower_only = 1;
    launch_gemm(g, nth, s);
    const int nb32 = (N + 31) / 32;
    switch (kind) {
      case GPRY_KERNEL_RBF: launch_grad<GPRY_KERNEL_RBF>(b, nth, nb32, s); break;
      case GPRY_KERNEL_MATERN15: launch_grad<GPRY_KERNEL_MATERN15>(b, nth, nb32, s); break;
      default: launch_grad<GPRY_KERNEL_MATERN25>(b, nth, nb32, s);
    }
    reduce_partials_kernel<<<dim3(P, nth), 256, 0, s>>>(b.partial, nb32 * nb32, P, b.grad);
    GPRY_CUDA(cudaGetLastError());
  }
}

// thetas per sub-batch: bounded by memory (3 Np^2 doubles each) and by what fills the GPU
static int sub_batch(int B, int Np, bool need_grad) {
  const double per_theta = (need_grad ? 3.0 : 2.0) * Np * (double)Np * 8.0;
  const int cap = (int)std::max(1.0, std::min(48.0, std::floor(20e9 / per_theta)));
  if (B <= cap) return B;
  // several sub-batches: the size whose panel launches (Np / 128 tiles per theta) waste the
  // least of their last wave on this GPU, counted over all sub-batches
  const int tiles_per_theta = Np / NB;
  int best = std::min(B, cap);
  double best_cost = 1e30;
  for (int bs = cap; bs >= std::max(1, cap / 2); bs--) {
    double rounds = 0;
    for (int done = 0; done < B; done += bs) {
      const int n = std::min(bs, B - done);
      rounds += (double)((n * tiles_per_theta + g_n_sm - 1) / g_n_sm);
    }
    if (rounds < best_cost) {
      best_cost = rounds;
      best = bs;
    }
  }
  return best;
}

void factorize_device(gpry_state* st, int kind, int N, int d, const double* X_train_t,
                      const double* noise2, const double* y_t, const double* theta, double* out_L,
                      double* out_V, double* out_alpha, double* out_logdet_half, int* info,
                      bool keep) {
  GPRY_CHECK_ARG(kind >= 0 && kind <= 2, "unknown kernel kind");
  GPRY_CHECK_ARG(N >= 1 && d >= 1 && d <= MAX_DIM, "need N >= 1 and 1 <= d <= 128");
  GPRY_CUDA(cudaSetDevice(st->device));
  st->f_valid = false;
  cudaStream_t s = 0;
  TrainBuffers b = carve(st, N, d, false, 1);
  upload_problem(b, X_train_t, noise2, y_t, s);
  factorize_batch(st, b, kind, theta, 1, false, s);
  int h_info = 0;
  double scal[2];
  GPRY_CUDA(cudaMemcpyAsync(&h_info, b.info, sizeof(int), cudaMemcpyDeviceToHost, s));
  GPRY_CUDA(cudaMemcpyAsync(scal, b.scal, 16, cudaMemcpyDeviceToHost, s));
  if (out_alpha) GPRY_CUDA(cudaMemcpyAsync(out_alpha, b.alpha, (size_t)N * 8, cudaMemcpyDeviceToHost, s));
  if (out_L || out_V) {
    st->tmp.reserve((size_t)N * N);
    int64_t tot = (int64_t)N * N;
    if (out_L) {
      extract_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(b.K, b.Np, N, 0, st->tmp.p);
      GPRY_CUDA(cudaGetLastError());
      GPRY_CUDA(cudaMemcpyAsync(out_L, st->tmp.p, tot * 8, cudaMemcpyDeviceToHost, s));
    }
    if (out_V) {
      extract_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(b.VT, b.Np, N, 1, st->tmp.p);
      GPRY_CUDA(cudaGetLastError());
      GPRY_CUDA(cudaMemcpyAsync(out_V, st->tmp.p, tot * 8, cudaMemcpyDeviceToHost, s));
    }
  }
  GPRY_CUDA(cudaStreamSynchronize(s));
  *info = h_info;
  if (out_logdet_half) *out_logdet_half = scal[0];
  if (keep && h_info == 0) {
    st->f_valid = true;
    st->f_N = N;
    st->f_d = d;
    st->f_kind = kind;
  }
}

// L and / or V of the factorisation kept resident by factorize_device(keep = true)
void factor_download_device(gpry_state* st, double* out_L, double* out_V) {
  if (!st->f_valid) throw GpryError{GPRY_ERR_STATE, "no device-resident factorization"};
  GPRY_CUDA(cudaSetDevice(st->device));
  cudaStream_t s = 0;
  const int N = st->f_N, Np = round_up(N, NB);
  const int64_t tot = (int64_t)N * N;
  st->tmp.reserve((size_t)tot);
  if (out_L) {
    extract_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(st->f_K.p, Np, N, 0, st->tmp.p);
    GPRY_CUDA(cudaGetLastError());
    GPRY_CUDA(cudaMemcpyAsync(out_L, st->tmp.p, tot * 8, cudaMemcpyDeviceToHost, s));
  }
  if (out_V) {
    extract_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(st->f_VT.p, Np, N, 1, st->tmp.p);
    GPRY_CUDA(cudaGetLastError());
    GPRY_CUDA(cudaMemcpyAsync(out_V, st->tmp.p, tot * 8, cudaMemcpyDeviceToHost, s));
  }
  GPRY_CUDA(cudaStreamSynchronize(s));
}


// ---------------------------------------------------------------------------------------
// Bordered append (gpr.py:996-1020 with unchanged theta / noise / preprocessing:
// the Kriging-believer lies of gp_acquisition.py:488-491).  The factorisation kept resident by
// factorize_device(keep) is extended by k points, one row at a time, in O(k N^2):
//   kvec = k(x_new, X[0..m)),  l = V kvec,  lambda^2 = k(x,x) + noise2 - |l|^2,
//   L' = [[L, 0], [l^T, lambda]],  V' = [[V, 0], [-(V^T l)^T / lambda, 1 / lambda]]
// then alpha_ = V'^T (V' y_).  Same N_pad only (the caller refactorises when the padded size
// changes); not positive definite -> info = failing order, the resident factor is dropped.
// ---------------------------------------------------------------------------------------
template <int KIND>
__global__ void append_krow_kernel(const double* __restrict__ T, int d, int m, double c,
                                   double diag, double* __restrict__ krow) {
  // krow[j] = c g(|T_m - T_j|) for j < m, krow[m] = diag
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j > m) return;
  if (j == m) {
    krow[m] = diag;
    return;
  }
  double r2 = 0.0;
  for (int q = 0; q < d; q++) {
    double df = T[(size_t)m * d + q] - T[(size_t)j * d + q];
    r2 = fma(df, df, r2);
  }
  krow[j] = c * stationary_value<KIND>(r2);
}
// Lrow (= row m of K on entry) <- [l, lambda]; scal[0] = lambda, info = m + 1 if not PD
__global__ void __launch_bounds__(256)
append_pivot_kernel(const double* __restrict__ l, int m, double* __restrict__ Lrow,
                    double* __restrict__ scal, int* __restrict__ info) {
  __shared__ double red[256];
  double s = 0.0;
  for (int j = threadIdx.x; j < m; j += 256) s = fma(l[j], l[j], s);
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  const double lam2 = Lrow[m] - red[0];
  __syncthreads();
  const bool ok = lam2 > 0.0 && *info == 0;
  const double lam = ok ? sqrt(lam2) : 1.0;
  for (int j = threadIdx.x; j < m; j += 256) Lrow[j] = l[j];
  if (threadIdx.x == 0) {
    Lrow[m] = lam;
    scal[0] = lam;
    if (!ok && *info == 0) *info = m + 1;
  }
}
// column m of VT (= row m of V'): VT[j][m] = -z[j] / lambda (j < m), VT[m][m] = 1 / lambda
__global__ void append_vcol_kernel(const double* __restrict__ z, int m, int Np,
                                   const double* __restrict__ scal, double* __restrict__ VT) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j > m) return;
  const double inv = 1.0 / scal[0];
  VT[(size_t)j * Np + m] = j == m ? inv : -z[j] * inv;
}

int factor_append_device(gpry_state* st, int k, const double* X_new_t, const double* noise2_new,
                         const double* y_all, const double* theta, double* out_alpha) {
  if (!st->f_valid) throw GpryError{GPRY_ERR_STATE, "no device-resident factorization"};
  GPRY_CUDA(cudaSetDevice(st->device));
  const int N = st->f_N, d = st->f_d, kind = st->f_kind, Np = round_up(N, NB), N2 = N + k;
  GPRY_CHECK_ARG(k >= 1 && round_up(N2, NB) == Np, "append: the padded size must not change");
  cudaStream_t s = 0;
  st->f_valid = false;
  TrainBuffers b = carve(st, N, d, false, 1);       // same pointers: sizes depend on Np only
  const double c = exp(theta[0]);
  std::vector<double> Tn((size_t)k * d);
  for (int i = 0; i < k; i++)
    for (int q = 0; q < d; q++) Tn[(size_t)i * d + q] = X_new_t[(size_t)i * d + q] / exp(theta[1 + q]);
  GPRY_CUDA(cudaMemcpyAsync(b.X + (size_t)N * d, X_new_t, (size_t)k * d * 8, cudaMemcpyHostToDevice, s));
  GPRY_CUDA(cudaMemcpyAsync(b.T + (size_t)N * d, Tn.data(), (size_t)k * d * 8, cudaMemcpyHostToDevice, s));
  GPRY_CUDA(cudaMemcpyAsync(b.noise2 + N, noise2_new, (size_t)k * 8, cudaMemcpyHostToDevice, s));
  GPRY_CUDA(cudaMemcpyAsync(b.y, y_all, (size_t)N2 * 8, cudaMemcpyHostToDevice, s));
  GPRY_CUDA(cudaMemsetAsync(b.info, 0, sizeof(int), s));
  st->tmp.reserve(2 * (size_t)Np);
  double* l = st->tmp.p;
  double* z = st->tmp.p + Np;
  for (int i = 0; i < k; i++) {
    const int m = N + i;
    double* Lrow = b.K + (size_t)m * Np;
    const double diag = c + noise2_new[i];
    const int nblk = (m + 1 + 255) / 256;
    switch (kind) {
      case GPRY_KERNEL_RBF:
        append_krow_kernel<GPRY_KERNEL_RBF><<<nblk, 256, 0, s>>>(b.T, d, m, c, diag, Lrow);
        break;
      case GPRY_KERNEL_MATERN15:
        append_krow_kernel<GPRY_KERNEL_MATERN15><<<nblk, 256, 0, s>>>(b.T, d, m, c, diag, Lrow);
        break;
      default:
        append_krow_kernel<GPRY_KERNEL_MATERN25><<<nblk, 256, 0, s>>>(b.T, d, m, c, diag, Lrow);
    }
    gemv_vt_t_kernel<<<dim3(Np / 32, 1), 256, 0, s>>>(b.VT, Np, m, Lrow, l);          // l = V kvec
    append_pivot_kernel<<<1, 256, 0, s>>>(l, m, Lrow, b.scal, b.info);
    gemv_vt_kernel<<<dim3((Np * 32 + 255) / 256, 1), 256, 0, s>>>(b.VT, Np, m, l, z);  // z = V^T l
    append_vcol_kernel<<<nblk, 256, 0, s>>>(z, m, Np, b.scal, b.VT);
    GPRY_CUDA(cudaGetLastError());
  }
  gemv_vt_t_kernel<<<dim3(Np / 32, 1), 256, 0, s>>>(b.VT, Np, N2, b.y, b.t);
  gemv_vt_kernel<<<dim3((Np * 32 + 255) / 256, 1), 256, 0, s>>>(b.VT, Np, N2, b.t, b.alpha);
  GPRY_CUDA(cudaGetLastError());
  int h_info = 0;
  GPRY_CUDA(cudaMemcpyAsync(&h_info, b.info, sizeof(int), cudaMemcpyDeviceToHost, s));
  if (out_alpha)
    GPRY_CUDA(cudaMemcpyAsync(out_alpha, b.alpha, (size_t)N2 * 8, cudaMemcpyDeviceToHost, s));
  GPRY_CUDA(cudaStreamSynchronize(s));
  if (h_info == 0) {
    st->f_valid = true;
    st->f_N = N2;
  }
  return h_info;
}

void lml_batched_device(gpry_state* st, int kind, int N, int d, const double* X_train_t,
                        const double* noise2, const double* y_t, const double* thetas, int B,
                        double* out_lml, double* out_grad, int* out_info) {
  GPRY_CHECK_ARG(kind >= 0 && kind <= 2, "unknown kernel kind");
  GPRY_CHECK_ARG(N >= 1 && d >= 1 && d <= MAX_DIM, "need N >= 1 and 1 <= d <= 128");
  GPRY_CUDA(cudaSetDevice(st->device));
  st->f_valid = false;
  cudaStream_t s = 0;
  const bool need_grad = out_grad != nullptr;
  const int P = d + 1;
  const int Bs = sub_batch(B, round_up(N, NB), need_grad);
  TrainBuffers b = carve(st, N, d, need_grad, Bs);
  upload_problem(b, X_train_t, noise2, y_t, s);
  std::vector<int> h_info(Bs);
  std::vector<double> h_scal(2 * (size_t)Bs), h_grad((size_t)(MAX_DIM + 8) * Bs);
  for (int i0 = 0; i0 < B; i0 += Bs) {
    const int nth = std::min(Bs, B - i0);
    factorize_batch(st, b, kind, thetas + (size_t)i0 * P, nth, need_grad, s);
    GPRY_CUDA(cudaMemcpyAsync(h_info.data(), b.info, (size_t)nth * sizeof(int),
                              cudaMemcpyDeviceToHost, s));
    GPRY_CUDA(cudaMemcpyAsync(h_scal.data(), b.scal, (size_t)nth * 16, cudaMemcpyDeviceToHost, s));
    if (need_grad)
      GPRY_CUDA(cudaMemcpyAsync(h_grad.data(), b.grad, (size_t)nth * (MAX_DIM + 8) * 8,
                                cudaMemcpyDeviceToHost, s));
    GPRY_CUDA(cudaStreamSynchronize(s));
    for (int i = 0; i < nth; i++) {
      const int gi = i0 + i;
      out_info[gi] = h_info[i];
      if (h_info[i] != 0) {   // sklearn:_gpr.py:592-593
        out_lml[gi] = -INFINITY;
        if (need_grad)
          for (int p = 0; p < P; p++) out_grad[(size_t)gi * P + p] = 0.0;
        continue;
      }
      // -0.5 y^T alpha - sum(log diag L) - N/2 log(2 pi)        (sklearn:_gpr.py:613-617)
      double lml = -0.5 * h_scal[2 * i + 1];
      lml -= h_scal[2 * i];
      lml -= N / 2.0 * log(2.0 * M_PI);
      out_lml[gi] = lml;
      if (need_grad)
        for (int p = 0; p < P; p++) out_grad[(size_t)gi * P + p] = h_grad[(size_t)i * (MAX_DIM + 8) + p];
    }
  }
}

}  // namespace gpry
