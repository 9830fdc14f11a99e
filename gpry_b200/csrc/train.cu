// Training-side hot path: kernel matrix, blocked FP64 Cholesky, L^-1, alpha, log marginal
// likelihood and its gradient.
//
// Reference arithmetic: gpr.py:1015-1017, 1453-1465 (_update_model / _kernel_inverse);
// sklearn:_gpr.py:584-651 (LML + gradient); kernel values and theta-gradients
// sklearn:kernels.py:1561-1584 (RBF), 1716-1771 (Matern), 964-969 (Product), 1283-1292
// (Constant).
//
// Everything works on matrices padded to Np = round_up(N, 128) with an identity block in the
// padding (unit diagonal, zero coupling), so every tile is full and the padded rows leave
// L, L^-1, alpha and log det untouched.
//
//   kmat_kernel        K = c g(r) + diag(noise2)                       (lower tiles)
//   potf2_inv_kernel   128 x 128 diagonal block: L_jj and W_jj = L_jj^-1 (one CTA: 32 x 32
//                      sub-blocks factored in registers by one warp, 32^3 products on DMMA)
//   gemm_nt_kernel     C (+)= alpha A B^T on FP64 tensor cores (DMMA.8x8x4), cp.async
//                      4-stage pipeline; used for the panel solve (x W_jj^T), the SYRK
//                      trailing update, the right-looking sweep that builds V^T = L^-T and
//                      K^-1 = V^T V (block-triangular k ranges skip the structural zeros)
//   lml_grad_kernel    fused trace contraction 1/2 sum_ij (a_i a_j - K^-1_ij) dK_ij/dtheta:
//                      kernel values and per-dimension distances are recomputed per pair,
//                      dK/dtheta (N x N x (1+d)) is never materialised
#include <math.h>

#include <vector>

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

// T = X_ / ell (row major N x d).  32 x 32 tile per CTA (lower tiles incl. diagonal).
template <int KIND>
__global__ void __launch_bounds__(256)
kmat_kernel(const double* __restrict__ T, int N, int d, int Np, double c,
            const double* __restrict__ noise2, double* __restrict__ K) {
  const int bi = blockIdx.y, bj = blockIdx.x;
  if (bj > bi) return;
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
// diagonal block: 128 x 128 Cholesky + triangular inverse in ONE CTA (8 warps).
//   A: the diagonal block (leading dimension ld), overwritten by L_jj (upper triangle zeroed)
//   W: dense 128 x 128 = L_jj^-1 (upper zeroed)
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

// One warp: Cholesky of the 32 x 32 block D (lane = row, row in registers, right-looking) and
// its inverse (lane = column, forward substitution); writes L (upper zeroed) back into D and
// the inverse into Winv.
__device__ __forceinline__ void factor_inv_32(double* __restrict__ D, double* __restrict__ Winv,
                                              int lane, int pivot_offset, int* __restrict__ info) {
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
    const double sq = sqrt(dkk);
    const double lk = (lane == k) ? sq : a[k] / sq;
    a[k] = lk;
#pragma unroll
    for (int j = k + 1; j < SB; j++) {
      double ljk = __shfl_sync(0xffffffffu, lk, j);
      if (lane >= j) a[j] = fma(-lk, ljk, a[j]);
    }
  }
  // inverse: lane c holds column c of W
  double w[SB];
#pragma unroll
  for (int i = 0; i < SB; i++) {
    double s = (i == lane) ? 1.0 : 0.0;
#pragma unroll
    for (int m = 0; m < i; m++) {
      double lim = __shfl_sync(0xffffffffu, a[m], i);
      s = fma(-lim, w[m], s);
    }
    double lii = __shfl_sync(0xffffffffu, a[i], i);
    w[i] = (i >= lane) ? s / lii : 0.0;
  }
#pragma unroll
  for (int k = 0; k < SB; k++) {
    D[lane * SBLD + k] = (k <= lane) ? a[k] : 0.0;
    Winv[k * SBLD + lane] = w[k];
  }
}

__global__ void __launch_bounds__(256)
potf2_inv_kernel(double* __restrict__ A, int ld, double* __restrict__ W, int row_offset,
                 int* __restrict__ info) {
  extern __shared__ double sh[];
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
      factor_inv_32(Lb + sbidx(kb, kb) * SB_DOUBLES, Wb + sbidx(kb, kb) * SB_DOUBLES, lane,
                    row_offset + kb * SB, info);
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
  }
}

// ---------------------------------------------------------------------------------------
// FP64 tensor-core GEMM, C (+)= alpha * A * B^T   (A: M x K, B: N x K, both k-contiguous)
// ---------------------------------------------------------------------------------------
constexpr int G_STAGES = 4;
constexpr int G_THREADS = 256;
enum { G_FILTER_ALL = 0, G_FILTER_LOWER = 1 };
enum { G_KLO_ZERO = 0, G_KLO_ROW = 1 };

struct GemmArgs {
  const double* A;
  const double* B;
  double* C;
  int lda, ldb, ldc;
  int M, N, K;          // multiples of 128, 128, 16
  double alpha;
  int accumulate;       // 0: C = alpha A B^T ; 1: C += alpha A B^T
  int filter;           // G_FILTER_LOWER: only tiles with tile_row >= tile_col
  int klo_mode;         // G_KLO_ROW: k starts at the tile's first row (A upper triangular)
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
  const int mt = blockIdx.y, nt = blockIdx.x;
  if (g.filter == G_FILTER_LOWER && mt < nt) return;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  GemmSmem& sm = *reinterpret_cast<GemmSmem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = mt * TILE_ROWS, n0 = nt * TILE_ROWS;
  const int k_lo = (g.klo_mode == G_KLO_ROW) ? m0 : 0;
  const int nk = (g.K - k_lo) / TILE_K;

  const double* Ag = g.A + (size_t)m0 * g.lda + k_lo;
  const double* Bg = g.B + (size_t)n0 * g.ldb + k_lo;
  auto load_stage = [&](int s, int kt) {
    const int k0 = kt * TILE_K;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      int q = tid + G_THREADS * i;        // 0..1023: 16-byte chunk id
      int row = q >> 3, kc = (q & 7) * 2;
      cp_async16(&sm.A[s][tile_elem(row, kc)], Ag + (size_t)row * g.lda + k0 + kc);
      cp_async16(&sm.B[s][tile_elem(row, kc)], Bg + (size_t)row * g.ldb + k0 + kc);
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
  // epilogue
  const int g8 = lane >> 2, t4 = lane & 3;
#pragma unroll
  for (int mi = 0; mi < 8; mi++)
#pragma unroll
    for (int ni = 0; ni < 4; ni++) {
      int row = m0 + rw * 64 + mi * 8 + g8;
      int col = n0 + cw * 32 + ni * 8 + 2 * t4;
      double2* cp = reinterpret_cast<double2*>(g.C + (size_t)row * g.ldc + col);
      double2 v = make_double2(g.alpha * acc[mi][ni][0], g.alpha * acc[mi][ni][1]);
      if (g.accumulate) {
        double2 o = *cp;
        v.x += o.x;
        v.y += o.y;
      }
      *cp = v;
    }
}

static void launch_gemm(const GemmArgs& g, cudaStream_t s) {
  if (g.M <= 0 || g.N <= 0 || g.K <= 0) return;
  dim3 grid(g.N / TILE_ROWS, g.M / TILE_ROWS);
  gemm_nt_kernel<<<grid, G_THREADS, sizeof(GemmSmem), s>>>(g);
  GPRY_CUDA(cudaGetLastError());
}

void gemm_nt(const double* A, int lda, const double* B, int ldb, double* C, int ldc, int M, int N,
             int K, double alpha, int accumulate, int lower_only, int klo_row, cudaStream_t s) {
  GemmArgs g{};
  g.A = A; g.lda = lda; g.B = B; g.ldb = ldb; g.C = C; g.ldc = ldc;
  g.M = M; g.N = N; g.K = K; g.alpha = alpha; g.accumulate = accumulate;
  g.filter = lower_only ? G_FILTER_LOWER : G_FILTER_ALL;
  g.klo_mode = klo_row ? G_KLO_ROW : G_KLO_ZERO;
  GPRY_CUDA(cudaFuncSetAttribute(gemm_nt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)sizeof(GemmSmem)));
  launch_gemm(g, s);
}

// ---------------------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------------------
__global__ void set_identity_kernel(double* __restrict__ A, int Np) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (int64_t)Np * Np) return;
  int r = (int)(e / Np), cidx = (int)(e % Np);
  A[e] = r == cidx ? 1.0 : 0.0;
}
__global__ void scale_rows_kernel(const double* __restrict__ X, int N, int d,
                                  const double* __restrict__ ell, double* __restrict__ T) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < N * d) T[e] = X[e] / ell[e % d];
}
// t = V y = VT^T y : CTA per 32 columns of VT, 8 row groups, rows j <= i; fixed-order sums
__global__ void __launch_bounds__(256)
gemv_vt_t_kernel(const double* __restrict__ VT, int Np, int N, const double* __restrict__ y,
                 double* __restrict__ t) {
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
__global__ void gemv_vt_kernel(const double* __restrict__ VT, int Np, int N,
                               const double* __restrict__ t, double* __restrict__ alpha) {
  int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (j >= Np) return;
  double s = 0.0;
  if (j < N)
    for (int i = j + lane; i < N; i += 32) s = fma(VT[(size_t)j * Np + i], t[i], s);
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) alpha[j] = s;
}
// out[0] = sum_i log L_ii, out[1] = y . alpha      (single block, fixed order)
__global__ void lml_scalars_kernel(const double* __restrict__ L, int Np, int N,
                                   const double* __restrict__ y, const double* __restrict__ alpha,
                                   double* __restrict__ out) {
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
__global__ void transpose_block_kernel(const double* __restrict__ W, double* __restrict__ dst,
                                       int ld) {
  // dst[a][b] = W[b][a], 128 x 128
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= 128 * 128) return;
  int a = e >> 7, b = e & 127;
  dst[(size_t)a * ld + b] = W[b * 128 + a];
}

// ---------------------------------------------------------------------------------------
// fused LML-gradient trace contraction
//   grad_p = 1/2 sum_ij (alpha_i alpha_j - Kinv_ij) dK_ij / dtheta_p   (sklearn:_gpr.py:647)
// 32 x 32 tile of pairs per CTA over the lower triangle; off-diagonal tiles count twice.
// partial[(tile)][p] written per CTA, reduced in fixed order by reduce_partials_kernel.
// ---------------------------------------------------------------------------------------
template <int KIND>
__global__ void __launch_bounds__(256)
lml_grad_kernel(const double* __restrict__ X, const double* __restrict__ T, int N, int d, int Np,
                double c, const double* __restrict__ ell, const double* __restrict__ alpha,
                const double* __restrict__ Kinv, double* __restrict__ partial, int P) {
  const int bi = blockIdx.y, bj = blockIdx.x;
  const int nb = gridDim.x;
  double* out = partial + ((size_t)bi * nb + bj) * P;
  extern __shared__ double sh[];
  double* Xi = sh;                      // [32][d+1]
  double* Xj = Xi + 32 * (d + 1);       // [32][d+1]
  double* Ti = Xj + 32 * (d + 1);       // [32][d+1]  X / ell
  double* Tj = Ti + 32 * (d + 1);       // [32][d+1]
  double* inv_l2 = Tj + 32 * (d + 1);   // [d]  ell^2 (divided by, as the reference does)
  double* red = inv_l2 + d;             // [8 warps][P]
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
  for (int k = tid; k < d; k += 256) inv_l2[k] = ell[k] * ell[k];   // length_scale**2
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
          double D = df * df / inv_l2[k];         // (xi - xj)**2 / length_scale**2
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
  // reduce g0 over the block
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
      s = fma(wk[q], df * df / inv_l2[k], s);
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

__global__ void reduce_partials_kernel(const double* __restrict__ partial, int n, int P,
                                       double* __restrict__ out) {
  // one block per p; fixed-order tree
  __shared__ double r[256];
  const int p = blockIdx.x;
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) s += partial[(size_t)i * P + p];
  r[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) r[threadIdx.x] += r[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[p] = r[0];
}

// ---------------------------------------------------------------------------------------
// host orchestration
// ---------------------------------------------------------------------------------------
struct TrainBuffers {
  int N, d, Np, nb;
  double *K, *VT, *W;         // Np x Np each (W: K^-1, only for the gradient)
  double *Winv;               // nb x 128 x 128
  double *X, *T, *noise2, *y, *ell, *alpha, *t, *scal, *partial, *grad;
  int* info;
  cudaStream_t stream;
};

constexpr int MAX_TRAIN_STREAMS = 4;

// shared problem: st->f_prob = [y Np][noise2 Np][X_ N*d][ell of each theta: B x MAX_DIM]
// per-set vec:    [alpha Np][t Np][T N*d][ell MAX_DIM][scal 8][grad MAX_DIM+8][info 2]
//                 [Winv nb*128*128][partial nb32*nb32*P]
static TrainBuffers carve(gpry_state* st, int set, int N, int d, bool need_grad, int B) {
  TrainBuffers b;
  b.N = N;
  b.d = d;
  b.Np = round_up(N, NB);
  b.nb = b.Np / NB;
  const size_t Np = b.Np;
  while ((int)st->f_sets.size() <= set) {
    TrainSet* ts = new TrainSet();
    GPRY_CUDA(cudaStreamCreateWithFlags(&ts->stream, cudaStreamNonBlocking));
    st->f_sets.push_back(ts);
  }
  TrainSet* ts = st->f_sets[set];
  ts->K.reserve(Np * Np);
  ts->VT.reserve(Np * Np);
  if (need_grad) ts->W.reserve(Np * Np);
  const int nb32 = (N + 31) / 32;
  const int P = d + 1;
  auto al = [](size_t x) { return (x + 31) / 32 * 32; };   // 256-byte aligned segments
  const size_t nXd = al((size_t)N * d);
  size_t n = 2 * Np + nXd + al(MAX_DIM) + 32 + al(MAX_DIM + 8) + 32 + (size_t)b.nb * NB * NB +
             (need_grad ? (size_t)nb32 * nb32 * P : 0) + 64;
  ts->vec.reserve(n);
  st->f_prob.reserve(2 * Np + nXd + (size_t)B * MAX_DIM);
  b.y = st->f_prob.p;
  b.noise2 = b.y + Np;
  b.X = b.noise2 + Np;
  double* p = ts->vec.p;
  b.alpha = p; p += Np;
  b.t = p; p += Np;
  b.T = p; p += nXd;
  b.ell = p; p += al(MAX_DIM);
  b.scal = p; p += 32;
  b.grad = p; p += al(MAX_DIM + 8);
  b.info = reinterpret_cast<int*>(p); p += 32;
  b.Winv = p; p += (size_t)b.nb * NB * NB;
  b.partial = p;
  b.K = ts->K.p;
  b.VT = ts->VT.p;
  b.W = need_grad ? ts->W.p : nullptr;
  b.stream = ts->stream;
  return b;
}

static void upload_problem(gpry_state* st, TrainBuffers& b, const double* X, const double* noise2,
                           const double* y, cudaStream_t s) {
  const size_t Np = b.Np;
  GPRY_CUDA(cudaMemsetAsync(b.y, 0, 2 * Np * 8, s));
  GPRY_CUDA(cudaMemcpyAsync(b.X, X, (size_t)b.N * b.d * 8, cudaMemcpyHostToDevice, s));
  GPRY_CUDA(cudaMemcpyAsync(b.noise2, noise2, (size_t)b.N * 8, cudaMemcpyHostToDevice, s));
  GPRY_CUDA(cudaMemcpyAsync(b.y, y, (size_t)b.N * 8, cudaMemcpyHostToDevice, s));
}

// exp(theta[1:]) of all B evaluations -> device (one pageable copy, before any kernel runs)
static double* upload_ells(gpry_state* st, const TrainBuffers& b, const double* thetas, int B,
                           cudaStream_t s) {
  const int P = b.d + 1;
  std::vector<double> ells((size_t)B * MAX_DIM, 1.0);
  for (int i = 0; i < B; i++)
    for (int k = 0; k < b.d; k++) ells[(size_t)i * MAX_DIM + k] = exp(thetas[(size_t)i * P + 1 + k]);
  double* dst = b.X + ((size_t)b.N * b.d + 31) / 32 * 32;
  GPRY_CUDA(cudaMemcpyAsync(dst, ells.data(), ells.size() * 8, cudaMemcpyHostToDevice, s));
  return dst;
}

template <int KIND>
static void launch_kmat(const TrainBuffers& b, double c, cudaStream_t s) {
  const int nb32 = b.Np / 32;
  size_t smem = 2 * 32 * (size_t)(b.d + 1) * 8;
  kmat_kernel<KIND><<<dim3(nb32, nb32), 256, smem, s>>>(b.T, b.N, b.d, b.Np, c, b.noise2, b.K);
  GPRY_CUDA(cudaGetLastError());
}

template <int KIND>
static void launch_grad(const TrainBuffers& b, double c, int nb32, cudaStream_t s) {
  const int P = b.d + 1;
  size_t smem = (4 * 32 * (size_t)(b.d + 1) + b.d + 8 * (size_t)P) * 8;
  lml_grad_kernel<KIND><<<dim3(nb32, nb32), 256, smem, s>>>(b.X, b.T, b.N, b.d, b.Np, c, b.ell, b.alpha,
                                                            b.W, b.partial, P);
  GPRY_CUDA(cudaGetLastError());
}

// K(theta) -> L (in K), VT = L^-T, alpha, scal = {sum log diag L, y.alpha}; optionally
// W = K^-1 and grad (device).  Returns nothing; info stays on the device.
// b.ell must already point at the device copy of exp(theta[1:]) for this evaluation.
static void factorize_on_device(gpry_state* st, TrainBuffers& b, int kind, const double* theta,
                                bool need_grad, cudaStream_t s) {
  const int N = b.N, d = b.d, Np = b.Np, nb = b.nb;
  const double c = exp(theta[0]);
  GPRY_CUDA(cudaMemsetAsync(b.info, 0, 8, s));
  scale_rows_kernel<<<(N * d + 255) / 256, 256, 0, s>>>(b.X, N, d, b.ell, b.T);
  GPRY_CUDA(cudaGetLastError());
  switch (kind) {
    case GPRY_KERNEL_RBF: launch_kmat<GPRY_KERNEL_RBF>(b, c, s); break;
    case GPRY_KERNEL_MATERN15: launch_kmat<GPRY_KERNEL_MATERN15>(b, c, s); break;
    default: launch_kmat<GPRY_KERNEL_MATERN25>(b, c, s);
  }
  {
    int64_t tot = (int64_t)Np * Np;
    set_identity_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(b.VT, Np);
    GPRY_CUDA(cudaGetLastError());
  }
  GPRY_CUDA(cudaFuncSetAttribute(gemm_nt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)sizeof(GemmSmem)));
  const size_t potf2_smem = 20 * (size_t)SB_DOUBLES * 8;
  GPRY_CUDA(cudaFuncSetAttribute(potf2_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)potf2_smem));
  for (int j = 0; j < nb; j++) {
    double* Ajj = b.K + (size_t)j * NB * (Np + 1);
    double* Wj = b.Winv + (size_t)j * NB * NB;
    potf2_inv_kernel<<<1, 256, potf2_smem, s>>>(Ajj, Np, Wj, j * NB, b.info);
    GPRY_CUDA(cudaGetLastError());
    const int rem = Np - (j + 1) * NB;
    if (rem > 0) {
      // panel: L[j+1:, j] = A[j+1:, j] W_jj^T        (in place)
      GemmArgs g{};
      g.A = b.K + (size_t)(j + 1) * NB * Np + (size_t)j * NB;
      g.lda = Np;
      g.B = Wj;
      g.ldb = NB;
      g.C = const_cast<double*>(g.A);
      g.ldc = Np;
      g.M = rem; g.N = NB; g.K = NB;
      g.alpha = 1.0; g.accumulate = 0; g.filter = G_FILTER_ALL; g.klo_mode = G_KLO_ZERO;
      launch_gemm(g, s);
      // trailing update: A[j+1:, j+1:] -= P P^T      (lower tiles)
      GemmArgs u{};
      u.A = g.A; u.lda = Np; u.B = g.A; u.ldb = Np;
      u.C = b.K + (size_t)(j + 1) * NB * (Np + 1);
      u.ldc = Np;
      u.M = rem; u.N = rem; u.K = NB;
      u.alpha = -1.0; u.accumulate = 1; u.filter = G_FILTER_LOWER; u.klo_mode = G_KLO_ZERO;
      launch_gemm(u, s);
    }
    // right-looking sweep for VT = L^-T on the (identity-initialised) array:
    //   VT[0:(j+1)B, jblock] = VT[0:(j+1)B, jblock] W_jj^T          (in place)
    //   VT[0:(j+1)B, r > j] -= VT[0:(j+1)B, jblock] L[r, jblock]^T
    {
      GemmArgs g{};
      g.A = b.VT + (size_t)j * NB;
      g.lda = Np;
      g.B = Wj;
      g.ldb = NB;
      g.C = b.VT + (size_t)j * NB;
      g.ldc = Np;
      g.M = (j + 1) * NB; g.N = NB; g.K = NB;
      g.alpha = 1.0; g.accumulate = 0; g.filter = G_FILTER_ALL; g.klo_mode = G_KLO_ZERO;
      launch_gemm(g, s);
      if (rem > 0) {
        GemmArgs u{};
        u.A = b.VT + (size_t)j * NB; u.lda = Np;
        u.B = b.K + (size_t)(j + 1) * NB * Np + (size_t)j * NB; u.ldb = Np;
        u.C = b.VT + (size_t)(j + 1) * NB; u.ldc = Np;
        u.M = (j + 1) * NB; u.N = rem; u.K = NB;
        u.alpha = -1.0; u.accumulate = 1; u.filter = G_FILTER_ALL; u.klo_mode = G_KLO_ZERO;
        launch_gemm(u, s);
      }
    }
  }
  // alpha = V^T (V y)
  gemv_vt_t_kernel<<<Np / 32, 256, 0, s>>>(b.VT, Np, N, b.y, b.t);
  GPRY_CUDA(cudaGetLastError());
  gemv_vt_kernel<<<(Np * 32 + 255) / 256, 256, 0, s>>>(b.VT, Np, N, b.t, b.alpha);
  GPRY_CUDA(cudaGetLastError());
  lml_scalars_kernel<<<1, 256, 0, s>>>(b.K, Np, N, b.y, b.alpha, b.scal);
  GPRY_CUDA(cudaGetLastError());
  if (need_grad) {
    // K^-1 = V^T V = VT VT^T (lower tiles; VT upper triangular: k >= tile row)
    GemmArgs g{};
    g.A = b.VT; g.lda = Np; g.B = b.VT; g.ldb = Np; g.C = b.W; g.ldc = Np;
    g.M = Np; g.N = Np; g.K = Np;
    g.alpha = 1.0; g.accumulate = 0; g.filter = G_FILTER_LOWER; g.klo_mode = G_KLO_ROW;
    launch_gemm(g, s);
    const int nb32 = (N + 31) / 32;
    switch (kind) {
      case GPRY_KERNEL_RBF: launch_grad<GPRY_KERNEL_RBF>(b, c, nb32, s); break;
      case GPRY_KERNEL_MATERN15: launch_grad<GPRY_KERNEL_MATERN15>(b, c, nb32, s); break;
      default: launch_grad<GPRY_KERNEL_MATERN25>(b, c, nb32, s);
    }
    reduce_partials_kernel<<<d + 1, 256, 0, s>>>(b.partial, nb32 * nb32, d + 1, b.grad);
    GPRY_CUDA(cudaGetLastError());
  }
}

void factorize_device(gpry_state* st, int kind, int N, int d, const double* X_train_t,
                      const double* noise2, const double* y_t, const double* theta, double* out_L,
                      double* out_V, double* out_alpha, double* out_logdet_half, int* info,
                      bool keep) {
  GPRY_CHECK_ARG(kind >= 0 && kind <= 2, "unknown kernel kind");
  GPRY_CHECK_ARG(N >= 1 && d >= 1 && d <= MAX_DIM, "need N >= 1 and 1 <= d <= 128");
  GPRY_CUDA(cudaSetDevice(st->device));
  st->f_valid = false;
  TrainBuffers b = carve(st, 0, N, d, false, 1);
  cudaStream_t s = b.stream;
  upload_problem(st, b, X_train_t, noise2, y_t, s);
  b.ell = upload_ells(st, b, theta, 1, s);
  factorize_on_device(st, b, kind, theta, false, s);
  int h_info[2] = {0, 0};
  double scal[2];
  GPRY_CUDA(cudaMemcpyAsync(h_info, b.info, 8, cudaMemcpyDeviceToHost, s));
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
  *info = h_info[0];
  if (out_logdet_half) *out_logdet_half = scal[0];
  if (keep && h_info[0] == 0) {
    st->f_valid = true;
    st->f_N = N;
    st->f_d = d;
    st->f_kind = kind;
  }
}

// B hyper-parameter vectors are evaluated round-robin on up to MAX_TRAIN_STREAMS streams with
// private work sets, so that the latency-bound diagonal-block kernels of one evaluation
// overlap the GEMMs of the others.  Per-theta results land in pinned host memory.
void lml_batched_device(gpry_state* st, int kind, int N, int d, const double* X_train_t,
                        const double* noise2, const double* y_t, const double* thetas, int B,
                        double* out_lml, double* out_grad, int* out_info) {
  GPRY_CHECK_ARG(kind >= 0 && kind <= 2, "unknown kernel kind");
  GPRY_CHECK_ARG(N >= 1 && d >= 1 && d <= MAX_DIM, "need N >= 1 and 1 <= d <= 128");
  GPRY_CUDA(cudaSetDevice(st->device));
  st->f_valid = false;
  const bool need_grad = out_grad != nullptr;
  const int P = d + 1;
  const int nset = std::min(B, MAX_TRAIN_STREAMS);
  std::vector<TrainBuffers> sets;
  for (int i = 0; i < nset; i++) sets.push_back(carve(st, i, N, d, need_grad, B));
  // f_prob may have been reallocated by a later carve: refresh the shared pointers
  for (auto& b : sets) {
    b.y = st->f_prob.p;
    b.noise2 = b.y + b.Np;
    b.X = b.noise2 + b.Np;
  }
  const size_t rec = (size_t)P + 4;   // per theta: [info][logdet_half][y.alpha][pad][grad P]
  if (st->f_pinned_cap < rec * B) {
    if (st->f_pinned) cudaFreeHost(st->f_pinned);
    st->f_pinned = nullptr;
    GPRY_CUDA(cudaMallocHost((void**)&st->f_pinned, rec * B * sizeof(double)));
    st->f_pinned_cap = rec * B;
  }
  if (!st->f_evt) GPRY_CUDA(cudaEventCreateWithFlags(&st->f_evt, cudaEventDisableTiming));
  upload_problem(st, sets[0], X_train_t, noise2, y_t, sets[0].stream);
  double* ells = upload_ells(st, sets[0], thetas, B, sets[0].stream);
  GPRY_CUDA(cudaEventRecord(st->f_evt, sets[0].stream));
  for (int i = 1; i < nset; i++) GPRY_CUDA(cudaStreamWaitEvent(sets[i].stream, st->f_evt, 0));
  for (int i = 0; i < B; i++) {
    TrainBuffers& b = sets[i % nset];
    cudaStream_t s = b.stream;
    b.ell = ells + (size_t)i * MAX_DIM;
    factorize_on_device(st, b, kind, thetas + (size_t)i * P, need_grad, s);
    double* r = st->f_pinned + rec * i;
    GPRY_CUDA(cudaMemcpyAsync(r, b.info, 8, cudaMemcpyDeviceToHost, s));
    GPRY_CUDA(cudaMemcpyAsync(r + 1, b.scal, 16, cudaMemcpyDeviceToHost, s));
    if (need_grad) GPRY_CUDA(cudaMemcpyAsync(r + 4, b.grad, P * 8, cudaMemcpyDeviceToHost, s));
  }
  for (int i = 0; i < nset; i++) GPRY_CUDA(cudaStreamSynchronize(sets[i].stream));
  for (int i = 0; i < B; i++) {
    const double* r = st->f_pinned + rec * i;
    const int h_info = *reinterpret_cast<const int*>(r);
    out_info[i] = h_info;
    if (h_info != 0) {   // sklearn:_gpr.py:592-593
      out_lml[i] = -INFINITY;
      if (need_grad)
        for (int p = 0; p < P; p++) out_grad[(size_t)i * P + p] = 0.0;
      continue;
    }
    // -0.5 y^T alpha - sum(log diag L) - N/2 log(2 pi)        (sklearn:_gpr.py:613-617)
    double lml = -0.5 * r[2];
    lml -= r[1];
    lml -= N / 2.0 * log(2.0 * M_PI);
    out_lml[i] = lml;
    if (need_grad)
      for (int p = 0; p < P; p++) out_grad[(size_t)i * P + p] = r[4 + p];
  }
}

}  // namespace gpry
