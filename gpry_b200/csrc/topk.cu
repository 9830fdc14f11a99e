// Ranked-pool pre-selection: the K' candidates with the largest acquisition value, in the
// order RankedPool.add(method="single sort acq") visits them (descending acq,
// gp_acquisition.py:1326-1333).  Exact, deterministic (ties: ascending index; NaN last).
//
// Each block bitonic-sorts 4096 (key, index) pairs in shared memory and keeps its best K';
// passes repeat on the survivors until one block remains.
#include "state.cuh"

namespace gpry {

constexpr int TK_E = 4096;        // elements per block
constexpr int TK_THREADS = 512;   // 8 elements per thread

__device__ __forceinline__ uint64_t sortable_key(double x) {
  if (x != x) return 0ull;   // NaN ranks below everything
  uint64_t b = (uint64_t)__double_as_longlong(x);
  return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double key_to_double(uint64_t k, double nan_value) {
  if (k == 0ull) return nan_value;
  uint64_t b = (k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k;
  return __longlong_as_double((long long)b);
}
// true if (ka, ia) must come BEFORE (kb, ib) in the output order
__device__ __forceinline__ bool before(uint64_t ka, int64_t ia, uint64_t kb, int64_t ib) {
  return ka > kb || (ka == kb && ia < ib);
}

__global__ void __launch_bounds__(TK_THREADS)
topk_block_kernel(const double* __restrict__ keys_in, const int64_t* __restrict__ idx_in,
                  int64_t n, int64_t idx_base, int Kp, double* __restrict__ keys_out,
                  int64_t* __restrict__ idx_out) {
  extern __shared__ __align__(16) unsigned char tk_smem[];
  uint64_t* sk = reinterpret_cast<uint64_t*>(tk_smem);
  int64_t* si = reinterpret_cast<int64_t*>(tk_smem + TK_E * 8);
  const int tid = threadIdx.x;
  const int64_t base = (int64_t)blockIdx.x * TK_E;
  for (int e = tid; e < TK_E; e += TK_THREADS) {
    int64_t g = base + e;
    if (g < n) {
      sk[e] = sortable_key(keys_in[g]);
      si[e] = idx_in ? idx_in[g] : idx_base + g;
    } else {
      sk[e] = 0ull;
      si[e] = INT64_MAX;
    }
  }
  __syncthreads();
  // bitonic sort, "before" order ascending in position
  for (int size = 2; size <= TK_E; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int e = tid; e < TK_E / 2; e += TK_THREADS) {
        int lo = 2 * e - (e & (stride - 1));   // index with the `stride` bit clear
        int hi = lo + stride;
        bool up = ((lo & size) == 0);          // this subsequence sorted in "before" order
        uint64_t ka = sk[lo], kb = sk[hi];
        int64_t ia = si[lo], ib = si[hi];
        bool swap = up ? before(kb, ib, ka, ia) : before(ka, ia, kb, ib);
        if (swap) {
          sk[lo] = kb; si[lo] = ib;
          sk[hi] = ka; si[hi] = ia;
        }
      }
      __syncthreads();
    }
  }
  const double nan_value = __longlong_as_double(0x7ff8000000000000ll);
  for (int e = tid; e < Kp; e += TK_THREADS) {
    keys_out[(int64_t)blockIdx.x * Kp + e] = key_to_double(sk[e], nan_value);
    idx_out[(int64_t)blockIdx.x * Kp + e] = si[e];
  }
}

// Returns n_out = min(Kp, M); *d_keys_out / *d_idx_out point at device arrays of Kp entries
// (entries beyond n_out have idx = INT64_MAX).
int64_t topk_device(gpry_state* st, const double* d_scores, int64_t M, int Kp, int64_t idx_base,
                    double** d_keys_out, int64_t** d_idx_out, cudaStream_t s) {
  GPRY_CHECK_ARG(Kp >= 1 && Kp <= MAX_TOPK, "Kp must be in [1, 2048]");
  GPRY_CHECK_ARG(M >= 1, "top-k of an empty pool");
  TimedScope ts(st, s, T_TOPK, 0);
  int64_t nblocks = (M + TK_E - 1) / TK_E;
  size_t cap = (size_t)nblocks * Kp;
  for (int b = 0; b < 2; b++) {
    st->tk_keys[b].reserve(cap);
    st->tk_idx[b].reserve(cap);
  }
  const size_t smem = (size_t)TK_E * 16;
  GPRY_CUDA(cudaFuncSetAttribute(topk_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)smem));
  const double* kin = d_scores;
  const int64_t* iin = nullptr;
  int64_t n = M;
  int cur = 0;
  while (true) {
    nblocks = (n + TK_E - 1) / TK_E;
    topk_block_kernel<<<(unsigned)nblocks, TK_THREADS, smem, s>>>(kin, iin, n, idx_base, Kp,
                                                              st->tk_keys[cur].p, st->tk_idx[cur].p);
    GPRY_CUDA(cudaGetLastError());
    st->n_launches += 1;
    if (nblocks == 1) break;
    kin = st->tk_keys[cur].p;
    iin = st->tk_idx[cur].p;
    n = nblocks * Kp;
    cur ^= 1;
  }
  *d_keys_out = st->tk_keys[cur].p;
  *d_idx_out = st->tk_idx[cur].p;
  return std::min<int64_t>(Kp, M);
}

__global__ void gather_topk_kernel(const int64_t* __restrict__ idx, int64_t n, int64_t idx_base,
                                   const double* __restrict__ X, int d,
                                   const double* __restrict__ mean, const double* __restrict__ std_,
                                   double* __restrict__ o_mean, double* __restrict__ o_std,
                                   double* __restrict__ o_X) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int64_t g = idx[i] - idx_base;
  if (o_mean) o_mean[i] = mean[g];
  if (o_std) o_std[i] = std_[g];
  if (o_X)
    for (int k = 0; k < d; k++) o_X[i * d + k] = X[g * d + k];
}

void gather_topk(gpry_state* st, const int64_t* d_idx, int64_t n, int64_t idx_base,
                 const double* dX, int d, const double* d_mean, const double* d_std,
                 double* o_mean, double* o_std, double* o_X, cudaStream_t s) {
  if (n <= 0) return;
  TimedScope ts(st, s, T_TOPK);
  gather_topk_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(d_idx, n, idx_base, dX, d, d_mean,
                                                                d_std, o_mean, o_std, o_X);
  GPRY_CUDA(cudaGetLastError());
}

}  // namespace gpry
