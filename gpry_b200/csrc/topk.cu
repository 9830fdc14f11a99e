// Ranked-pool pre-selection: the K' candidates with the largest acquisition value, in the
// order RankedPool.add(method="single sort acq") visits them (descending acq,
// gp_acquisition.py:1326-1333).  Exact, deterministic (ties: ascending index; NaN last).
//
// Each block bitonic-sorts 4096 (key, index) pairs in shared memory and keeps its best K';
// passes repeat on the survivors until one block remains.
#include "state.cuh"

#include <algorithm>

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

// ---------------------------------------------------------------------------------------
// Streaming selection (gpry_predict_logexp_topk).  finish_select_kernel (predict.cu) appends
// the record (acq, global index, mean, std) of every candidate whose key is >= the current
// threshold tau to the record buffer with warp-aggregated atomics; tau = key of the K'-th best
// record seen so far (0 = everything passes until K' records exist).  A compaction sorts the
// buffer exactly (block bitonic sorts of 4096 (key, index, position) triples, repeated on the
// per-block survivors), moves the K' best -- in output order -- to the front of the other
// buffer and raises tau.  The per-candidate arrays mean/std/acq[M] are never written.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TK_THREADS)
topk_rec_kernel(const double* __restrict__ keys_in, const int64_t* __restrict__ gidx_in,
                const int* __restrict__ pos_in, const unsigned long long* __restrict__ count_ptr,
                int64_t n, int Kp, double* __restrict__ keys_out, int64_t* __restrict__ gidx_out,
                int* __restrict__ pos_out, const unsigned long long* __restrict__ done_flag) {
  if (done_flag && *done_flag != 0ull) return;
  extern __shared__ __align__(16) unsigned char tk_smem[];
  uint64_t* sk = reinterpret_cast<uint64_t*>(tk_smem);
  int64_t* si = reinterpret_cast<int64_t*>(tk_smem + TK_E * 8);
  int* sp = reinterpret_cast<int*>(tk_smem + TK_E * 16);
  const int tid = threadIdx.x;
  const int64_t base = (int64_t)blockIdx.x * TK_E;
  if (count_ptr) n = min(n, (int64_t)*count_ptr);
  const double nan_value = __longlong_as_double(0x7ff8000000000000ll);
  if (base >= n) {        // nothing in this block's range: empties
    for (int e = tid; e < Kp; e += TK_THREADS) {
      keys_out[(int64_t)blockIdx.x * Kp + e] = nan_value;
      gidx_out[(int64_t)blockIdx.x * Kp + e] = INT64_MAX;
      pos_out[(int64_t)blockIdx.x * Kp + e] = -1;
    }
    return;
  }
  for (int e = tid; e < TK_E; e += TK_THREADS) {
    int64_t g = base + e;
    if (g < n && gidx_in[g] != INT64_MAX) {
      sk[e] = sortable_key(keys_in[g]);
      si[e] = gidx_in[g];
      sp[e] = pos_in ? pos_in[g] : (int)g;
    } else {
      sk[e] = 0ull;
      si[e] = INT64_MAX;
      sp[e] = -1;
    }
  }
  __syncthreads();
  for (int size = 2; size <= TK_E; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int e = tid; e < TK_E / 2; e += TK_THREADS) {
        int lo = 2 * e - (e & (stride - 1));
        int hi = lo + stride;
        bool up = ((lo & size) == 0);
        uint64_t ka = sk[lo], kb = sk[hi];
        int64_t ia = si[lo], ib = si[hi];
        bool swap = up ? before(kb, ib, ka, ia) : before(ka, ia, kb, ib);
        if (swap) {
          int pa = sp[lo], pb = sp[hi];
          sk[lo] = kb; si[lo] = ib; sp[lo] = pb;
          sk[hi] = ka; si[hi] = ia; sp[hi] = pa;
        }
      }
      __syncthreads();
    }
  }
  for (int e = tid; e < Kp; e += TK_THREADS) {
    keys_out[(int64_t)blockIdx.x * Kp + e] = key_to_double(sk[e], nan_value);
    gidx_out[(int64_t)blockIdx.x * Kp + e] = si[e];
    pos_out[(int64_t)blockIdx.x * Kp + e] = sp[e];
  }
}

// Fast path of a compaction: at most 4096 records in the buffer (the usual case once the
// threshold is set) -> one block sorts them, writes the K' best to the other buffer and updates
// count / threshold; ctl[3] tells the general chain behind it whether it still has to run.
__global__ void __launch_bounds__(TK_THREADS)
select_compact_small_kernel(const double* __restrict__ a_in, const int64_t* __restrict__ i_in,
                            const double* __restrict__ m_in, const double* __restrict__ s_in,
                            int Kp, double* __restrict__ a_out, int64_t* __restrict__ i_out,
                            double* __restrict__ m_out, double* __restrict__ s_out,
                            unsigned long long* __restrict__ ctl) {
  extern __shared__ __align__(16) unsigned char tk_smem[];
  uint64_t* sk = reinterpret_cast<uint64_t*>(tk_smem);
  int64_t* si = reinterpret_cast<int64_t*>(tk_smem + TK_E * 8);
  int* sp = reinterpret_cast<int*>(tk_smem + TK_E * 16);
  const int tid = threadIdx.x;
  const unsigned long long count = ctl[0];
  if (count > (unsigned long long)TK_E) {
    if (tid == 0) ctl[3] = 0ull;
    return;
  }
  const int n = (int)count;
  for (int e = tid; e < TK_E; e += TK_THREADS) {
    if (e < n) {
      sk[e] = sortable_key(a_in[e]);
      si[e] = i_in[e];
      sp[e] = e;
    } else {
      sk[e] = 0ull;
      si[e] = INT64_MAX;
      sp[e] = -1;
    }
  }
  __syncthreads();
  // smallest power of two >= n is enough to sort (the tail is all "empty", already last)
  int span = 2;
  while (span < n) span <<= 1;
  for (int size = 2; size <= span; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int e = tid; e < span / 2; e += TK_THREADS) {
        int lo = 2 * e - (e & (stride - 1));
        int hi = lo + stride;
        bool up = ((lo & size) == 0);
        uint64_t ka = sk[lo], kb = sk[hi];
        int64_t ia = si[lo], ib = si[hi];
        bool swap = up ? before(kb, ib, ka, ia) : before(ka, ia, kb, ib);
        if (swap) {
          int pa = sp[lo], pb = sp[hi];
          sk[lo] = kb; si[lo] = ib; sp[lo] = pb;
          sk[hi] = ka; si[hi] = ia; sp[hi] = pa;
        }
      }
      __syncthreads();
    }
  }
  for (int e = tid; e < Kp; e += TK_THREADS) {
    const int p = sp[e];
    if (p >= 0) {
      a_out[e] = a_in[p];
      m_out[e] = m_in[p];
      s_out[e] = s_in[p];
      i_out[e] = si[e];
    } else {
      i_out[e] = INT64_MAX;
    }
  }
  if (tid == 0) {
    const unsigned long long kept = count < (unsigned long long)Kp ? count : (unsigned long long)Kp;
    ctl[0] = kept;
    ctl[1] = kept == (unsigned long long)Kp ? sk[Kp - 1] : 0ull;
    ctl[3] = 1ull;
  }
}

// the K' best (sorted) -> front of the other record buffer; count = min(count, K'); tau
__global__ void __launch_bounds__(1024)
select_gather_kernel(const double* __restrict__ keys, const int64_t* __restrict__ gidx,
                     const int* __restrict__ pos, int Kp, const double* __restrict__ a_in,
                     const double* __restrict__ m_in, const double* __restrict__ s_in,
                     double* __restrict__ a_out, int64_t* __restrict__ i_out,
                     double* __restrict__ m_out, double* __restrict__ s_out,
                     unsigned long long* __restrict__ ctl) {
  if (ctl[3] != 0ull) return;        // the single-block fast path has done it
  const unsigned long long count = ctl[0];
  for (int e = threadIdx.x; e < Kp; e += blockDim.x) {
    const int p = pos[e];
    if (p >= 0) {
      a_out[e] = a_in[p];
      m_out[e] = m_in[p];
      s_out[e] = s_in[p];
      i_out[e] = gidx[e];
    } else {
      i_out[e] = INT64_MAX;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned long long kept = count < (unsigned long long)Kp ? count : (unsigned long long)Kp;
    ctl[0] = kept;
    ctl[1] = kept == (unsigned long long)Kp ? sortable_key(keys[Kp - 1]) : 0ull;
  }
}

void select_begin(gpry_state* st, int Kp, int chunk_cands, cudaStream_t s) {
  SelectRun& r = st->sel;
  r.on = true;
  r.Kp = Kp;
  r.cur = 0;
  r.pending = 0;
  r.first = true;              // compact right after the first chunk: sets the threshold early
  r.max_pending = 64;
  r.chunk_cands = chunk_cands;
  r.cap = Kp + r.max_pending * chunk_cands;
  for (int b = 0; b < 2; b++) {
    st->sel_acq[b].reserve(r.cap);
    st->sel_mean[b].reserve(r.cap);
    st->sel_std[b].reserve(r.cap);
    st->sel_idx[b].reserve(r.cap);
  }
  st->sel_ctl.reserve(4);
  GPRY_CUDA(cudaMemsetAsync(st->sel_ctl.p, 0, 4 * sizeof(unsigned long long), s));
}

void select_compact(gpry_state* st, cudaStream_t s) {
  SelectRun& r = st->sel;
  if (r.pending == 0) return;
  TimedScope ts(st, s, T_TOPK, 0);
  const int Kp = r.Kp;
  const int o = r.cur ^ 1;
  const size_t smem = (size_t)TK_E * 20;
  GPRY_CUDA(cudaFuncSetAttribute(select_compact_small_kernel,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  select_compact_small_kernel<<<1, TK_THREADS, smem, s>>>(
      st->sel_acq[r.cur].p, st->sel_idx[r.cur].p, st->sel_mean[r.cur].p, st->sel_std[r.cur].p, Kp,
      st->sel_acq[o].p, st->sel_idx[o].p, st->sel_mean[o].p, st->sel_std[o].p, st->sel_ctl.p);
  GPRY_CUDA(cudaGetLastError());
  st->n_launches += 1;
  // general chain (no-ops when the fast path did it): sized by what can be in the buffer
  int64_t n = std::min<int64_t>(r.cap, (int64_t)Kp + (int64_t)r.pending * r.chunk_cands);
  int64_t nblocks = (n + TK_E - 1) / TK_E;
  for (int b = 0; b < 2; b++) {
    st->tk_keys[b].reserve((size_t)nblocks * Kp);
    st->tk_idx[b].reserve((size_t)nblocks * Kp);
    st->tk_pos[b].reserve((size_t)nblocks * Kp);
  }
  GPRY_CUDA(cudaFuncSetAttribute(topk_rec_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)smem));
  const double* kin = st->sel_acq[r.cur].p;
  const int64_t* iin = st->sel_idx[r.cur].p;
  const int* pin = nullptr;
  const unsigned long long* cnt = st->sel_ctl.p;
  int lvl = 0;
  while (true) {
    nblocks = (n + TK_E - 1) / TK_E;
    topk_rec_kernel<<<(unsigned)nblocks, TK_THREADS, smem, s>>>(
        kin, iin, pin, cnt, n, Kp, st->tk_keys[lvl].p, st->tk_idx[lvl].p, st->tk_pos[lvl].p,
        st->sel_ctl.p + 3);
    GPRY_CUDA(cudaGetLastError());
    st->n_launches += 1;
    kin = st->tk_keys[lvl].p;
    iin = st->tk_idx[lvl].p;
    pin = st->tk_pos[lvl].p;
    cnt = nullptr;
    if (nblocks == 1) break;
    n = nblocks * Kp;
    lvl ^= 1;
  }
  select_gather_kernel<<<1, 1024, 0, s>>>(kin, iin, pin, Kp, st->sel_acq[r.cur].p,
                                          st->sel_mean[r.cur].p, st->sel_std[r.cur].p,
                                          st->sel_acq[o].p, st->sel_idx[o].p, st->sel_mean[o].p,
                                          st->sel_std[o].p, st->sel_ctl.p);
  GPRY_CUDA(cudaGetLastError());
  st->n_launches += 1;
  r.cur = o;
  r.pending = 0;
  r.first = false;
}

// final compaction; returns the number of records (<= K'), which sit sorted at the front of
// sel_*[st->sel.cur].  Synchronises the stream.
int64_t select_finish(gpry_state* st, cudaStream_t s) {
  SelectRun& r = st->sel;
  r.pending = std::max(r.pending, 1);
  select_compact(st, s);
  unsigned long long h[3] = {0, 0, 0};
  GPRY_CUDA(cudaMemcpyAsync(h, st->sel_ctl.p, sizeof(h), cudaMemcpyDeviceToHost, s));
  GPRY_CUDA(cudaStreamSynchronize(s));
  r.on = false;
  if (h[2] != 0)
    throw GpryError{GPRY_ERR_STATE, "selection buffer overflow (internal error)"};
  return (int64_t)h[0];
}

// ---------------------------------------------------------------------------------------
// merge of all-gathered survivor records (comm.cu): rows of R doubles whose first two entries are
// (acq, index as int64 bits) -> the Kq best, as sorted (key, index, row number) triples
// ---------------------------------------------------------------------------------------
__global__ void split_records_kernel(const double* __restrict__ rec, int n, int R,
                                     double* __restrict__ keys, int64_t* __restrict__ gidx) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  keys[i] = rec[(size_t)i * R];
  gidx[i] = __double_as_longlong(rec[(size_t)i * R + 1]);
}

void merge_records(gpry_state* st, const double* rec, int n, int R, int Kq, double** keys,
                   int64_t** gidx, int** pos, cudaStream_t s) {
  GPRY_CHECK_ARG(Kq >= 1 && Kq <= TK_E, "merge: too many records requested");
  st->mg_keys.reserve((size_t)n);
  st->mg_idx.reserve((size_t)n);
  split_records_kernel<<<(n + 255) / 256, 256, 0, s>>>(rec, n, R, st->mg_keys.p, st->mg_idx.p);
  GPRY_CUDA(cudaGetLastError());
  int64_t m = n;
  int64_t nblocks = (m + TK_E - 1) / TK_E;
  for (int b = 0; b < 2; b++) {
    st->tk_keys[b].reserve((size_t)nblocks * Kq);
    st->tk_idx[b].reserve((size_t)nblocks * Kq);
    st->tk_pos[b].reserve((size_t)nblocks * Kq);
  }
  const size_t smem = (size_t)TK_E * 20;
  GPRY_CUDA(cudaFuncSetAttribute(topk_rec_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)smem));
  const double* kin = st->mg_keys.p;
  const int64_t* iin = st->mg_idx.p;
  const int* pin = nullptr;
  int lvl = 0;
  while (true) {
    nblocks = (m + TK_E - 1) / TK_E;
    topk_rec_kernel<<<(unsigned)nblocks, TK_THREADS, smem, s>>>(
        kin, iin, pin, nullptr, m, Kq, st->tk_keys[lvl].p, st->tk_idx[lvl].p, st->tk_pos[lvl].p,
        nullptr);
    GPRY_CUDA(cudaGetLastError());
    kin = st->tk_keys[lvl].p;
    iin = st->tk_idx[lvl].p;
    pin = st->tk_pos[lvl].p;
    if (nblocks == 1) break;
    m = nblocks * Kq;
    lvl ^= 1;
  }
  *keys = st->tk_keys[lvl].p;
  *gidx = st->tk_idx[lvl].p;
  *pos = st->tk_pos[lvl].p;
}

__global__ void gather_rows_kernel(const int64_t* __restrict__ idx, int64_t n, int64_t idx_base,
                                   const double* __restrict__ X, int d, double* __restrict__ o_X) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int64_t g = idx[i] - idx_base;
  for (int k = 0; k < d; k++) o_X[i * d + k] = X[g * d + k];
}
void gather_rows(gpry_state* st, const int64_t* d_idx, int64_t n, int64_t idx_base,
                 const double* dX, int d, double* o_X, cudaStream_t s) {
  if (n <= 0 || !o_X) return;
  gather_rows_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(d_idx, n, idx_base, dX, d, o_X);
  GPRY_CUDA(cudaGetLastError());
  st->n_launches += 1;
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
