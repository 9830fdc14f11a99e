// The two exchange steps of the multi-GPU path, on the library's own NCCL communicator
// (bcast_state, allgather_topk):
//
//   gpry_bcast_state     the model a rank holds after a refit (X_train_, alpha_, V_ = L^-1,
//                        kernel / pre-processor scalars) -> every rank's device state, GPU to
//                        GPU over NVLink.  Replaces the pickled-regressor broadcast of the
//                        reference (run.py:749-756 _share_gpr -> mpi.bcast) for what the
//                        scoring ranks need.
//   gpry_allgather_topk  per-GPU survivor records -> all ranks, merged on the device into the
//                        K' best of the union (same order on every rank).  Replaces the five
//                        gathers + re-ranking input + bcast of gp_acquisition.py:1148-1191.
//
// NCCL is bound at run time with dlopen("libnccl.so.2"): a process that has imported PyTorch
// gets the copy PyTorch loaded (one NCCL per process), anything else the system library; the
// shared object itself has no link-time dependency on NCCL.  Only the 128-byte unique id has to
// travel through a host channel (MPI bcast in the reference's world, torch.distributed here).
#include <dlfcn.h>

#include <algorithm>
#include <cstring>
#include <mutex>
#include <vector>

#include "state.cuh"

namespace gpry {

namespace {

// the slice of nccl.h this file needs (ABI stable across NCCL 2.x)
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;          // 0 = ncclSuccess
constexpr int kNcclInt8 = 0;       // ncclInt8 / ncclChar
constexpr int kNcclFloat64 = 8;    // ncclFloat64 / ncclDouble

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi& nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
      api.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (api.handle) break;
    }
    if (!api.handle) return;
    auto sym = [&](const char* n) { return dlsym(api.handle, n); };
    api.GetVersion = (decltype(api.GetVersion))sym("ncclGetVersion");
    api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
    api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
    api.Broadcast = (decltype(api.Broadcast))sym("ncclBroadcast");
    api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
    api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
    api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
    api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
  });
  if (!api.handle || !api.GetUniqueId || !api.CommInitRank || !api.Broadcast || !api.AllGather)
    throw GpryError{GPRY_ERR_STATE, "NCCL (libnccl.so.2) could not be loaded"};
  return api;
}

void nccl_check(ncclResult_t r, const char* what) {
  if (r == 0) return;
  const char* txt = nccl().GetErrorString ? nccl().GetErrorString(r) : "?";
  throw GpryError{GPRY_ERR_CUDA, std::string("NCCL error in ") + what + ": " + txt};
}

ncclComm_t comm_of(gpry_state* st) {
  if (!st->comm) throw GpryError{GPRY_ERR_STATE, "no communicator: call gpry_comm_init first"};
  return (ncclComm_t)st->comm;
}

// header of a broadcast model: scalars, then x_min / x_width / ell as in prm_dev
constexpr int HDR_SCALARS = 16;
constexpr int HDR_DOUBLES = HDR_SCALARS + 3 * MAX_DIM;

// survivor records packed as rows of (4 + d) doubles: acq, index (int64 bits), mean, std, x[d]
__global__ void pack_records_kernel(const double* __restrict__ acq, const int64_t* __restrict__ idx,
                                    const double* __restrict__ mean, const double* __restrict__ sd,
                                    const double* __restrict__ X, int n, int Kp, int d,
                                    double* __restrict__ rec) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Kp) return;
  double* r = rec + (size_t)i * (4 + d);
  if (i < n) {
    r[0] = acq[i];
    r[1] = __longlong_as_double(idx[i]);
    r[2] = mean ? mean[i] : 0.0;
    r[3] = sd ? sd[i] : 0.0;
    for (int k = 0; k < d; k++) r[4 + k] = X ? X[(size_t)i * d + k] : 0.0;
  } else {      // padding of a short list: ranks below everything
    r[0] = __longlong_as_double(0x7ff8000000000000ll);
    r[1] = __longlong_as_double(INT64_MAX);
    r[2] = r[3] = 0.0;
    for (int k = 0; k < d; k++) r[4 + k] = 0.0;
  }
}
// out[e] = record pos[e] of the union, e < Kp; next = acquisition value of the first record left out
__global__ void unpack_records_kernel(const double* __restrict__ rec, const int* __restrict__ pos,
                                      const int64_t* __restrict__ gidx, int Kp, int d,
                                      double* __restrict__ acq, int64_t* __restrict__ idx,
                                      double* __restrict__ mean, double* __restrict__ sd,
                                      double* __restrict__ X, double* __restrict__ tail) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e > Kp) return;
  const int p = pos[e];
  const bool live = p >= 0 && gidx[e] != INT64_MAX;
  if (e == Kp) {       // first record beyond the K' best: the bound NORA's exactness test needs
    tail[1] = live ? rec[(size_t)p * (4 + d)] : -INFINITY;
    return;
  }
  if (!live) {
    if (idx) idx[e] = INT64_MAX;
    return;
  }
  const double* r = rec + (size_t)p * (4 + d);
  if (acq) acq[e] = r[0];
  if (idx) idx[e] = gidx[e];
  if (mean) mean[e] = r[2];
  if (sd) sd[e] = r[3];
  if (X)
    for (int k = 0; k < d; k++) X[(size_t)e * d + k] = r[4 + k];
}
__global__ void count_live_kernel(const int64_t* __restrict__ gidx, int Kp, double* __restrict__ tail) {
  __shared__ int cnt;
  if (threadIdx.x == 0) cnt = 0;
  __syncthreads();
  int c = 0;
  for (int e = threadIdx.x; e < Kp; e += blockDim.x) c += gidx[e] != INT64_MAX;
  atomicAdd(&cnt, c);
  __syncthreads();
  if (threadIdx.x == 0) tail[0] = (double)cnt;
}

}  // namespace

void comm_destroy(gpry_state* st);

void comm_unique_id(void* out128) {
  ncclUniqueId id;
  nccl_check(nccl().GetUniqueId(&id), "ncclGetUniqueId");
  memcpy(out128, id.internal, 128);
}

void comm_init(gpry_state* st, const void* id128, int rank, int nranks) {
  GPRY_CHECK_ARG(nranks >= 1 && rank >= 0 && rank < nranks, "bad rank / nranks");
  GPRY_CUDA(cudaSetDevice(st->device));
  comm_destroy(st);
  ncclUniqueId id;
  memcpy(id.internal, id128, 128);
  ncclComm_t c = nullptr;
  nccl_check(nccl().CommInitRank(&c, nranks, id, rank), "ncclCommInitRank");
  st->comm = c;
  st->comm_owner = true;
  st->comm_rank = rank;
  st->comm_size = nranks;
}

// dst uses src's communicator (same process, same GPU) without owning it
void comm_share(gpry_state* dst, gpry_state* src) {
  GPRY_CHECK_ARG(dst->device == src->device, "communicators are per GPU");
  if (!src->comm) throw GpryError{GPRY_ERR_STATE, "the source state has no communicator"};
  if (dst == src) return;
  comm_destroy(dst);
  dst->comm = src->comm;
  dst->comm_owner = false;
  dst->comm_rank = src->comm_rank;
  dst->comm_size = src->comm_size;
}

void comm_destroy(gpry_state* st) {
  if (!st->comm) return;
  cudaSetDevice(st->device);
  if (st->comm_owner) nccl().CommDestroy((ncclComm_t)st->comm);
  st->comm_owner = true;
  st->comm = nullptr;
  st->comm_size = 1;
  st->comm_rank = 0;
}

int comm_nccl_version() {
  int v = 0;
  if (nccl().GetVersion) nccl().GetVersion(&v);
  return v;
}

// Two NCCL communicators must never have kernels in flight at the same time on a GPU (the host
// process usually owns one of its own, e.g. PyTorch's): every collective of this file first
// waits for everything queued on the device, and returns only when its own work is done.
static void quiesce_device() { GPRY_CUDA(cudaDeviceSynchronize()); }

void bcast_state(gpry_state* st, int root, cudaStream_t s) {
  ncclComm_t comm = comm_of(st);
  GPRY_CHECK_ARG(root >= 0 && root < st->comm_size, "bad root");
  GPRY_CUDA(cudaSetDevice(st->device));
  quiesce_device();
  const bool am_root = st->comm_rank == root;
  if (am_root && !st->loaded)
    throw GpryError{GPRY_ERR_STATE, "bcast_state: the root has no model uploaded"};
  // 1. header
  st->cm_hdr.reserve(HDR_DOUBLES);
  std::vector<double> h(HDR_DOUBLES, 0.0);
  if (am_root) {
    h[0] = st->kind; h[1] = st->N; h[2] = st->d; h[3] = st->c; h[4] = st->y_mean;
    h[5] = st->y_std; h[6] = st->clip_hi; h[7] = st->has_V ? 1.0 : 0.0;
    GPRY_CUDA(cudaMemcpyAsync(st->cm_hdr.p, h.data(), HDR_SCALARS * 8, cudaMemcpyHostToDevice, s));
    GPRY_CUDA(cudaMemcpyAsync(st->cm_hdr.p + HDR_SCALARS, st->prm_dev.p, 3 * MAX_DIM * 8,
                              cudaMemcpyDeviceToDevice, s));
  }
  nccl_check(nccl().Broadcast(st->cm_hdr.p, st->cm_hdr.p, HDR_DOUBLES, kNcclFloat64, root, comm, s),
             "ncclBroadcast(header)");
  GPRY_CUDA(cudaMemcpyAsync(h.data(), st->cm_hdr.p, HDR_DOUBLES * 8, cudaMemcpyDeviceToHost, s));
  GPRY_CUDA(cudaStreamSynchronize(s));
  const int kind = (int)h[0], N = (int)h[1], d = (int)h[2];
  const bool has_V = h[7] != 0.0;
  const int Np = round_up(N, TILE_ROWS);
  // 2. payload: X_train_ (N x d), alpha_ (Np, zero padded), V row-major (Np x Np, zero padded)
  double *xt, *al, *vr = nullptr;
  if (am_root) {
    xt = st->Xt.p; al = st->alpha.p; vr = st->Vrm.p;
  } else {
    st->cm_recv.reserve((size_t)N * d + Np);
    xt = st->cm_recv.p; al = xt + (size_t)N * d;
    if (has_V) {
      st->tmp.reserve((size_t)Np * Np);
      vr = st->tmp.p;
    }
  }
  if (nccl().GroupStart) nccl_check(nccl().GroupStart(), "ncclGroupStart");
  nccl_check(nccl().Broadcast(xt, xt, (size_t)N * d, kNcclFloat64, root, comm, s), "ncclBroadcast(X)");
  nccl_check(nccl().Broadcast(al, al, (size_t)Np, kNcclFloat64, root, comm, s), "ncclBroadcast(alpha)");
  if (has_V)
    nccl_check(nccl().Broadcast(vr, vr, (size_t)Np * Np, kNcclFloat64, root, comm, s),
               "ncclBroadcast(V)");
  if (nccl().GroupEnd) nccl_check(nccl().GroupEnd(), "ncclGroupEnd");
  if (am_root) {
    GPRY_CUDA(cudaStreamSynchronize(s));
    return;
  }
  // 3. the receivers build their state from the device copies (no N^2 host traffic)
  std::vector<double> Xt_host((size_t)N * d);
  GPRY_CUDA(cudaMemcpyAsync(Xt_host.data(), xt, (size_t)N * d * 8, cudaMemcpyDeviceToHost, s));
  GPRY_CUDA(cudaStreamSynchronize(s));
  st->clf_on = false;
  upload_model(st, kind, N, d, Xt_host.data(), nullptr, nullptr, vr, nullptr, Np, al, h[3],
               h.data() + HDR_SCALARS + 2 * MAX_DIM, h.data() + HDR_SCALARS,
               h.data() + HDR_SCALARS + MAX_DIM, h[4], h[5], h[6]);
}

// All ranks: n_local (<= Kp) records in, the Kp best of the union out (sorted: descending acq,
// ascending index).  *n_out = records returned; *next_acq = best acquisition value NOT returned.
void allgather_topk(gpry_state* st, int n_local, int Kp, int d, const double* acq,
                    const int64_t* idx, const double* mean, const double* sd, const double* X,
                    bool in_dev, bool out_dev, double* o_acq, int64_t* o_idx, double* o_mean,
                    double* o_sd, double* o_X, int64_t* n_out, double* next_acq, cudaStream_t s) {
  ncclComm_t comm = comm_of(st);
  GPRY_CHECK_ARG(Kp >= 1 && Kp <= MAX_TOPK && n_local >= 0 && n_local <= Kp, "bad K' / n_local");
  GPRY_CHECK_ARG(d >= 0 && d <= MAX_DIM, "bad d");
  GPRY_CHECK_ARG(n_local == 0 || (acq && idx), "acq / idx missing");
  GPRY_CUDA(cudaSetDevice(st->device));
  quiesce_device();
  const int W = st->comm_size, R = 4 + d;
  const size_t rec_local = (size_t)Kp * R, n_union = (size_t)W * Kp;
  st->cm_send.reserve(rec_local + (size_t)Kp * (3 + d) + Kp);
  st->cm_recv.reserve(n_union * R);
  const double *dacq = acq, *dmean = mean, *dsd = sd, *dX = X;
  const int64_t* didx = idx;
  if (!in_dev && n_local > 0) {      // stage the host records
    double* stage = st->cm_send.p + rec_local;
    auto up = [&](const void* src, size_t n8) {
      double* dst = stage;
      GPRY_CUDA(cudaMemcpyAsync(dst, src, n8 * 8, cudaMemcpyHostToDevice, s));
      stage += n8;
      return dst;
    };
    dacq = up(acq, n_local);
    didx = reinterpret_cast<const int64_t*>(up(idx, n_local));
    if (mean) dmean = up(mean, n_local);
    if (sd) dsd = up(sd, n_local);
    if (X) dX = up(X, (size_t)n_local * d);
  }
  pack_records_kernel<<<(Kp + 127) / 128, 128, 0, s>>>(dacq, didx, dmean, dsd, dX, n_local, Kp, d,
                                                      st->cm_send.p);
  GPRY_CUDA(cudaGetLastError());
  nccl_check(nccl().AllGather(st->cm_send.p, st->cm_recv.p, rec_local, kNcclFloat64, comm, s),
             "ncclAllGather(survivors)");
  // merge: exact top-(Kp + 1) of the union by (acq desc, index asc)
  double* keys;
  int64_t* gidx;
  int* pos;
  merge_records(st, st->cm_recv.p, (int)n_union, R, Kp + 1, &keys, &gidx, &pos, s);
  st->small.reserve((size_t)Kp * (4 + d) + 2 * MAX_DIM + 8);
  double* tail = st->small.p + (size_t)Kp * (4 + d) + 2 * MAX_DIM;      // [n_live, next_acq]
  double* g_acq = out_dev ? o_acq : st->small.p;
  double* g_mean = out_dev ? o_mean : st->small.p + Kp;
  double* g_sd = out_dev ? o_sd : st->small.p + 2 * (size_t)Kp;
  double* g_X = out_dev ? o_X : st->small.p + 4 * (size_t)Kp;
  int64_t* g_idx = out_dev ? o_idx : reinterpret_cast<int64_t*>(st->small.p + 3 * (size_t)Kp);
  unpack_records_kernel<<<(Kp + 1 + 127) / 128, 128, 0, s>>>(
      st->cm_recv.p, pos, gidx, Kp, d, o_acq ? g_acq : nullptr, g_idx, o_mean ? g_mean : nullptr,
      o_sd ? g_sd : nullptr, o_X ? g_X : nullptr, tail);
  GPRY_CUDA(cudaGetLastError());
  count_live_kernel<<<1, 256, 0, s>>>(gidx, Kp, tail);
  GPRY_CUDA(cudaGetLastError());
  double h_tail[2];
  GPRY_CUDA(cudaMemcpyAsync(h_tail, tail, 16, cudaMemcpyDeviceToHost, s));
  GPRY_CUDA(cudaStreamSynchronize(s));
  const int64_t n = (int64_t)h_tail[0];
  if (!out_dev) {
    if (o_acq) GPRY_CUDA(cudaMemcpyAsync(o_acq, g_acq, n * 8, cudaMemcpyDeviceToHost, s));
    if (o_idx) GPRY_CUDA(cudaMemcpyAsync(o_idx, g_idx, n * 8, cudaMemcpyDeviceToHost, s));
    if (o_mean) GPRY_CUDA(cudaMemcpyAsync(o_mean, g_mean, n * 8, cudaMemcpyDeviceToHost, s));
    if (o_sd) GPRY_CUDA(cudaMemcpyAsync(o_sd, g_sd, n * 8, cudaMemcpyDeviceToHost, s));
    if (o_X) GPRY_CUDA(cudaMemcpyAsync(o_X, g_X, n * d * 8, cudaMemcpyDeviceToHost, s));
    GPRY_CUDA(cudaStreamSynchronize(s));
  }
  if (n_out) *n_out = n;
  if (next_acq) *next_acq = h_tail[1];
  st->n_launches += 4;
}

}  // namespace gpry
