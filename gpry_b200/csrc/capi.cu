// extern "C" surface of libgpry_b200.so (see include/gpry_b200.h for the contract).
#include <string.h>

#include <algorithm>
#include <mutex>

#include "state.cuh"

#include <cstdlib>

namespace gpry {

static thread_local std::string g_last_error;
void set_last_error(const std::string& msg) { g_last_error = msg; }

TimedScope::TimedScope(gpry_state* st_, cudaStream_t s_, int cat_, int launches)
    : st(st_), s(s_), cat(cat_) {
  st->n_launches += launches;
  if (!st->profiling) return;
  auto get = [&]() {
    cudaEvent_t e;
    if (!st->pool.empty()) {
      e = st->pool.back();
      st->pool.pop_back();
    } else {
      GPRY_CUDA(cudaEventCreate(&e));
    }
    return e;
  };
  e0 = get();
  e1 = get();
  cudaEventRecord(e0, s);
}
TimedScope::~TimedScope() {
  if (!e0) return;
  cudaEventRecord(e1, s);
  st->pending.push_back(EventPair{cat, e0, e1});
}
void resolve_timings(gpry_state* st) {
  for (auto& p : st->pending) {
    float ms = 0.f;
    if (cudaEventSynchronize(p.e1) == cudaSuccess &&
        cudaEventElapsedTime(&ms, p.e0, p.e1) == cudaSuccess)
      st->t_ms[p.cat] += ms;
    st->pool.push_back(p.e0);
    st->pool.push_back(p.e1);
  }
  st->pending.clear();
}

template <typename F>
static int guarded(F&& f) {
  try {
    f();
    return GPRY_OK;
  } catch (const GpryError& e) {
    set_last_error(e.msg);
    cudaGetLastError();
    return e.code;
  } catch (const std::exception& e) {
    set_last_error(std::string("exception: ") + e.what());
    return GPRY_ERR_ARG;
  }
}

// stage a host input on the device (or pass a device pointer through)
static const double* stage_in(gpry_state* st, const double* X, size_t n, bool on_device,
                              DevBuf<double>& buf, cudaStream_t s) {
  if (on_device) return X;
  buf.reserve(n);
  TimedScope ts(st, s, T_H2D, 0);
  GPRY_CUDA(cudaMemcpyAsync(buf.p, X, n * sizeof(double), cudaMemcpyHostToDevice, s));
  return buf.p;
}

// Host candidates: enqueue the H2D copy in blocks on a dedicated copy stream and run the
// pipeline block by block behind the matching event, so that only the first block's copy is
// exposed.  Device candidates: one call.
static void run_block(gpry_state* st, const double* dXb, int64_t off, int64_t n, bool want_var,
                      bool want_acq, double zeta, double sigma_n, double y_max, double* dm,
                      double* ds, double* da, int64_t idx_offset, cudaStream_t s) {
  if (st->sel.on) {     // selection: masks are applied inside finish_select (predict.cu)
    st->sel.gbase = idx_offset + off;
    st->sel.lbase = off;
    st->sel.clf_dec = nullptr;
    if (st->clf_on) {
      st->clf_dec.reserve((size_t)n);
      predict_pipeline(st->clf, dXb, n, true, false, false, 0, 0, 0, st->clf_dec.p, nullptr,
                       nullptr, s);
      st->sel.clf_dec = st->clf_dec.p;
    }
  }
  predict_pipeline(st, dXb, n, dm != nullptr, want_var, want_acq, zeta, sigma_n, y_max,
                   dm ? dm + off : nullptr, ds ? ds + off : nullptr, da ? da + off : nullptr, s);
}

static void run_pipeline_blocks(gpry_state* st, const double* X, int64_t M, bool x_dev,
                                bool want_var, bool want_acq, double zeta, double sigma_n,
                                double y_max, double* dm, double* ds, double* da,
                                const double** dX_out, cudaStream_t s, int64_t idx_offset = 0) {
  const int d = st->d;
  const int64_t block = (int64_t)56 * 2 * st->n_sm * TILE_ROWS;   // multiple of the chunk size
  if (x_dev) {
    *dX_out = X;
    // (selection: block by block as well, so that the classifier scratch stays block sized)
    for (int64_t off = 0; off < M; off += block)
      run_block(st, X + off * d, off, std::min(block, M - off), want_var, want_acq, zeta, sigma_n,
                y_max, dm, ds, da, idx_offset, s);
    return;
  }
  st->Xdev.reserve((size_t)M * d);
  *dX_out = st->Xdev.p;
  const int nblocks = (int)((M + block - 1) / block);
  if (nblocks <= 1) {
    {
      TimedScope ts(st, s, T_H2D, 0);
      GPRY_CUDA(cudaMemcpyAsync(st->Xdev.p, X, (size_t)M * d * 8, cudaMemcpyHostToDevice, s));
    }
    run_block(st, st->Xdev.p, 0, M, want_var, want_acq, zeta, sigma_n, y_max, dm, ds, da,
              idx_offset, s);
    return;
  }
  if (!st->copy_stream)
    GPRY_CUDA(cudaStreamCreateWithFlags(&st->copy_stream, cudaStreamNonBlocking));
  if (!st->call_start)
    GPRY_CUDA(cudaEventCreateWithFlags(&st->call_start, cudaEventDisableTiming));
  while ((int)st->copy_events.size() < nblocks) {
    cudaEvent_t e;
    GPRY_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    st->copy_events.push_back(e);
  }
  // the staging buffer may still be read by work queued earlier on the compute stream
  GPRY_CUDA(cudaEventRecord(st->call_start, s));
  GPRY_CUDA(cudaStreamWaitEvent(st->copy_stream, st->call_start, 0));
  auto enqueue_copy = [&](int b) {
    const int64_t off = b * block, n = std::min(block, M - off);
    GPRY_CUDA(cudaMemcpyAsync(st->Xdev.p + off * d, X + off * d, (size_t)n * d * 8,
                              cudaMemcpyHostToDevice, st->copy_stream));
    GPRY_CUDA(cudaEventRecord(st->copy_events[b], st->copy_stream));
  };
  // The copy of block b + 1 is enqueued right AFTER the kernels of block b: from pageable host
  // memory cudaMemcpyAsync returns only once the block has been staged, and while the host
  // waits there the GPU works through block b, so the transfer stays hidden for pageable as
  // for pinned buffers; only the copy of block 0 is exposed.
  enqueue_copy(0);
  for (int b = 0; b < nblocks; b++) {
    const int64_t off = b * block, n = std::min(block, M - off);
    GPRY_CUDA(cudaStreamWaitEvent(s, st->copy_events[b], 0));
    run_block(st, st->Xdev.p + off * d, off, n, want_var, want_acq, zeta, sigma_n, y_max, dm, ds,
              da, idx_offset, s);
    if (b + 1 < nblocks) enqueue_copy(b + 1);
  }
}

}  // namespace gpry

using namespace gpry;

extern "C" {

int gpry_abi_version(void) { return GPRY_ABI_VERSION; }
const char* gpry_last_error(void) { return g_last_error.c_str(); }

int gpry_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return -1;
  }
  return n;
}

int gpry_state_create(int device, gpry_state** out) {
  return guarded([&] {
    GPRY_CHECK_ARG(out != nullptr, "out is NULL");
    int n = 0;
    GPRY_CUDA(cudaGetDeviceCount(&n));
    GPRY_CHECK_ARG(device >= 0 && device < n, "no such CUDA device");
    GPRY_CUDA(cudaSetDevice(device));
    cudaDeviceProp p;
    GPRY_CUDA(cudaGetDeviceProperties(&p, device));
    if (p.major != 10)
      throw GpryError{GPRY_ERR_ARG, std::string("gpry_b200 is built for sm_100a only; device is ") +
                                        p.name};
    gpry_state* st = new gpry_state();
    st->device = device;
    st->n_sm = p.multiProcessorCount;
    if (const char* e = getenv("GPRY_B200_CONTRACT"))      // "fp64" / "int8": A/B switch
      st->contract_mode = (std::string(e) == "fp64") ? GPRY_CONTRACT_FP64
                          : (std::string(e) == "int8_1pass") ? GPRY_CONTRACT_INT8_1PASS
                                                             : GPRY_CONTRACT_INT8;
    *out = st;
  });
}

int gpry_state_destroy(gpry_state* st) {
  return guarded([&] {
    if (!st) return;
    cudaSetDevice(st->device);
    resolve_timings(st);
    for (auto e : st->pool) cudaEventDestroy(e);
    for (auto e : st->copy_events) cudaEventDestroy(e);
    if (st->call_start) cudaEventDestroy(st->call_start);
    if (st->copy_stream) cudaStreamDestroy(st->copy_stream);
    st->prm_dev.release(); st->T.release(); st->Xt.release(); st->alpha.release();
    st->Vt.release(); st->Ks.release(); st->meanp.release(); st->ssqp.release();
    st->Xdev.release(); st->o_mean.release(); st->o_std.release(); st->o_acq.release();
    for (int b = 0; b < 2; b++) { st->tk_keys[b].release(); st->tk_idx[b].release(); }
    st->tmp.release(); st->small.release(); st->Vrm.release(); st->trust.release();
    st->pc_U.release(); st->pc_Ks.release(); st->pc_UT.release(); st->pc_G.release();
    st->VTrm.release(); st->gr_out.release(); st->clf_dec.release();
    for (int b = 0; b < 2; b++) { st->sel_acq[b].release(); st->sel_mean[b].release(); st->sel_std[b].release(); st->sel_idx[b].release(); st->tk_pos[b].release(); }
    st->sel_ctl.release(); st->excl.release();
    st->oz_probe.release(); st->oz_Ksl.release(); st->oz_Vs.release(); st->oz_scale.release(); st->oz_rb.release(); st->oz_park.release();
    comm_destroy(st);
    st->cm_hdr.release(); st->cm_send.release(); st->cm_recv.release(); st->mg_keys.release();
    st->mg_idx.release();
    if (st->clf) gpry_state_destroy(st->clf);
    st->f_K.release(); st->f_VT.release(); st->f_W.release(); st->f_TT.release();
    st->f_Winv.release(); st->f_misc.release(); st->f_prob.release();
    delete st;
  });
}

int gpry_state_upload(gpry_state* st, int kind, int N, int d, const double* X_train_t,
                      const double* alpha_, const double* V, double c, const double* ell,
                      const double* x_min, const double* x_width, double y_mean, double y_std,
                      double clip_hi) {
  return guarded([&] {
    GPRY_CHECK_ARG(st && X_train_t && alpha_ && ell, "NULL argument");
    st->clf_on = false;
    upload_model(st, kind, N, d, X_train_t, alpha_, V, nullptr, nullptr, 0, nullptr, c, ell, x_min,
                 x_width, y_mean, y_std, clip_hi);
  });
}

int gpry_state_adopt_factorization(gpry_state* st, double c, const double* ell,
                                   const double* x_min, const double* x_width, double y_mean,
                                   double y_std, double clip_hi) {
  return guarded([&] {
    GPRY_CHECK_ARG(st && ell, "NULL argument");
    if (!st->f_valid)
      throw GpryError{GPRY_ERR_STATE, "no device-resident factorization to adopt"};
    st->clf_on = false;
    const int N = st->f_N, d = st->f_d;
    const int Np = round_up(N, TILE_ROWS);
    std::vector<double> Xt((size_t)N * d);
    // train.cu layouts: f_prob = [y Np][noise2 Np][X_ N*d]; f_VT = V^T; f_misc = [alpha Np]...
    GPRY_CUDA(cudaMemcpy(Xt.data(), st->f_prob.p + 2 * (size_t)Np, (size_t)N * d * 8,
                         cudaMemcpyDeviceToHost));
    upload_model(st, st->f_kind, N, d, Xt.data(), nullptr, nullptr, nullptr, st->f_VT.p, Np,
                 st->f_misc.p /* alpha_ */, c, ell, x_min, x_width, y_mean, y_std, clip_hi);
  });
}

int gpry_set_trust_region(gpry_state* st, int d, const double* lower, const double* upper,
                          double value) {
  return guarded([&] {
    GPRY_CHECK_ARG(st != nullptr, "state is NULL");
    if (!lower || !upper) {
      st->trust_on = false;
      return;
    }
    GPRY_CHECK_ARG(d >= 1 && d <= MAX_DIM, "bad dimensionality");
    GPRY_CUDA(cudaSetDevice(st->device));
    std::vector<double> h(2 * MAX_DIM, 0.0);
    for (int k = 0; k < d; k++) {
      h[k] = lower[k];
      h[MAX_DIM + k] = upper[k];
    }
    st->trust.reserve(2 * MAX_DIM);
    GPRY_CUDA(cudaMemcpy(st->trust.p, h.data(), 2 * MAX_DIM * 8, cudaMemcpyHostToDevice));
    st->trust_value = value;
    st->trust_on = true;
  });
}

int gpry_set_contract_mode(gpry_state* st, int mode) {
  return guarded([&] {
    GPRY_CHECK_ARG(st != nullptr, "state is NULL");
    GPRY_CHECK_ARG(mode >= GPRY_CONTRACT_FP64 && mode <= GPRY_CONTRACT_INT8_1PASS, "unknown mode");
    st->contract_mode = mode;
  });
}

int gpry_set_contract_guard(gpry_state* st, int enable) {
  return guarded([&] {
    GPRY_CHECK_ARG(st != nullptr, "state is NULL");
    st->oz_guard = enable != 0;
    st->oz_checked = false;
  });
}

int gpry_contract_info(gpry_state* st, double* out8) {
  return guarded([&] {
    GPRY_CHECK_ARG(st && out8, "NULL argument");
    if (!st->loaded) throw GpryError{GPRY_ERR_STATE, "no model uploaded into this state"};
    GPRY_CUDA(cudaSetDevice(st->device));
    const bool eligible = st->contract_mode != 0 && st->has_V && st->Npad >= 512 && st->Npad <= 16384;
    if (eligible) {        // evaluate the guard now if no scoring call has done it yet
      ozaki_prepare(st, 0);
      ozaki_validate(st, 0);
    }
    out8[0] = st->contract_mode;
    out8[1] = (eligible && ozaki_supported(st)) ? st->contract_mode : GPRY_CONTRACT_FP64;
    out8[2] = eligible ? st->oz_est_sigma : 0.0;
    out8[3] = eligible ? st->oz_bound_worst : 0.0;
    out8[4] = eligible ? st->oz_probe_err : -1.0;
    out8[5] = OZ_TOLERANCE;
    out8[6] = st->oz_guard ? 1.0 : 0.0;
    out8[7] = 0.0;
  });
}

int gpry_int8_peak(gpry_state* st, double* out_tops) {
  return guarded([&] {
    GPRY_CHECK_ARG(st != nullptr && out_tops != nullptr, "NULL argument");
    *out_tops = ozaki_int8_peak_tops(st);
  });
}

int gpry_int8_peak_sustained(gpry_state* st, double seconds, double* out_tops) {
  return guarded([&] {
    GPRY_CHECK_ARG(st != nullptr && out_tops != nullptr, "NULL argument");
    GPRY_CHECK_ARG(seconds > 0.0 && seconds <= 30.0, "seconds must be in (0, 30]");
    *out_tops = ozaki_int8_peak_sustained_tops(st, seconds);
  });
}

int gpry_set_mask_value(gpry_state* st, double value) {
  return guarded([&] {
    GPRY_CHECK_ARG(st != nullptr, "state is NULL");
    st->trust_value = value;
  });
}

int gpry_set_classifier(gpry_state* st, int n_sv, int d, const double* sv, const double* dual_coef,
                        double intercept, double gamma) {
  return guarded([&] {
    GPRY_CHECK_ARG(st != nullptr, "state is NULL");
    if (!sv || !dual_coef) {
      st->clf_on = false;
      return;
    }
    if (!st->loaded) throw GpryError{GPRY_ERR_STATE, "upload the model before its classifier"};
    GPRY_CHECK_ARG(n_sv >= 1 && d == st->d, "classifier: need n_sv >= 1 and the model's d");
    GPRY_CHECK_ARG(gamma > 0.0, "classifier: gamma must be > 0");
    if (!st->clf) {
      st->clf = new gpry_state();
      st->clf->device = st->device;
      st->clf->n_sm = st->n_sm;
    }
    // exp(-gamma r^2) = exp(-r^2 / (2 ell^2)) with ell = 1 / sqrt(2 gamma); same (min, width)
    // transform as the model, decision = 1 * sum_i coef_i k_i + intercept, never clipped
    std::vector<double> prm(3 * MAX_DIM);
    GPRY_CUDA(cudaSetDevice(st->device));
    GPRY_CUDA(cudaMemcpy(prm.data(), st->prm_dev.p, 3 * MAX_DIM * 8, cudaMemcpyDeviceToHost));
    std::vector<double> ell(d, 1.0 / sqrt(2.0 * gamma));
    upload_model(st->clf, GPRY_KERNEL_RBF, n_sv, d, sv, dual_coef, nullptr, nullptr, nullptr, 0,
                 nullptr, 1.0, ell.data(), prm.data(), prm.data() + MAX_DIM, intercept, 1.0,
                 INFINITY);
    st->clf_on = true;
  });
}

int gpry_classify(gpry_state* st, const double* X, int64_t M, int where, double* out_decision,
                  void* stream) {
  return guarded([&] {
    GPRY_CHECK_ARG(st != nullptr, "state is NULL");
    if (!st->clf_on) throw GpryError{GPRY_ERR_STATE, "no classifier set on this state"};
    GPRY_CHECK_ARG(M >= 0, "M < 0");
    if (M == 0) return;
    GPRY_CHECK_ARG(X != nullptr && out_decision != nullptr, "NULL argument");
    GPRY_CUDA(cudaSetDevice(st->device));
    cudaStream_t s = (cudaStream_t)stream;
    const bool x_dev = where & GPRY_X_ON_DEVICE, o_dev = where & GPRY_OUT_ON_DEVICE;
    const double* dX = stage_in(st, X, (size_t)M * st->d, x_dev, st->Xdev, s);
    double* dd = out_decision;
    if (!o_dev) {
      st->clf_dec.reserve((size_t)M);
      dd = st->clf_dec.p;
    }
    predict_pipeline(st->clf, dX, M, true, false, false, 0, 0, 0, dd, nullptr, nullptr, s);
    if (!o_dev) {
      GPRY_CUDA(cudaMemcpyAsync(out_decision, dd, (size_t)M * 8, cudaMemcpyDeviceToHost, s));
      GPRY_CUDA(cudaStreamSynchronize(s));
    }
  });
}

int gpry_state_info(const gpry_state* st, int* N, int* d, int* kind) {
  if (!st || !st->loaded) {
    set_last_error("no model uploaded into this state");
    return GPRY_ERR_STATE;
  }
  if (N) *N = st->N;
  if (d) *d = st->d;
  if (kind) *kind = st->kind;
  return GPRY_OK;
}

static void predict_common(gpry_state* st, const double* X, int64_t M, bool want_mean,
                           bool want_std, bool want_acq, double zeta, double sigma_n, double y_max,
                           int where, double* out_mean, double* out_std, double* out_acq,
                           cudaStream_t s) {
  GPRY_CHECK_ARG(st != nullptr, "state is NULL");
  if (!st->loaded) throw GpryError{GPRY_ERR_STATE, "no model uploaded into this state"};
  GPRY_CHECK_ARG(M >= 0, "M < 0");
  if (M == 0) return;
  GPRY_CHECK_ARG(X != nullptr, "X is NULL");
  GPRY_CUDA(cudaSetDevice(st->device));
  const bool x_dev = where & GPRY_X_ON_DEVICE, o_dev = where & GPRY_OUT_ON_DEVICE;
  double *dm = nullptr, *ds = nullptr, *da = nullptr;
  if (want_mean && out_mean) {
    if (o_dev) dm = out_mean; else { st->o_mean.reserve(M); dm = st->o_mean.p; }
  }
  if (want_std && out_std) {
    if (o_dev) ds = out_std; else { st->o_std.reserve(M); ds = st->o_std.p; }
  }
  if (want_acq && out_acq) {
    if (o_dev) da = out_acq; else { st->o_acq.reserve(M); da = st->o_acq.p; }
  }
  const bool need_var = (ds != nullptr) || (da != nullptr);
  const double* dX = nullptr;
  run_pipeline_blocks(st, X, M, x_dev, need_var, da != nullptr, zeta, sigma_n, y_max, dm, ds, da,
                      &dX, s);
  apply_classifier(st, dX, M, dm, ds, da, s);
  apply_trust_region(st, dX, M, dm, da, s);
  if (!o_dev) {
    TimedScope ts(st, s, T_D2H, 0);
    if (dm) GPRY_CUDA(cudaMemcpyAsync(out_mean, dm, M * 8, cudaMemcpyDeviceToHost, s));
    if (ds) GPRY_CUDA(cudaMemcpyAsync(out_std, ds, M * 8, cudaMemcpyDeviceToHost, s));
    if (da) GPRY_CUDA(cudaMemcpyAsync(out_acq, da, M * 8, cudaMemcpyDeviceToHost, s));
  }
  if (!o_dev || st->profiling) GPRY_CUDA(cudaStreamSynchronize(s));
  if (st->profiling) resolve_timings(st);
}

int gpry_predict(gpry_state* st, const double* X, int64_t M, int what, int where,
                 double* out_mean, double* out_std, void* stream) {
  return guarded([&] {
    predict_common(st, X, M, what & GPRY_WANT_MEAN, what & GPRY_WANT_STD, false, 0, 0, 0, where,
                   out_mean, out_std, nullptr, (cudaStream_t)stream);
  });
}

int gpry_predict_logexp(gpry_state* st, const double* X, int64_t M, double zeta, double sigma_n,
                        double y_max, int where, double* out_mean, double* out_std,
                        double* out_acq, void* stream) {
  return guarded([&] {
    predict_common(st, X, M, true, true, true, zeta, sigma_n, y_max, where, out_mean, out_std,
                   out_acq, (cudaStream_t)stream);
  });
}

int gpry_set_excluded(gpry_state* st, const int64_t* rows, int n) {
  return guarded([&] {
    GPRY_CHECK_ARG(st != nullptr, "state is NULL");
    GPRY_CHECK_ARG(n >= 0 && (n == 0 || rows != nullptr), "bad skip list");
    for (int i = 1; i < n; i++) GPRY_CHECK_ARG(rows[i - 1] < rows[i], "skip list must be sorted");
    st->n_excl = n;
    if (n == 0) return;
    GPRY_CUDA(cudaSetDevice(st->device));
    st->excl.reserve((size_t)n);
    GPRY_CUDA(cudaMemcpy(st->excl.p, rows, (size_t)n * 8, cudaMemcpyHostToDevice));
  });
}

int gpry_predict_logexp_topk(gpry_state* st, const double* X, int64_t M, double zeta,
                             double sigma_n, double y_max, int Kp, int64_t idx_offset, int where,
                             double* out_acq, int64_t* out_idx, double* out_mean,
                             double* out_std, double* out_X, int64_t* n_out, void* stream) {
  return guarded([&] {
    GPRY_CHECK_ARG(st != nullptr && n_out != nullptr, "NULL argument");
    if (!st->loaded) throw GpryError{GPRY_ERR_STATE, "no model uploaded into this state"};
    GPRY_CHECK_ARG(Kp >= 1 && Kp <= MAX_TOPK, "Kp must be in [1, 2048]");
    *n_out = 0;
    if (M <= 0) return;
    cudaStream_t s = (cudaStream_t)stream;
    GPRY_CUDA(cudaSetDevice(st->device));
    const bool x_dev = where & GPRY_X_ON_DEVICE, o_dev = where & GPRY_OUT_ON_DEVICE;
    const int d = st->d;
    const cudaMemcpyKind kind = o_dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
    const double* dX = nullptr;
    // Streaming selection: every chunk's finish kernel appends only the records that can
    // still be among the Kp best (topk.cu); mean / std / acq [M] never exist.
    const int64_t tiles = (M + TILE_ROWS - 1) / TILE_ROWS;
    const int chunk_cands = (int)std::min<int64_t>(tiles, 2 * st->n_sm) * TILE_ROWS;
    select_begin(st, Kp, chunk_cands, s);
    try {
      run_pipeline_blocks(st, X, M, x_dev, true, true, zeta, sigma_n, y_max, nullptr, nullptr,
                          nullptr, &dX, s, idx_offset);
    } catch (...) {
      st->sel.on = false;
      throw;
    }
    const int64_t n = select_finish(st, s);          // synchronises: n is data dependent
    const int b = st->sel.cur;
    {
      TimedScope ts(st, s, T_D2H, 0);
      if (out_acq) GPRY_CUDA(cudaMemcpyAsync(out_acq, st->sel_acq[b].p, n * 8, kind, s));
      if (out_idx) GPRY_CUDA(cudaMemcpyAsync(out_idx, st->sel_idx[b].p, n * 8, kind, s));
      if (out_mean) GPRY_CUDA(cudaMemcpyAsync(out_mean, st->sel_mean[b].p, n * 8, kind, s));
      if (out_std) GPRY_CUDA(cudaMemcpyAsync(out_std, st->sel_std[b].p, n * 8, kind, s));
      if (out_X) {
        st->small.reserve((size_t)Kp * d + 2 * MAX_DIM);
        double* g_X = o_dev ? out_X : st->small.p;
        gather_rows(st, st->sel_idx[b].p, n, idx_offset, dX, d, g_X, s);
        if (!o_dev) GPRY_CUDA(cudaMemcpyAsync(out_X, g_X, n * d * 8, kind, s));
      }
    }
    *n_out = n;
    if (!o_dev || st->profiling) GPRY_CUDA(cudaStreamSynchronize(s));
    if (st->profiling) resolve_timings(st);
  });
}

int gpry_topk(gpry_state* st, const double* scores, int64_t M, int Kp, int where,
              double* out_scores, int64_t* out_idx, int64_t* n_out, void* stream) {
  return guarded([&] {
    GPRY_CHECK_ARG(st && scores && n_out, "NULL argument");
    *n_out = 0;
    if (M <= 0) return;
    cudaStream_t s = (cudaStream_t)stream;
    GPRY_CUDA(cudaSetDevice(st->device));
    const bool x_dev = where & GPRY_X_ON_DEVICE, o_dev = where & GPRY_OUT_ON_DEVICE;
    const double* dS = stage_in(st, scores, (size_t)M, x_dev, st->o_acq, s);
    double* d_keys;
    int64_t* d_idx;
    int64_t n = topk_device(st, dS, M, Kp, 0, &d_keys, &d_idx, s);
    const cudaMemcpyKind kind = o_dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
    if (out_scores) GPRY_CUDA(cudaMemcpyAsync(out_scores, d_keys, n * 8, kind, s));
    if (out_idx) GPRY_CUDA(cudaMemcpyAsync(out_idx, d_idx, n * 8, kind, s));
    *n_out = n;
    if (!o_dev || st->profiling) GPRY_CUDA(cudaStreamSynchronize(s));
    if (st->profiling) resolve_timings(st);
  });
}

int gpry_mean_grad(gpry_state* st, const double* x, double* out_grad) {
  return guarded([&] {
    GPRY_CHECK_ARG(st && x && out_grad, "NULL argument");
    mean_grad_device(st, x, out_grad);
  });
}

int gpry_predict_grad(gpry_state* st, const double* X, int M, double* out_mean, double* out_std,
                      double* out_grad_mean, double* out_grad_std) {
  return guarded([&] {
    GPRY_CHECK_ARG(st != nullptr, "state is NULL");
    predict_grad_device(st, X, M, out_mean, out_std, out_grad_mean, out_grad_std);
  });
}

int gpry_std_grad(gpry_state* st, const double* x, double* out_grad, double* out_std) {
  return guarded([&] {
    GPRY_CHECK_ARG(st && x && out_grad, "NULL argument");
    std_grad_device(st, x, out_grad, out_std);
  });
}

int gpry_posterior_cov(gpry_state* st, const double* X, int Ka, int where, double* out_cov,
                       void* stream) {
  return guarded([&] {
    GPRY_CHECK_ARG(st && X && out_cov, "NULL argument");
    if (!st->loaded) throw GpryError{GPRY_ERR_STATE, "no model uploaded into this state"};
    cudaStream_t s = (cudaStream_t)stream;
    GPRY_CUDA(cudaSetDevice(st->device));
    const bool x_dev = where & GPRY_X_ON_DEVICE, o_dev = where & GPRY_OUT_ON_DEVICE;
    const double* dX = stage_in(st, X, (size_t)Ka * st->d, x_dev, st->Xdev, s);
    double* d_out = out_cov;
    if (!o_dev) {
      st->o_acq.reserve((size_t)Ka * Ka);
      d_out = st->o_acq.p;
    }
    posterior_cov_device(st, dX, Ka, d_out, s);
    if (!o_dev)
      GPRY_CUDA(cudaMemcpyAsync(out_cov, d_out, (size_t)Ka * Ka * 8, cudaMemcpyDeviceToHost, s));
    if (!o_dev || st->profiling) GPRY_CUDA(cudaStreamSynchronize(s));
    if (st->profiling) resolve_timings(st);
  });
}

int gpry_kernel_cross(gpry_state* st, int kind, int d, const double* theta, const double* X,
                      int M, const double* Y, int N, double* out) {
  return guarded([&] {
    GPRY_CHECK_ARG(st && theta && X && Y && out, "NULL argument");
    kernel_cross_device(st, kind, d, theta, X, M, Y, N, out);
  });
}

int gpry_kernel_gradient_x(gpry_state* st, const double* x_t, double* out) {
  return guarded([&] {
    GPRY_CHECK_ARG(st && x_t && out, "NULL argument");
    kernel_gradx_device(st, x_t, out);
  });
}

int gpry_factorize(gpry_state* st, int kind, int N, int d, const double* X_train_t,
                   const double* noise2, const double* y_t, const double* theta, double* out_L,
                   double* out_V, double* out_alpha, double* out_logdet_half, int* info,
                   int keep_on_device) {
  return guarded([&] {
    GPRY_CHECK_ARG(st && X_train_t && noise2 && y_t && theta && info, "NULL argument");
    factorize_device(st, kind, N, d, X_train_t, noise2, y_t, theta, out_L, out_V, out_alpha,
                     out_logdet_half, info, keep_on_device != 0);
  });
}

int gpry_factor_append(gpry_state* st, int k, const double* X_new_t, const double* noise2_new,
                       const double* y_t_all, const double* theta, double* out_alpha, int* info) {
  return guarded([&] {
    GPRY_CHECK_ARG(st && X_new_t && noise2_new && y_t_all && theta && info, "NULL argument");
    *info = factor_append_device(st, k, X_new_t, noise2_new, y_t_all, theta, out_alpha);
  });
}

int gpry_factor_download(gpry_state* st, double* out_L, double* out_V) {
  return guarded([&] {
    GPRY_CHECK_ARG(st != nullptr, "state is NULL");
    factor_download_device(st, out_L, out_V);
  });
}

int gpry_lml_batched(gpry_state* st, int kind, int N, int d, const double* X_train_t,
                     const double* noise2, const double* y_t, const double* thetas, int B,
                     double* out_lml, double* out_grad, int* out_info) {
  return guarded([&] {
    GPRY_CHECK_ARG(st && X_train_t && noise2 && y_t && thetas && out_lml && out_info,
                   "NULL argument");
    GPRY_CHECK_ARG(B >= 1, "B < 1");
    lml_batched_device(st, kind, N, d, X_train_t, noise2, y_t, thetas, B, out_lml, out_grad,
                       out_info);
  });
}

int gpry_comm_unique_id(void* out128) {
  return guarded([&] {
    GPRY_CHECK_ARG(out128 != nullptr, "NULL argument");
    comm_unique_id(out128);
  });
}

int gpry_comm_init(gpry_state* st, const void* id128, int rank, int nranks) {
  return guarded([&] {
    GPRY_CHECK_ARG(st && id128, "NULL argument");
    comm_init(st, id128, rank, nranks);
  });
}

int gpry_comm_share(gpry_state* dst, gpry_state* src) {
  return guarded([&] {
    GPRY_CHECK_ARG(dst && src, "NULL argument");
    comm_share(dst, src);
  });
}

int gpry_comm_destroy(gpry_state* st) {
  return guarded([&] {
    GPRY_CHECK_ARG(st != nullptr, "state is NULL");
    comm_destroy(st);
  });
}

int gpry_comm_info(const gpry_state* st, int* rank, int* nranks, int* nccl_version) {
  return guarded([&] {
    GPRY_CHECK_ARG(st != nullptr, "state is NULL");
    if (rank) *rank = st->comm_rank;
    if (nranks) *nranks = st->comm ? st->comm_size : 0;
    if (nccl_version) *nccl_version = comm_nccl_version();
  });
}

int gpry_bcast_state(gpry_state* st, int root, void* stream) {
  return guarded([&] {
    GPRY_CHECK_ARG(st != nullptr, "state is NULL");
    bcast_state(st, root, (cudaStream_t)stream);
  });
}

int gpry_allgather_topk(gpry_state* st, int n_local, int Kp, int d, const double* acq,
                        const int64_t* idx, const double* mean, const double* std_,
                        const double* X, int where, double* out_acq, int64_t* out_idx,
                        double* out_mean, double* out_std, double* out_X, int64_t* n_out,
                        double* next_acq, void* stream) {
  return guarded([&] {
    GPRY_CHECK_ARG(st != nullptr, "state is NULL");
    allgather_topk(st, n_local, Kp, d, acq, idx, mean, std_, X, where & GPRY_X_ON_DEVICE,
                   where & GPRY_OUT_ON_DEVICE, out_acq, out_idx, out_mean, out_std, out_X, n_out,
                   next_acq, (cudaStream_t)stream);
  });
}

int gpry_set_profiling(gpry_state* st, int enable) {
  if (!st) return GPRY_ERR_ARG;
  st->profiling = enable != 0;
  return GPRY_OK;
}

int gpry_get_timings(gpry_state* st, double* out8, int reset) {
  return guarded([&] {
    GPRY_CHECK_ARG(st && out8, "NULL argument");
    cudaSetDevice(st->device);
    resolve_timings(st);
    for (int i = 0; i < T_NCATS; i++) out8[i] = st->t_ms[i];
    out8[6] = st->n_launches;
    out8[7] = st->n_contract_launches;
    if (reset) {
      for (int i = 0; i < T_NCATS; i++) st->t_ms[i] = 0;
      st->n_launches = 0;
      st->n_contract_launches = 0;
    }
  });
}

}  // extern "C"
