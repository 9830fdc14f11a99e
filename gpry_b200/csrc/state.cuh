// Device-resident model state and scratch management for the gpry_b200 library.
#pragma once
#include <vector>

#include "common.cuh"

namespace gpry {

constexpr int MAX_DIM = 128;      // supported input dimensionality
constexpr int MAX_DIM_REG = 32;   // dimensionality handled with coordinates in registers
constexpr int MAX_TOPK = 2048;    // largest K' of the fused ranking

// per-dimension affine transform + length scales, passed by value to kernels
struct XformParams {
  double x_min[MAX_DIM_REG];
  double x_width[MAX_DIM_REG];
  double ell[MAX_DIM_REG];
};

enum TimingCat { T_BUILD = 0, T_CONTRACT = 1, T_FINISH = 2, T_TOPK = 3, T_H2D = 4, T_D2H = 5,
                 T_NCATS = 6 };

template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t cap = 0;  // elements
  void reserve(size_t n) {
    if (n <= cap) return;
    if (p) GPRY_CUDA(cudaFree(p));
    p = nullptr;
    cap = 0;
    size_t want = n + n / 8;
    cudaError_t e = cudaMalloc((void**)&p, want * sizeof(T));
    if (e != cudaSuccess) {
      cudaGetLastError();
      want = n;
      e = cudaMalloc((void**)&p, want * sizeof(T));
    }
    if (e != cudaSuccess) {
      cudaGetLastError();
      throw GpryError{GPRY_ERR_NOMEM, "device allocation of " +
                                          std::to_string(want * sizeof(T)) + " bytes failed"};
    }
    cap = want;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};

struct EventPair {
  int cat;
  cudaEvent_t e0, e1;
};

// Host-side bookkeeping of one gpry_predict_logexp_topk call (the selection runs chunk by chunk
// inside predict_pipeline; see topk.cu)
struct SelectRun {
  bool on = false;
  int Kp = 0, cap = 0, cur = 0;
  bool first = true;          // no compaction yet in this run
  int chunk_cands = 0;        // largest chunk (upper bound of the records one chunk can append)
  int pending = 0;            // chunks appended since the last compaction
  int max_pending = 0;        // compaction period (the buffer holds K' + that many chunks)
  int64_t gbase = 0;          // global index (idx_offset included) of row 0 of the current block
  int64_t lbase = 0;          // row number in the pool of row 0 of the current block
  const double* clf_dec = nullptr;   // classifier decisions of the current block, or NULL
};

}  // namespace gpry

struct gpry_state {
  int device = 0;
  int n_sm = 148;
  bool loaded = false;

  // model
  int kind = 0, N = 0, d = 0;
  int DP = 0;     // d padded to a multiple of 4
  int Npad = 0;   // N padded to a multiple of 128
  int nJ = 0;     // row blocks of V (Npad / 128)
  int nKT = 0;    // k-tiles per candidate tile (Npad / 16)
  double c = 1.0, y_mean = 0.0, y_std = 1.0, clip_hi = 0.0;
  // optional trust region applied to the mean on the device (gpr.py:1104-1109, 1200-1201)
  bool has_V = true;                     // false: mean-only state (classifier)
  // INT8 split of the variance contraction (ozaki.cu); contract_mode: 0 = FP64 DMMA, 1 = INT8
  int contract_mode = 1;
  bool oz_valid = false;                 // digits of V / row scales / split lists are current
  // guard of the INT8 split (ozaki.cu error model + probe comparison with the FP64 kernel):
  // a model whose estimated or probed error exceeds the tolerance takes the FP64 contraction
  bool oz_guard = true, oz_checked = false, oz_ok = true;
  double oz_bound_worst = 0.0, oz_est_sigma = 0.0, oz_probe_err = -1.0;
  int oz_splits = 1, oz_max_rb = 0, oz_rows = 0;
  gpry::DevBuf<double> oz_probe;         // probe candidates + their std from both kernels
  gpry::DevBuf<double> oz_park;          // per-SM scratch of the two-pass kernel (L2 resident)
  gpry::DevBuf<uint8_t> oz_Ksl, oz_Vs;   // digits of the K* chunk / of V
  gpry::DevBuf<double> oz_scale;         // [Npad] c 2^e_j 2^-12, then [Npad] 2^e_j
  gpry::DevBuf<int> oz_rb;               // row blocks per split + counts
  uint8_t* oz_Ksl_cur = nullptr;         // outputs of the build launch in flight
  double* meanp_cur = nullptr;
  gpry_state* clf = nullptr;             // infinities classifier: a mean-only sub-state whose
  bool clf_on = false;                   //   'mean' is the SVC decision function (svm.py:308-346)
  gpry::DevBuf<double> clf_dec;          // decision values of the current call
  bool trust_on = false;
  double trust_value = 0.0;
  gpry::DevBuf<double> trust;            // [2][MAX_DIM] lower, upper (un-transformed)
  gpry::XformParams prm;                 // valid when d <= MAX_DIM_REG
  gpry::DevBuf<double> prm_dev;          // [3][MAX_DIM] x_min, x_width, ell (generic-d path)
  gpry::DevBuf<double> T;                // [Npad][DP]  X_train_ / ell   (zero padded)
  gpry::DevBuf<double> Xt;               // [N][d]      X_train_ (un-scaled, gradient kernel)
  gpry::DevBuf<double> alpha;            // [Npad]
  gpry::DevBuf<double> Vt;               // tiled lower block triangle of V = L^-1
  gpry::DevBuf<double> Vrm;              // [Npad][Npad] row-major V, zero padded (posterior cov)
  gpry::DevBuf<double> pc_U, pc_Ks, pc_UT, pc_G;   // posterior-covariance scratch
  gpry::DevBuf<double> VTrm, gr_out;     // [Npad][Npad] row-major V^T (lazy), batched-gradient outputs
  bool vtrm_valid = false;

  // scratch (grown on demand, reused across calls)
  gpry::DevBuf<double> Ks;               // [chunk_tiles][nKT] K* tiles
  gpry::DevBuf<double> meanp;            // [JS][chunk_cands]
  gpry::DevBuf<double> ssqp;             // [row_splits][chunk_cands]
  gpry::DevBuf<double> Xdev;             // staged candidates (host input)
  gpry::DevBuf<double> o_mean, o_std, o_acq;   // per-candidate outputs (device)
  gpry::DevBuf<double> tk_keys[2];
  gpry::DevBuf<int64_t> tk_idx[2];
  gpry::DevBuf<int> tk_pos[2];
  // streaming selection of gpry_predict_logexp_topk (topk.cu): the records (acq, global index,
  // mean, std) of the candidates that can still be among the K' best; two buffers (compaction
  // copies the sorted K' best of one to the front of the other); sel_ctl = [count, threshold
  // key, overflow flag]
  gpry::DevBuf<double> sel_acq[2], sel_mean[2], sel_std[2];
  gpry::DevBuf<int64_t> sel_idx[2];
  gpry::DevBuf<unsigned long long> sel_ctl;
  gpry::SelectRun sel;                   // host side of the selection in flight
  gpry::DevBuf<int64_t> excl;            // sorted local row numbers skipped by the ranking
  int n_excl = 0;
  // NCCL communicator owned by the library (comm.cu) + exchange buffers
  void* comm = nullptr;
  bool comm_owner = true;
  int comm_rank = 0, comm_size = 1;
  gpry::DevBuf<double> cm_hdr, cm_send, cm_recv, mg_keys;
  gpry::DevBuf<int64_t> mg_idx;
  gpry::DevBuf<double> tmp;              // upload staging (raw V etc.)
  gpry::DevBuf<double> small;            // small outputs (gradient, top-k records)

  // training-side residency (train.cu): shared problem + batched work buffers
  gpry::DevBuf<double> f_K, f_VT, f_W, f_TT, f_Winv, f_misc;
  gpry::DevBuf<double> f_prob;           // [y Np][noise2 Np][X_ N*d]
  int f_N = 0, f_d = 0, f_kind = -1;
  bool f_valid = false;

  // host-input pipelining: H2D copies of later blocks overlap the compute of earlier ones
  cudaStream_t copy_stream = nullptr;
  std::vector<cudaEvent_t> copy_events;
  cudaEvent_t call_start = nullptr;

  // profiling
  bool profiling = false;
  std::vector<gpry::EventPair> pending;
  std::vector<cudaEvent_t> pool;
  double t_ms[gpry::T_NCATS] = {0, 0, 0, 0, 0, 0};
  double n_launches = 0, n_contract_launches = 0;
};

namespace gpry {

// RAII timing scope: records an event pair on `s` when profiling is on; counts launches.
struct TimedScope {
  gpry_state* st;
  cudaStream_t s;
  int cat;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  TimedScope(gpry_state* st_, cudaStream_t s_, int cat_, int launches = 1);
  ~TimedScope();
};
void resolve_timings(gpry_state* st);

// predict.cu
void predict_pipeline(gpry_state* st, const double* dX, int64_t M, bool want_mean,
                      bool want_var, bool want_acq, double zeta, double sigma_n, double y_max,
                      double* d_mean, double* d_std, double* d_acq, cudaStream_t s);
void mean_grad_device(gpry_state* st, const double* x_host, double* out_host);
void predict_grad_device(gpry_state* st, const double* hX, int M, double* h_mean, double* h_std,
                         double* h_gmean, double* h_gstd);
void apply_classifier(gpry_state* st, const double* dX, int64_t M, double* d_mean, double* d_std,
                      double* d_acq, cudaStream_t s);
void apply_trust_region(gpry_state* st, const double* dX, int64_t M, double* d_mean,
                        double* d_acq, cudaStream_t s);
void upload_model(gpry_state* st, int kind, int N, int d, const double* X_train_t,
                  const double* alpha_, const double* V_host, const double* V_dev_rowmajor,
                  const double* VT_dev_rowmajor, int ldV, const double* alpha_dev, double c,
                  const double* ell, const double* x_min, const double* x_width, double y_mean,
                  double y_std, double clip_hi);
void posterior_cov_device(gpry_state* st, const double* dX, int Ka, double* d_out, cudaStream_t s);
void kernel_cross_device(gpry_state* st, int kind, int d, const double* theta, const double* hX,
                         int M, const double* hY, int N, double* h_out);
void kernel_gradx_device(gpry_state* st, const double* x_host, double* out_host);
void std_grad_device(gpry_state* st, const double* x_host, double* out_grad, double* out_std);
// topk.cu
void select_begin(gpry_state* st, int Kp, int chunk_cands, cudaStream_t s);
void select_compact(gpry_state* st, cudaStream_t s);
int64_t select_finish(gpry_state* st, cudaStream_t s);
void merge_records(gpry_state* st, const double* rec, int n, int R, int Kq, double** keys,
                   int64_t** gidx, int** pos, cudaStream_t s);
// comm.cu
void comm_unique_id(void* out128);
void comm_init(gpry_state* st, const void* id128, int rank, int nranks);
void comm_destroy(gpry_state* st);
void comm_share(gpry_state* dst, gpry_state* src);
int comm_nccl_version();
void bcast_state(gpry_state* st, int root, cudaStream_t s);
void allgather_topk(gpry_state* st, int n_local, int Kp, int d, const double* acq,
                    const int64_t* idx, const double* mean, const double* sd, const double* X,
                    bool in_dev, bool out_dev, double* o_acq, int64_t* o_idx, double* o_mean,
                    double* o_sd, double* o_X, int64_t* n_out, double* next_acq, cudaStream_t s);
void gather_rows(gpry_state* st, const int64_t* d_idx, int64_t n, int64_t idx_base,
                 const double* dX, int d, double* o_X, cudaStream_t s);
int64_t topk_device(gpry_state* st, const double* d_scores, int64_t M, int Kp, int64_t idx_base,
                    double** d_keys_out, int64_t** d_idx_out, cudaStream_t s);
void gather_topk(gpry_state* st, const int64_t* d_idx, int64_t n, int64_t idx_base,
                 const double* dX, int d, const double* d_mean, const double* d_std,
                 double* o_mean, double* o_std, double* o_X, cudaStream_t s);
// train.cu
void gemm_nt(const double* A, int lda, const double* B, int ldb, double* C, int ldc, int M, int N,
             int K, double alpha, int accumulate, int lower_only, int klo_row, cudaStream_t s);
void factorize_device(gpry_state* st, int kind, int N, int d, const double* X_train_t,
                      const double* noise2, const double* y_t, const double* theta,
                      double* out_L, double* out_V, double* out_alpha, double* out_logdet_half,
                      int* info, bool keep);
bool ozaki_supported(const gpry_state* st);
void ozaki_validate(gpry_state* st, cudaStream_t s);
constexpr double OZ_TOLERANCE = 1e-10;   // on the variance, in units of max(var, y_std^2)
double ozaki_int8_peak_tops(gpry_state* st);
double ozaki_int8_peak_sustained_tops(gpry_state* st, double seconds);
void ozaki_prepare(gpry_state* st, cudaStream_t s);
size_t ozaki_kslices_bytes(const gpry_state* st, int tiles);
void ozaki_contract(gpry_state* st, const uint8_t* Ksl, int tiles, int chunk_cands, cudaStream_t s);
int factor_append_device(gpry_state* st, int k, const double* X_new_t, const double* noise2_new,
                         const double* y_all, const double* theta, double* out_alpha);
void factor_download_device(gpry_state* st, double* out_L, double* out_V);
void lml_batched_device(gpry_state* st, int kind, int N, int d, const double* X_train_t,
                        const double* noise2, const double* y_t, const double* thetas, int B,
                        double* out_lml, double* out_grad, int* out_info);

}  // namespace gpry
