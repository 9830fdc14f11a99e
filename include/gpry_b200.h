/*
 * gpry_b200 -- C ABI of the B200-native GP surrogate hot path (drop-in boundary).
 *
 * One shared library (gpry_b200/libgpry_b200.so, built by __graft_entry__.build()) exports
 * the entry points below.  Plain C: opaque handle, raw pointers, sizes, int status codes.
 * No torch types.  Every pointer argument may be a HOST pointer or a DEVICE pointer on the
 * state's GPU, as told by the GPRY_*_ON_DEVICE bits of `where`.  All floating point is
 * IEEE binary64; indices are int64.  All calls are synchronous with respect to the host
 * unless stated; they enqueue work on `stream` (a cudaStream_t passed as void*, NULL = the
 * legacy default stream) and, for host outputs, synchronize that stream before returning.
 *
 * Each entry point replaces a piece of the reference GPry 3.0.0 (Python; file:line are
 * into the reference tree, "sklearn:" into scikit-learn 1.9.0 gaussian_process/):
 *
 *   gpry_state_upload        state produced by GaussianProcessRegressor._update_model /
 *                            _kernel_inverse (gpr.py:996-1020, 1453-1465): X_train_, alpha_,
 *                            V_ = L^-1, kernel_ hyper-parameters, Normalize_bounds /
 *                            Normalize_y scalars (preprocessing.py:380, 620, 630)
 *   gpry_predict             GaussianProcessRegressor.predict(return_std) and predict_std
 *                            arithmetic, gpr.py:1176-1227, 1325-1347 (kernel call
 *                            sklearn:kernels.py:971,1278,1569-1570,1720-1729)
 *   gpry_predict_logexp      the above + LogExp.f, acquisition_functions.py:1068-1074, as
 *                            evaluated by NORA (mpi.py:182-218 compute_y_parallel,
 *                            gp_acquisition.py:1049-1051, 1110-1125)
 *   gpry_predict_logexp_topk the above + the descending-acquisition pre-ranking that
 *                            RankedPool.add(method="single sort acq") starts from
 *                            (gp_acquisition.py:1326-1333); only the K' best leave the GPU
 *   gpry_mean_grad           predict(return_mean_grad) for one point, gpr.py:1236-1242 with
 *                            Kernel.gradient_x kernels.py:257-278, 363-432, 687-699
 *   gpry_factorize           _update_model + _kernel_inverse: K = k(X_,X_) + diag(alpha),
 *                            L_ = chol(K), V_ = L^-1, alpha_ = K^-1 y_  (gpr.py:1015-1017,
 *                            1456-1465)
 *   gpry_lml_batched         log_marginal_likelihood(theta, eval_gradient) for a batch of
 *                            theta, gpr.py:876-881 -> sklearn:_gpr.py:541-656 with kernel
 *                            gradients sklearn:kernels.py:964-969, 1283-1292, 1581-1584,
 *                            1752-1771
 *
 * Return value: 0 on success; GPRY_ERR_* (<0) on failure, message via gpry_last_error().
 * Non-positive-definite matrices are reported through `info` outputs (LAPACK convention:
 * info = k > 0 means the leading minor of order k is not positive definite), never as an
 * error code: the Python layer turns them into numpy.linalg.LinAlgError (gpr.py:1458-1464)
 * or into (-inf, 0) (sklearn:_gpr.py:592-593).
 */
#ifndef GPRY_B200_H
#define GPRY_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GPRY_ABI_VERSION 1

/* kernel kinds: ConstantKernel * {RBF, Matern(nu=1.5), Matern(nu=2.5)}, anisotropic */
#define GPRY_KERNEL_RBF 0
#define GPRY_KERNEL_MATERN15 1
#define GPRY_KERNEL_MATERN25 2

/* `where` bits */
#define GPRY_X_ON_DEVICE 1   /* input candidates are a device pointer   */
#define GPRY_OUT_ON_DEVICE 2 /* output arrays are device pointers       */

/* `what` bits for gpry_predict */
#define GPRY_WANT_MEAN 1
#define GPRY_WANT_STD 2

#define GPRY_OK 0
#define GPRY_ERR_CUDA -1
#define GPRY_ERR_ARG -2
#define GPRY_ERR_STATE -3 /* state has no uploaded model */
#define GPRY_ERR_NOMEM -4

typedef struct gpry_state gpry_state;

int gpry_abi_version(void);
const char* gpry_last_error(void);

/* Number of CUDA devices visible; <0 on error. */
int gpry_device_count(void);

int gpry_state_create(int device, gpry_state** out);
int gpry_state_destroy(gpry_state* st);

/*
 * Upload a fitted model.  X_train_t is the TRANSFORMED training set (N x d, row major),
 * alpha_ (N), V (N x N row major, lower triangular = L^-1; the strict upper triangle is
 * ignored; V = NULL uploads a mean-only model: any std / acquisition request then fails with
 * GPRY_ERR_STATE).  c = constant_value, ell[d] = length scales.  Candidates are transformed on the
 * device as ((x - x_min) / x_width) / ell  (identity: x_min = 0, x_width = 1).  Outputs are
 * de-normalised as mean * y_std + y_mean, std * y_std (identity: 0, 1) and the mean is
 * clipped above at clip_hi (+inf = no clipping).  All pointers are host pointers.
 */
int gpry_state_upload(gpry_state* st, int kind, int N, int d, const double* X_train_t,
                      const double* alpha_, const double* V, double c, const double* ell,
                      const double* x_min, const double* x_width, double y_mean,
                      double y_std, double clip_hi);

/* Same, but V / alpha_ are taken from the device-resident result of the last
 * gpry_factorize(..., keep_on_device=1) on this state (no N^2 host round trip). */
int gpry_state_adopt_factorization(gpry_state* st, double c, const double* ell,
                                   const double* x_min, const double* x_width,
                                   double y_mean, double y_std, double clip_hi);

/* Trust region applied ON THE DEVICE: candidates outside [lower, upper] (un-transformed, d
 * entries each, host pointers, bounds inclusive) get mean = value (gpr.py:1104-1109, 1200-1201
 * with tools.py:263-287) and, in the acquisition entry points, acq = -inf (LogExp of a
 * non-finite mean, acquisition_functions.py:983-992); the std is not touched.
 * lower = upper = NULL switches it off.  Stays set across uploads. */
int gpry_set_trust_region(gpry_state* st, int d, const double* lower, const double* upper,
                          double value);

/*
 * How the variance contraction sum_j (sum_k V_jk k*_ik)^2 of gpr.py:1204-1208 is evaluated:
 *   GPRY_CONTRACT_FP64            FP64 tensor cores (DMMA.8x8x4)
 *   GPRY_CONTRACT_INT8 (default)  Ozaki split: both operands rounded to 55-bit fixed point
 *                                 (k* / c to 2^-54, V_jk to 2^(e_j - 54), e_j the exponent of the
 *                                 row maximum) and written as 7 balanced int8 digits; the 28 digit
 *                                 products of groups p + q <= 6 run on the INT8 tensor cores
 *                                 (tcgen05.mma kind::i8, exact int32 accumulators in TMEM) and are
 *                                 recombined in FP64.  FP64-EQUIVALENT, not exact: per row product
 *                                 the error is <= 7.02 c 2^e_j n_j 2^-54 in the worst case and
 *                                 ~2.2 c 2^-55 sqrt((|V_j|^2 + 4^e_j n_j) / 3) statistically
 *                                 (n_j = j + 1 terms); d var <= 2 sqrt(c) max_j |d w_j|.
 *                                 Eligible: 512 <= N_pad <= 16384 and more than 64 candidates per
 *                                 call; other calls use FP64.
 * GUARD (on by default): once per uploaded model the statistical estimate is evaluated and 512
 * probe candidates (training points + draws from their bounding box) are scored with BOTH
 * kernels; if the estimate exceeds 1e-10 or the probes differ by more than 1e-10 / 16 (variance,
 * in units of max(var, y_std^2)) the model takes the FP64 kernel.  gpry_contract_info reports
 * what was decided: out8 = [mode requested, mode in use, statistical estimate, worst-case bound,
 * probe difference (-1: not probed), tolerance, guard on, 0].  gpry_set_contract_guard(st, 0)
 * switches the guard off (measurements of the raw INT8 path).
 * The environment variable GPRY_B200_CONTRACT=fp64|int8|int8_1pass sets the initial mode of new
 * states.
 */
#define GPRY_CONTRACT_FP64 0
#define GPRY_CONTRACT_INT8 1       /* two passes over the digit groups, 128 x 128 x 32 MMAs */
#define GPRY_CONTRACT_INT8_1PASS 2 /* one pass, 128 x 64 x 32 MMAs (all 7 groups in TMEM at once) */
int gpry_set_contract_mode(gpry_state* st, int mode);
int gpry_set_contract_guard(gpry_state* st, int enable);
int gpry_contract_info(gpry_state* st, double* out8);

/* Measured issue rate (TOPS, 2 ops per multiply-add) of tcgen05.mma kind::i8 at its best shape
 * (128 x 256 x 32) on operands resident in shared memory, all SMs busy: the roofline
 * denominator of the INT8 contraction.  Takes a few milliseconds. */
int gpry_int8_peak(gpry_state* st, double* out_tops);
/* The same measurement launched back to back for `seconds` (<= 30): the rate sustained under the
 * board's power cap, i.e. the denominator for a kernel timed inside a long step. */
int gpry_int8_peak_sustained(gpry_state* st, double seconds, double* out_tops);

/* The value written by the two masks (GaussianProcessRegressor.minus_inf_value, read at call
 * time by the reference: gpr.py:1145, 1201; gp_acquisition.py:788-792 changes it temporarily). */
int gpry_set_mask_value(gpry_state* st, double value);

/*
 * Infinities classifier evaluated ON THE DEVICE (svm.py:308-346 SVM.predict = sign of the
 * decision function of a two-class RBF SVC; called by predict / predict_std on every pool,
 * gpr.py:1136-1174, 1300-1318):  decision(x) = sum_i dual_coef[i] exp(-gamma |x_ - sv_i|^2)
 * + intercept, with x_ the transformed candidate (the (x_min, x_width) of the uploaded model)
 * and sv (n_sv x d, row major) given in that transformed space.  While set, rows with
 * decision <= 0 come out of gpry_predict / gpry_predict_logexp / gpry_predict_logexp_topk
 * with mean = mask value, std = 0, acq = -inf, so the mask never crosses PCIe.  Host pointers;
 * sv = NULL switches it off; every gpry_state_upload / adopt_factorization clears it.
 */
int gpry_set_classifier(gpry_state* st, int n_sv, int d, const double* sv,
                        const double* dual_coef, double intercept, double gamma);

/* Decision values for M un-transformed candidates (host or device per `where`). */
int gpry_classify(gpry_state* st, const double* X, int64_t M, int where, double* out_decision,
                  void* stream);

/* Query what is loaded: N, d, kind (any may be NULL). Returns GPRY_ERR_STATE if empty. */
int gpry_state_info(const gpry_state* st, int* N, int* d, int* kind);

/* Posterior mean and/or std at M candidates X (M x d row major, un-transformed). */
int gpry_predict(gpry_state* st, const double* X, int64_t M, int what, int where,
                 double* out_mean, double* out_std, void* stream);

/* mean, std and acq = 2 zeta (mean - y_max) + log(sqrt(max(std^2 - sigma_n^2, 0))).
 * Any of the three outputs may be NULL. */
int gpry_predict_logexp(gpry_state* st, const double* X, int64_t M, double zeta,
                        double sigma_n, double y_max, int where, double* out_mean,
                        double* out_std, double* out_acq, void* stream);

/*
 * Rows that gpry_predict_logexp_topk must leave out of the ranking: NORA's already proposed
 * points, which the reference deletes from the MC sample before scoring it
 * (gp_acquisition.py:1037-1047).  rows: n strictly increasing row numbers within the pool X of
 * the following calls (host pointer); n = 0 clears the list.  Stays set until changed.
 */
int gpry_set_excluded(gpry_state* st, const int64_t* rows, int n);

/*
 * Fused scoring + ranking: the Kp candidates with the largest acquisition value, sorted by
 * descending acq (ties: ascending index; NaN ranks last).  idx are positions in X plus
 * idx_offset (so that shards report global indices).  out_X (Kp x d) may be NULL.
 * *n_out = min(Kp, M - rows skipped by gpry_set_excluded).  Output arrays are host or device
 * per GPRY_OUT_ON_DEVICE and must hold Kp entries; n_out is always a host pointer.
 * The finishing kernel of every chunk of candidates keeps only the records whose acquisition
 * value can still be among the Kp best (warp ballot + one atomic per warp against the running
 * Kp-th best value; exact block sorts compact the survivors): the per-candidate mean / std /
 * acquisition arrays are never written, only ranked candidates leave the GPU.  The call
 * synchronises `stream` once (the number of survivors is data dependent).
 */
int gpry_predict_logexp_topk(gpry_state* st, const double* X, int64_t M, double zeta,
                             double sigma_n, double y_max, int Kp, int64_t idx_offset,
                             int where, double* out_acq, int64_t* out_idx,
                             double* out_mean, double* out_std, double* out_X,
                             int64_t* n_out, void* stream);

/* Top-Kp of an arbitrary device/host score array (used to merge shard results). */
int gpry_topk(gpry_state* st, const double* scores, int64_t M, int Kp, int where,
              double* out_scores, int64_t* out_idx, int64_t* n_out, void* stream);

/* d mean / d x_ at ONE un-transformed point x (d) -> out_grad (d), host pointers. The
 * gradient is w.r.t. the transformed coordinate, scaled by y_std (gpr.py:1237-1242). */
int gpry_mean_grad(gpry_state* st, const double* x, double* out_grad);

/* d std / d x_ at ONE un-transformed point x (d) -> out_grad (d); out_std (1, may be NULL)
 * receives the std at x.  gpr.py:1247-1261: -(kstar^T V^T V dkstar_dx) / sqrt(var_), multiplied by
 * y_std twice as the reference does; zero if the variance vanishes.  Host pointers. */
int gpry_std_grad(gpry_state* st, const double* x, double* out_grad, double* out_std);

/* Batched version of the two calls above for the acquisition optimiser's many starts
 * (gp_acquisition.py:280-390 runs them one L-BFGS-B at a time; acquisition_functions.py:
 * 967-1007 consumes mean, std and both gradients): X (M x d, un-transformed, 1 <= M <= 8192),
 * out_mean / out_std (M), out_grad_mean / out_grad_std (M x d); any output may be NULL.  Same
 * conventions per row as gpry_predict (no trust region), gpry_mean_grad and gpry_std_grad.
 * Host pointers. */
int gpry_predict_grad(gpry_state* st, const double* X, int M, double* out_mean, double* out_std,
                      double* out_grad_mean, double* out_grad_std);

/*
 * Posterior covariance (normalised units, prior variance c on the diagonal minus the explained
 * part, NO noise term) among Ka <= 8192 candidates X (Ka x d, un-transformed):
 *   Sigma = k(Xa, Xa) - (V K*a^T)^T (V K*a^T),   out_cov: Ka x Ka row major.
 * This is what RankedPool's Kriging-believer conditioning (gp_acquisition.py:1464-1494,
 * 1522-1555, 1598-1670: deepcopy + append_to_data + O((N+i)^3) refit per cached model) needs:
 * var(a | pool P) = Sigma_aa - Sigma_aP (Sigma_PP + noise2 I)^-1 Sigma_Pa.
 */
int gpry_posterior_cov(gpry_state* st, const double* X, int Ka, int where, double* out_cov,
                       void* stream);

/* k_theta(X, Y) for already-transformed host arrays X (M x d), Y (N x d) -> out (M x N):
 * Product.__call__ (sklearn:kernels.py:971).  API completeness; needs no uploaded model. */
int gpry_kernel_cross(gpry_state* st, int kind, int d, const double* theta, const double* X,
                      int M, const double* Y, int N, double* out);

/* d k(x, X_train_) / d x_ at one TRANSFORMED point x_t (d) -> out (N x d):
 * Kernel.gradient_x (kernels.py:257-278, 363-432, 687-699) of the uploaded model. */
int gpry_kernel_gradient_x(gpry_state* st, const double* x_t, double* out);

/*
 * K = k_theta(X_, X_) + diag(noise2); L = chol(K); V = L^-1; alpha_ = K^-1 y_.
 * theta = [log c, log ell_1..ell_d].  Host pointers; out_L / out_V (N x N row major, lower,
 * upper triangle zeroed) and out_alpha (N) may be NULL.  *info = 0 or the order of the
 * first non-positive leading minor.  If keep_on_device != 0 the factor stays resident in
 * `st` for gpry_state_adopt_factorization.  out_logdet_half = sum(log(diag L)).
 */
int gpry_factorize(gpry_state* st, int kind, int N, int d, const double* X_train_t,
                   const double* noise2, const double* y_t, const double* theta,
                   double* out_L, double* out_V, double* out_alpha,
                   double* out_logdet_half, int* info, int keep_on_device);

/*
 * Extends the factorisation kept resident by gpry_factorize(..., keep_on_device = 1) by k points
 * appended at the end of the training set, in O(k N^2) instead of O(N^3): what
 * _update_model (gpr.py:996-1020) amounts to when theta, the noise of the old points and the
 * pre-processors are unchanged -- the Kriging-believer lies of gp_acquisition.py:488-491.
 * X_new_t (k x d, transformed), noise2_new (k), y_t_all (N + k, all targets), theta as given to
 * gpry_factorize; out_alpha (N + k, may be NULL).  The padded size round_up(N, 128) must not
 * change (GPRY_ERR_ARG otherwise: refactorise).  *info > 0: not positive definite, the resident
 * factorisation is dropped.  Follow with gpry_state_adopt_factorization.  Host pointers.
 */
int gpry_factor_append(gpry_state* st, int k, const double* X_new_t, const double* noise2_new,
                       const double* y_t_all, const double* theta, double* out_alpha, int* info);

/* L and / or V (N x N row major, lower, upper triangle zeroed; either may be NULL) of the
 * factorisation kept resident by the last gpry_factorize(..., keep_on_device = 1) on this
 * state: lets the host attributes L_ / V_ (gpr.py:1456-1457) be fetched only when something
 * reads them.  GPRY_ERR_STATE if no factorisation is resident.  Host pointers. */
int gpry_factor_download(gpry_state* st, double* out_L, double* out_V);

/*
 * Log marginal likelihood (and its gradient w.r.t. theta if out_grad != NULL) for B
 * hyper-parameter vectors thetas (B x (1+d)).  out_lml (B), out_grad (B x (1+d)),
 * out_info (B; >0 = not positive definite: lml = -inf, grad = 0).  Host pointers.
 */
int gpry_lml_batched(gpry_state* st, int kind, int N, int d, const double* X_train_t,
                     const double* noise2, const double* y_t, const double* thetas, int B,
                     double* out_lml, double* out_grad, int* out_info);

/*
 * Multi-GPU exchange steps on an NCCL communicator owned by the library (one process per GPU;
 * NCCL is bound at run time with dlopen("libnccl.so.2"), so a process that imported PyTorch
 * shares PyTorch's copy).  They replace the reference's mpi4py traffic for this path:
 *
 *   gpry_comm_unique_id   rank 0 creates the 128-byte NCCL id; ship it to the other ranks by any
 *                         host channel (the reference's own MPI.COMM_WORLD.bcast, mpi.py:53-59, or
 *                         torch.distributed) ...
 *   gpry_comm_init        ... and every rank joins with (id, rank, nranks) on its state's GPU.
 *   gpry_bcast_state      the model uploaded into `root`'s state (X_train_, alpha_, V_ = L^-1,
 *                         kernel and pre-processor scalars) goes GPU -> GPU over NVLink into every
 *                         other rank's state, which is then ready for gpry_predict*: what the
 *                         scoring ranks need of the pickled regressor that run.py:749-756
 *                         (_share_gpr -> mpi.bcast) ships once per refit.  The device-side
 *                         classifier and trust region are not part of it (set them per rank).
 *   gpry_allgather_topk   every rank contributes n_local <= Kp survivor records (acq, idx, mean,
 *                         std, X[d]; host or device per GPRY_X_ON_DEVICE) and receives the Kp best
 *                         of the union, sorted by (descending acq, ascending idx), identical on
 *                         all ranks: gp_acquisition.py:1148-1171 (_gather_pools, five gathers) and
 *                         the bcast of :1190 in one ncclAllGather + a device merge.  *n_out =
 *                         records returned; *next_acq = the best acquisition value NOT returned
 *                         (-inf if none): the bound NORA's exact pre-selection test needs.
 *                         mean / std / X (and their outputs) may be NULL.
 */
int gpry_comm_unique_id(void* out128);
int gpry_comm_init(gpry_state* st, const void* id128, int rank, int nranks);
/* dst borrows src's communicator (same process and GPU; src must outlive its use): lets every
 * model state of a process use one communicator. */
int gpry_comm_share(gpry_state* dst, gpry_state* src);
int gpry_comm_destroy(gpry_state* st);
/* rank, size (0 = no communicator) and the NCCL version bound (e.g. 22809); any may be NULL */
int gpry_comm_info(const gpry_state* st, int* rank, int* nranks, int* nccl_version);
int gpry_bcast_state(gpry_state* st, int root, void* stream);
int gpry_allgather_topk(gpry_state* st, int n_local, int Kp, int d, const double* acq,
                        const int64_t* idx, const double* mean, const double* std_,
                        const double* X, int where, double* out_acq, int64_t* out_idx,
                        double* out_mean, double* out_std, double* out_X, int64_t* n_out,
                        double* next_acq, void* stream);

/* Device timings (ms, CUDA events on the launching stream) accumulated since the last
 * reset, per stage: [0] kstar build, [1] variance contraction, [2] finish/acquisition,
 * [3] top-k, [4] h2d, [5] d2h, [6] kernel launches counted, [7] contraction launches.
 * Enable with gpry_set_profiling(st, 1): adds event records around every launch. */
int gpry_set_profiling(gpry_state* st, int enable);
int gpry_get_timings(gpry_state* st, double* out8, int reset);

#ifdef __cplusplus
}
#endif
#endif /* GPRY_B200_H */
