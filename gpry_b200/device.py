"""
``DeviceGP``: one GP model resident on one B200, driven through the C ABI.

This is the thin layer between the reference-shaped Python classes (``gpry_b200.gpr``,
``gpry_b200.gp_acquisition``) and ``libgpry_b200.so``.  Inputs / outputs are numpy arrays
(host path: the library does the H2D / D2H copies) or torch CUDA tensors (device path: raw
device pointers are handed over, nothing is copied).
"""
import ctypes as C
import threading

import numpy as np

from . import _lib
from ._lib import KERNEL_KINDS, MAX_TOPK, check, as_f64, ptr


def _is_torch_cuda(a):
    return hasattr(a, "data_ptr") and getattr(a, "is_cuda", False)


def _stream_ptr(stream):
    if stream is None:
        return None
    if isinstance(stream, int):
        return C.c_void_p(stream)
    return C.c_void_p(stream.cuda_stream)  # torch.cuda.Stream


_workspaces = {}
_workspaces_lock = threading.Lock()


def workspace(device=0):
    """Process-wide DeviceGP whose big training-side buffers are shared by every regressor
    instance of this process (factorizations, LML, stand-alone kernel evaluations)."""
    with _workspaces_lock:
        if device not in _workspaces:
            _workspaces[device] = DeviceGP(device)
        return _workspaces[device]


class DeviceGP:
    """Opaque device state + the hot-path calls.  Not picklable by design: the owning
    ``GaussianProcessRegressor`` keeps it outside of its pickled/deep-copied attributes and
    rebuilds it lazily (SURVEY.md section 5 "Checkpoint / resume")."""

    def __init__(self, device=0):
        self._lib = _lib.load_library()
        h = C.c_void_p()
        check(self._lib.gpry_state_create(int(device), C.byref(h)))
        self._h = h
        self.device = int(device)
        self.N = self.d = 0
        self.kind = None

    def close(self):
        if getattr(self, "_h", None):
            self._lib.gpry_state_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __reduce__(self):
        raise TypeError("DeviceGP holds device memory and cannot be pickled")

    # ------------------------------------------------------------------ model upload
    def upload(self, kind, X_train_, alpha_, V_, c, ell, x_min=None, x_width=None,
               y_mean=0.0, y_std=1.0, clip_hi=np.inf):
        X_train_ = as_f64(X_train_)
        N, d = X_train_.shape
        alpha_ = as_f64(alpha_, (N,))
        V_ = None if V_ is None else as_f64(V_, (N, N))     # None: mean-only model
        ell = as_f64(np.broadcast_to(ell, (d,)))
        x_min = None if x_min is None else as_f64(x_min, (d,))
        x_width = None if x_width is None else as_f64(x_width, (d,))
        check(self._lib.gpry_state_upload(
            self._h, KERNEL_KINDS[kind], N, d, ptr(X_train_), ptr(alpha_), ptr(V_), float(c),
            ptr(ell), ptr(x_min), ptr(x_width), float(y_mean), float(y_std), float(clip_hi)))
        self.N, self.d, self.kind = N, d, kind

    def adopt_factorization(self, c, ell, x_min=None, x_width=None, y_mean=0.0, y_std=1.0,
                            clip_hi=np.inf):
        d = self._f_d
        ell = as_f64(np.broadcast_to(ell, (d,)))
        x_min = None if x_min is None else as_f64(x_min, (d,))
        x_width = None if x_width is None else as_f64(x_width, (d,))
        check(self._lib.gpry_state_adopt_factorization(
            self._h, float(c), ptr(ell), ptr(x_min), ptr(x_width), float(y_mean),
            float(y_std), float(clip_hi)))
        self.N, self.d, self.kind = self._f_N, d, self._f_kind

    def set_contract_mode(self, mode, guard=True):
        """"int8" (Ozaki split into 7 int8 digits on the INT8 tensor cores, FP64-equivalent;
        default) or "fp64" (DMMA) for the variance contraction of large pools.  With the guard
        on (default) a model whose estimated or probed INT8 error exceeds the tolerance takes
        the FP64 kernel anyway (see ``contract_info``)."""
        check(self._lib.gpry_set_contract_mode(self._h, {"fp64": 0, "int8": 1, "int8_1pass": 2}[mode]))
        check(self._lib.gpry_set_contract_guard(self._h, 1 if guard else 0))

    def contract_info(self):
        """What the INT8 guard decided for the uploaded model."""
        out = np.zeros(8)
        check(self._lib.gpry_contract_info(self._h, ptr(out)))
        names = {0: "fp64", 1: "int8", 2: "int8_1pass"}
        return {"requested": names[int(out[0])], "in_use": names[int(out[1])],
                "estimate_sigma": out[2], "bound_worst_case": out[3],
                "probe_diff": None if out[4] < 0 else out[4], "tolerance": out[5],
                "guard": bool(out[6])}

    def int8_peak_tops(self, seconds=None):
        """Measured tcgen05 INT8 MMA issue rate of this GPU (TOPS): roofline denominator.  A 2 ms
        burst by default; ``seconds``: launched back to back for that long (sustained under the
        power cap)."""
        out = C.c_double(0.0)
        if seconds is None:
            check(self._lib.gpry_int8_peak(self._h, C.byref(out)))
        else:
            check(self._lib.gpry_int8_peak_sustained(self._h, float(seconds), C.byref(out)))
        return out.value

    def set_mask_value(self, value):
        """The value masked rows get for the mean (``minus_inf_value``)."""
        check(self._lib.gpry_set_mask_value(self._h, float(value)))

    def set_classifier(self, spec=None):
        """Device-side infinities classifier: ``spec = (support vectors (n_sv, d) in the
        transformed space, dual_coef (n_sv,), intercept, gamma)`` or None to clear.  Cleared
        by every ``upload`` / ``adopt_factorization``."""
        if spec is None:
            check(self._lib.gpry_set_classifier(self._h, 0, 0, None, None, 0.0, 1.0))
            return
        sv, coef, intercept, gamma = spec
        sv = as_f64(sv)
        coef = as_f64(coef, (sv.shape[0],))
        check(self._lib.gpry_set_classifier(self._h, sv.shape[0], sv.shape[1], ptr(sv), ptr(coef),
                                            float(intercept), float(gamma)))

    def classify(self, X, stream=None):
        """Decision values of the device-side classifier (> 0: finite)."""
        X, M, where = self._prep_X(X)
        if where & _lib.X_ON_DEVICE:
            import torch
            out = torch.empty(M, dtype=torch.float64, device=X.device)
            where |= _lib.OUT_ON_DEVICE
        else:
            out = np.empty(M)
        check(self._lib.gpry_classify(self._h, ptr(X), M, where, ptr(out), _stream_ptr(stream)))
        return out

    def set_trust_region(self, bounds=None, value=-np.inf):
        """Device-side trust region (mean, and acq in the acquisition calls): (d, 2) bounds or
        None to clear."""
        if bounds is None:
            check(self._lib.gpry_set_trust_region(self._h, 0, None, None, 0.0))
            return
        b = as_f64(bounds)
        lo, hi = np.ascontiguousarray(b[:, 0]), np.ascontiguousarray(b[:, 1])
        check(self._lib.gpry_set_trust_region(self._h, b.shape[0], ptr(lo), ptr(hi),
                                              float(value)))

    # ------------------------------------------------------------------ candidate side
    def _prep_X(self, X):
        if self.kind is None:
            raise _lib.GpryB200Error("no model uploaded into this DeviceGP")
        if _is_torch_cuda(X):
            if X.dtype.is_floating_point and X.element_size() == 8 and X.is_contiguous():
                return X, int(X.shape[0]), _lib.X_ON_DEVICE
            raise ValueError("device candidates must be a contiguous float64 tensor")
        if hasattr(X, "data_ptr"):  # torch CPU tensor (e.g. pinned): use its memory directly
            if not (X.is_contiguous() and X.element_size() == 8):
                raise ValueError("host tensor candidates must be contiguous float64")
            return X, int(X.shape[0]), 0
        X = as_f64(X)
        if X.ndim != 2 or X.shape[1] != self.d:
            raise ValueError(f"X must be (M, {self.d}), got {X.shape}")
        return X, X.shape[0], 0

    def predict(self, X, return_mean=True, return_std=False, stream=None, out=None):
        """mean and/or std.  numpy in -> numpy out; torch CUDA in -> torch CUDA out."""
        X, M, where = self._prep_X(X)
        what = (_lib.WANT_MEAN if return_mean else 0) | (_lib.WANT_STD if return_std else 0)
        if where & _lib.X_ON_DEVICE:
            import torch
            mean = torch.empty(M, dtype=torch.float64, device=X.device) if return_mean else None
            std = torch.empty(M, dtype=torch.float64, device=X.device) if return_std else None
            where |= _lib.OUT_ON_DEVICE
        else:
            mean = np.empty(M) if return_mean else None
            std = np.empty(M) if return_std else None
        check(self._lib.gpry_predict(self._h, ptr(X), M, what, where, ptr(mean), ptr(std),
                                     _stream_ptr(stream)))
        return mean, std

    def predict_logexp(self, X, zeta, sigma_n, y_max, stream=None):
        X, M, where = self._prep_X(X)
        if where & _lib.X_ON_DEVICE:
            import torch
            outs = [torch.empty(M, dtype=torch.float64, device=X.device) for _ in range(3)]
            where |= _lib.OUT_ON_DEVICE
        else:
            outs = [np.empty(M) for _ in range(3)]
        check(self._lib.gpry_predict_logexp(self._h, ptr(X), M, float(zeta), float(sigma_n),
                                            float(y_max), where, ptr(outs[0]), ptr(outs[1]),
                                            ptr(outs[2]), _stream_ptr(stream)))
        return tuple(outs)

    def set_excluded(self, rows=None):
        """Row numbers (sorted, within the pool given to the next ``predict_logexp_topk``
        calls) that are skipped by the ranking; ``None`` clears the list."""
        if rows is None or len(rows) == 0:
            check(self._lib.gpry_set_excluded(self._h, None, 0))
            return
        rows = np.ascontiguousarray(np.sort(np.asarray(rows, dtype=np.int64)))
        check(self._lib.gpry_set_excluded(self._h, ptr(rows), len(rows)))

    def predict_logexp_topk(self, X, zeta, sigma_n, y_max, Kp, idx_offset=0, stream=None,
                            device_out=False, want_X=True, exclude=None):
        """The Kp best candidates by LogExp: (acq, idx, mean, std, X) sorted by descending
        acq.  Only these records leave the GPU (``device_out``: they stay in torch tensors).
        ``exclude``: rows of X left out of the ranking."""
        X, M, where = self._prep_X(X)
        self.set_excluded(exclude)
        Kp = int(min(Kp, MAX_TOPK))
        n_out = C.c_int64(0)
        d = self.d
        if device_out:
            import torch
            dev = X.device if _is_torch_cuda(X) else torch.device("cuda", self.device)
            acq = torch.empty(Kp, dtype=torch.float64, device=dev)
            idx = torch.empty(Kp, dtype=torch.int64, device=dev)
            mean = torch.empty(Kp, dtype=torch.float64, device=dev)
            std = torch.empty(Kp, dtype=torch.float64, device=dev)
            Xo = torch.empty((Kp, d), dtype=torch.float64, device=dev) if want_X else None
            where |= _lib.OUT_ON_DEVICE
        else:
            acq, mean, std = np.empty(Kp), np.empty(Kp), np.empty(Kp)
            idx = np.empty(Kp, dtype=np.int64)
            Xo = np.empty((Kp, d)) if want_X else None
        check(self._lib.gpry_predict_logexp_topk(
            self._h, ptr(X), M, float(zeta), float(sigma_n), float(y_max), Kp, int(idx_offset),
            where, ptr(acq), ptr(idx), ptr(mean), ptr(std), ptr(Xo), C.byref(n_out),
            _stream_ptr(stream)))
        n = n_out.value
        return (acq[:n], idx[:n], mean[:n], std[:n], None if Xo is None else Xo[:n])

    def topk(self, scores, Kp, stream=None):
        Kp = int(min(Kp, MAX_TOPK))
        n_out = C.c_int64(0)
        if _is_torch_cuda(scores):
            import torch
            M = scores.numel()
            vals = torch.empty(Kp, dtype=torch.float64, device=scores.device)
            idx = torch.empty(Kp, dtype=torch.int64, device=scores.device)
            where = _lib.X_ON_DEVICE | _lib.OUT_ON_DEVICE
        else:
            scores = as_f64(scores)
            M = scores.size
            vals, idx, where = np.empty(Kp), np.empty(Kp, dtype=np.int64), 0
        check(self._lib.gpry_topk(self._h, ptr(scores), M, Kp, where, ptr(vals), ptr(idx),
                                  C.byref(n_out), _stream_ptr(stream)))
        return vals[:n_out.value], idx[:n_out.value]

    def mean_grad(self, x):
        x = as_f64(x).reshape(-1)
        if x.size != self.d:
            raise ValueError(f"x must have {self.d} entries")
        out = np.empty(self.d)
        check(self._lib.gpry_mean_grad(self._h, ptr(x), ptr(out)))
        return out

    def std_grad(self, x):
        """(d std/dx_ (d,), std) at one un-transformed point (gpr.py:1247-1261)."""
        x = as_f64(x).reshape(-1)
        if x.size != self.d:
            raise ValueError(f"x must have {self.d} entries")
        out = np.empty(self.d)
        std = C.c_double(0.0)
        check(self._lib.gpry_std_grad(self._h, ptr(x), ptr(out), C.byref(std)))
        return out, std.value

    def predict_grad(self, X, return_std_grad=True):
        """Batched mean, std, d mean/dx_ and d std/dx_ for the rows of a host array X
        (M <= 8192); the per-row conventions of ``predict``, ``mean_grad`` and ``std_grad``."""
        X = as_f64(X)
        if X.ndim != 2 or X.shape[1] != self.d:
            raise ValueError(f"X must have shape (M, {self.d})")
        M = X.shape[0]
        mean, std = np.empty(M), np.empty(M)
        gm = np.empty((M, self.d))
        gs = np.empty((M, self.d)) if return_std_grad else None
        check(self._lib.gpry_predict_grad(self._h, ptr(X), M, ptr(mean), ptr(std), ptr(gm),
                                          ptr(gs) if gs is not None else None))
        return mean, std, gm, gs

    def posterior_cov(self, X, stream=None):
        """Posterior covariance (normalised units, no noise) among the rows of X (Ka <= 8192)."""
        X, Ka, where = self._prep_X(X)
        if where & _lib.X_ON_DEVICE:
            import torch
            out = torch.empty((Ka, Ka), dtype=torch.float64, device=X.device)
            where |= _lib.OUT_ON_DEVICE
        else:
            out = np.empty((Ka, Ka))
        check(self._lib.gpry_posterior_cov(self._h, ptr(X), Ka, where, ptr(out),
                                           _stream_ptr(stream)))
        return out

    def kernel_cross(self, kind, theta, X_, Y_):
        """k_theta(X_, Y_) for transformed host arrays (no uploaded model needed)."""
        X_, Y_ = as_f64(np.atleast_2d(X_)), as_f64(np.atleast_2d(Y_))
        theta = as_f64(theta, (X_.shape[1] + 1,))
        out = np.empty((X_.shape[0], Y_.shape[0]))
        check(self._lib.gpry_kernel_cross(self._h, KERNEL_KINDS[kind], X_.shape[1], ptr(theta),
                                          ptr(X_), X_.shape[0], ptr(Y_), Y_.shape[0], ptr(out)))
        return out

    def kernel_gradient_x(self, x_t):
        """d k(x_, X_train_)/d x_ (N x d) of the uploaded model at a transformed point."""
        x_t = as_f64(x_t).reshape(-1)
        out = np.empty((self.N, self.d))
        check(self._lib.gpry_kernel_gradient_x(self._h, ptr(x_t), ptr(out)))
        return out

    # ------------------------------------------------------------------ training side
    def factorize(self, kind, X_train_, noise2, y_train_, theta, want_L=True, want_V=True,
                  keep_on_device=False):
        """K = k_theta(X_,X_) + diag(noise2) -> (L_, V_, alpha_, sum(log diag L), info)."""
        X_train_ = as_f64(X_train_)
        N, d = X_train_.shape
        noise2 = as_f64(np.broadcast_to(noise2, (N,)))
        y_train_ = as_f64(y_train_, (N,))
        theta = as_f64(theta, (d + 1,))
        L = np.empty((N, N)) if want_L else None
        V = np.empty((N, N)) if want_V else None
        alpha_ = np.empty(N)
        logdet_half = C.c_double(0.0)
        info = C.c_int(0)
        check(self._lib.gpry_factorize(
            self._h, KERNEL_KINDS[kind], N, d, ptr(X_train_), ptr(noise2), ptr(y_train_),
            ptr(theta), ptr(L), ptr(V), ptr(alpha_), C.byref(logdet_half), C.byref(info),
            1 if keep_on_device else 0))
        if keep_on_device and info.value == 0:
            self._f_N, self._f_d, self._f_kind = N, d, kind
        return L, V, alpha_, logdet_half.value, info.value

    def factor_append(self, X_new_, noise2_new, y_all_, theta):
        """Extend the resident factorisation by the rows of X_new_ -> (alpha_, info)."""
        X_new_ = np.atleast_2d(as_f64(X_new_))
        k, d = X_new_.shape
        noise2_new = as_f64(np.broadcast_to(noise2_new, (k,)))
        N2 = self._f_N + k
        y_all_ = as_f64(y_all_, (N2,))
        theta = as_f64(theta, (d + 1,))
        alpha_ = np.empty(N2)
        info = C.c_int(0)
        check(self._lib.gpry_factor_append(self._h, k, ptr(X_new_), ptr(noise2_new), ptr(y_all_),
                                           ptr(theta), ptr(alpha_), C.byref(info)))
        if info.value == 0:
            self._f_N = N2
        return alpha_, info.value

    def factor_download(self, want_L=True, want_V=True):
        """(L_, V_) of the factorisation kept on the device by ``factorize(keep_on_device=True)``."""
        N = self._f_N
        L = np.empty((N, N)) if want_L else None
        V = np.empty((N, N)) if want_V else None
        check(self._lib.gpry_factor_download(self._h, ptr(L), ptr(V)))
        return L, V

    def lml_batched(self, kind, X_train_, noise2, y_train_, thetas, eval_gradient=True):
        X_train_ = as_f64(X_train_)
        N, d = X_train_.shape
        noise2 = as_f64(np.broadcast_to(noise2, (N,)))
        y_train_ = as_f64(y_train_, (N,))
        thetas = np.atleast_2d(as_f64(thetas))
        B = thetas.shape[0]
        if thetas.shape[1] != d + 1:
            raise ValueError(f"thetas must be (B, {d + 1})")
        lml = np.empty(B)
        grad = np.empty((B, d + 1)) if eval_gradient else None
        info = np.zeros(B, dtype=np.int32)
        check(self._lib.gpry_lml_batched(
            self._h, KERNEL_KINDS[kind], N, d, ptr(X_train_), ptr(noise2), ptr(y_train_),
            ptr(thetas), B, ptr(lml), ptr(grad), ptr(info)))
        return lml, grad, info

    # ------------------------------------------------------------------ multi-GPU exchange
    def comm_unique_id(self):
        """128-byte NCCL id (create on rank 0, ship to the others over any host channel)."""
        buf = C.create_string_buffer(128)
        check(self._lib.gpry_comm_unique_id(buf))
        return buf.raw

    def comm_init(self, uid, rank, nranks):
        check(self._lib.gpry_comm_init(self._h, C.c_char_p(uid), int(rank), int(nranks)))

    def comm_share(self, other):
        """Use ``other``'s communicator (same process, same GPU)."""
        check(self._lib.gpry_comm_share(self._h, other._h))

    def comm_info(self):
        r, n, v = C.c_int(0), C.c_int(0), C.c_int(0)
        check(self._lib.gpry_comm_info(self._h, C.byref(r), C.byref(n), C.byref(v)))
        return {"rank": r.value, "size": n.value, "nccl_version": v.value}

    def bcast_state(self, root=0, stream=None):
        """The model uploaded on ``root`` -> this state on every rank, GPU to GPU (NCCL)."""
        check(self._lib.gpry_bcast_state(self._h, int(root), _stream_ptr(stream)))
        info = [C.c_int(0), C.c_int(0), C.c_int(0)]
        check(self._lib.gpry_state_info(self._h, *[C.byref(i) for i in info]))
        self.N, self.d = info[0].value, info[1].value
        self.kind = {v: k for k, v in KERNEL_KINDS.items()}[info[2].value]

    def allgather_topk(self, acq, idx, mean, std, X, Kp, d=None, stream=None):
        """Every rank's survivor records -> the Kp best of the union (numpy arrays, the same on
        every rank) and the best acquisition value left out: (acq, idx, mean, std, X, next)."""
        n = len(acq)
        on_dev = _is_torch_cuda(acq)
        d = (X.shape[1] if X is not None else 0) if d is None else d
        if not on_dev:
            acq, mean, std = as_f64(acq), as_f64(mean), as_f64(std)
            idx = np.ascontiguousarray(idx, dtype=np.int64)
            X = as_f64(X)
        Kp = int(Kp)
        o_acq, o_mean, o_std = np.empty(Kp), np.empty(Kp), np.empty(Kp)
        o_idx, o_X = np.empty(Kp, dtype=np.int64), np.empty((Kp, d))
        n_out, nxt = C.c_int64(0), C.c_double(0.0)
        check(self._lib.gpry_allgather_topk(
            self._h, n, Kp, d, ptr(acq), ptr(idx), ptr(mean), ptr(std), ptr(X),
            _lib.X_ON_DEVICE if on_dev else 0, ptr(o_acq), ptr(o_idx), ptr(o_mean), ptr(o_std),
            ptr(o_X), C.byref(n_out), C.byref(nxt), _stream_ptr(stream)))
        k = n_out.value
        return o_acq[:k], o_idx[:k], o_mean[:k], o_std[:k], o_X[:k], nxt.value

    # ------------------------------------------------------------------ profiling
    def set_profiling(self, enable=True):
        check(self._lib.gpry_set_profiling(self._h, 1 if enable else 0))

    def timings(self, reset=True):
        out = np.zeros(8)
        check(self._lib.gpry_get_timings(self._h, ptr(out), 1 if reset else 0))
        keys = ["build_ms", "contract_ms", "finish_ms", "topk_ms", "h2d_ms", "d2h_ms",
                "launches", "contract_launches"]
        return dict(zip(keys, out.tolist()))
