"""
CPU ORACLE for the GPry GP-surrogate hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

This file is a numpy/scipy restatement of the reference algorithm (GPry 3.0.0 at
``/root/reference`` plus the scikit-learn 1.9.0 kernel / LML arithmetic that GPry inherits;
scikit-learn is an un-vendored, un-pinned dependency of the reference: ``pyproject.toml:34-37``).
It exists so that the CUDA path can be checked on a machine where the reference itself is
not present (the GPU box).  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  The product
package ``gpry_b200`` never imports it and has no CPU fallback.

Parity status: PINNED.  ``oracle/gen_golden.py`` imports the real reference (in the build
container) and writes ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks every
function below against those vectors.

Every function cites the reference lines it follows.  ``sklearn:`` means
``site-packages/sklearn/gaussian_process/`` of scikit-learn 1.9.0.

The operation ORDER deliberately mirrors the reference (divide-then-difference distances,
``np.full`` constant kernel times the stationary kernel, ``dtrmm`` + ``einsum`` variance...),
both because floating-point parity depends on it and because this module doubles as the
"port" CPU baseline whose cost profile should look like the reference's.
"""
from copy import deepcopy
from math import sqrt

import numpy as np
from scipy.linalg import cholesky, solve_triangular, cho_solve
from scipy.linalg.blas import dtrmm
from scipy.spatial.distance import cdist, pdist, squareform

KINDS = ("rbf", "matern15", "matern25")


# --------------------------------------------------------------------------------------
# Pre-processing  (gpry/preprocessing.py)
# --------------------------------------------------------------------------------------
def normalize_bounds_transform(X, bounds):
    """``Normalize_bounds.transform`` -- preprocessing.py:380.  ``bounds`` is (d, 2)."""
    bounds = np.asarray(bounds)
    return (X - bounds[:, 0]) / (bounds[:, 1] - bounds[:, 0])


def normalize_y_fit(y):
    """``Normalize_y.fit`` -- preprocessing.py:561,586: population mean / std of finite y."""
    y = np.asarray(y)
    y = y[np.isfinite(y)]
    return np.mean(y), np.std(y)


# --------------------------------------------------------------------------------------
# Kernels: ConstantKernel * {RBF, Matern(nu=1.5), Matern(nu=2.5)}, anisotropic
#   theta = [log c, log l_1 .. log l_d]   (Product concatenation, sklearn:kernels.py:739-753)
# --------------------------------------------------------------------------------------
def split_theta(theta):
    theta = np.asarray(theta, dtype=float)
    return float(np.exp(theta[0])), np.exp(theta[1:])


def _stationary(kind, dists):
    """g(r) given ``dists`` = sq. distance (rbf) or distance (matern)."""
    if kind == "rbf":  # sklearn:kernels.py:1570
        return np.exp(-0.5 * dists)
    if kind == "matern15":  # sklearn:kernels.py:1724-1726
        K = dists * sqrt(3)
        return (1.0 + K) * np.exp(-K)
    if kind == "matern25":  # sklearn:kernels.py:1727-1729
        K = dists * sqrt(5)
        return (1.0 + K + K ** 2 / 3.0) * np.exp(-K)
    raise ValueError(f"unknown kernel kind {kind!r}")


def kernel_cross(kind, theta, X, Y):
    """k(X, Y), shapes (M,d),(N,d) -> (M,N).

    Product.__call__ sklearn:kernels.py:971 = ConstantKernel (np.full, :1278) * stationary
    kernel on ``cdist(X / l, Y / l)`` (:1569 sqeuclidean for RBF, :1720 euclidean for Matern).
    """
    c, ell = split_theta(theta)
    metric = "sqeuclidean" if kind == "rbf" else "euclidean"
    dists = cdist(X / ell, Y / ell, metric=metric)
    K2 = _stationary(kind, dists)
    K1 = np.full((X.shape[0], Y.shape[0]), c)
    return K1 * K2


def kernel_diag(kind, theta, X):
    """k(x, x) = c  -- Product.diag sklearn:kernels.py:973-990, Constant.diag :1298-1322."""
    c, _ = split_theta(theta)
    return np.full(X.shape[0], c) * np.ones(X.shape[0])


def kernel_train(kind, theta, X, eval_gradient=False):
    """k(X, X) (no noise term) and optionally dK/dtheta of shape (N, N, 1 + d).

    Uses pdist + squareform + unit diagonal (sklearn:kernels.py:1561-1566 RBF, 1716-1744
    Matern); gradients sklearn:kernels.py:1581-1584 (RBF anisotropic), :1752-1771 (Matern),
    stacked by Product (:964-969) after the ConstantKernel gradient (:1283-1292).
    """
    c, ell = split_theta(theta)
    N = X.shape[0]
    if kind == "rbf":
        dists = pdist(X / ell, metric="sqeuclidean")
    else:
        dists = pdist(X / ell, metric="euclidean")
    K2 = squareform(_stationary(kind, dists))
    np.fill_diagonal(K2, 1)
    K1 = np.full((N, N), c)
    K = K1 * K2
    if not eval_gradient:
        return K
    D = (X[:, np.newaxis, :] - X[np.newaxis, :, :]) ** 2 / (ell ** 2)
    if kind == "rbf":
        K2_gradient = D * K2[..., np.newaxis]
    elif kind == "matern15":
        K2_gradient = 3 * D * np.exp(-np.sqrt(3 * D.sum(-1)))[..., np.newaxis]
    else:
        tmp = np.sqrt(5 * D.sum(-1))[..., np.newaxis]
        K2_gradient = 5.0 / 3.0 * D * (tmp + 1) * np.exp(-tmp)
    K1_gradient = np.full((N, N, 1), c)
    K_gradient = np.dstack((K1_gradient * K2[:, :, np.newaxis],
                            K2_gradient * K1[:, :, np.newaxis]))
    return K, K_gradient


def kernel_gradient_x(kind, theta, x, X_train):
    """d k(x, X_train) / d x  -> (N, d), w.r.t. the *transformed* coordinate.

    Product.gradient_x kernels.py:687-699 with ConstantKernel.gradient_x = 0 (:608-609) and
    RBF :257-278, Matern nu=1.5 :363-393, nu=2.5 :395-432.
    """
    c, ell = split_theta(theta)
    x = np.asarray(x, dtype=float)
    diff = x - X_train
    diff /= ell
    if kind == "rbf":
        e = np.sum(diff ** 2, axis=1)
        e *= -0.5
        e = np.exp(e)
        e = np.expand_dims(e, axis=1)
        e *= -1
        grad = e * diff
        grad /= ell
    else:
        dist_sq = np.sum(diff ** 2, axis=1)
        dist = np.sqrt(dist_sq)
        if kind == "matern15":
            sqrt_3_dist = sqrt(3) * dist
            f = np.expand_dims(1 + sqrt_3_dist, axis=1)
            sqrt_3_by_dist = np.zeros_like(dist)
            nzd = dist != 0.0
            sqrt_3_by_dist[nzd] = sqrt(3) / dist[nzd]
            f_grad = diff / ell
            f_grad *= np.expand_dims(sqrt_3_by_dist, axis=1)
            g = np.expand_dims(np.exp(-sqrt_3_dist), axis=1)
            f = 1 - f
            grad = g * f_grad * f
        else:
            sqrt_5_dist = sqrt(5) * dist
            f2 = (5.0 / 3.0) * dist_sq
            f2 += sqrt_5_dist
            f2 += 1
            f = np.expand_dims(f2, axis=1)
            inv = np.zeros_like(dist)
            nzd = dist != 0.0
            inv[nzd] = 1.0 / dist[nzd]
            inv *= sqrt(5)
            inv = np.expand_dims(inv, axis=1)
            diff = diff / ell
            f1_grad = inv * diff
            f2_grad = (10.0 / 3.0) * diff
            f_grad = f1_grad + f2_grad
            g = np.expand_dims(np.exp(-sqrt_5_dist), axis=1)
            g_grad = -g * f1_grad
            grad = f * g_grad + g * f_grad
    # Product rule with the constant kernel: k1(x, X)[:, None] * k2.gradient_x + 0
    return np.full((X_train.shape[0], 1), c) * grad


# --------------------------------------------------------------------------------------
# Fit state:  K + diag(noise) -> L, V = L^-1, alpha_   (gpr.py:1015-1017, 1453-1465)
# --------------------------------------------------------------------------------------
class GPState:
    """Plain container of what ``GaussianProcessRegressor`` holds after ``_update_model``.

    kind, theta            kernel (``kernel_``)
    bounds (d,2) or None   ``Normalize_bounds`` (None = DummyPreprocessor)
    y_mean, y_std          ``Normalize_y.mean_/std_`` (0, 1 = no y preprocessing)
    X_train_ (N,d), y_train_ (N,), noise2 (N,) = ``alpha`` (gpr.py:747)
    L_, V_ (N,N), alpha_ (N,)
    y_train (N,) untransformed (for the clip and ``y_max``); noise_level; clip_factor
    """

    def __init__(self, kind, theta, X_train, y_train, bounds=None, normalize_y=True,
                 noise_level=1e-2, clip_factor=1.1, y_mean=None, y_std=None):
        assert kind in KINDS
        self.kind = kind
        self.theta = np.array(theta, dtype=float)
        self.bounds = None if bounds is None else np.array(bounds, dtype=float)
        self.X_train = np.array(X_train, dtype=float)
        self.y_train = np.array(y_train, dtype=float)
        self.noise_level = noise_level
        self.clip_factor = clip_factor
        if normalize_y:
            if y_mean is None:
                y_mean, y_std = normalize_y_fit(self.y_train)
            self.y_mean, self.y_std = float(y_mean), float(y_std)
        else:
            self.y_mean, self.y_std = 0.0, 1.0
        self.normalize_y = normalize_y
        self.X_train_ = self.transform_X(self.X_train)
        self.y_train_ = (self.y_train - self.y_mean) / self.y_std if normalize_y \
            else self.y_train.copy()
        # gpr.py:711-715,747: noise_level_ = noise / std_ ; alpha = noise_level_**2
        nl = np.full(len(self.y_train), noise_level) if np.isscalar(noise_level) \
            else np.asarray(noise_level, dtype=float)
        self.noise2 = ((nl / self.y_std) if normalize_y else nl) ** 2
        self.update_model()

    @property
    def d(self):
        return self.X_train.shape[1]

    @property
    def y_max(self):  # gpr.py:399-402
        return np.max(self.y_train)

    def transform_X(self, X):
        return X if self.bounds is None else normalize_bounds_transform(X, self.bounds)

    def update_model(self):
        """gpr.py:1015-1017 then ``_kernel_inverse`` :1453-1465."""
        K = kernel_train(self.kind, self.theta, self.X_train_)
        K[np.diag_indices_from(K)] += self.noise2
        self.L_ = cholesky(K, lower=True)
        self.V_ = solve_triangular(self.L_, np.eye(self.L_.shape[0]), lower=True)
        self.alpha_ = cho_solve((self.L_, True), self.y_train_)

    def appended(self, X_new, y_new):
        """A copy of this state with lie points appended, hyper-parameters and
        pre-processors unchanged: ``RankedPool.cache_model`` gp_acquisition.py:1550-1553
        (``deepcopy`` + ``append_to_data(..., fit_gpr=False, fit_classifier=False)``)."""
        new = deepcopy(self)
        new.X_train = np.append(self.X_train, np.atleast_2d(X_new), axis=0)
        new.y_train = np.append(self.y_train, np.atleast_1d(y_new))
        new.X_train_ = new.transform_X(new.X_train)
        new.y_train_ = (new.y_train - new.y_mean) / new.y_std if new.normalize_y \
            else new.y_train.copy()
        nl = np.full(len(new.y_train), new.noise_level)
        new.noise2 = ((nl / new.y_std) if new.normalize_y else nl) ** 2
        new.update_model()
        return new


# --------------------------------------------------------------------------------------
# Predict  (gpr.py:1176-1266; predict_std :1325-1347)
# --------------------------------------------------------------------------------------
def predict(st, X, return_std=False, return_mean_grad=False, return_std_grad=False):
    """Posterior mean [, std [, d mean/dx [, d std/dx]]] at un-transformed X (M, d).

    No infinities classifier / trust region: those are host-side masks applied around this
    arithmetic (gpr.py:1136-1174, 1196-1201) and are mirrored in ``gpry_b200.gpr``.
    """
    X = np.asarray(X, dtype=float)
    if return_std_grad and not (return_std and return_mean_grad):
        raise ValueError("Not returning std_gradient without returning the std and the "
                         "mean grad.")
    if X.shape[0] != 1 and (return_mean_grad or return_std_grad):
        raise ValueError("Mean grad and std grad not implemented for n_samples > 1")
    X_ = st.transform_X(X)
    K_trans = kernel_cross(st.kind, st.theta, X_, st.X_train_)        # gpr.py:1179
    y_mean_ = K_trans.dot(st.alpha_)                                   # :1180
    y_mean = y_mean_ * st.y_std + st.y_mean if st.normalize_y else y_mean_   # :1185
    if st.clip_factor is not None:                                     # :1187-1195
        y_mean = np.clip(y_mean, None,
                         st.clip_factor * max(st.y_train)
                         - (st.clip_factor - 1) * min(st.y_train))
    out = [y_mean]
    if return_std:
        M = dtrmm(1., st.V_, K_trans.T, lower=True)                    # :1204
        y_var = kernel_diag(st.kind, st.theta, X_)                     # :1207
        y_var -= np.einsum("ji,ji->i", M, M, optimize=True)            # :1208
        y_var[y_var < 0] = 0.0                                         # :1214-1219
        y_std_ = np.sqrt(y_var)
        y_std = y_std_ * st.y_std if st.normalize_y else y_std_        # :1225-1227
        out.append(y_std)
    if return_mean_grad:
        grad = kernel_gradient_x(st.kind, st.theta, X_[0], st.X_train_)   # :1237
        grad_mean = np.dot(grad.T, st.alpha_)                          # :1238
        if st.normalize_y:
            grad_mean = grad_mean * st.y_std                           # :1240-1242
        out.append(grad_mean)
        if return_std_grad:                                            # :1247-1261
            grad_std = np.zeros(X_.shape[1])
            if not np.allclose(y_std, grad_std):
                grad_std = -np.dot(K_trans, np.dot(st.V_.T.dot(st.V_), grad))[0] / y_std_
                if st.normalize_y:
                    grad_std = grad_std * st.y_std * st.y_std          # applied twice
            out.append(grad_std)
    return out[0] if len(out) == 1 else tuple(out)


def predict_std(st, X):
    """gpr.py:1325-1347 (no mean, no trust region)."""
    X_ = st.transform_X(np.asarray(X, dtype=float))
    K_trans = kernel_cross(st.kind, st.theta, X_, st.X_train_)
    M = dtrmm(1., st.V_, K_trans.T, lower=True)
    y_var = kernel_diag(st.kind, st.theta, X_)
    y_var -= np.einsum("ji,ji->i", M, M, optimize=True)
    y_var[y_var < 0] = 0.0
    y_std = np.sqrt(y_var)
    return y_std * st.y_std if st.normalize_y else y_std


# --------------------------------------------------------------------------------------
# LogExp acquisition  (acquisition_functions.py:933-934, 974-992, 1068-1074)
# --------------------------------------------------------------------------------------
def auto_zeta(d, scaling=0.85):
    return d ** (-scaling)


def logexp_f(mu, std, baseline, noise_level, zeta):
    """``LogExp.f`` acquisition_functions.py:1068-1074 (static)."""
    with np.errstate(divide="ignore"):
        return (2 * zeta * (mu - baseline) +
                np.log(np.sqrt(np.clip(std ** 2. - noise_level ** 2., 0., None))))


def logexp_call(mu, std, baseline, noise_level, zeta):
    """``BaseLogExp.__call__`` value branch, acquisition_functions.py:974-992: ``f`` where
    ``std**2 - sigma_n**2 > 0`` and ``mu`` finite, ``-inf`` elsewhere."""
    noise_var = np.mean(noise_level) if np.iterable(noise_level) else noise_level
    var = std ** 2 - noise_var ** 2.
    mask = (var > 0) & np.isfinite(mu)
    values = np.zeros_like(std)
    if np.any(mask):
        values[mask] = logexp_f(mu[mask], std[mask], baseline, noise_var, zeta)
    if np.any(~mask):
        values[~mask] = -np.inf
    return values


def predict_logexp(st, X, zeta=None):
    """predict(return_std) + ``LogExp.f`` exactly as NORA evaluates it
    (mpi.py:195 then gp_acquisition.py:1049-1051,1123-1124)."""
    zeta = auto_zeta(st.d) if zeta is None else zeta
    mu, std = predict(st, X, return_std=True)
    acq = logexp_f(mu, std, st.y_max, st.noise_level, zeta)
    return mu, std, acq


# --------------------------------------------------------------------------------------
# Log marginal likelihood + gradient  (sklearn:_gpr.py:584-651 via gpr.py:876-881)
# --------------------------------------------------------------------------------------
def log_marginal_likelihood(kind, theta, X_train_, y_train_, noise2, eval_gradient=False):
    theta = np.asarray(theta, dtype=float)
    if eval_gradient:
        K, K_gradient = kernel_train(kind, theta, X_train_, eval_gradient=True)
    else:
        K = kernel_train(kind, theta, X_train_)
    K[np.diag_indices_from(K)] += noise2
    try:
        L = cholesky(K, lower=True, check_finite=False)
    except np.linalg.LinAlgError:
        return (-np.inf, np.zeros_like(theta)) if eval_gradient else -np.inf
    y_train = y_train_[:, np.newaxis]
    alpha = cho_solve((L, True), y_train, check_finite=False)
    lml = -0.5 * np.einsum("ik,ik->k", y_train, alpha)
    lml -= np.log(np.diag(L)).sum()
    lml -= K.shape[0] / 2 * np.log(2 * np.pi)
    lml = lml.sum(axis=-1)
    if not eval_gradient:
        return lml
    inner_term = np.einsum("ik,jk->ijk", alpha, alpha)
    K_inv = cho_solve((L, True), np.eye(K.shape[0]), check_finite=False)
    inner_term -= K_inv[..., np.newaxis]
    grad = 0.5 * np.einsum("ijl,jik->kl", inner_term, K_gradient)
    return lml, grad.sum(axis=-1)


# --------------------------------------------------------------------------------------
# Ranked pool with Kriging-believer conditioning (gp_acquisition.py:1194-1670)
# --------------------------------------------------------------------------------------
class RankedPool:
    """Restatement of ``RankedPool`` for ``add(method="single sort acq")`` and ``"bulk"``.

    ``st`` is a :class:`GPState`; ``acq_func(y, sigma)`` the partial of ``LogExp.f``
    (gp_acquisition.py:1049-1051).  Conditioned models are ``st.appended(pool.X[:i+1],
    pool.y[:i+1])`` (``cache_model`` :1522-1555).
    """

    def __init__(self, size, st, acq_func):
        self._st, self._acq_func = st, acq_func
        self.X = np.zeros((size + 1, st.d))
        self.y = np.zeros(size + 1)
        self.acq_cond = np.full(size + 1, -np.inf)
        self.sigma = np.zeros(size + 1)
        self.acq = np.zeros(size + 1)
        self.idx = np.full(size + 1, -1, dtype=np.int64)   # bookkeeping only (not in ref)
        self.st_cond = [None] * (size + 1)
        self.cache_counter = 0

    def __len__(self):
        return len(self.y) - 1

    @property
    def min_acq(self):  # :1238-1247
        return self.acq_cond[len(self) - 1]

    def cache_model(self, i):  # :1522-1555
        if i < 0:
            return self._st
        self.st_cond[i] = self._st.appended(self.X[:i + 1], self.y[:i + 1])
        self.cache_counter += 1
        return self.st_cond[i]

    def add(self, X, y, sigma, acq, method="single sort acq", idx=None):  # :1290-1335
        X = np.atleast_2d(X)
        idx = np.arange(len(X)) if idx is None else np.asarray(idx)
        if method == "bulk":
            self.add_bulk(X, y, sigma, acq, idx)
            return
        i_sort = np.argsort(acq)[::-1] if "sort acq" in method else range(len(X))
        for i in i_sort:
            self.add_one(X[i], y[i], sigma[i], acq[i], idx[i])

    def add_bulk(self, X, y, sigma, acq, idx, i_start=0):  # :1337-1390
        if i_start == 0:
            acq_cond = np.asarray(acq)
        else:
            st = self.cache_model(i_start - 1)
            sigma_cond = predict_std(st, X)
            acq_cond = self._acq_func(y, sigma_cond)
        if acq_cond.size == 0:
            return
        i_max = np.argmax(acq_cond)
        acq_cond_max = acq_cond[i_max]
        if acq_cond_max == np.inf:
            return
        self.X[i_start], self.y[i_start] = X[i_max], y[i_max]
        self.sigma[i_start], self.acq[i_start] = sigma[i_max], acq[i_max]
        self.acq_cond[i_start], self.idx[i_start] = acq_cond_max, idx[i_max]
        if i_start == len(self) - 1:
            return
        keep = np.logical_not(acq_cond == -np.inf)
        keep[i_max] = False
        self.add_bulk(X[keep], y[keep], sigma[keep], acq[keep], idx[keep],
                      i_start=i_start + 1)

    def add_one(self, X, y, sigma, acq, idx=-1):  # :1392-1520
        if acq <= self.min_acq:
            return
        if np.isnan(acq):
            raise ValueError(f"Acquisition function value not a number: {acq}")
        X = np.atleast_2d(X)
        n = len(self)
        i_new_last = n
        acq_cond = acq
        while True:
            try:
                i_new = n - next(i for i in range(n) if self.acq_cond[-(i + 2)] >= acq_cond)
            except StopIteration:
                i_new = 0
            if i_new in [0, i_new_last, n]:
                break
            sigma_cond = predict_std(self.st_cond[i_new - 1], X)[0]
            acq_cond = min(acq_cond, self._acq_func(y, sigma_cond))
            i_new_last = i_new
        if i_new >= n:
            return
        for pool, value in [(self.X, X), (self.y, y), (self.sigma, sigma),
                            (self.acq, acq), (self.acq_cond, acq_cond), (self.idx, idx)]:
            pool[i_new + 1:] = pool[i_new:-1]
            pool[i_new] = value
        assert self.acq_cond[i_new] > -np.inf
        self.sort(i_new + 1)
        self.acq_cond[-1] = -np.inf

    def sort(self, i_start=0):  # :1598-1670
        if i_start >= len(self):
            return
        upper = self.cache_model(i_start - 1)
        if self.acq_cond[i_start] == -np.inf:
            return
        try:
            i_1st_inf = next(i for i, ac in enumerate(self.acq_cond) if ac == -np.inf)
        except StopIteration:
            i_1st_inf = len(self) + 1
        sigma_cond = predict_std(upper, self.X[i_start:i_1st_inf])
        acq_cond = np.clip(self._acq_func(self.y[i_start:i_1st_inf], sigma_cond), None,
                           np.inf if i_start == 0 else self.acq_cond[i_start - 1])
        j_sort = np.argsort(-acq_cond)
        if acq_cond[j_sort[0]] == -np.inf:
            self.acq_cond[i_start:i_1st_inf] = -np.inf
            return
        i_sort_partial = i_start + j_sort
        for arr in (self.X, self.y, self.sigma, self.acq, self.idx):
            arr[i_start:i_1st_inf] = arr[i_sort_partial]
        self.acq_cond[i_start:i_1st_inf] = acq_cond[j_sort]
        self.sort(i_start + 1)


def ranked_pool_select(st, X, y, sigma, acq, n_points, zeta=None, method="single sort acq"):
    """What ``NORA.multi_add`` returns for one process given a scored pool
    (gp_acquisition.py:1141-1146, 1097-1100): ``(idx, X_pool, y_pool, acq_pool)``."""
    zeta = auto_zeta(st.d) if zeta is None else zeta

    def acq_func(y_, sigma_):
        return logexp_f(y_, sigma_, st.y_max, st.noise_level, zeta)

    pool = RankedPool(n_points, st, acq_func)
    pool.add(X, y, sigma, acq, method=method)
    n_full = n_points
    for i, a in enumerate(pool.acq_cond[:-1]):   # copy(drop_empty=True) :1583-1595
        if a == -np.inf:
            n_full = i
            break
    sel = slice(0, n_full)
    acq_pool = acq_func(pool.y[sel], pool.sigma[sel])
    return pool.idx[sel].copy(), pool.X[sel].copy(), pool.y[sel].copy(), acq_pool


# --------------------------------------------------------------------------------------
# Synthetic workloads -- shared by tests and bench
# --------------------------------------------------------------------------------------
def synthetic_problem(N, d, seed=1234, ell=None):
    """X_train ~ U(0,1)^{N x d}; y = -1/2 |(x-0.5)/0.15|^2; bounds [0,1]^d; theta fixed."""
    rng = np.random.default_rng(seed)
    X = rng.uniform(size=(N, d))
    y = -0.5 * np.sum(((X - 0.5) / 0.15) ** 2, axis=1)
    if ell is None:
        ell = 0.5 if d <= 8 else (1.0 if d <= 16 else 1.5)
    theta = np.log(np.concatenate([[1.0], np.full(d, ell)]))
    bounds = np.array([[0.0, 1.0]] * d)
    return X, y, theta, bounds
