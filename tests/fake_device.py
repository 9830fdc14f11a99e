"""CPU stand-in for ``gpry_b200.device.DeviceGP`` (TEST INFRASTRUCTURE ONLY): the same method
surface, arithmetic by the oracle.  It lets the host logic of ``gpry_b200.gpr`` and of the
in-place GPry patch (model-update decisions, lazy ``L_``/``V_``, masks, pickling) run on a
machine without a GPU.  Never imported by the product package."""
import numpy as np
from scipy.linalg import cholesky, cho_solve, solve_triangular

from oracle import gp_oracle as orc


class FakeDeviceGP:
    calls = []          # (method, detail) log shared by all instances, for the tests

    def __init__(self, device=0):
        self.device = device
        self.kind = None
        self.N = self.d = 0
        self._f = None
        self._trust = None
        self._mask_value = -np.inf
        self._clf = None
        self.contract_mode = "int8"

    # ------------------------------------------------------------------ training side
    def _factor(self, kind, X_, noise2, y_, theta):
        K = orc.kernel_cross(kind, theta, X_, X_)
        K[np.diag_indices_from(K)] += noise2
        try:
            L = cholesky(K, lower=True)
        except np.linalg.LinAlgError:
            return None
        V = solve_triangular(L, np.eye(len(L)), lower=True)
        return dict(kind=kind, X_=np.array(X_), noise2=np.array(noise2), y_=np.array(y_),
                    theta=np.array(theta), L=L, V=V, alpha_=cho_solve((L, True), y_))

    def factorize(self, kind, X_, noise2, y_, theta, want_L=True, want_V=True, keep_on_device=False):
        FakeDeviceGP.calls.append(("factorize", len(y_)))
        noise2 = np.broadcast_to(noise2, (len(y_),))
        f = self._factor(kind, np.asarray(X_, float), noise2, np.asarray(y_, float), theta)
        self._f = None
        if f is None:
            return None, None, np.zeros(len(y_)), 0.0, 1
        if keep_on_device:
            self._f = f
            self._f_N = len(y_)
        return (f["L"] if want_L else None, f["V"] if want_V else None, f["alpha_"],
                float(np.log(np.diag(f["L"])).sum()), 0)

    def factor_append(self, X_new_, noise2_new, y_all_, theta):
        f = self._f
        assert f is not None and np.array_equal(theta, f["theta"])
        k = len(np.atleast_2d(X_new_))
        FakeDeviceGP.calls.append(("factor_append", k))
        X_ = np.vstack([f["X_"], np.atleast_2d(X_new_)])
        noise2 = np.concatenate([f["noise2"], np.broadcast_to(noise2_new, (k,))])
        g = self._factor(f["kind"], X_, noise2, np.asarray(y_all_, float), theta)
        self._f = g
        if g is None:
            return np.zeros(len(y_all_)), 1
        self._f_N = len(y_all_)
        return g["alpha_"], 0

    def factor_download(self, want_L=True, want_V=True):
        FakeDeviceGP.calls.append(("factor_download", self._f_N))
        return (self._f["L"].copy() if want_L else None, self._f["V"].copy() if want_V else None)

    def lml_batched(self, kind, X_, noise2, y_, thetas, eval_gradient=True):
        thetas = np.atleast_2d(thetas)
        noise2 = np.broadcast_to(noise2, (len(y_),))
        lml = np.empty(len(thetas))
        grad = np.zeros_like(thetas)
        info = np.zeros(len(thetas), dtype=np.int32)
        for i, th in enumerate(thetas):
            out = orc.log_marginal_likelihood(kind, th, X_, y_, noise2, eval_gradient=True)
            lml[i], grad[i] = out
            info[i] = 0 if np.isfinite(lml[i]) else 1
        return lml, (grad if eval_gradient else None), info

    # ------------------------------------------------------------------ model state
    def upload(self, kind, X_, alpha_, V_, c, ell, x_min=None, x_width=None, y_mean=0.0,
               y_std=1.0, clip_hi=np.inf):
        FakeDeviceGP.calls.append(("upload", len(alpha_)))
        self.kind, self.N, self.d = kind, len(alpha_), X_.shape[1]
        self._m = dict(X_=np.array(X_), alpha_=np.array(alpha_), V=None if V_ is None else np.array(V_),
                       theta=np.log(np.concatenate([[c], np.broadcast_to(ell, (X_.shape[1],))])),
                       x_min=np.zeros(self.d) if x_min is None else np.array(x_min),
                       x_width=np.ones(self.d) if x_width is None else np.array(x_width),
                       y_mean=y_mean, y_std=y_std, clip_hi=clip_hi, c=c)
        self._clf = None

    def adopt_factorization(self, c, ell, x_min=None, x_width=None, y_mean=0.0, y_std=1.0,
                            clip_hi=np.inf):
        f = self._f
        FakeDeviceGP.calls.append(("adopt", len(f["alpha_"])))
        n = len(FakeDeviceGP.calls)
        self.upload(f["kind"], f["X_"], f["alpha_"], f["V"], c, ell, x_min, x_width, y_mean, y_std,
                    clip_hi)
        del FakeDeviceGP.calls[n:]

    def set_contract_mode(self, mode):
        self.contract_mode = mode

    def set_mask_value(self, value):
        self._mask_value = value

    def set_trust_region(self, bounds=None, value=-np.inf):
        self._trust = None if bounds is None else np.array(bounds, dtype=float)
        self._mask_value = value

    def set_classifier(self, spec=None):
        self._clf = spec

    # ------------------------------------------------------------------ candidate side
    def _xt(self, X):
        m = self._m
        return (np.asarray(X, float) - m["x_min"]) / m["x_width"]

    def classify(self, X):
        sv, coef, intercept, gamma = self._clf
        X_ = self._xt(X)
        d2 = ((X_[:, None, :] - sv[None]) ** 2).sum(-1)
        return np.exp(-gamma * d2) @ coef + intercept

    def _mean_var(self, X, want_var):
        m = self._m
        Ks = orc.kernel_cross(self.kind, m["theta"], self._xt(X), m["X_"])
        mean = np.minimum(Ks @ m["alpha_"] * m["y_std"] + m["y_mean"], m["clip_hi"])
        var = None
        if want_var:
            W = m["V"] @ Ks.T
            var = np.maximum(m["c"] - np.einsum("ji,ji->i", W, W), 0.0)
        return mean, var

    def _masks(self, X, mean, std, acq):
        if self._clf is not None:
            bad = self.classify(X) <= 0
            if mean is not None:
                mean[bad] = self._mask_value
            if std is not None:
                std[bad] = 0.0
            if acq is not None:
                acq[bad] = -np.inf
        if self._trust is not None:
            out = ~np.all((X >= self._trust[:, 0]) & (X <= self._trust[:, 1]), axis=1)
            if mean is not None:
                mean[out] = self._mask_value
            if acq is not None:
                acq[out] = -np.inf

    def predict(self, X, return_mean=True, return_std=False, stream=None, out=None):
        X = np.asarray(X, float)
        mean, var = self._mean_var(X, return_std)
        std = np.sqrt(var) * self._m["y_std"] if return_std else None
        self._masks(X, mean if return_mean else None, std, None)
        return (mean if return_mean else None), std

    def predict_logexp(self, X, zeta, sigma_n, y_max, stream=None):
        X = np.asarray(X, float)
        mean, var = self._mean_var(X, True)
        std = np.sqrt(var) * self._m["y_std"]
        with np.errstate(divide="ignore"):
            acq = 2 * zeta * (mean - y_max) + np.log(np.sqrt(np.clip(std ** 2 - sigma_n ** 2, 0, None)))
        self._masks(X, mean, std, acq)
        return mean, std, acq

    def predict_logexp_topk(self, X, zeta, sigma_n, y_max, Kp, idx_offset=0, stream=None,
                            device_out=False, want_X=True, exclude=None):
        mean, std, acq = self.predict_logexp(X, zeta, sigma_n, y_max)
        key = np.where(np.isnan(acq), -np.inf, acq)
        order = np.lexsort((np.arange(len(acq)), -key))
        if exclude is not None and len(exclude):
            order = order[~np.isin(order, exclude)]
        order = order[:Kp]
        return (acq[order], order.astype(np.int64) + idx_offset, mean[order], std[order],
                np.asarray(X)[order] if want_X else None)

    def _oracle_state(self):
        raise NotImplementedError

    def mean_grad(self, x):
        m = self._m
        g = orc.kernel_gradient_x(self.kind, m["theta"], self._xt(np.atleast_2d(x))[0], m["X_"])
        return g.T @ m["alpha_"] * m["y_std"]

    def std_grad(self, x):
        m = self._m
        X_ = self._xt(np.atleast_2d(x))
        Ks = orc.kernel_cross(self.kind, m["theta"], X_, m["X_"])
        g = orc.kernel_gradient_x(self.kind, m["theta"], X_[0], m["X_"])
        w = m["V"] @ Ks[0]
        var = m["c"] - w @ w
        if var <= 0:
            return np.zeros(self.d), 0.0
        return -((m["V"].T @ w) @ g) / np.sqrt(var) * m["y_std"] ** 2, np.sqrt(var) * m["y_std"]

    def predict_grad(self, X, return_std_grad=True):
        X = np.asarray(X, float)
        mean, var = self._mean_var(X, True)
        gm = np.array([self.mean_grad(x) for x in X])
        gs = np.array([self.std_grad(x)[0] for x in X]) if return_std_grad else None
        return mean, np.sqrt(var) * self._m["y_std"], gm, gs

    def posterior_cov(self, X, stream=None):
        m = self._m
        X_ = self._xt(X)
        Ks = orc.kernel_cross(self.kind, m["theta"], X_, m["X_"])
        U = m["V"] @ Ks.T
        return orc.kernel_cross(self.kind, m["theta"], X_, X_) - U.T @ U

    def close(self):
        pass


def install(monkeypatch):
    """Route ``gpry_b200.gpr`` (and everything built on it) to the fake device."""
    import gpry_b200.gpr as gpr_mod
    import gpry_b200.device as dev_mod
    FakeDeviceGP.calls.clear()
    ws = {}

    def workspace(device=0):
        return ws.setdefault(device, FakeDeviceGP(device))
    import gpry_b200.svm as svm_mod
    monkeypatch.setattr(svm_mod, "DeviceGP", FakeDeviceGP)
    monkeypatch.setattr(gpr_mod, "DeviceGP", FakeDeviceGP)
    monkeypatch.setattr(gpr_mod, "workspace", workspace)
    monkeypatch.setattr(dev_mod, "workspace", workspace)
    return FakeDeviceGP
