"""
Ranked-pool batch acquisition with the interface of ``gpry.gp_acquisition``
(reference gp_acquisition.py:525-1191 ``NORA``, :1194-1670 ``RankedPool``).

The reference scores an MC sample with ``predict`` (mpi.py:182-218), applies ``LogExp.f``,
and ranks it with Kriging-believer (KB) conditioning: the acquisition value of pool slot i
uses the std of a GP augmented with the slots above it, which the reference obtains by
deep-copying the regressor and re-factorising it for every cached model
(``cache_model`` :1522-1555, ~50 % of the ranking time by its own account :1530-1532).

Here:
* scoring + LogExp + descending pre-ranking is ONE fused device pass that returns only the
  K' best candidates (``GaussianProcessRegressor.predict_logexp_topk``);
* KB conditioning uses the posterior covariance among those survivors, computed once on the
  GPU (``gpry_posterior_cov``): var(a | P) = S_aa - S_aP (S_PP + noise I)^-1 S_Pa -- the
  same quantity as the reference's refit, without the O((N+i)^3) refactorisations;
* the pre-selection is exact: a candidate whose *unconditioned* acquisition is not above the
  final last-slot value can never enter the pool (conditioning only lowers acquisition
  values, :1482-1488, and ``add_one`` drops ``acq <= min_acq`` :1432), so K' is enlarged
  until ``acq[K'-th] <= pool.min_acq``.

The control flow of ``RankedPool`` (``add`` / ``add_one`` / ``sort`` / ``add_bulk``) follows
the reference step by step so that the resulting pool is the same.
"""
from copy import deepcopy
from functools import partial

import numpy as np

from .acquisition_functions import LogExp


class _RefitConditioner:
    """Faithful conditioned models: deepcopy + append_to_data(fit_gpr=False) per cached
    model (gp_acquisition.py:1547-1555), i.e. a device re-factorisation each."""

    def __init__(self, gpr):
        self.gpr = gpr
        self.models = {}

    def cache(self, i, pool_X, pool_y, pool_idx=None):
        if i < 0:
            return
        m = deepcopy(self.gpr)
        m.append_to_data(pool_X[:i + 1], pool_y[:i + 1], fit_gpr=False, fit_classifier=False)
        self.models[i] = m

    def std(self, i, X, idx=None):
        model = self.gpr if i < 0 else self.models[i]
        return model.predict_std(np.atleast_2d(X), validate=False)


class _CovConditioner:
    """Conditioning from the posterior covariance among a fixed candidate set (device)."""

    def __init__(self, gpr, X_candidates):
        self.gpr = gpr
        dev = gpr._device_state()
        self.S = np.asarray(dev.posterior_cov(np.ascontiguousarray(X_candidates)))
        py = gpr.preprocessing_y
        self.y_std = getattr(py, "std_", 1.0) if hasattr(py, "std_") else 1.0
        nl = gpr.noise_level
        if np.iterable(nl):
            raise NotImplementedError("KB conditioning needs a scalar noise_level (as NORA, "
                                      "gp_acquisition.py:1049-1051)")
        self.noise2 = (nl / self.y_std) ** 2     # alpha of the appended lie points
        self.cached = {}                          # i -> (P indices, Cholesky-solved block)

    def cache(self, i, pool_X, pool_y, pool_idx):
        if i < 0:
            return
        P = np.asarray(pool_idx[:i + 1], dtype=int)
        A = self.S[np.ix_(P, P)] + self.noise2 * np.eye(len(P))
        self.cached[i] = (P, np.linalg.cholesky(A))

    def std(self, i, X, idx):
        idx = np.atleast_1d(np.asarray(idx, dtype=int))
        var = self.S[idx, idx].copy()
        if i >= 0:
            P, Lc = self.cached[i]
            B = np.linalg.solve(Lc, self.S[np.ix_(P, idx)])     # |P| x n
            var -= np.einsum("ij,ij->j", B, B)
        var[var < 0] = 0.0
        return np.sqrt(var) * self.y_std


class RankedPool:
    """Ranked pool of proposals with KB conditioning (gp_acquisition.py:1194-1670).

    ``conditioning="refit"`` reproduces the reference's cached re-fitted models; ``"cov"``
    (used by :class:`NORA`) needs ``add`` to be called once with the whole candidate batch."""

    def __init__(self, size, gpr, acq_func, verbose=1, conditioning="refit"):
        self._gpr = gpr
        self._acq_func = acq_func
        self.verbose = verbose
        self.conditioning = conditioning
        self.X = np.zeros((size + 1, gpr.d))
        self.y = np.zeros(size + 1)
        self.acq_cond = np.full(size + 1, -np.inf)
        self.sigma = np.zeros(size + 1)
        self.acq = np.zeros(size + 1)
        self.idx = np.full(size + 1, -1, dtype=np.int64)   # position in the batch given to add
        self._cond = _RefitConditioner(gpr)
        self.cache_counter = 0

    def __len__(self):
        return len(self.y) - 1

    @property
    def min_acq(self):
        return self.acq_cond[len(self) - 1]

    def log(self, level=None, msg=""):
        if level is None or level <= self.verbose:
            print(msg)

    def reset_cache(self):
        self._cond = _RefitConditioner(self._gpr)

    def cache_model(self, i):
        if i >= 0:
            self._cond.cache(i, self.X, self.y, self.idx)
            self.cache_counter += 1

    # ---------------------------------------------------------------- add
    def add(self, X, y=None, sigma=None, acq=None, method="single sort acq"):
        """gp_acquisition.py:1290-1335."""
        X = np.atleast_2d(X)
        if y is not None:
            y = np.atleast_1d(y)
        if sigma is not None:
            sigma = np.atleast_1d(sigma)
        if y is None:
            y, sigma = self._gpr.predict(X, return_std=True, validate=False)
        elif sigma is None:
            sigma = self._gpr.predict_std(X, validate=False)
        if acq is None:
            acq = self._acq_func(y, sigma)
        idx = np.arange(len(X))
        if self.conditioning == "cov":
            self._cond = _CovConditioner(self._gpr, X)
        if method.lower() == "bulk":
            self.add_bulk(X, y, sigma, acq, idx)
        elif method.lower().startswith("single"):
            i_sort = None
            if "sort" in method.lower():
                i_sort = np.argsort({"acq": acq, "y": y}[method.lower().split()[-1]])[::-1]
            for i in (i_sort if i_sort is not None else range(len(X))):
                self.add_one(X[i], y[i], sigma[i], acq[i], idx=idx[i])
        else:
            raise ValueError(f"Algorithm '{method}' not known.")

    def add_bulk(self, X, y, sigma, acq, idx, i_start=0):
        """gp_acquisition.py:1337-1390."""
        if i_start == 0:
            acq_cond = acq if isinstance(acq, np.ndarray) else np.array(acq)
        else:
            self.cache_model(i_start - 1)
            sigma_cond = self._cond.std(i_start - 1, X, idx)
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
        self.add_bulk(X[keep], y[keep], sigma[keep], acq[keep], idx[keep], i_start=i_start + 1)

    def add_one(self, X, y=None, sigma=None, acq=None, acq_nan_is_null=False, idx=-1):
        """gp_acquisition.py:1392-1520."""
        if acq is not None and acq <= self.min_acq:
            return
        X = np.atleast_2d(X)
        if y is None:
            y, sigma = self._gpr.predict(X, return_std=True, validate=False)
            y, sigma = y[0], sigma[0]
        if sigma is None:
            sigma = self._gpr.predict_std(X, validate=False)[0]
        if acq is None:
            acq = self._acq_func(y, sigma)
        if acq <= self.min_acq:
            return
        if np.isnan(acq):
            if not acq_nan_is_null:
                raise ValueError(f"Acquisition function value not a number: {acq}")
            acq = -np.inf
        n = len(self)
        i_new_last = n
        acq_cond = acq
        while True:
            # provisional position from the bottom; '>=' keeps -inf from climbing (:1469-1474)
            i_new = 0
            for i in range(n):
                if self.acq_cond[-(i + 2)] >= acq_cond:
                    i_new = n - i
                    break
            if i_new in [0, i_new_last, n]:
                break
            sigma_cond = self._cond.std(i_new - 1, X, idx)[0]
            acq_cond = min(acq_cond, self._acq_func(y, sigma_cond))
            i_new_last = i_new
        if i_new >= n:
            return
        for pool, value in [(self.X, X), (self.y, y), (self.sigma, sigma), (self.acq, acq),
                            (self.acq_cond, acq_cond), (self.idx, idx)]:
            pool[i_new + 1:] = pool[i_new:-1]
            pool[i_new] = value
        assert self.acq_cond[i_new] > -np.inf
        self.sort(i_new + 1)
        self.acq_cond[-1] = -np.inf

    def sort(self, i_start=0):
        """gp_acquisition.py:1598-1670."""
        if i_start >= len(self):
            return
        self.cache_model(i_start - 1)
        if self.acq_cond[i_start] == -np.inf:
            return
        i_1st_inf = len(self) + 1
        for i, ac in enumerate(self.acq_cond):
            if ac == -np.inf:
                i_1st_inf = i
                break
        sl = slice(i_start, i_1st_inf)
        sigma_cond = self._cond.std(i_start - 1, self.X[sl], self.idx[sl])
        acq_cond = np.clip(self._acq_func(self.y[sl], sigma_cond), None,
                           np.inf if i_start == 0 else self.acq_cond[i_start - 1])
        j_sort = np.argsort(-acq_cond)
        if acq_cond[j_sort[0]] == -np.inf:
            self.acq_cond[sl] = -np.inf
            return
        i_sort_partial = i_start + j_sort
        for arr in (self.X, self.y, self.sigma, self.acq, self.idx):
            arr[sl] = arr[i_sort_partial]
        self.acq_cond[sl] = acq_cond[j_sort]
        self.sort(i_start + 1)

    # ---------------------------------------------------------------- copies
    def __getstate__(self):
        return self.__deepcopy__().__dict__

    def __deepcopy__(self, memo=None):
        new = self.__class__.__new__(self.__class__)
        new.__dict__ = {k: deepcopy(v) for k, v in self.__dict__.items()
                        if k not in ("_gpr", "_acq_func", "_cond")}
        return new

    def copy(self, drop_empty=False):
        """gp_acquisition.py:1575-1596."""
        c = deepcopy(self)
        if drop_empty:
            for i, a in enumerate(c.acq_cond[:-1]):
                if a == -np.inf:
                    for name in ("X", "y", "acq_cond", "sigma", "acq", "idx"):
                        setattr(c, name, getattr(c, name)[:i])
                    break
        return c


def ranked_pool_from_scores(gpr, X, y, sigma, acq, n_points, acq_func, method="single sort acq",
                            conditioning="cov"):
    """``RankedPool(n_points).add(X, y, sigma, acq)`` -> (idx, X, y, sigma, acq_cond, min_acq,
    full) for a (small) scored candidate set."""
    pool = RankedPool(n_points, gpr=gpr, acq_func=acq_func, verbose=0, conditioning=conditioning)
    with np.errstate(divide="ignore"):
        pool.add(X, y, sigma, acq, method=method)
    return pool


class NORA:
    """Batch acquisition from a ranked MC pool (gp_acquisition.py:525-1191).

    The MC sample of the GP mean normally comes from an external nested sampler (PolyChord /
    UltraNest / nessai: out of scope here); it can be handed in through ``X_mc``, or drawn
    with the reference's test sampler ``sampler="uniform"`` (1000 d uniform points,
    gp_acquisition.py:676-682, 748-758).  Everything from the scoring on runs on the GPU.

    With ``torch.distributed`` initialised (one process per GPU) the pool is sharded by
    stride as ``mpi.step_split`` does (mpi.py:105-115), every rank scores its shard, the
    per-rank survivor lists are all-gathered over NCCL and every rank ranks the same merged
    list (replacing the 5 gathers + bcast of gp_acquisition.py:1148-1191).
    """

    def __init__(self, bounds, acq_func=None, mc_every=1, sampler="uniform", nsamples=None,
                 kprime=256, verbose=1):
        self.bounds = np.asarray(bounds, dtype=float)
        d = self.bounds.shape[0]
        self.acq_func = acq_func if acq_func is not None else LogExp(dimension=d)
        self.mc_every = mc_every
        self.mc_every_i = 0
        self.sampler = sampler
        self.nsamples = nsamples if nsamples is not None else 1000 * d
        self.kprime = kprime
        self.verbose = verbose
        self._X_mc = None
        self._X_already_proposed = np.empty((0, d))
        self.pool = None
        self.last_kprime = None

    def do_MC_sample(self, gpr, bounds=None, rng=None):
        if self.sampler != "uniform":
            raise NotImplementedError("only the 'uniform' test sampler is built in; pass X_mc "
                                      "from your nested sampler")
        b = self.bounds if bounds is None else np.asarray(bounds)
        rng = np.random.default_rng(rng) if not isinstance(rng, np.random.Generator) else rng
        return rng.uniform(b[:, 0], b[:, 1], size=(self.nsamples, b.shape[0]))

    def multi_add(self, gpr, n_points=1, bounds=None, rng=None, force_resample=False,
                  X_mc=None):
        """gp_acquisition.py:971-1108 -> (X_pool, y_pool, acq_pool), identical on all ranks."""
        from . import parallel
        if not (isinstance(n_points, int) and n_points > 0):
            raise ValueError(f"n_points should be int > 0, got {n_points}")
        mc_this_time = not bool(self.mc_every_i % self.mc_every) or force_resample \
            or self._X_mc is None or X_mc is not None
        if mc_this_time:
            if X_mc is None:
                X_mc = self.do_MC_sample(gpr, bounds=bounds, rng=rng) \
                    if parallel.is_main_process() else None
                X_mc = parallel.bcast(X_mc)
            self._X_mc = np.ascontiguousarray(X_mc, dtype=float)
            self._X_already_proposed = np.empty((0, gpr.d))
        self.mc_every_i += 1
        X_all = self._X_mc
        if self._X_already_proposed.size > 0:   # gp_acquisition.py:1037-1047
            used = {row.tobytes() for row in self._X_already_proposed}
            keep = np.array([row.tobytes() not in used for row in X_all])
            X_all = X_all[keep]
        zeta = self.acq_func.zeta
        noise = gpr.noise_level
        acq_func = partial(self.acq_func.f, baseline=gpr.y_max, noise_level=noise, zeta=zeta)
        self.acq_func_y_sigma = acq_func
        # shard by stride (mpi.py:114-115), score + pre-rank on this rank's GPU
        rank, size = parallel.rank(), parallel.size()
        this_X = np.ascontiguousarray(X_all[rank::size])
        Kp = max(self.kprime, 4 * n_points)
        while True:
            Kp_eff = min(Kp, 2048)
            if len(this_X):
                a, i, m, s, Xs = gpr.predict_logexp_topk(this_X, zeta, Kp_eff)
            else:
                a, i, m, s, Xs = (np.empty(0), np.empty(0, dtype=np.int64), np.empty(0),
                                  np.empty(0), np.empty((0, gpr.d)))
            local_cut = float(a[-1]) if (len(this_X) > Kp_eff and len(a)) else -np.inf
            i = i * size + rank       # position in the un-sharded sample
            a, i, m, s, Xs = parallel.allgather_survivors(a, i, m, s, Xs)
            order = np.lexsort((i, -a))
            a, i, m, s, Xs = a[order], i[order], m[order], s[order], Xs[order]
            pool = ranked_pool_from_scores(gpr, Xs, m, s, a, n_points, acq_func)
            # Exactness of the pre-selection (module docstring): every candidate a truncated
            # shard did NOT send has acq <= that shard's smallest survivor; if the largest such
            # bound is <= the last-slot conditioned acq of a full pool, none of them could
            # have entered.
            cut = parallel.max_scalar(local_cut)
            if cut == -np.inf or (np.isfinite(pool.min_acq) and cut <= pool.min_acq) \
                    or Kp_eff >= 2048:
                if cut > -np.inf and not (np.isfinite(pool.min_acq) and cut <= pool.min_acq):
                    import warnings
                    warnings.warn("ranked pool: pre-selection bound not met at K'=2048; the "
                                  "pool may differ from a ranking of the full sample")
                break
            Kp *= 2
        self.last_kprime = Kp_eff
        self.pool = pool
        merged = pool.copy(drop_empty=True)
        X_pool, y_pool = merged.X[:n_points], merged.y[:n_points]
        with np.errstate(divide="ignore"):
            acq_pool = acq_func(y_pool, merged.sigma[:n_points])
        self._X_already_proposed = np.concatenate([self._X_already_proposed, X_pool])
        return X_pool, y_pool, acq_pool
