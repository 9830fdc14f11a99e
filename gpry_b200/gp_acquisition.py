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
from time import perf_counter

import numpy as np

from .acquisition_functions import LogExp
from .gpr import is_in_bounds


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
    UltraNest / nessai: out of scope here); it can be handed in through ``X_mc``, drawn with
    the reference's test sampler ``sampler="uniform"`` (1000 d uniform points,
    gp_acquisition.py:676-682, 748-758), or drawn from the surrogate posterior itself with
    ``sampler="ensemble"``: the batched GPU ensemble sampler of ``gpry_b200.mc`` (``nsamples``
    walkers, ``mc_steps`` updates).  Everything from the scoring on runs on the GPU.

    With ``torch.distributed`` initialised (one process per GPU) the pool is sharded by
    stride as ``mpi.step_split`` does (mpi.py:105-115), every rank scores its shard, the
    per-rank survivor lists are all-gathered over NCCL and every rank ranks the same merged
    list (replacing the 5 gathers + bcast of gp_acquisition.py:1148-1191).
    """

    def __init__(self, bounds, acq_func=None, mc_every=1, sampler="uniform", nsamples=None,
                 kprime=256, verbose=1, mc_steps=100):
        self.bounds = np.asarray(bounds, dtype=float)
        d = self.bounds.shape[0]
        self.acq_func = acq_func if acq_func is not None else LogExp(dimension=d)
        self.mc_every = mc_every
        self.mc_every_i = 0
        self.sampler = sampler
        self.nsamples = nsamples if nsamples is not None else 1000 * d      # (sampler=None: the
        # caller always hands the MC sample in, through X_mc / X_shard)
        self.mc_steps = mc_steps
        self.kprime = kprime
        self.verbose = verbose
        self._X_shard = None          # this rank's strided shard of the current MC sample
        self._pool_given = None       # the object the sample came from (identity = same sample)
        self._X_already_proposed = np.empty((0, d))
        self._idx_already_proposed = np.empty(0, dtype=np.int64)
        self.pool = None
        self.last_kprime = None

    def do_MC_sample(self, gpr, bounds=None, rng=None):
        b = self.bounds if bounds is None else np.asarray(bounds)
        rng = np.random.default_rng(rng) if not isinstance(rng, np.random.Generator) else rng
        if self.sampler == "ensemble":
            from .mc import ensemble_sample
            res = ensemble_sample(gpr, bounds=b, n_walkers=self.nsamples + self.nsamples % 2,
                                  n_steps=self.mc_steps, seed=int(rng.integers(2 ** 31)),
                                  X_init="training")
            self.last_mc = res
            return res.X[:self.nsamples]
        if self.sampler != "uniform":
            raise NotImplementedError("built-in samplers: 'uniform' (the reference's test "
                                      "sampler) and 'ensemble'; or pass X_mc from your own")
        return rng.uniform(b[:, 0], b[:, 1], size=(self.nsamples, b.shape[0]))

    # ------------------------------------------------------------------ NS front-ends
    @staticmethod
    def logp_function(gpr, flavour="ultranest"):
        """The surrogate log-posterior closure the reference hands to its nested samplers
        (gp_acquisition.py:770, 784-793): ``logp(X)`` for ``X`` of shape (d,) or (M, d).
        UltraNest calls it VECTORISED (ns_interfaces.py:448 ``"vectorized": True``), i.e. with
        whole batches of live-point proposals: each call is one mean-only device pass
        (``gpry_predict(what=MEAN)``, masks applied on the GPU).

        ``"ultranest"``: masked rows (classifier / trust region) come back as ``-1e-300``
        instead of ``-inf`` -- the reference swaps ``gpr.minus_inf_value`` around the call
        (:788-792) because UltraNest cannot digest infinities; ``"polychord"`` / ``"nessai"``:
        one point per call, returns a float (:770)."""
        if flavour == "ultranest":
            def logp(X):
                prev = gpr.minus_inf_value
                gpr.minus_inf_value = -1e-300
                try:
                    return gpr.predict(np.atleast_2d(X), return_std=False, validate=False)
                finally:
                    gpr.minus_inf_value = prev
            return logp
        if flavour in ("polychord", "nessai"):
            return lambda X: gpr.predict(np.atleast_2d(X), return_std=False, validate=False)[0]
        raise ValueError(f"unknown nested-sampler flavour {flavour!r}")

    # ------------------------------------------------------------------ the MC pool
    def _set_pool(self, gpr, bounds, rng, force_resample, X_mc, X_shard):
        """Decides whether this call works on a new MC sample (gp_acquisition.py:1016-1027) and
        leaves this rank's strided shard in ``self._X_shard`` (row i of the shard is row
        ``i * size + rank`` of the whole sample, mpi.py:114-115).  Nothing is pickled and the
        whole sample never has to exist on one rank:

        * ``X_shard``: every rank hands in its own shard (pageable numpy array, or a float64
          torch CUDA tensor already resident on the rank's GPU) -- no communication at all;
        * ``X_mc``: the whole sample, given on every rank (each keeps ``X_mc[rank::size]``), or
          on rank 0 only (then the array is broadcast as a tensor, not as a pickled object);
        * neither: rank 0 draws the sample with the built-in sampler and it is broadcast.
        """
        from . import parallel
        rank, size = parallel.rank(), parallel.size()
        given = X_shard if X_shard is not None else X_mc
        same = given is not None and given is self._pool_given
        new_sample = (force_resample or self._X_shard is None
                      or (given is not None and not same)
                      or (given is None and not bool(self.mc_every_i % self.mc_every)))
        if not new_sample:
            return False
        if X_shard is not None:
            shard = X_shard
        else:
            if X_mc is None and self.sampler is not None:
                X_mc = self.do_MC_sample(gpr, bounds=bounds, rng=rng) \
                    if parallel.is_main_process() else None
            have = parallel.allgather(X_mc is not None) if size > 1 else [X_mc is not None]
            if not have[0]:
                raise ValueError("no MC sample: pass X_shard on every rank, or X_mc on every "
                                 "rank or on rank 0")
            if not all(have):
                X_mc = parallel.bcast_array(X_mc)
            shard = X_mc[rank::size] if size > 1 else X_mc
        if not hasattr(shard, "is_cuda"):
            shard = np.ascontiguousarray(shard, dtype=float)
        self._X_shard = shard
        self._pool_given = given
        self._X_already_proposed = np.empty((0, gpr.d))
        self._idx_already_proposed = np.empty(0, dtype=np.int64)
        return True

    def multi_add(self, gpr, n_points=1, bounds=None, rng=None, force_resample=False,
                  X_mc=None, X_shard=None):
        """gp_acquisition.py:971-1108 -> (X_pool, y_pool, acq_pool), identical on all ranks.

        One acquisition step: every rank scores its shard of the MC sample on its GPU
        (mean, std, LogExp and the exact top-K' in one pass; rows proposed since the sample
        was drawn are skipped on the device by index, :1037-1047), the per-rank survivor
        lists are all-gathered, and every rank runs the same Kriging-believer ranking on
        the best K' of the union -- which equals a single-process ranking of the whole
        sample (see the module docstring for why the pre-selection is exact)."""
        from . import parallel
        if not (isinstance(n_points, int) and n_points > 0):
            raise ValueError(f"n_points should be int > 0, got {n_points}")
        rng = parallel.get_random_generator(rng)
        self.last_new_sample = self._set_pool(gpr, bounds, rng, force_resample, X_mc, X_shard)
        self.mc_every_i += 1
        rank, size = parallel.rank(), parallel.size()
        this_X = self._X_shard
        n_this = int(this_X.shape[0])
        zeta = self.acq_func.zeta
        noise = gpr.noise_level
        acq_func = partial(self.acq_func.f, baseline=gpr.y_max, noise_level=noise, zeta=zeta)
        self.acq_func_y_sigma = acq_func
        done = self._idx_already_proposed
        skip = np.sort(done[done % size == rank] // size) if len(done) else None
        n_live = n_this - (0 if skip is None else len(skip))
        Kp = max(self.kprime, 4 * n_points)
        timing = {"score_s": 0.0, "exchange_s": 0.0, "rank_s": 0.0}
        while True:
            Kp_eff = min(Kp, 2048)
            t0 = perf_counter()
            if n_live > 0:
                a, i, m, s, Xs = gpr.predict_logexp_topk(this_X, zeta, Kp_eff, exclude=skip)
                a, i, m, s, Xs = (_to_numpy(v) for v in (a, i, m, s, Xs))
            else:
                a, i, m, s, Xs = (np.empty(0), np.empty(0, dtype=np.int64), np.empty(0),
                                  np.empty(0), np.empty((0, gpr.d)))
            t1 = perf_counter()
            local_cut = float(a[-1]) if (n_live > Kp_eff and len(a)) else -np.inf
            i = i * size + rank       # position in the un-sharded sample
            # Only the best K' of the union are ranked (keeps the posterior covariance at
            # K' x K' whatever the number of ranks); the first row left out bounds what any
            # dropped row could have scored, like a shard's own cut
            a, i, m, s, Xs, dropped = parallel.merge_survivors(a, i, m, s, Xs, Kp_eff)
            local_cut = max(local_cut, dropped)
            t2 = perf_counter()
            pool = ranked_pool_from_scores(gpr, Xs, m, s, a, n_points, acq_func)
            t3 = perf_counter()
            timing["score_s"] += t1 - t0
            timing["exchange_s"] += t2 - t1
            timing["rank_s"] += t3 - t2
            # Exactness of the pre-selection (module docstring): every candidate that was NOT
            # ranked has acq <= cut; if cut <= the last-slot conditioned acq of a full pool,
            # none of them could have entered.
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
        self.last_timing = timing
        self.pool = pool
        merged = pool.copy(drop_empty=True)
        X_pool, y_pool = merged.X[:n_points], merged.y[:n_points]
        with np.errstate(divide="ignore"):
            acq_pool = acq_func(y_pool, merged.sigma[:n_points])
        self._X_already_proposed = np.concatenate([self._X_already_proposed, X_pool])
        self._idx_already_proposed = np.concatenate(
            [self._idx_already_proposed, i[merged.idx[:len(X_pool)]]])
        self.last_pool_idx = i[merged.idx[:len(X_pool)]]
        return X_pool, y_pool, acq_pool


def _to_numpy(v):
    if v is None or isinstance(v, np.ndarray):
        return v
    return v.detach().cpu().numpy()


def _number_times_d(value, d, varname):
    """"5d" -> 5 * d, "d" -> d, numbers pass through (tools.py:185-240 ``get_Xnumber``)."""
    if isinstance(value, str):
        if "d" not in value:
            raise ValueError(f"'{varname}' must be a number or a string like '5d', got {value!r}")
        factor, power = value.split("d")
        factor = float(factor) if factor else 1.0
        power = float(power.lstrip("^")) if power else 1.0
        return int(factor * d ** power)
    return int(value)


class BatchOptimizer:
    """Acquisition by direct multi-start optimisation of the acquisition function with
    Kriging-believer lies between the points of a batch (gp_acquisition.py:121-520).

    The reference runs ``n_restarts_optimizer`` scipy L-BFGS-B one after the other, every
    objective call being a one-point ``predict(return_std, return_mean_grad,
    return_std_grad)`` (:316-336, 503-511), and scores the start-point proposals one by one
    (:364-379).  Here the proposals of all restarts are scored in one device pass and the
    restarts advance in lock-step: each round of objective calls is ONE batched
    ``gpry_predict_grad`` (mean, std and both gradients for every active restart).  Start
    selection, the optimiser, the choice of the best optimum, the lie and the append follow
    the reference; with ``torch.distributed`` initialised the restarts are split over the
    ranks as mpi.split_number_for_parallel_processes does (:454-458) and every rank returns
    the same result.
    """

    def __init__(self, bounds, preprocessing_X=None, verbose=1, acq_func="LogExp",
                 proposer=None, acq_optimizer="fmin_l_bfgs_b", n_restarts_optimizer="5d",
                 n_repeats_propose=10):
        from .proposal import Proposer, PartialProposer, CentroidsProposer
        self.bounds_ = np.asarray(bounds, dtype=float)
        self.n_d = self.bounds_.shape[0]
        self.preprocessing_X = preprocessing_X
        self.verbose = verbose
        if isinstance(acq_func, str):
            if acq_func.lower() != "logexp":
                raise ValueError(f"Unknown acquisition function {acq_func!r}; 'LogExp' is built in")
            acq_func = LogExp(dimension=self.n_d)
        elif isinstance(acq_func, dict):
            name, args = next(iter(acq_func.items()))
            if name.lower() != "logexp":
                raise ValueError(f"Unknown acquisition function {name!r}; 'LogExp' is built in")
            acq_func = LogExp(dimension=self.n_d, **(args or {}))
        self.acq_func = acq_func
        if proposer is None:
            self.proposer = PartialProposer(self.bounds_, CentroidsProposer(self.bounds_))
        else:
            if not isinstance(proposer, Proposer):
                raise TypeError("'proposer' must be a Proposer instance. "
                                f"Got {proposer} of type {type(proposer)}.")
            self.proposer = proposer
            self.proposer.update_bounds(self.bounds_)
        has_gradient = getattr(self.acq_func, "hasgradient", True)
        if acq_optimizer == "auto":
            self.acq_optimizer = "fmin_l_bfgs_b" if has_gradient else "sampling"
        elif isinstance(acq_optimizer, str):
            if acq_optimizer == "fmin_l_bfgs_b" and not has_gradient:
                raise ValueError("In order to use the 'fmin_l_bfgs_b' optimizer the acquisition "
                                 f"function needs to be able to return gradients. Got {acq_func}")
            if acq_optimizer not in ("fmin_l_bfgs_b", "sampling"):
                raise ValueError("Supported internal optimizers are 'auto', 'lbfgs' or "
                                 f"'sampling', got {acq_optimizer}")
            self.acq_optimizer = acq_optimizer
        else:
            self.acq_optimizer = acq_optimizer
        self.n_restarts_optimizer = _number_times_d(n_restarts_optimizer, self.n_d,
                                                    "n_restarts_optimizer")
        self.n_repeats_propose = n_repeats_propose

    # ------------------------------------------------------------------ coordinates
    def _to_opt(self, X):
        return X if self.preprocessing_X is None else self.preprocessing_X.transform(X)

    def _from_opt(self, X):
        return X if self.preprocessing_X is None else self.preprocessing_X.inverse_transform(X)

    def _opt_bounds(self, bounds):
        if self.preprocessing_X is None:
            return bounds
        return self.preprocessing_X.transform_bounds(bounds)

    # ------------------------------------------------------------------ objective
    def _objective_batch(self, gpr):
        """(-acq, -grad) for the rows of X given in the optimiser's coordinates
        (gp_acquisition.py:316-336)."""
        def batch(X):
            X = self._from_opt(np.atleast_2d(np.asarray(X, dtype=float)))
            acq, grad = self.acq_func(X, gpr, eval_gradient=True)
            return -1 * np.atleast_1d(acq), -1 * np.atleast_2d(grad)
        return batch

    # ------------------------------------------------------------------ starting points
    def _starting_points(self, gpr, i_restarts, bounds, rng):
        """One start per restart index (gp_acquisition.py:346-390): index 0 starts from the
        last in-bounds training point, the others from the best of the first
        ``n_repeats_propose + 1`` proposals with a finite acquisition value, out of at most
        ``10 d n_restarts_optimizer`` tries.  Returns (x0 (n, d) un-transformed, value (n,),
        optimise (n,) bool): a restart that found no finite value is not optimised."""
        n = len(i_restarts)
        d = self.n_d
        need = self.n_repeats_propose + 1
        n_tries = 10 * d * self.n_restarts_optimizer
        x0 = np.empty((n, d))
        value = np.full(n, -np.inf)
        optimise = np.ones(n, dtype=bool)
        n_found = np.zeros(n, dtype=int)
        tries = np.zeros(n, dtype=int)
        proposing = np.ones(n, dtype=bool)
        for r, i in enumerate(i_restarts):
            if i == 0:
                x0[r] = next(X for X in gpr.X_train[::-1] if np.all(is_in_bounds(X, bounds)))
                proposing[r] = False
        while True:
            todo = [r for r in range(n)
                    if proposing[r] and n_found[r] < need and tries[r] < n_tries]
            if not todo:
                break
            per = [min(need, n_tries - tries[r]) for r in todo]
            X = self.proposer.get_batch(int(np.sum(per)), rng=rng)
            vals = np.atleast_1d(self.acq_func(X, gpr))
            off = 0
            for r, m in zip(todo, per):
                for j in range(off, off + m):
                    if n_found[r] == need:
                        break
                    tries[r] += 1
                    if np.isfinite(vals[j]):
                        n_found[r] += 1
                        if n_found[r] == 1 or vals[j] > value[r]:
                            value[r], x0[r] = vals[j], X[j]
                    elif n_found[r] == 0:
                        x0[r], value[r] = X[j], vals[j]   # the last draw, should none be finite
                off += m
        optimise[proposing & (n_found == 0)] = False
        return x0, value, optimise

    # ------------------------------------------------------------------ optimisation
    def _optimize_all(self, gpr, x0_opt, opt_bounds):
        batch = self._objective_batch(gpr)
        if self.acq_optimizer == "fmin_l_bfgs_b":
            from .lockstep import lockstep_minimize
            return lockstep_minimize(batch, x0_opt, opt_bounds)
        out = []
        for x0 in x0_opt:
            if self.acq_optimizer == "sampling":
                def value_only(x):
                    X = self._from_opt(np.atleast_2d(x))
                    return -1 * float(np.atleast_1d(self.acq_func(X, gpr))[0])
                import scipy.optimize
                res = scipy.optimize.minimize(value_only, x0, method="Powell", bounds=opt_bounds)
                out.append((res.x, res.fun))
            elif callable(self.acq_optimizer):
                def obj_func(x, eval_gradient=False):
                    f, g = batch(np.atleast_2d(x))
                    return (f[0], g[0]) if eval_gradient else f[0]
                out.append(tuple(self.acq_optimizer(obj_func, x0, bounds=opt_bounds)))
            else:
                raise ValueError("Unknown optimizer %s." % self.acq_optimizer)
        return out

    def optimize_acquisition_function(self, gpr, i, bounds=None, rng=None):
        """One restart (gp_acquisition.py:262-390) -> (x_opt in optimiser coordinates, -acq)."""
        rng = np.random.default_rng(rng) if not isinstance(rng, np.random.Generator) else rng
        use_bounds = self.bounds_ if bounds is None else np.asarray(bounds, dtype=float)
        self.proposer.update(gpr)
        self.proposer.update_bounds(use_bounds)
        x0, value, optimise = self._starting_points(gpr, [i], use_bounds, rng)
        if not optimise[0]:
            return self._to_opt(x0)[0], -1 * value[0]
        x, f = self._optimize_all(gpr, self._to_opt(x0), self._opt_bounds(use_bounds))[0]
        return x, f

    def multi_add(self, gpr, n_points=1, bounds=None, rng=None, force_resample=False):
        """gp_acquisition.py:392-501 -> (X (n_points, d), y_lies, acq values), the same on
        every rank."""
        from . import parallel
        if not (isinstance(n_points, int) and n_points > 0):
            raise ValueError(f"n_points should be int > 0, got {n_points}")
        rng = np.random.default_rng(rng) if not isinstance(rng, np.random.Generator) else rng
        use_bounds = self.bounds_ if bounds is None else np.asarray(bounds, dtype=float)
        opt_bounds = self._opt_bounds(use_bounds)
        X_opts = np.empty((n_points, gpr.d))
        y_lies = np.empty(n_points)
        acq_vals = np.empty(n_points)
        gpr_ = parallel.bcast(deepcopy(gpr) if parallel.is_main_process() else None)
        n_per = parallel.split_number_for_parallel_processes(self.n_restarts_optimizer)
        n_this = n_per[parallel.rank()]
        i_first = int(sum(n_per[:parallel.rank()]))
        for ipoint in range(n_points):
            self.proposer.update(gpr_)
            self.proposer.update_bounds(use_bounds)
            proposal_X = np.empty((n_this, gpr_.d))
            acq_X = np.empty(n_this)
            if n_this:
                x0, value, optimise = self._starting_points(
                    gpr_, list(range(i_first, i_first + n_this)), use_bounds, rng)
                x0_opt = self._to_opt(x0)
                proposal_X[:] = x0_opt
                acq_X[:] = -1 * value
                run = np.flatnonzero(optimise)
                if len(run):
                    for r, (x, f) in zip(run, self._optimize_all(gpr_, x0_opt[run], opt_bounds)):
                        proposal_X[r], acq_X[r] = x, f
            if parallel.multiple_processes():
                parts = parallel.allgather((proposal_X, acq_X))
                proposal_X = np.concatenate([p[0] for p in parts])
                acq_X = np.concatenate([p[1] for p in parts])
            # gp_acquisition.py:471-497, identical on every rank (same gathered arrays)
            max_pos = np.argmin(acq_X) if np.any(np.isfinite(acq_X)) else len(acq_X) - 1
            X_opt = self._from_opt(np.array([proposal_X[max_pos]]))
            acq_val = -1 * acq_X[max_pos]
            y_lie = gpr_.predict(X_opt)
            if ipoint < n_points - 1:
                lie_noise_level = (np.array([np.mean(gpr_.noise_level)])
                                   if np.iterable(gpr_.noise_level) else None)
                gpr_.append_to_data(X_opt, y_lie, noise_level=lie_noise_level, fit_gpr=False,
                                    fit_classifier=False)
            X_opts[ipoint], y_lies[ipoint], acq_vals[ipoint] = X_opt[0], y_lie[0], acq_val
        gpr.n_eval = gpr_.n_eval
        return X_opts, y_lies, acq_vals
