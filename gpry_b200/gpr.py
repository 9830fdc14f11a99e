"""
``GaussianProcessRegressor`` with the interface of ``gpry.gpr.GaussianProcessRegressor``
(reference gpr.py:27-1488) whose arithmetic runs on a B200 through ``libgpry_b200.so``.

What stays on the host (exactly as in the reference, O(N) or control flow): data
bookkeeping in ``append_to_data`` (:577-753), pre-processor scalars, the infinities
classifier and trust-region masks around ``predict`` (:1104-1174, 1196-1201), evaluation
counters (:1088, 880, 1296), the L-BFGS-B driver of ``fit_gpr_hyperparameters`` (:883-994).

What runs on the GPU: ``predict`` / ``predict_std`` arithmetic (:1176-1227, 1325-1347),
``_update_model`` + ``_kernel_inverse`` (:996-1020, 1453-1465), ``log_marginal_likelihood``
and its gradient (:876-881 -> sklearn _gpr.py:541-656), the single-point mean gradient
(:1236-1242).  There is no CPU fallback for any of these.

The instance stays picklable / deep-copyable like the reference's (:1354-1433, io.py): all
fitted attributes (``X_train_, y_train_, alpha, alpha_, L_, V_, kernel_`` ...) are numpy
arrays; the device handle lives outside the copied state and is rebuilt lazily.
"""
import os
import warnings
from copy import deepcopy
from numbers import Number
from operator import itemgetter

import numpy as np
import scipy.optimize

from .lockstep import lockstep_minimize
from .device import DeviceGP, workspace
from .kernels import ConstantKernel as C, RBF, Matern, Product
from .preprocessing import DummyPreprocessor


def default_device():
    """One process per GPU: LOCAL_RANK (torchrun) or the current torch device, else 0."""
    if "LOCAL_RANK" in os.environ:
        return int(os.environ["LOCAL_RANK"])
    try:
        import torch
        if torch.cuda.is_available() and torch.cuda.is_initialized():
            return torch.cuda.current_device()
    except ImportError:
        pass
    return 0


def is_in_bounds(points, bounds):
    """tools.py:263-287."""
    points = np.atleast_2d(points)
    return np.all((points >= bounds[:, 0]) & (points <= bounds[:, 1]), axis=1)


def shrink_bounds(bounds, samples, factor=1):
    """tools.py:308-361."""
    bounds = np.atleast_2d(bounds)
    samples = np.atleast_2d(samples)
    out = np.empty(shape=bounds.shape, dtype=float)
    out[:, 0] = samples.min(axis=0)
    out[:, 1] = samples.max(axis=0)
    width = out[:, 1] - out[:, 0]
    delta = (factor - 1) / 2 * width
    out[:, 0] -= delta
    out[:, 1] += delta
    out[:, 0] = np.array([out[:, 0], bounds[:, 0]]).max(axis=0)
    out[:, 1] = np.array([out[:, 1], bounds[:, 1]]).min(axis=0)
    return out


def check_random_state(seed):
    """tools.py:134-145 (Generators pass through, else sklearn's ``check_random_state``:
    None -> numpy's global RandomState, int -> RandomState(seed), RandomState -> itself)."""
    if isinstance(seed, (np.random.Generator, np.random.RandomState)):
        return seed
    if seed is None or seed is np.random:
        return np.random.mtrand._rand
    if isinstance(seed, (int, np.integer)):
        return np.random.RandomState(seed)
    raise ValueError(f"{seed!r} cannot be used to seed a numpy.random.RandomState instance")


def delta_logp_of_1d_nstd(n1, d):
    """tools.py:100-118."""
    from scipy.stats import chi2
    from scipy.special import erfc
    return 0.5 * chi2.isf(erfc(n1 / np.sqrt(2)), d)


class GaussianProcessRegressor:
    """See the module docstring; constructor arguments as gpr.py:265-271 plus ``device``."""

    def __init__(self, kernel="RBF", output_scale_prior=[1e-2, 1e3],
                 length_scale_prior=[1e-3, 1e1], noise_level=1e-2, clip_factor=1.1,
                 optimizer="fmin_l_bfgs_b", n_restarts_optimizer=0,
                 preprocessing_X=None, preprocessing_y=None,
                 account_for_inf="SVM", inf_threshold="20s", keep_min_finite=None,
                 trust_region_factor=None, trust_region_nstd=None,
                 bounds=None, random_state=None, verbose=1, device=None, contraction=None):
        self.n_last_appended = 0
        self.n_last_appended_finite = 0
        self.newly_appended_for_inv = 0
        self.preprocessing_X = DummyPreprocessor if preprocessing_X is None else preprocessing_X
        self.preprocessing_y = DummyPreprocessor if preprocessing_y is None else preprocessing_y
        self.noise_level = noise_level
        if clip_factor is not None and clip_factor < 1:
            raise ValueError("'clip_factor' must be >= 1, or None for no clippling.")
        self.clip_factor = clip_factor
        self.n_eval = 0
        self.n_eval_loglike = 0
        self.verbose = verbose
        self.inf_value = np.inf
        self.minus_inf_value = -np.inf
        self._fitted = False
        if bounds is None:
            raise ValueError("'bounds' (prior bounds, shape (d, 2)) are required")
        self.bounds = np.asarray(bounds, dtype=float)
        self.trust_bounds = None
        self.trust_region_factor = trust_region_factor
        self.trust_region_nstd = trust_region_nstd
        self.optimizer = optimizer
        self.n_restarts_optimizer = n_restarts_optimizer
        self.random_state = random_state
        self.inf_threshold = inf_threshold
        # Infinities classifier (gpr.py:296-303): "SVM" = gpry_b200.svm.SVM, trained on the host
        # like the reference's, evaluated on the device inside the predict calls.  Any other
        # object with the interface of gpry.svm.SVM (fit / predict / _is_finite_raw) can be
        # plugged in and is then called on the host, as in the reference.
        if isinstance(account_for_inf, str) and account_for_inf.lower() == "svm":
            from .svm import SVM
            self.infinities_classifier = SVM(random_state=random_state, device=device)
        elif account_for_inf is False or account_for_inf is None:
            self.infinities_classifier = None
        else:
            self.infinities_classifier = account_for_inf
        # Auto-construct inbuilt kernels (gpr.py:329-363)
        if isinstance(kernel, str):
            kernel = {kernel: {}}
        if isinstance(kernel, dict):
            if len(kernel) != 1:
                raise ValueError("'kernel' must be a single-key dict.")
            kernel_name = list(kernel)[0]
            kernel_args = kernel[kernel_name] or {}
            self.bounds_ = self.preprocessing_X.transform_bounds(self.bounds)
            try:
                length_corr_kernel = {"rbf": RBF, "matern": Matern}[kernel_name.lower()]
            except KeyError as excpt:
                raise ValueError("Currently only 'RBF' and 'Matern' are supported as "
                                 f"standard kernels. Got '{kernel_name}'.") from excpt
            output_scale_init = np.sqrt(output_scale_prior[0] * output_scale_prior[1])
            length_scale_init = np.sqrt(length_scale_prior[0] * length_scale_prior[1])
            kernel = (
                C(output_scale_init ** 2,
                  [output_scale_prior[0] ** 2, output_scale_prior[1] ** 2])
                * length_corr_kernel([length_scale_init] * self.d, length_scale_prior,
                                     prior_bounds=self.bounds_, **kernel_args))
        if not isinstance(kernel, Product):
            raise NotImplementedError("kernel must be 'RBF', 'Matern', {'Matern': {'nu': ..}} "
                                      "or ConstantKernel * RBF/Matern")
        self.kernel = kernel
        self.alpha = noise_level ** 2.
        d = self.d
        self.X_train, self.y_train = np.empty((0, d)), np.empty((0,))
        self.X_train_, self.y_train_ = None, None
        self.X_train_all, self.y_train_all = np.empty((0, d)), np.empty((0,))
        self.X_train_all_, self.y_train_all_ = None, None
        self.noise_level_ = None
        self.kernel_ = None
        self.keep_min_finite = keep_min_finite if keep_min_finite is not None else max(2, d)
        self._diff_threshold = None
        if self.infinities_classifier is not None:
            if isinstance(inf_threshold, str) and inf_threshold.endswith("s"):
                self._diff_threshold = delta_logp_of_1d_nstd(float(inf_threshold[:-1]), d)
            else:
                self._diff_threshold = float(inf_threshold)
        self.device = default_device() if device is None else int(device)
        # variance contraction of large pools: None = library default ("int8": exact integer
        # split on the INT8 tensor cores), "fp64" = FP64 tensor cores, "int8_1pass"
        self.contraction = contraction
        self._dev = None          # DeviceGP holding the predict state (never pickled)
        self._dev_dirty = True
        self._clf_bound = None    # (id, version) of the classifier bound to the device state
        self._clf_const = None    # True / False when the classifier answers without an SVC
        self._factor_resident = False   # L_ / V_ live on the device only (see the properties)

    # ------------------------------------------------------------------ pickling / copies
    _NOT_COPIED = ("_dev", "_dev_dirty", "_clf_bound", "_clf_const", "_factor_resident",
                   "_fact_sig")

    def __getstate__(self):
        self._materialize_factor()      # L_ / V_ may still live on the device only
        state = {k: v for k, v in self.__dict__.items() if k not in self._NOT_COPIED}
        return state

    def __setstate__(self, state):
        self.__dict__.update(state)
        if "LOCAL_RANK" in os.environ:
            # device ordinals are per process: a regressor broadcast from another rank
            # (mpi.bcast of the reference, gp_acquisition.py:453) must use this rank's GPU
            self.device = default_device()
        self._dev = None
        self._dev_dirty = True
        self._factor_resident = False
        self._clf_bound = None
        self._clf_const = None

    def __deepcopy__(self, memo):
        """Same observable result as gpr.py:1354-1433 (a fresh instance carrying copies of
        the data and fitted attributes); the device state is rebuilt lazily by the copy."""
        new = self.__class__.__new__(self.__class__)
        memo[id(self)] = new
        self._materialize_factor()
        for k, v in self.__dict__.items():
            if k in self._NOT_COPIED:
                continue
            new.__dict__[k] = deepcopy(v, memo)
        new._dev = None
        new._dev_dirty = True
        new._clf_bound = None
        new._clf_const = None
        new._factor_resident = False
        return new

    # ------------------------------------------------------------------ lazy L_ / V_
    # gpr.py:1456-1457 keeps L_ and V_ = L^-1 as dense host arrays.  Here the factorisation is
    # produced and consumed on the device (``gpry_factorize(keep_on_device)`` ->
    # ``gpry_state_adopt_factorization``); the two N x N host copies (2 x 128 MB at N = 4000,
    # most of the cost of a model update) are fetched only when something reads the
    # attributes: a pickle, a deep copy, a plot, a test.
    def _materialize_factor(self):
        if self.__dict__.get("_factor_resident") and self.__dict__.get("_V") is None:
            self._L, self._V = self._dev.factor_download()   # the device keeps its copy

    @property
    def L_(self):
        self._materialize_factor()
        return self.__dict__.get("_L")

    @L_.setter
    def L_(self, value):
        self._L = value
        self._factor_resident = False

    @property
    def V_(self):
        self._materialize_factor()
        return self.__dict__.get("_V")

    @V_.setter
    def V_(self, value):
        self._V = value
        self._factor_resident = False

    # ------------------------------------------------------------------ properties
    @property
    def d(self):
        if self.bounds is None:
            return self.X_train.shape[1]
        return self.bounds.shape[0]

    @property
    def y_max(self):
        return np.max(getattr(self, "y_train", [self.minus_inf_value]))

    @property
    def n(self):
        return len(getattr(self, "y_train", []))

    n_finite = n

    @property
    def n_total(self):
        if self.infinities_classifier:
            return self.infinities_classifier.n or self.n
        return self.n

    @property
    def fitted(self):
        return self._fitted

    @property
    def last_appended_finite(self):
        return (np.copy(self.X_train[-self.n_last_appended_finite:]),
                np.copy(self.y_train[-self.n_last_appended_finite:]))

    @property
    def last_appended(self):
        if self.infinities_classifier is None:
            return self.last_appended_finite
        return (np.copy(self.X_train_all[-self.n_last_appended:]),
                np.copy(self.y_train_all[-self.n_last_appended:]))

    @property
    def scales(self):
        return (self.preprocessing_y.inverse_transform_scale(
            np.sqrt(self.kernel_.k1.constant_value)),
            tuple(self.preprocessing_X.inverse_transform_scale(self.kernel_.k2.length_scale)))

    def is_finite(self, y):
        if self.infinities_classifier is None:
            return np.full(shape=len(y), fill_value=True)
        return self.infinities_classifier.is_finite(self.preprocessing_y.transform(y))

    def update_trust_region(self):
        """gpr.py:554-575."""
        if self.trust_region_factor is None:
            return
        if self.trust_region_nstd is None:
            use_X = self.X_train
        else:
            nstd = self.trust_region_nstd
            use_X = np.empty(shape=(0, self.X_train.shape[1]))
            while len(use_X) < min(self.d, self.n):
                use_X = self.X_train[np.where(
                    max(self.y_train) - self.y_train < delta_logp_of_1d_nstd(nstd, self.d))]
                nstd = nstd + 0.1
        self.trust_bounds = shrink_bounds(self.bounds, use_X, factor=self.trust_region_factor)

    # ------------------------------------------------------------------ data
    def _validate_noise_level(self, noise_level, n_train):
        """gpr.py:755-785."""
        if n_train == 0 and noise_level is not None:
            raise ValueError("noise_level must be None if not fitting to new points.")
        if np.iterable(noise_level):
            noise_level = np.atleast_1d(noise_level)
            if noise_level.shape[0] != n_train:
                raise ValueError("noise_level must be an array with same number of entries "
                                 f"as y, but len(n)={noise_level.shape[0]} != len(y)={n_train})")
        elif isinstance(noise_level, Number):
            if np.iterable(self.noise_level):
                noise_level = np.full(fill_value=noise_level, shape=(n_train,))
        elif noise_level is None:
            if np.iterable(self.noise_level):
                raise ValueError("Need to pass non-null noise_level (scalar or array) because "
                                 "concrete values were given earlier for the training points.")
        else:
            raise ValueError("noise_level needs to be an iterable, number or None. "
                             f"Got type(noise_level)={type(noise_level)}")
        return noise_level

    def _update_noise_level(self, noise_level):
        """gpr.py:787-817."""
        if np.iterable(noise_level):
            if not np.iterable(self.noise_level):
                self.noise_level = np.full(fill_value=self.noise_level,
                                           shape=(len(self.y_train_all) - len(noise_level),))
            self.noise_level = np.append(self.noise_level, noise_level, axis=0)
        elif isinstance(noise_level, Number):
            if not np.isclose(noise_level, self.noise_level):
                self.noise_level = noise_level

    @staticmethod
    def _diff_threshold_if_keep_n_finite(y, n, reference_diff_threshold, epsilon=1e-6):
        """gpr.py:1475-1488."""
        if n is None or n <= 1:
            return reference_diff_threshold
        y_sorted = np.sort(y)
        difference_to_nth_point = y_sorted[-1] - y_sorted[-min(n, len(y_sorted))]
        return max(reference_diff_threshold, difference_to_nth_point + epsilon)

    @staticmethod
    def _fit_request(fit_gpr, fit_classifier):
        """Normalises the ``fit_gpr`` argument (gpr.py:648-668) -> (kwargs or None, refit
        classifier?).  A hyper-parameter fit implies a classifier (and pre-processor) refit."""
        if fit_gpr is False:
            return None, bool(fit_classifier)
        if fit_gpr is True:
            return {}, True
        if isinstance(fit_gpr, str) and fit_gpr == "simple":
            return {"simple": True}, True
        if isinstance(fit_gpr, dict):
            return deepcopy(fit_gpr), True
        raise ValueError("`fit_gpr` needs to be bool, 'simple', or a dict of args for the "
                         f"`fit_gpr_hyperparameters` method. Got {fit_gpr}.")

    def _finite_mask(self):
        """Which of all points added so far enter the GPR (gpr.py:689-704); the rest is left
        to the infinities classifier.  Returns (mask, threshold used or None)."""
        if self.infinities_classifier is None:
            return np.ones(len(self.y_train_all), dtype=bool), None
        threshold = self._diff_threshold_if_keep_n_finite(
            self.y_train_all, self.keep_min_finite, self._diff_threshold)
        return self.infinities_classifier._is_finite_raw(self.y_train_all, threshold), threshold

    def append_to_data(self, X, y, noise_level=None, fit_gpr=True, fit_classifier=True):
        """Adds points and updates the model: gpr.py:577-753.  ``fit_gpr``: True / dict /
        'simple' re-fit the hyper-parameters (GPU: batched LML + gradient), False keeps them
        and only re-factorises (GPU: Cholesky, L^-1, alpha).  ``X = y = None`` re-fits without
        new points.  Host-side bookkeeping is as in the reference."""
        fit_kwargs, refit_classifier = self._fit_request(fit_gpr, fit_classifier)
        refit_only = X is None and y is None
        if refit_only:
            if noise_level is not None:
                raise ValueError("Cannot give a noise level if X and y are not given.")
            X, y = np.empty((0, self.d)), np.empty((0,))
        elif X is None or y is None:
            raise ValueError("If passing X, y needs to be passed too, and viceversa.")
        X = np.atleast_2d(np.asarray(X, dtype=float))
        y = np.atleast_1d(np.asarray(y, dtype=float))
        new_noise = self._validate_noise_level(noise_level, len(y))
        # 1. raw bookkeeping
        self.n_last_appended = len(y)
        self.X_train_all = np.append(self.X_train_all, X, axis=0)
        self.y_train_all = np.append(self.y_train_all, y)
        self._update_noise_level(new_noise)
        # 2. finite selection, pre-processors (refit together with the classifier), transforms
        finite, threshold = self._finite_mask()
        if refit_classifier:
            self.preprocessing_X.fit(self.X_train_all[finite], self.y_train_all[finite])
            self.preprocessing_y.fit(self.X_train_all[finite], self.y_train_all[finite])
        self.X_train_all_ = self.preprocessing_X.transform(self.X_train_all)
        self.y_train_all_ = self.preprocessing_y.transform(self.y_train_all)
        noise_all = (np.full(len(self.y_train_all_), self.noise_level)
                     if isinstance(self.noise_level, Number) else self.noise_level)
        self.noise_level_ = self.preprocessing_y.transform_scale(noise_all)
        # 3. classifier (lives in the transformed space: gets all points every time)
        if self.infinities_classifier is not None and refit_classifier:
            predicted = self.infinities_classifier.fit(
                self.X_train_all_, self.y_train_all_,
                self.preprocessing_y.transform_scale(threshold))
            assert np.array_equal(finite, predicted), \
                "Infinities classifier miss-classified at least 1 point."
        n_new = self.n_last_appended
        if self.infinities_classifier is None:
            self.n_last_appended_finite = n_new
        else:   # NB: the reference slices [-n_new:], i.e. ALL points when n_new == 0 (:735-737)
            self.n_last_appended_finite = int(np.sum(finite[-n_new:]))
        if not self.n_last_appended_finite and not (refit_only and fit_kwargs is not None):
            return self        # nothing new for the GPR (gpr.py:738-741)
        # 4. GPR training set in both spaces, then the device work
        self.X_train = np.copy(self.X_train_all[finite])
        self.y_train = np.copy(self.y_train_all[finite])
        self.X_train_ = self.preprocessing_X.transform(self.X_train)
        self.y_train_ = self.preprocessing_y.transform(self.y_train)
        self.alpha = self.noise_level_[finite] ** 2       # NB: not alpha_ (gpr.py:747)
        self.newly_appended_for_inv = self.n_last_appended_finite
        self._dev_dirty = True
        if fit_kwargs is not None:
            self.fit_gpr_hyperparameters(**fit_kwargs)
        else:
            self._update_model()
        self.update_trust_region()
        return self

    # ------------------------------------------------------------------ kernel plumbing
    def _kernel_spec(self, kernel=None):
        kernel = self.kernel_ if kernel is None else kernel
        if not kernel.theta_is_standard(self.d):
            raise NotImplementedError(
                "the B200 path needs theta = [log c, log l_1..l_d] (free constant, free "
                "anisotropic length scales), as GPry's auto-constructed kernels have")
        return kernel.device_spec(self.d)

    # ------------------------------------------------------------------ LML + fit
    def log_marginal_likelihood(self, theta=None, eval_gradient=False, clone_kernel=True):
        """gpr.py:876-881 -> sklearn _gpr.py:541-656, evaluated by gpry_lml_batched."""
        self.n_eval_loglike += 1
        if theta is None:
            if eval_gradient:
                raise ValueError("Gradient can only be evaluated for theta!=None")
            return self.log_marginal_likelihood_value_
        theta = np.asarray(theta, dtype=float)
        if clone_kernel:
            kernel = self.kernel_.clone_with_theta(theta)
        else:
            kernel = self.kernel_
            kernel.theta = theta
        kind, _, _ = self._kernel_spec(kernel)
        lml, grad, info = workspace(self.device).lml_batched(
            kind, self.X_train_, self.alpha, self.y_train_, theta[None, :],
            eval_gradient=eval_gradient)
        if eval_gradient:
            return float(lml[0]), grad[0]
        return float(lml[0])

    def log_marginal_likelihood_batch(self, thetas, eval_gradient=True):
        """LML (and gradient) for several theta at once (one device call; evaluations overlap
        on the GPU).  Counts ``len(thetas)`` evaluations.  Does not touch ``kernel_``."""
        thetas = np.atleast_2d(np.asarray(thetas, dtype=float))
        self.n_eval_loglike += len(thetas)
        kind, _, _ = self._kernel_spec()
        lml, grad, info = workspace(self.device).lml_batched(
            kind, self.X_train_, self.alpha, self.y_train_, thetas, eval_gradient=eval_gradient)
        return (lml, grad) if eval_gradient else lml

    def fit_gpr_hyperparameters(self, simple=False, start_from_current=True, n_restarts=None,
                                hyperparameter_bounds=None, lockstep=True):
        """gpr.py:883-994.  Same restarts, same starting points (``rng.uniform`` over the
        log-bounds in loop order), same scipy L-BFGS-B driver per restart.  With
        ``lockstep=True`` (default) the restarts advance together and every round of
        objective evaluations is ONE batched device call; each restart sees the function
        values it would see alone up to round-off (see ``gpry_b200.lockstep``), so it ends at
        the same optimum to the optimiser's tolerance."""
        if simple:
            start_from_current = True
            n_restarts = 1
        if not self._fitted:
            start_from_current = False
        if n_restarts is None:
            n_restarts = self.n_restarts_optimizer
        if self.kernel_ is None:
            self.kernel_ = deepcopy(self.kernel)
        no_optimizer = self.optimizer is None
        no_hyperparams = self.kernel.n_dims == 0
        no_restarts = n_restarts <= 0
        if no_optimizer or no_hyperparams or no_restarts:
            reasons = []
            if no_optimizer:
                reasons += ["no optimizer has been specified"]
            if no_hyperparams:
                reasons += ["the kernel has no hyperparamenters"]
            if no_restarts:
                reasons += ["the number of optimizer restarts requested is 0."]
            warnings.warn(f"Hyper-parameters not (re)fit. Reason(s): {'; '.join(reasons)}.")
            self.log_marginal_likelihood_value_ = self.log_marginal_likelihood(
                self.kernel_.theta, clone_kernel=False)
            self._update_model()
            return self
        if hyperparameter_bounds is None:
            hyperparameter_bounds = self.kernel_.bounds
        if n_restarts - int(start_from_current):
            if not np.isfinite(hyperparameter_bounds).all():
                raise ValueError("There is at least one optimizer run the requires sampling "
                                 "from the hyperparameters' prior, but it has not finite "
                                 "density, because not all bounds are finite.")
        self._rng = check_random_state(self.random_state)
        theta_initials = []
        for iteration in range(n_restarts):
            if iteration == 0 and start_from_current:
                theta_initials.append(np.array(self.kernel_.theta))
            else:
                theta_initials.append(self._rng.uniform(hyperparameter_bounds[:, 0],
                                                        hyperparameter_bounds[:, 1]))
        if lockstep and n_restarts > 1 and self.optimizer == "fmin_l_bfgs_b":
            optima = self._lockstep_optimization(theta_initials, hyperparameter_bounds)
        else:
            def obj_func(theta, eval_gradient=True):
                if eval_gradient:
                    lml, grad = self.log_marginal_likelihood(theta, eval_gradient=True,
                                                             clone_kernel=False)
                    return -lml, -grad
                return -self.log_marginal_likelihood(theta, clone_kernel=False)
            optima = [self._constrained_optimization(obj_func, th0, hyperparameter_bounds)
                      for th0 in theta_initials]
        lml_values = list(map(itemgetter(1), optima))
        self.log_marginal_likelihood_value_ = -np.min(lml_values)
        self.kernel_.theta = optima[np.argmin(lml_values)][0]
        self.newly_appended_for_inv = max(self.newly_appended_for_inv, 1)
        self._update_model()
        self._fitted = True
        return self

    def _constrained_optimization(self, obj_func, initial_theta, bounds):
        """gpr.py:1435-1451."""
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            if self.optimizer == "fmin_l_bfgs_b":
                opt_res = scipy.optimize.minimize(obj_func, initial_theta, method="L-BFGS-B",
                                                  jac=True, bounds=bounds)
                return opt_res.x, opt_res.fun
            if callable(self.optimizer):
                return self.optimizer(obj_func, initial_theta, bounds=bounds)
            raise ValueError("Unknown optimizer %s." % self.optimizer)

    def _lockstep_optimization(self, theta_initials, bounds):
        """One scipy L-BFGS-B per restart, their objective calls evaluated together by
        ``log_marginal_likelihood_batch`` (see ``gpry_b200.lockstep``)."""
        def batch(thetas):
            lml, grad = self.log_marginal_likelihood_batch(thetas, eval_gradient=True)
            return -lml, -grad
        return lockstep_minimize(batch, theta_initials, bounds)

    # ------------------------------------------------------------------ model update
    def _update_model(self):
        """gpr.py:996-1020: K = k(X_,X_) + diag(alpha); Cholesky; L^-1; alpha_ (on the GPU)."""
        if self.newly_appended_for_inv < 1:
            warnings.warn("No new points have been appended to the model.")
            return self
        self._kernel_inverse()
        self.newly_appended_for_inv = 0
        return self

    def _kernel_inverse(self, kernel=None):
        """gpr.py:1453-1465.  The reference receives the assembled matrix; here the kernel
        matrix is built on the device from ``kernel_``, so the argument is ignored."""
        kind, _, _ = self._kernel_spec()
        N = len(self.y_train_)
        if self._dev is None:
            self._dev = DeviceGP(self.device)
        noise2 = np.ascontiguousarray(np.broadcast_to(self.alpha, (N,)), dtype=float)
        theta = np.array(self.kernel_.theta, dtype=float)
        alpha_, info = None, 0
        sig = self.__dict__.get("_fact_sig")
        # Bordered append instead of a refactorisation when only rows were added at the end
        # (same theta, same transformed old points, same noise on them): the Kriging-believer
        # lies (gp_acquisition.py:488-491), appends with fit_classifier=False, ...
        if (self._factor_resident and sig is not None and sig["kind"] == kind
                and 0 < N - sig["N"] <= 64 and -(-N // 128) == -(-sig["N"] // 128)
                and np.array_equal(sig["theta"], theta)
                and np.array_equal(sig["X_"], self.X_train_[:sig["N"]])
                and np.array_equal(sig["noise2"], noise2[:sig["N"]])):
            self._factor_resident = False
            alpha_, info = self._dev.factor_append(self.X_train_[sig["N"]:], noise2[sig["N"]:],
                                                   self.y_train_, theta)
            self.n_appends_without_refactor = self.__dict__.get("n_appends_without_refactor", 0) + 1
        else:
            self._factor_resident = False  # whatever was resident is overwritten from here on
            _, _, alpha_, _, info = self._dev.factorize(
                kind, self.X_train_, noise2, self.y_train_, theta, want_L=False, want_V=False,
                keep_on_device=True)
        if info != 0:
            # nothing of the previous (smaller) factorisation may survive next to the new
            # training set: a later predict must fail with "no valid factorisation"
            self._fact_sig = None
            self._L = self._V = None
            self.alpha_ = None
            self._factor_resident = False
            self._dev_dirty = True
            raise np.linalg.LinAlgError(
                "The kernel, %s, is not returning a positive definite matrix. Try gradually "
                "increasing the 'noise_level' parameter of your GaussianProcessRegressor "
                "estimator." % self.kernel_,
                f"{info}-th leading minor of the array is not positive definite")
        self._fact_sig = {"kind": kind, "N": N, "theta": theta, "X_": np.array(self.X_train_),
                          "noise2": noise2}
        self._L = self._V = None           # fetched from the device on first access
        self.alpha_ = alpha_
        self._factor_resident = True
        self._dev_dirty = True

    def drop_resident_factor(self):
        """Forget the device-resident factorisation: the next ``_update_model`` factorises from
        scratch instead of bordering the old factor."""
        self._factor_resident = False
        self._fact_sig = None

    # ------------------------------------------------------------------ device state
    def _device_state(self):
        """Uploads (lazily, once per model change) what predict needs."""
        if self._dev is None:
            self._dev = DeviceGP(self.device)
            self._dev_dirty = True
        if self.__dict__.get("contraction"):
            self._dev.set_contract_mode(self.contraction)
        if self._dev_dirty:
            kind, c, ell = self._kernel_spec()
            px, py = self.preprocessing_X, self.preprocessing_y
            if hasattr(px, "bounds_min"):
                x_min, x_width = px.bounds_min, px.bounds_max - px.bounds_min
            elif px is DummyPreprocessor or isinstance(px, DummyPreprocessor):
                x_min = x_width = None
            else:
                raise NotImplementedError("X pre-processor must be Normalize_bounds or None")
            if hasattr(py, "mean_"):
                y_mean, y_std = py.mean_, py.std_
            elif py is DummyPreprocessor or isinstance(py, DummyPreprocessor):
                y_mean, y_std = 0.0, 1.0
            else:
                raise NotImplementedError("y pre-processor must be Normalize_y or None")
            clip_hi = np.inf
            if self.clip_factor is not None:   # gpr.py:1187-1195
                clip_hi = (self.clip_factor * max(self.y_train)
                           - (self.clip_factor - 1) * min(self.y_train))
            if self._factor_resident:      # straight from the factorisation, no N^2 transfer
                self._dev.adopt_factorization(c, ell, x_min, x_width, y_mean, y_std, clip_hi)
            else:
                if self.V_ is None:
                    raise ValueError("the model has no valid factorisation (not fit to data "
                                     "yet, or the last factorisation failed)")
                self._dev.upload(kind, self.X_train_, self.alpha_, self.V_, c, ell, x_min,
                                 x_width, y_mean, y_std, clip_hi)
            self._dev_dirty = False
            self._clf_bound = None          # an upload clears the device-side classifier
        clf = self.infinities_classifier
        if self._classifier_on_device():
            tag = (id(clf), clf.version)
            if self._clf_bound != tag:
                spec = clf.device_spec()
                self._clf_const = spec[1] if spec[0] == "all" else None
                self._dev.set_classifier(None if spec[0] == "all" else spec[1:])
                self._clf_bound = tag
        elif self._clf_bound is not None:
            self._dev.set_classifier(None)
            self._clf_bound = self._clf_const = None
        return self._dev

    def _classifier_on_device(self):
        """True if the infinities classifier can be evaluated by the library (``gpry_b200.svm``
        or anything else exposing ``device_spec``) and has been trained."""
        clf = self.infinities_classifier
        return clf is not None and hasattr(clf, "device_spec") and clf.y_train is not None

    def _set_masks(self, dev, trust=True):
        """Per-call device masks: ``minus_inf_value`` is read at call time (gp_acquisition.py:
        788-792 changes it temporarily), the trust region can be switched off per call."""
        dev.set_mask_value(self.minus_inf_value)
        dev.set_trust_region(self.trust_bounds if trust else None, self.minus_inf_value)

    # ------------------------------------------------------------------ predict
    @staticmethod
    def _as_2d(X, validate):
        if validate:
            X = np.asarray(X, dtype=float)
            if X.ndim != 2:
                raise ValueError(f"Expected 2D array, got {X.ndim}D array instead")
            if not np.all(np.isfinite(X)):
                raise ValueError("Input contains NaN or infinity")
        return X

    def predict(self, X, return_std=False, return_cov=False, return_mean_grad=False,
                return_std_grad=False, validate=True, ignore_trust_region=False):
        """gpr.py:1022-1273."""
        self.n_eval += len(X)
        # NB: ``return_cov`` is accepted and ignored, exactly as in the reference (its body
        # never reads the argument, gpr.py:1022-1273)
        if return_std_grad and not (return_std and return_mean_grad):
            raise ValueError("Not returning std_gradient without returning the std and the "
                             "mean grad.")
        if X.shape[0] != 1 and (return_mean_grad or return_std_grad):
            raise ValueError("Mean grad and std grad not implemented for n_samples > 1")
        X = self._as_2d(X, validate)
        impose_trust_region = self.trust_bounds is not None and not ignore_trust_region
        if self.X_train_ is None:
            return self._predict_from_prior(X, return_std, return_mean_grad, return_std_grad,
                                            impose_trust_region)
        clf = self.infinities_classifier
        clf_on_device = self._classifier_on_device()
        host_clf = clf is not None and not clf_on_device
        # the masks are applied on the device (no O(M d) host pass over large pools) unless
        # the classifier is a foreign host object: then rows are re-packed on the host anyway
        i_outside_trust = None
        if impose_trust_region and host_clf:
            i_outside_trust = np.logical_not(is_in_bounds(X, self.trust_bounds))
        finite = None
        dev = None
        n_samples, n_dims = X.shape
        all_infinite = False
        if host_clf:   # gpr.py:1136-1174
            X = np.copy(X)
            y_mean_full = np.ones(n_samples)
            y_std_full = np.zeros(n_samples)
            grad_mean_full = np.ones((n_samples, n_dims))
            grad_std_full = np.zeros((n_samples, n_dims))
            X_ = self.preprocessing_X.transform(X)
            finite = clf.predict(np.ascontiguousarray(X_), validate=validate)
            all_infinite = bool(np.all(~finite))
        elif clf_on_device:
            dev = self._device_state()
            all_infinite = self._clf_const is False
            if not all_infinite and self._clf_const is None and return_mean_grad:
                all_infinite = not dev.classify(X)[0] > 0      # one point (checked above)
        if all_infinite:
            y_mean = np.ones(n_samples) * self.minus_inf_value
            out = [y_mean]
            if return_std:
                out.append(np.zeros(n_samples))
            if return_mean_grad:
                out.append(np.ones((n_samples, n_dims)) * self.inf_value)
            if return_std_grad:
                out.append(np.zeros((n_samples, n_dims)))
            return out[0] if len(out) == 1 else tuple(out)
        if host_clf:
            y_mean_full[~finite] = self.minus_inf_value
            grad_mean_full[~finite] = self.inf_value
            X = X[finite]
        dev = self._device_state() if dev is None else dev
        self._set_masks(dev, trust=impose_trust_region and not host_clf)
        y_mean, y_std = dev.predict(X, return_mean=True, return_std=return_std)
        if finite is not None:
            y_mean_full[finite] = y_mean
            y_mean = y_mean_full
        if i_outside_trust is not None:
            y_mean[i_outside_trust] = self.minus_inf_value
        if return_std:
            if finite is not None:
                y_std_full[finite] = y_std
                y_std = y_std_full
            if not return_mean_grad:
                return y_mean, y_std
        if return_mean_grad:
            grad_mean = dev.mean_grad(X[0])
            if finite is not None:
                grad_mean_full[finite] = grad_mean
                grad_mean = grad_mean_full
            if return_std_grad:   # gpr.py:1247-1266
                grad_std = np.zeros(X.shape[1])
                if not np.allclose(y_std, grad_std):
                    grad_std, _ = dev.std_grad(X[0])
                    if finite is not None:
                        grad_std_full[finite] = grad_std
                        grad_std = grad_std_full
                return y_mean, y_std, grad_mean, grad_std
            if return_std:
                return y_mean, y_std, grad_mean
            return y_mean, grad_mean
        return y_mean

    def _predict_from_prior(self, X, return_std, return_mean_grad, return_std_grad,
                            impose_trust_region):
        """Never fit to data: the GP prior, gpr.py:1111-1133 (zero mean, std = sqrt(k(x, x)),
        zero gradients).  The reference's own guard is ``hasattr(self, "X_train_")``, which its
        constructor makes always true (gpr.py:371), so there the branch is dead and an unfitted
        ``predict`` fails inside the kernel call; here the documented behaviour is served."""
        X = np.asarray(X, dtype=float)
        y_mean = np.zeros(X.shape[0])
        if impose_trust_region:
            y_mean[np.logical_not(is_in_bounds(X, self.trust_bounds))] = self.minus_inf_value
        out = [y_mean]
        if return_std:
            out.append(np.sqrt(self.kernel.diag(X)))
        if return_mean_grad:
            out.append(np.zeros_like(X))
            if return_std and return_std_grad:
                out.append(np.zeros_like(X))
        return out[0] if len(out) == 1 else tuple(out)

    def predict_grad_batch(self, X, return_std_grad=True, validate=True,
                           ignore_trust_region=False):
        """Row-wise equivalent of ``predict(x, return_std=True, return_mean_grad=True,
        return_std_grad=True)`` (gpr.py:1022-1273, which accepts one point per call) for M
        points in one device pass: ``(mean (M,), std (M,), grad_mean (M, d), grad_std (M, d))``.
        Same conventions per row: gradients w.r.t. the transformed coordinate, rows the
        classifier calls infinite get ``(minus_inf_value, 0, inf_value, 0)``, the trust region
        touches the mean only, ``grad_std = 0`` where ``std`` is numerically zero."""
        self.n_eval += len(X)
        X = self._as_2d(X, validate)
        n_samples, n_dims = X.shape
        finite = None
        if self.infinities_classifier is not None:
            X_ = self.preprocessing_X.transform(X)
            finite = np.asarray(self.infinities_classifier.predict(
                np.ascontiguousarray(X_), validate=validate), dtype=bool)
        y_mean = np.full(n_samples, self.minus_inf_value, dtype=float)
        y_std = np.zeros(n_samples)
        grad_mean = np.full((n_samples, n_dims), self.inf_value, dtype=float)
        grad_std = np.zeros((n_samples, n_dims))
        rows = np.arange(n_samples) if finite is None else np.flatnonzero(finite)
        dev = self._device_state()
        for lo in range(0, len(rows), 8192):
            sel = rows[lo:lo + 8192]
            m, sd, gm, gs = dev.predict_grad(X[sel], return_std_grad=return_std_grad)
            y_mean[sel], y_std[sel], grad_mean[sel] = m, sd, gm
            if return_std_grad:
                gs[np.abs(sd) <= 1e-8] = 0.0       # ``np.allclose(y_std, 0)``, gpr.py:1248
                grad_std[sel] = gs
        if self.trust_bounds is not None and not ignore_trust_region:
            y_mean[np.logical_not(is_in_bounds(X, self.trust_bounds))] = self.minus_inf_value
        if return_std_grad:
            return y_mean, y_std, grad_mean, grad_std
        return y_mean, y_std, grad_mean

    def predict_std(self, X, validate=True):
        """gpr.py:1275-1352 (no trust region)."""
        self.n_eval += len(X)
        X = self._as_2d(X, validate)
        if self.X_train_ is None:       # GP prior, gpr.py:1303-1305
            return np.sqrt(self.kernel.diag(np.asarray(X, dtype=float)))
        clf = self.infinities_classifier
        finite = None
        if clf is not None and not self._classifier_on_device():
            X = np.copy(X)
            n_samples = X.shape[0]
            y_std_full = np.zeros(n_samples)
            X_ = self.preprocessing_X.transform(X)
            finite = clf.predict(np.ascontiguousarray(X_), validate=validate)
            if np.all(~finite):
                return np.zeros(n_samples)
            X = X[finite]
        dev = self._device_state()
        if self._clf_const is False:
            return np.zeros(X.shape[0])
        self._set_masks(dev, trust=False)
        _, y_std = dev.predict(X, return_mean=False, return_std=True)
        if finite is not None:
            y_std_full[finite] = y_std
            y_std = y_std_full
        return y_std

    # ------------------------------------------------------------------ fused fast paths
    def _scalar_noise(self, noise_level):
        noise_level = self.noise_level if noise_level is None else noise_level
        return float(np.mean(noise_level)) if np.iterable(noise_level) else noise_level

    def _host_classifier_rows(self, X):
        """Rows a foreign (host-side) classifier calls finite, or None if there is nothing to
        do on the host (no classifier, or one the library evaluates itself)."""
        clf = self.infinities_classifier
        if clf is None or self._classifier_on_device():
            return None
        if hasattr(X, "is_cuda"):
            raise NotImplementedError("a host-side classifier cannot mask device-resident "
                                      "candidates; use account_for_inf='SVM' (gpry_b200.svm)")
        X_ = self.preprocessing_X.transform(np.asarray(X, dtype=float))
        return np.flatnonzero(clf.predict(np.ascontiguousarray(X_), validate=False))

    def predict_logexp(self, X, zeta, noise_level=None, stream=None):
        """mean, std and LogExp acquisition in one device pass (NORA's scoring: mpi.py:182-218
        calls ``predict`` -- with its classifier and trust-region masks -- and
        gp_acquisition.py:1049-1051, 1123-1124 apply ``LogExp.f``): masked rows come out as
        ``(minus_inf_value, 0 | std, -inf)``.  Counts ``len(X)`` evaluations."""
        self.n_eval += len(X)
        noise_level = self._scalar_noise(noise_level)
        dev = self._device_state()
        self._set_masks(dev, trust=self.trust_bounds is not None)
        rows = self._host_classifier_rows(X)
        if rows is None:
            mean, std, acq = dev.predict_logexp(X, zeta, noise_level, self.y_max, stream=stream)
            if self._clf_const is False:
                mean[:], std[:], acq[:] = self.minus_inf_value, 0.0, -np.inf
            return mean, std, acq
        M = len(X)
        mean, std = np.full(M, self.minus_inf_value, dtype=float), np.zeros(M)
        acq = np.full(M, -np.inf)
        if len(rows):
            mean[rows], std[rows], acq[rows] = dev.predict_logexp(
                np.ascontiguousarray(np.asarray(X, dtype=float)[rows]), zeta, noise_level,
                self.y_max, stream=stream)
        return mean, std, acq

    def predict_logexp_topk(self, X, zeta, Kp, noise_level=None, idx_offset=0, stream=None,
                            device_out=False, want_X=True, exclude=None):
        """Fused scoring + descending-acquisition pre-ranking: only the Kp best candidates
        leave the GPU.  Returns (acq, idx, mean, std, X); masks as in ``predict_logexp``.
        ``exclude``: sorted row numbers of X that must not be returned (NORA's already
        proposed points, gp_acquisition.py:1037-1047); they are skipped on the device."""
        self.n_eval += len(X)
        noise_level = self._scalar_noise(noise_level)
        dev = self._device_state()
        self._set_masks(dev, trust=self.trust_bounds is not None)
        rows = self._host_classifier_rows(X)
        if rows is None:
            out = dev.predict_logexp_topk(X, zeta, noise_level, self.y_max, Kp,
                                          idx_offset=idx_offset, stream=stream,
                                          device_out=device_out, want_X=want_X, exclude=exclude)
            if self._clf_const is False:
                out[0][:], out[2][:], out[3][:] = -np.inf, self.minus_inf_value, 0.0
            return out
        if exclude is not None and len(exclude):
            rows = np.setdiff1d(rows, np.asarray(exclude, dtype=np.int64))
        Xf = np.ascontiguousarray(np.asarray(X, dtype=float)[rows])
        if not len(rows):
            return (np.empty(0), np.empty(0, dtype=np.int64), np.empty(0), np.empty(0),
                    np.empty((0, self.d)) if want_X else None)
        a, i, m, sd, Xs = dev.predict_logexp_topk(Xf, zeta, noise_level, self.y_max, Kp,
                                                  idx_offset=0, stream=stream, want_X=want_X)
        return a, rows[i] + idx_offset, m, sd, Xs
