"""
Infinities classifier with the interface of ``gpry.svm.SVM`` whose PREDICTION runs on the GPU
(reference svm.py:20-346).

The reference classifies every candidate pool on the host (``SVC.predict`` inside
``GaussianProcessRegressor.predict`` / ``predict_std``, gpr.py:1136-1174, 1300-1318): a
single-threaded libsvm pass of cost O(M n_sv d) that, once the GP arithmetic is on the GPU,
is the slowest stage of scoring a large pool.  The decision function of the two-class RBF SVC,

    f(x_) = sum_i dual_coef_i exp(-gamma |x_ - sv_i|^2) + intercept,     finite  <=>  f > 0,

is a mean-only GP "prediction" with the support vectors as training set, so it runs through
the same device kernel (``kstar_build``): either fused into the regressor's calls
(``gpry_set_classifier``: the mask is applied on the device and never crosses PCIe) or on
its own (``SVM.predict``, a mean-only device state).  Training stays what it is in the
reference: scikit-learn's SVC (libsvm) on the <= few thousand training points, on the host.
"""
import warnings

import numpy as np
from sklearn.svm import SVC

from .device import DeviceGP


class SVM:
    """Two-class (finite / infinite) RBF support-vector classifier (svm.py:20-346): same
    constructor defaults (C = 1e7, gamma = "scale"), ``fit(X, y, diff_threshold)``,
    ``predict(X, validate)``, ``is_finite``, ``_is_finite_raw``, ``abs_threshold``, ``d``,
    ``n`` and the attributes ``X_train, y_train, y_finite, all_finite,
    at_least_one_finite``.  Operates in the transformed space, like the reference's."""

    def __init__(self, C=1e7, gamma="scale", tol=0.001, cache_size=200, max_iter=-1,
                 random_state=None, device=None):
        self.C, self.gamma, self.tol = C, gamma, tol
        self.cache_size, self.max_iter = cache_size, max_iter
        if isinstance(random_state, np.random.Generator):   # SVC wants an int / RandomState
            random_state = int(random_state.integers(2 ** 31))
        self.random_state = random_state
        self.device = device
        self.X_train = None
        self.y_train = None
        self.y_finite = None
        self.at_least_one_finite = False
        self.all_finite = False
        self.diff_threshold = None
        self._max_y = None
        self._spec = None
        self._dev = None
        self._dev_dirty = True
        self.version = 0          # bumped by every fit: tells the regressor to re-bind it

    # ------------------------------------------------------------------ bookkeeping
    @property
    def d(self):
        if self.X_train is None:
            raise ValueError("You need to add some data before determining its dimension.")
        return self.X_train.shape[1]

    @property
    def n(self):
        return 0 if self.y_train is None else len(self.y_train)

    @property
    def abs_threshold(self):
        return self._max_y - self.diff_threshold

    @staticmethod
    def _is_finite_raw(y, diff_threshold, max_y=None):
        """svm.py:273-295: threshold check (not a prediction); NaN and +-inf are not finite."""
        if max_y is None:
            max_y = np.max(y)
        return np.greater_equal(y, max_y - diff_threshold) & np.isfinite(y)

    def is_finite(self, y):
        if self.y_train is None:
            raise ValueError("Cannot do anything: the SVM has not been trained yet!")
        return self._is_finite_raw(y, self.diff_threshold, self._max_y)

    # ------------------------------------------------------------------ training (host)
    def fit(self, X, y, diff_threshold):
        """svm.py:227-271 -> the boolean classification of the training points."""
        self.X_train = np.copy(X)
        self.y_train = np.copy(y)
        self.version += 1
        self._spec = None
        self._dev_dirty = True
        if np.all(self.y_train == -np.inf):
            self.at_least_one_finite = False
            self.y_finite = np.full(len(X), False)
            return self.y_finite
        self.at_least_one_finite = True
        self.diff_threshold = diff_threshold
        self._max_y = max(self.y_train)
        self.y_finite = self._is_finite_raw(self.y_train, self.diff_threshold, max_y=self._max_y)
        if np.all(self.y_finite):
            self.all_finite = True
            return self.y_finite
        self.all_finite = False
        svc = SVC(C=self.C, kernel="rbf", gamma=self.gamma, tol=self.tol,
                  cache_size=self.cache_size, max_iter=self.max_iter,
                  random_state=self.random_state)
        svc.fit(self.X_train, self.y_finite)
        # binary SVC: decision_function = dual_coef_ . k + intercept_ > 0 <=> classes_[1] = True
        assert list(svc.classes_) == [False, True]
        self._spec = (np.ascontiguousarray(svc.support_vectors_, dtype=float),
                      np.ascontiguousarray(svc.dual_coef_[0], dtype=float),
                      float(svc.intercept_[0]), float(svc._gamma))
        return self.y_finite

    # ------------------------------------------------------------------ prediction (device)
    def device_spec(self):
        """What ``gpry_set_classifier`` needs, or the constant answer when no SVC was needed:
        ``("all", True | False)`` or ``("svc", sv, dual_coef, intercept, gamma)``."""
        if self.y_train is None:
            raise ValueError("The SVM has not been trained yet.")
        if self.all_finite:
            return ("all", True)
        if not self.at_least_one_finite:
            return ("all", False)
        return ("svc",) + self._spec

    def decision_function(self, X):
        """f(x_) for transformed points (host array or CUDA tensor), evaluated on the GPU."""
        if self._spec is None:
            raise ValueError("no SVC has been fit (all points finite, all infinite, or no data)")
        if self._dev is None:
            from .gpr import default_device
            self._dev = DeviceGP(default_device() if self.device is None else self.device)
            self._dev_dirty = True
        if self._dev_dirty:
            sv, coef, intercept, gamma = self._spec
            self._dev.upload("rbf", sv, coef, None, 1.0, 1.0 / np.sqrt(2.0 * gamma),
                             y_mean=intercept, y_std=1.0)
            self._dev_dirty = False
        return self._dev.predict(X, return_mean=True, return_std=False)[0]

    def predict(self, X, validate=True):
        """svm.py:308-346: True where a finite posterior is predicted."""
        if self.y_train is None:
            raise ValueError("The SVM has not been trained yet.")
        if validate:
            X = np.atleast_2d(X)
        if self.all_finite:
            return np.full(len(X), True)
        if not self.at_least_one_finite:
            warnings.warn("Only -inf points added to the classifier so far. "
                          "Returning False unconditionally.")
            return np.full(len(X), False)
        return self.decision_function(X) > 0

    # ------------------------------------------------------------------ copies / pickles
    def __getstate__(self):
        state = dict(self.__dict__)
        state["_dev"] = None
        state["_dev_dirty"] = True
        return state

    def __setstate__(self, state):
        self.__dict__.update(state)
        import os
        if "LOCAL_RANK" in os.environ:
            self.device = None       # device ordinals are per process: use this rank's GPU

    def __deepcopy__(self, memo):
        from copy import deepcopy
        new = self.__class__.__new__(self.__class__)
        for k, v in self.__getstate__().items():
            setattr(new, k, deepcopy(v, memo))
        return new
