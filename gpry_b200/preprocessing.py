"""
The two pre-processors ``gpry.run.Runner`` uses (run.py:318-319), mirrored from
``gpry/preprocessing.py``: ``Normalize_bounds`` (:311-411) and ``Normalize_y`` (:528-630), plus
``DummyPreprocessor`` (:29).  They are O(N) host scalars; on the device they are fused into
the kernels as per-dimension (min, width) and (mean_, std_).
"""
import numpy as np


class DummyPreprocessor:
    """Identity (reference: used when ``preprocessing_X/y=None``, gpr.py:275-278)."""
    is_linear = True
    fitted = True

    @staticmethod
    def fit(X, y):
        pass

    @staticmethod
    def transform_bounds(bounds):
        return bounds

    @staticmethod
    def transform(X):
        return X

    @staticmethod
    def inverse_transform(X):
        return X

    @staticmethod
    def transform_scale(s):
        return s

    @staticmethod
    def inverse_transform_scale(s):
        return s


class Normalize_bounds:
    """X -> (X - min) / (max - min)   (preprocessing.py:349-411)."""

    def __init__(self, bounds):
        self.update_bounds(bounds)
        self.fitted = True

    def update_bounds(self, bounds):
        bounds = np.asarray(bounds, dtype=float)
        self.bounds = bounds
        self.bounds_min = bounds[:, 0]
        self.bounds_max = bounds[:, 1]
        if np.any(self.bounds_min > self.bounds_max):
            raise ValueError("The bounds must be in dimension-wise order min->max")

    def transform_bounds(self, bounds):
        out = np.ones_like(bounds, dtype=float)
        out[:, 0] = 0
        return out

    def fit(self, X, y):
        pass

    def transform(self, X):
        return (X - self.bounds_min) / (self.bounds_max - self.bounds_min)

    def inverse_transform(self, X):
        return (X * (self.bounds_max - self.bounds_min)) + self.bounds_min

    def inverse_transform_scale(self, X):
        return X * (self.bounds_max - self.bounds_min)


class Normalize_y:
    """y -> (y - mean_) / std_ with population std of the finite y (preprocessing.py:528-630)."""
    is_linear = True

    def __init__(self):
        self.mean_ = None
        self.std_ = None

    @property
    def fitted(self):
        return self.mean_ is not None and self.std_ is not None

    def fit(self, X, y):
        y = np.asarray(y)
        y = y[np.isfinite(y)]
        self.mean_, self.std_ = np.mean(y), np.std(y)

    def _check(self):
        if not self.fitted:
            raise TypeError("mean_ and std_ have not been fit before")

    def transform(self, y):
        self._check()
        return (y - self.mean_) / self.std_

    def inverse_transform(self, y):
        self._check()
        return (y * self.std_) + self.mean_

    def transform_scale(self, scale):
        self._check()
        return scale / self.std_

    def inverse_transform_scale(self, scale):
        self._check()
        return scale * self.std_
