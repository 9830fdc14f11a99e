"""
Kernel descriptors mirroring ``gpry.kernels`` (reference kernels.py:213-432, 601-609, 681-699)
for the kernels ``gpry.gpr.GaussianProcessRegressor`` auto-constructs (gpr.py:344-363):

    ConstantKernel(c, c_bounds) * RBF(length_scale[d], bounds)
    ConstantKernel(c, c_bounds) * Matern(length_scale[d], bounds, nu in {1.5, 2.5})

In the reference these classes inherit ``__call__`` / ``diag`` / theta-gradients from
scikit-learn and evaluate them with numpy.  Here they are *descriptors*: they own the
hyper-parameters (``theta`` = log values, ``bounds`` = log bounds, exactly the layout of the
reference: [log c, log l_1 .. log l_d]) and the arithmetic runs in the CUDA library (fused
into predict / log-marginal-likelihood; ``__call__`` and ``gradient_x`` are provided through
dedicated device kernels for API completeness).  There is no CPU evaluation path.
"""
from collections import namedtuple

import numpy as np


class Hyperparameter(namedtuple("Hyperparameter",
                                ("name", "value_type", "bounds", "max_length", "n_elements",
                                 "fixed", "dynamic"))):
    """Same fields as the reference's namedtuple (kernels.py:26-115)."""
    __slots__ = ()

    def __new__(cls, name, value_type, bounds, max_length=None, n_elements=1, fixed=None,
                dynamic=None):
        if not isinstance(bounds, str) or bounds not in ("fixed", "dynamic"):
            bounds = np.atleast_2d(bounds)
            if n_elements > 1:
                if bounds.shape[0] == 1:
                    bounds = np.repeat(bounds, n_elements, 0)
                elif bounds.shape[0] != n_elements:
                    raise ValueError(f"Bounds on {name} should have either 1 or {n_elements} "
                                     f"dimensions. Given are {bounds.shape[0]}")
        if fixed is None:
            fixed = isinstance(bounds, str) and bounds == "fixed"
        if dynamic is None:
            dynamic = isinstance(bounds, str) and bounds == "dynamic"
        return super().__new__(cls, name, value_type, bounds, max_length, n_elements, fixed,
                               dynamic)


class Kernel:
    """Base: operator overloading and the theta / bounds protocol (kernels.py:117-191)."""

    def __mul__(self, b):
        return Product(self, b if isinstance(b, Kernel) else ConstantKernel(b))

    def __rmul__(self, b):
        return Product(b if isinstance(b, Kernel) else ConstantKernel(b), self)

    @property
    def n_dims(self):
        return self.theta.shape[0]

    @property
    def requires_vector_input(self):
        return True

    def clone_with_theta(self, theta):
        import copy
        new = copy.deepcopy(self)
        new.theta = theta
        return new

    def __eq__(self, other):
        return type(self) is type(other) and repr(self) == repr(other) and \
            np.array_equal(self.theta, other.theta)

    __hash__ = None


class ConstantKernel(Kernel):
    """k(x, x') = constant_value (sklearn:kernels.py:1244-1322)."""
    kind = "constant"

    def __init__(self, constant_value=1.0, constant_value_bounds=(1e-5, 1e5)):
        self.constant_value = float(constant_value)
        self.constant_value_bounds = constant_value_bounds

    @property
    def hyperparameters(self):
        return [Hyperparameter("constant_value", "numeric", self.constant_value_bounds, None)]

    @property
    def fixed(self):
        return isinstance(self.constant_value_bounds, str) and \
            self.constant_value_bounds == "fixed"

    @property
    def theta(self):
        return np.array([]) if self.fixed else np.log([self.constant_value])

    @theta.setter
    def theta(self, theta):
        if not self.fixed:
            self.constant_value = float(np.exp(np.asarray(theta, dtype=float)[0]))

    @property
    def bounds(self):
        if self.fixed:
            return np.empty((0, 2))
        return np.log(np.atleast_2d(self.constant_value_bounds).astype(float))

    def __repr__(self):
        return "{0:.3g}**2".format(np.sqrt(self.constant_value))


class _LengthScaleKernel(Kernel):
    def __init__(self, length_scale=1.0, length_scale_bounds=(1e-5, 1e5), prior_bounds=None):
        self.length_scale = np.array(length_scale, dtype=float) if np.iterable(length_scale) \
            else float(length_scale)
        self.length_scale_bounds = length_scale_bounds
        self.prior_bounds = prior_bounds
        if isinstance(length_scale_bounds, str) and length_scale_bounds == "dynamic":
            if prior_bounds is None:
                raise TypeError("Prior bounds are required if the hyperparameter bounds are "
                                "set to 'dynamic'.")
            pb = np.asarray(prior_bounds)
            self.max_length = pb[:, 1] - pb[:, 0]      # kernels.py:240-242
        else:
            self.max_length = None

    @property
    def anisotropic(self):
        return np.iterable(self.length_scale) and len(self.length_scale) > 1

    @property
    def fixed(self):
        return isinstance(self.length_scale_bounds, str) and self.length_scale_bounds == "fixed"

    @property
    def hyperparameters(self):
        n = len(self.length_scale) if self.anisotropic else 1
        return [Hyperparameter("length_scale", "numeric", self.length_scale_bounds,
                               self.max_length, n)]

    @property
    def theta(self):
        return np.array([]) if self.fixed else np.log(np.atleast_1d(self.length_scale))

    @theta.setter
    def theta(self, theta):
        if self.fixed:
            return
        val = np.exp(np.asarray(theta, dtype=float))
        self.length_scale = val.copy() if self.anisotropic else float(val[0])

    @property
    def bounds(self):
        """Log bounds; 'dynamic' handling as kernels.py:157-191."""
        if self.fixed:
            return np.empty((0, 2))
        n = len(self.length_scale) if self.anisotropic else 1
        if isinstance(self.length_scale_bounds, str) and self.length_scale_bounds == "dynamic":
            ls = np.atleast_1d(self.length_scale)
            out = []
            for t in range(n):
                ml = None if self.max_length is None else self.max_length[t if self.anisotropic
                                                                          else 0]
                out.append([ls[t] * 1e-3, ls[t] * 100.] if ml is None
                           else [ml * 1e-3, ml * 100.])
            return np.log(np.array(out, dtype=float))
        b = np.atleast_2d(self.length_scale_bounds).astype(float)
        if b.shape[0] == 1 and n > 1:
            b = np.repeat(b, n, 0)
        return np.log(b)


class RBF(_LengthScaleKernel):
    """exp(-1/2 |x/l - x'/l|^2)  (sklearn:kernels.py:1530-1584; gradient_x kernels.py:257-278)."""
    kind = "rbf"

    def __repr__(self):
        ls = np.atleast_1d(self.length_scale)
        return "RBF(length_scale=[{0}])".format(", ".join(map("{0:.3g}".format, ls)))


class Matern(_LengthScaleKernel):
    """Matern nu in {1.5, 2.5} (sklearn:kernels.py:1685-1771; gradient_x kernels.py:326-432)."""

    def __init__(self, length_scale=1.0, length_scale_bounds=(1e-5, 1e5), nu=1.5,
                 prior_bounds=None):
        super().__init__(length_scale, length_scale_bounds, prior_bounds)
        if nu not in (1.5, 2.5):
            raise ValueError("the B200 path implements Matern for nu = 1.5 and 2.5 only "
                             f"(got nu={nu})")
        self.nu = nu

    @property
    def kind(self):
        return "matern15" if self.nu == 1.5 else "matern25"

    def __repr__(self):
        ls = np.atleast_1d(self.length_scale)
        return "Matern(length_scale=[{0}], nu={1:.3g})".format(
            ", ".join(map("{0:.3g}".format, ls)), self.nu)


class Product(Kernel):
    """k1 * k2 with theta = [k1.theta, k2.theta] (sklearn:kernels.py:739-753, 936-990)."""

    def __init__(self, k1, k2):
        self.k1, self.k2 = k1, k2

    @property
    def hyperparameters(self):
        r = [Hyperparameter("k1__" + h.name, *h[1:]) for h in self.k1.hyperparameters]
        r += [Hyperparameter("k2__" + h.name, *h[1:]) for h in self.k2.hyperparameters]
        return r

    @property
    def theta(self):
        return np.append(self.k1.theta, self.k2.theta)

    @theta.setter
    def theta(self, theta):
        theta = np.asarray(theta, dtype=float)
        n1 = self.k1.n_dims
        self.k1.theta = theta[:n1]
        self.k2.theta = theta[n1:]

    @property
    def bounds(self):
        b1, b2 = self.k1.bounds, self.k2.bounds
        if b1.size == 0:
            return b2
        if b2.size == 0:
            return b1
        return np.vstack((b1, b2))

    def __repr__(self):
        return "{0} * {1}".format(self.k1, self.k2)

    # ---- what the device library needs --------------------------------------------------
    def device_spec(self, d):
        """(kind, c, ell[d]) if this is ConstantKernel * {RBF, Matern}; else raises."""
        k1, k2 = self.k1, self.k2
        if isinstance(k2, ConstantKernel):
            k1, k2 = k2, k1
        if not (isinstance(k1, ConstantKernel) and isinstance(k2, _LengthScaleKernel)):
            raise NotImplementedError(
                "the B200 path supports ConstantKernel * RBF/Matern only (the kernels GPry "
                f"constructs itself, gpr.py:344-363); got {self!r}")
        ell = np.broadcast_to(np.asarray(k2.length_scale, dtype=float), (d,)).copy()
        return k2.kind, float(k1.constant_value), ell

    def theta_is_standard(self, d):
        """True if theta is exactly [log c, log l_1..l_d] (nothing fixed, anisotropic)."""
        return self.theta.shape[0] == d + 1 and isinstance(self.k1, ConstantKernel)

    def __call__(self, X, Y=None, eval_gradient=False):
        """k(X, Y) evaluated on the GPU (Product.__call__, sklearn:kernels.py:971).  The
        (N, N, 1+d) theta-gradient tensor is never built by this package: use
        ``GaussianProcessRegressor.log_marginal_likelihood(theta, eval_gradient=True)``."""
        if eval_gradient:
            raise NotImplementedError(
                "dK/dtheta is fused into the log-marginal-likelihood kernel on the device and "
                "never materialised")
        from .device import workspace
        from .gpr import default_device
        X = np.atleast_2d(np.asarray(X, dtype=float))
        d = X.shape[1]
        kind, c, ell = self.device_spec(d)
        theta = np.log(np.concatenate([[c], ell]))
        K = workspace(default_device()).kernel_cross(kind, theta, X, X if Y is None else Y)
        if Y is None:
            np.fill_diagonal(K, c)   # pdist/squareform path: unit diagonal times c (:1566)
        return K

    def diag(self, X):
        """k(x, x) = c (sklearn:kernels.py:973-990, 1298-1322)."""
        kind, c, _ = self.device_spec(np.atleast_2d(X).shape[1])
        return np.full(np.atleast_2d(X).shape[0], c)


C = ConstantKernel
