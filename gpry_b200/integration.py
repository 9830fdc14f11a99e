"""
In-place binding of ``libgpry_b200.so`` into an installed GPry.

``patch_gpry()`` replaces the numeric methods of ``gpry.gpr.GaussianProcessRegressor`` --
``predict``, ``predict_std``, ``log_marginal_likelihood``, ``_kernel_inverse``,
``_update_model`` -- with the device-backed ones of ``gpry_b200.gpr`` (the two classes use the
same attribute names on purpose), turns ``L_`` / ``V_`` into lazily fetched properties, teaches it to read hyper-parameters from the scikit-learn
based kernel objects GPry builds (gpr.py:353-363), and keeps the device handle out of
pickles / deep copies.  Everything else (``append_to_data`` bookkeeping, SVM, trust region,
``Runner``, NORA's sampler, convergence criteria, I/O) stays GPry's own code.
``unpatch_gpry()`` restores the originals.
"""
import numpy as np

from . import gpr as _mirror
from .device import DeviceGP  # noqa: F401  (re-exported for the patched methods)

_ORIGINALS = {}

_PATCHED = ("predict", "predict_std", "predict_grad_batch", "log_marginal_likelihood",
            "log_marginal_likelihood_batch", "_kernel_inverse", "_update_model", "_device_state",
            "_kernel_spec", "_as_2d", "predict_logexp", "predict_logexp_topk", "_classifier_on_device",
            "_set_masks", "_materialize_factor", "_host_classifier_rows", "_scalar_noise", "L_", "V_",
            "__getstate__", "__setstate__")

# private state of the device-backed methods, with the values a fresh object starts from
_DEFAULTS = {"_dev": None, "_dev_dirty": True, "_clf_bound": None, "_clf_const": None,
             "_factor_resident": False, "_fact_sig": None}


def _ensure(self):
    """Objects created before patching, or by GPry's own ``__deepcopy__`` (which re-runs
    ``__init__`` and copies a fixed attribute list, gpr.py:1354-1433), lack this state."""
    d = self.__dict__
    for k, v in _DEFAULTS.items():
        if k not in d:
            d[k] = v
    if "device" not in d:
        d["device"] = _mirror.default_device()
    # a host copy assigned as a plain attribute before patching
    for public, private in (("L_", "_L"), ("V_", "_V")):
        if public in d:
            d[private] = d.pop(public)


def _sklearn_kernel_spec(self, kernel=None):
    """(kind, c, ell[d]) from ``ConstantKernel * RBF`` / ``ConstantKernel * Matern`` objects of
    gpry.kernels (scikit-learn subclasses); theta must be [log c, log l_1..l_d]."""
    kernel = self.kernel_ if kernel is None else kernel
    k1, k2 = kernel.k1, kernel.k2
    name = type(k2).__name__
    if type(k1).__name__ != "ConstantKernel" or name not in ("RBF", "Matern"):
        raise NotImplementedError(f"the B200 path supports ConstantKernel * RBF/Matern, got {kernel}")
    if name == "RBF":
        kind = "rbf"
    else:
        kind = {1.5: "matern15", 2.5: "matern25"}.get(k2.nu)
        if kind is None:
            raise NotImplementedError("Matern is implemented for nu = 1.5 and 2.5")
    d = self.d
    if kernel.theta.shape[0] != d + 1:
        raise NotImplementedError("the B200 path needs theta = [log c, log l_1..l_d]")
    ell = np.broadcast_to(np.asarray(k2.length_scale, dtype=float), (d,)).copy()
    return kind, float(k1.constant_value), ell


def _getstate(self):
    _ensure(self)
    _mirror.GaussianProcessRegressor._materialize_factor(self)   # L_ / V_ may be device-only
    return {k: v for k, v in self.__dict__.items() if k not in _DEFAULTS}


def _setstate(self, state):
    self.__dict__.update(state)
    self.__dict__.update(_DEFAULTS)
    self.__dict__.pop("device", None)         # device ordinals are per process
    _ensure(self)


def _device_state(self):
    _ensure(self)
    return _mirror.GaussianProcessRegressor._device_state(self)


def _kernel_inverse(self, kernel=None):
    _ensure(self)
    return _mirror.GaussianProcessRegressor._kernel_inverse(self, kernel)


def _materialize_factor(self):
    _ensure(self)
    return _mirror.GaussianProcessRegressor._materialize_factor(self)


def _lml(self, theta=None, eval_gradient=False, clone_kernel=True):
    _ensure(self)
    return _mirror.GaussianProcessRegressor.log_marginal_likelihood(
        self, theta, eval_gradient=eval_gradient, clone_kernel=clone_kernel)


def _factor_property(private):
    def getter(self):
        _materialize_factor(self)
        return self.__dict__.get(private)

    def setter(self, value):
        _ensure(self)
        if isinstance(value, np.ndarray) and value.dtype == object and value.ndim == 0:
            value = None          # GPry's __deepcopy__ does np.copy(self.V_) even before a fit
        self.__dict__[private] = value
        self.__dict__["_factor_resident"] = False
    return property(getter, setter)


def patch_gpry(gpry_module=None):
    """Patches ``gpry.gpr.GaussianProcessRegressor`` in place; returns the patched class."""
    if gpry_module is None:
        import gpry as gpry_module
    cls = gpry_module.gpr.GaussianProcessRegressor
    if _ORIGINALS:
        return cls
    M = _mirror.GaussianProcessRegressor
    replacements = {
        "predict": M.predict, "predict_std": M.predict_std,
        "log_marginal_likelihood": _lml,
        "log_marginal_likelihood_batch": M.log_marginal_likelihood_batch,
        "_kernel_inverse": _kernel_inverse, "_update_model": M._update_model,
        "_device_state": _device_state, "_kernel_spec": _sklearn_kernel_spec,
        "_as_2d": staticmethod(M._as_2d), "predict_logexp": M.predict_logexp,
        "predict_logexp_topk": M.predict_logexp_topk, "predict_grad_batch": M.predict_grad_batch,
        "_classifier_on_device": M._classifier_on_device, "_set_masks": M._set_masks,
        "_materialize_factor": _materialize_factor,
        "_host_classifier_rows": M._host_classifier_rows, "_scalar_noise": M._scalar_noise,
        # L_ / V_ stay on the device until something reads them (gpr.py:1456-1457 keeps them
        # as dense host arrays)
        "L_": _factor_property("_L"), "V_": _factor_property("_V"),
        "__getstate__": _getstate, "__setstate__": _setstate,
    }
    for name, fn in replacements.items():
        _ORIGINALS[name] = cls.__dict__.get(name, None)
        setattr(cls, name, fn)
    return cls


def unpatch_gpry(gpry_module=None):
    if gpry_module is None:
        import gpry as gpry_module
    cls = gpry_module.gpr.GaussianProcessRegressor
    for name, fn in _ORIGINALS.items():
        if fn is None:
            delattr(cls, name)
        else:
            setattr(cls, name, fn)
    _ORIGINALS.clear()
