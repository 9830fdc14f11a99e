"""
ctypes binding of ``libgpry_b200.so`` (the C ABI declared in ``include/gpry_b200.h``).

There is deliberately NO fallback: if the shared library has not been built
(``python -c "import __graft_entry__ as g; g.build()"`` or ``make -C gpry_b200/csrc``) or no
B200 is visible, the first call raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# (GPRY_B200_LIB: another build of the same library, for A/B measurements)
LIB_PATH = os.environ.get("GPRY_B200_LIB") or os.path.join(_HERE, "libgpry_b200.so")

KERNEL_KINDS = {"rbf": 0, "matern15": 1, "matern25": 2}
X_ON_DEVICE, OUT_ON_DEVICE = 1, 2
WANT_MEAN, WANT_STD = 1, 2
MAX_TOPK = 2048

_c_double_p = C.POINTER(C.c_double)
_c_int64_p = C.POINTER(C.c_int64)
_c_int_p = C.POINTER(C.c_int)

# name -> (restype, argtypes); every symbol include/gpry_b200.h declares
SIGNATURES = {
    "gpry_abi_version": (C.c_int, []),
    "gpry_last_error": (C.c_char_p, []),
    "gpry_device_count": (C.c_int, []),
    "gpry_state_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "gpry_state_destroy": (C.c_int, [C.c_void_p]),
    "gpry_state_upload": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_double, C.c_double, C.c_double]),
    "gpry_state_adopt_factorization": (C.c_int, [C.c_void_p, C.c_double, C.c_void_p,
                                                 C.c_void_p, C.c_void_p, C.c_double,
                                                 C.c_double, C.c_double]),
    "gpry_set_trust_region": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_double]),
    "gpry_predict_grad": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p]),
    "gpry_set_contract_mode": (C.c_int, [C.c_void_p, C.c_int]),
    "gpry_set_contract_guard": (C.c_int, [C.c_void_p, C.c_int]),
    "gpry_contract_info": (C.c_int, [C.c_void_p, C.c_void_p]),
    "gpry_int8_peak": (C.c_int, [C.c_void_p, C.c_void_p]),
    "gpry_int8_peak_sustained": (C.c_int, [C.c_void_p, C.c_double, C.c_void_p]),
    "gpry_set_mask_value": (C.c_int, [C.c_void_p, C.c_double]),
    "gpry_set_classifier": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                      C.c_double, C.c_double]),
    "gpry_classify": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p,
                                C.c_void_p]),
    "gpry_factor_append": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_void_p]),
    "gpry_factor_download": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "gpry_state_info": (C.c_int, [C.c_void_p, _c_int_p, _c_int_p, _c_int_p]),
    "gpry_predict": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int,
                               C.c_void_p, C.c_void_p, C.c_void_p]),
    "gpry_predict_logexp": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_double,
                                      C.c_double, C.c_double, C.c_int, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p]),
    "gpry_predict_logexp_topk": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_double,
                                           C.c_double, C.c_double, C.c_int, C.c_int64, C.c_int,
                                           C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                           C.c_void_p, _c_int64_p, C.c_void_p]),
    "gpry_set_excluded": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "gpry_topk": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p,
                            C.c_void_p, _c_int64_p, C.c_void_p]),
    "gpry_mean_grad": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "gpry_std_grad": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "gpry_posterior_cov": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                     C.c_void_p]),
    "gpry_kernel_cross": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                    C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
    "gpry_kernel_gradient_x": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "gpry_factorize": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                 _c_double_p, _c_int_p, C.c_int]),
    "gpry_lml_batched": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                   C.c_void_p, C.c_void_p]),
    "gpry_comm_unique_id": (C.c_int, [C.c_void_p]),
    "gpry_comm_init": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "gpry_comm_share": (C.c_int, [C.c_void_p, C.c_void_p]),
    "gpry_comm_destroy": (C.c_int, [C.c_void_p]),
    "gpry_comm_info": (C.c_int, [C.c_void_p, _c_int_p, _c_int_p, _c_int_p]),
    "gpry_bcast_state": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "gpry_allgather_topk": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                      _c_int64_p, _c_double_p, C.c_void_p]),
    "gpry_set_profiling": (C.c_int, [C.c_void_p, C.c_int]),
    "gpry_get_timings": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
}

_lib = None


class GpryB200Error(RuntimeError):
    """An error reported by libgpry_b200.so."""


def load_library():
    """Loads (once) and returns the ctypes handle, with all prototypes installed."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GpryB200Error(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` (needs nvcc).  gpry_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.gpry_abi_version() != 1:
        raise GpryB200Error("libgpry_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(code):
    if code != 0:
        msg = load_library().gpry_last_error()
        raise GpryB200Error(f"libgpry_b200 error {code}: {msg.decode() if msg else '?'}")


def as_f64(a, shape=None):
    """C-contiguous float64 view/copy of ``a`` (host)."""
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None and a.shape != shape:
        raise ValueError(f"expected shape {shape}, got {a.shape}")
    return a


def ptr(a):
    """void* of a numpy array, a torch tensor (host or device), an int address, or None."""
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    if isinstance(a, np.ndarray):
        return C.c_void_p(a.ctypes.data)
    if hasattr(a, "data_ptr"):  # torch.Tensor
        return C.c_void_p(a.data_ptr())
    raise TypeError(f"cannot take a pointer of {type(a)}")
