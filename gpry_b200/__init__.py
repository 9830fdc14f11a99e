"""
gpry_b200 -- B200-native (sm_100a) GP surrogate hot path behind GPry's Python API.

Only the hot path of GPry lives here: batched posterior mean / std,
LogExp acquisition, ranked-pool pre-selection, log-marginal-likelihood + gradient, all
computed by hand-written CUDA kernels in ``libgpry_b200.so`` (C ABI: include/gpry_b200.h).
There is no CPU fallback.
"""
__version__ = "0.1.0"

from ._lib import GpryB200Error, load_library, LIB_PATH  # noqa: F401
from .device import DeviceGP  # noqa: F401
