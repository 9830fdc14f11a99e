import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def golden_names():
    return sorted(os.path.splitext(os.path.basename(p))[0]
                  for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz"))
                  if not os.path.basename(p).startswith(("lml_nonpd", "fit_", "loop_", "config_")))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    g = {k: z[k] for k in z.files}
    for k in ("kind",):
        g[k] = str(g[k])
    for k in ("N", "d", "M", "seed"):
        g[k] = int(g[k])
    g["normalize"] = bool(g["normalize"])
    g["noise_level"] = float(g["noise_level"])
    g["zeta"] = float(g["zeta"])
    g["clip_factor"] = float(g["clip_factor"]) if "clip_factor" in g else 1.1
    if "X_train" not in g:  # regenerate from the seed exactly as oracle/gen_golden.py does
        rng = np.random.default_rng(g["seed"])
        lo, hi = g["bounds"][:, 0], g["bounds"][:, 1]
        U = rng.uniform(size=(g["N"], g["d"]))
        g["X_train"] = lo + U * (hi - lo)
        g["y_train"] = -0.5 * np.sum(((U - 0.5) / 0.15) ** 2, axis=1)
    return g


def golden_pool_candidates(g):
    lo, hi = g["bounds"][:, 0], g["bounds"][:, 1]
    return lo + np.random.default_rng(int(g["pool_seed"])).uniform(
        size=(int(g["pool_M"]), g["d"])) * (hi - lo)


def oracle_state(g):
    from oracle import gp_oracle as orc
    return orc.GPState(g["kind"], g["theta"], g["X_train"], g["y_train"],
                       bounds=g["bounds"] if g["normalize"] else None,
                       normalize_y=g["normalize"], noise_level=g["noise_level"],
                       clip_factor=g["clip_factor"])


def scaled_err(a, b, scale):
    """max |a-b| / max(|b|, scale) elementwise -> scalar (the tolerance form of DESIGN.md)."""
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    both_inf = np.isinf(a) & np.isinf(b) & (np.sign(a) == np.sign(b))
    den = np.maximum(np.abs(b), scale)
    err = np.where(both_inf, 0.0, np.abs(a - b) / den)
    return float(np.max(err)) if err.size else 0.0


@pytest.fixture(params=golden_names())
def golden(request):
    return load_golden(request.param)
