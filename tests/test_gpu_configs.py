"""GPU parity AT THE SIZES of BASELINE.json's configs, against vectors minted by the real
reference on bench.py's own synthetic workload (oracle/gen_golden.py ``config_cases``):

  C  N_train = 2000, d = 12: mean / std / LogExp of 20000 candidates, both contractions, and
     the ranked pool of NORA (reference RankedPool, 12 points) through ``NORA.multi_add``
  D  N_train = 4000, d = 20: LML + gradient at the restart points (RBF and Matern-5/2)
  E  N_train = 2000, d = 16: mean only (surrogate-MCMC proposals)

Tolerances are the north star's (rel. 1e-10 in the scaled form of DESIGN.md)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR, scaled_err
from oracle import gp_oracle as orc

pytestmark = pytest.mark.gpu
TOL = 1e-10


def bench_gpr(N, d, kind="RBF", **kw):
    from gpry_b200.gpr import GaussianProcessRegressor
    from gpry_b200.preprocessing import Normalize_bounds, Normalize_y
    from copy import deepcopy
    X, y, theta, bounds = orc.synthetic_problem(N, d)
    gpr = GaussianProcessRegressor(kernel=kind, bounds=bounds, noise_level=1e-2,
                                   preprocessing_X=Normalize_bounds(bounds),
                                   preprocessing_y=Normalize_y(), account_for_inf=None,
                                   verbose=0, **kw)
    gpr.kernel_ = deepcopy(gpr.kernel)
    gpr.kernel_.theta = theta
    gpr.append_to_data(X, y, fit_gpr=False)
    return gpr


@pytest.mark.parametrize("contraction", ["int8", "fp64"])
def test_config_c_scores(contraction):
    z = np.load(os.path.join(GOLDEN_DIR, "config_c_n2000_d12.npz"))
    N, d, M = int(z["N"]), int(z["d"]), int(z["M"])
    gpr = bench_gpr(N, d, contraction=contraction)
    Xc = np.random.default_rng(int(z["cand_seed"])).uniform(size=(M, d))
    sy = float(z["y_std"])
    mean, std = gpr.predict(Xc, return_std=True)
    assert scaled_err(mean, z["mean"], sy) < TOL
    assert scaled_err(std ** 2, z["std"] ** 2, sy ** 2) < TOL
    m2, s2, acq = gpr.predict_logexp(Xc, float(z["zeta"]))
    ok = z["std"] ** 2 - 1e-4 > 1e-6 * sy ** 2
    assert np.array_equal(np.isfinite(acq), np.isfinite(z["acq"]))
    assert scaled_err(acq[ok], z["acq"][ok], 1.0) < 1e-9
    assert scaled_err(np.diag(gpr.L_), z["L_diag"], 1.0) < TOL
    assert scaled_err(gpr.alpha_[:16], z["alpha_head"], np.abs(z["alpha_head"]).max()) < 1e-9


def test_config_c_ranked_pool():
    from gpry_b200.acquisition_functions import LogExp
    from gpry_b200.gp_acquisition import NORA
    z = np.load(os.path.join(GOLDEN_DIR, "config_c_n2000_d12.npz"))
    N, d, M = int(z["N"]), int(z["d"]), int(z["M"])
    gpr = bench_gpr(N, d)
    Xc = np.random.default_rng(int(z["cand_seed"])).uniform(size=(M, d))
    n_points = int(z["pool_n_points"])
    nora = NORA(gpr.bounds, acq_func=LogExp(zeta=float(z["zeta"])), kprime=1024)
    X_pool, y_pool, acq_pool = nora.multi_add(gpr, n_points=n_points, X_shard=Xc)
    assert np.array_equal(nora.last_pool_idx, z["pool_idx"])
    assert np.array_equal(X_pool, Xc[z["pool_idx"]])
    assert scaled_err(y_pool, z["pool_y"], float(z["y_std"])) < TOL
    assert scaled_err(nora.pool.acq_cond[:n_points], z["pool_acq_cond"], 1.0) < 1e-7


def test_config_d_lml():
    from gpry_b200.device import workspace
    z = np.load(os.path.join(GOLDEN_DIR, "config_d_n4000_d20.npz"))
    N, d = int(z["N"]), int(z["d"])
    X, y, _, _ = orc.synthetic_problem(N, d, seed=int(z["seed"]))
    y_mean, y_std = orc.normalize_y_fit(y)
    y_ = (y - y_mean) / y_std
    noise2 = np.full(N, (float(z["noise_level"]) / y_std) ** 2)
    for kind in ("rbf", "matern25"):
        lml, grad, info = workspace(0).lml_batched(kind, X, noise2, y_, z[f"thetas_{kind}"])
        assert np.all(info == 0)
        for k in range(len(lml)):
            assert abs(lml[k] - z[f"lml_{kind}"][k]) <= TOL * abs(z[f"lml_{kind}"][k]), (kind, k)
            g_ref = z[f"grad_{kind}"][k]
            assert scaled_err(grad[k], g_ref, np.abs(g_ref).max()) < TOL, (kind, k)


def test_config_e_mean_only():
    z = np.load(os.path.join(GOLDEN_DIR, "config_e_n2000_d16.npz"))
    N, d, M = int(z["N"]), int(z["d"]), int(z["M"])
    gpr = bench_gpr(N, d)
    Xc = np.random.default_rng(int(z["cand_seed"])).uniform(size=(M, d))
    mean = gpr.predict(Xc)
    assert scaled_err(mean, z["mean"], float(z["y_std"])) < TOL
