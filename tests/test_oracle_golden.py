"""The CPU oracle (oracle/gp_oracle.py) pinned against golden vectors produced by the REAL
reference (oracle/gen_golden.py).  CPU-only; runs in the `-m "not gpu"` suite."""
import os

import numpy as np
import pytest

from conftest import (GOLDEN_DIR, golden_pool_candidates, oracle_state, scaled_err)
from oracle import gp_oracle as orc

# The oracle repeats the reference's operations in the same order with the same libraries,
# so agreement is at round-off level; the bound is far below the 1e-10 product tolerance.
TOL = 1e-12


def test_predict_mean_std(golden):
    g = golden
    st = oracle_state(g)
    mean, std = orc.predict(st, g["Xc"], return_std=True)
    sy = float(g["y_std"])
    assert scaled_err(mean, g["mean"], sy) < TOL
    assert scaled_err(std, g["std"], 1e-3 * sy) < 1e-9   # sqrt near the clamp amplifies
    assert scaled_err(std ** 2, g["std"] ** 2, sy ** 2) < TOL
    assert scaled_err(orc.predict_std(st, g["Xc"]), g["std_only"], 1e-3 * sy) < 1e-9
    assert scaled_err(orc.predict(st, g["Xc"]), g["mean_only"], sy) < TOL


def test_state(golden):
    g = golden
    st = oracle_state(g)
    assert abs(st.y_mean - float(g["y_mean"])) <= 1e-15 * max(1, abs(float(g["y_mean"])))
    assert abs(st.y_std - float(g["y_std"])) <= 1e-15 * abs(float(g["y_std"]))
    assert scaled_err(st.alpha_, g["alpha_"], np.abs(g["alpha_"]).max()) < 1e-11
    N = g["N"]
    assert scaled_err(st.V_[[0, N // 2, N - 1]], g["V_rows"], np.abs(g["V_rows"]).max()) < 1e-11
    assert scaled_err(np.diag(st.L_), g["L_diag"], 1.0) < 1e-12


def test_logexp(golden):
    g = golden
    st = oracle_state(g)
    mean, std, acq = orc.predict_logexp(st, g["Xc"], zeta=g["zeta"])
    fin = np.isfinite(g["acq_f"])
    assert np.array_equal(np.isfinite(acq), fin)
    # acq = 2 zeta (mu - ymax) + 0.5 log(var - noise^2): compare on the var scale
    assert scaled_err(acq[fin], g["acq_f"][fin], 1.0) < 1e-8
    call = orc.logexp_call(mean, std, st.y_max, st.noise_level, g["zeta"])
    assert np.array_equal(np.isfinite(call), np.isfinite(g["acq_call"]))
    f2 = np.isfinite(call)
    assert scaled_err(call[f2], g["acq_call"][f2], 1.0) < 1e-8


def test_gradients(golden):
    g = golden
    st = oracle_state(g)
    m, s, gm, gs = orc.predict(st, g["Xc"][:1], return_std=True, return_mean_grad=True,
                               return_std_grad=True)
    assert scaled_err(gm, g["grad_mean"], np.abs(g["grad_mean"]).max()) < 1e-11
    assert scaled_err(gs, g["grad_std"], np.abs(g["grad_std"]).max()) < 1e-8


def test_lml(golden):
    g = golden
    if "lml" not in g:
        pytest.skip("no LML in this fixture")
    st = oracle_state(g)
    for th, v, gr in zip(g["lml_thetas"], g["lml"], g["lml_grad"]):
        lml, grad = orc.log_marginal_likelihood(g["kind"], th, st.X_train_, st.y_train_,
                                                st.noise2, eval_gradient=True)
        assert abs(lml - v) <= 1e-12 * abs(v)
        assert scaled_err(grad, gr, np.abs(gr).max()) < 1e-11
        assert orc.log_marginal_likelihood(g["kind"], th, st.X_train_, st.y_train_,
                                           st.noise2) == pytest.approx(v, rel=1e-12)


def test_lml_nonpd():
    z = np.load(os.path.join(GOLDEN_DIR, "lml_nonpd.npz"))
    lml, grad = orc.log_marginal_likelihood("rbf", z["theta"], z["X_train_"], z["y_train_"],
                                            z["noise2"], eval_gradient=True)
    assert lml == -np.inf and float(z["lml"]) == -np.inf
    assert np.array_equal(grad, np.zeros(3)) and np.array_equal(z["grad"], np.zeros(3))


@pytest.mark.parametrize("method", ["single sort acq", "bulk"])
def test_ranked_pool(golden, method):
    g = golden
    if "pool_M" not in g:
        pytest.skip("no ranked pool in this fixture")
    st = oracle_state(g)
    Xp = golden_pool_candidates(g)
    y, sigma, acq = orc.predict_logexp(st, Xp, zeta=g["zeta"])
    assert scaled_err(np.sort(acq)[::-1][:64], g["pool_acq_top"], 1.0) < 1e-8
    idx, Xs, ys, acqs = orc.ranked_pool_select(st, Xp, y, sigma, acq,
                                               int(g["pool_n_points"]), zeta=g["zeta"],
                                               method=method)
    tag = method.replace(" ", "_")
    assert np.array_equal(idx, g[f"pool_idx_{tag}"])
    assert scaled_err(ys, g[f"pool_y_{tag}"], float(g["y_std"])) < 1e-12


# ---------------------------------------------------------------------------------------
# BASELINE.json config sizes (reference-minted on bench.py's synthetic workload;
# oracle/gen_golden.py config_cases): the oracle is pinned where the benchmarks run
# ---------------------------------------------------------------------------------------
def test_config_c_scores_and_pool():
    z = np.load(os.path.join(GOLDEN_DIR, "config_c_n2000_d12.npz"))
    N, d, M = int(z["N"]), int(z["d"]), int(z["M"])
    X, y, theta, bounds = orc.synthetic_problem(N, d, seed=int(z["seed"]))
    assert np.array_equal(theta, z["theta"])
    st = orc.GPState("rbf", theta, X, y, bounds=bounds)
    Xc = np.random.default_rng(int(z["cand_seed"])).uniform(size=(M, d))
    sy = float(z["y_std"])
    sub = slice(0, 6000)
    mean, std, acq = orc.predict_logexp(st, Xc[sub], zeta=float(z["zeta"]))
    assert scaled_err(mean, z["mean"][sub], sy) < TOL
    assert scaled_err(std ** 2, z["std"][sub] ** 2, sy ** 2) < TOL
    assert scaled_err(acq, z["acq"][sub], 1.0) < 1e-9
    assert scaled_err(np.diag(st.L_), z["L_diag"], 1.0) < 1e-12
    # the ranked pool from the reference's own scores (the oracle's refits at N = 2000)
    idx, Xs, ys, acqs = orc.ranked_pool_select(st, Xc, z["mean"], z["std"], z["acq"],
                                               int(z["pool_n_points"]), zeta=float(z["zeta"]))
    assert np.array_equal(idx, z["pool_idx"])


def test_config_d_lml():
    z = np.load(os.path.join(GOLDEN_DIR, "config_d_n4000_d20.npz"))
    N, d = int(z["N"]), int(z["d"])
    X, y, _, _ = orc.synthetic_problem(N, d, seed=int(z["seed"]))
    y_mean, y_std = orc.normalize_y_fit(y)
    y_ = (y - y_mean) / y_std
    noise2 = np.full(N, (float(z["noise_level"]) / y_std) ** 2)
    for kind, pick in (("rbf", 1), ("matern25", 0)):      # (all of them: GPU test_gpu_configs.py)
        for th, v, gr in zip(z[f"thetas_{kind}"][pick:pick + 1], z[f"lml_{kind}"][pick:],
                             z[f"grad_{kind}"][pick:]):
            lml, grad = orc.log_marginal_likelihood(kind, th, X, y_, noise2, eval_gradient=True)
            assert abs(lml - v) <= 1e-11 * abs(v), (kind, lml, v)
            assert scaled_err(grad, gr, np.abs(gr).max()) < 1e-10


def test_config_e_mean_only():
    z = np.load(os.path.join(GOLDEN_DIR, "config_e_n2000_d16.npz"))
    N, d, M = int(z["N"]), int(z["d"]), int(z["M"])
    X, y, theta, bounds = orc.synthetic_problem(N, d, seed=int(z["seed"]))
    st = orc.GPState("rbf", theta, X, y, bounds=bounds)
    Xc = np.random.default_rng(int(z["cand_seed"])).uniform(size=(M, d))
    mean = orc.predict(st, Xc[:8000])
    assert scaled_err(mean, z["mean"][:8000], float(z["y_std"])) < TOL
