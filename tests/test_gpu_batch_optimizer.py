"""batched predict gradients and the lock-step acquisition optimiser.

The reference evaluates gradients one point per call (gpr.py:1095-1097), so the batched entry
is pinned row by row against the oracle's one-point path (itself pinned to the reference's
golden vectors), against the one-point C entries, and -- for the optimiser -- against running
scipy's L-BFGS-B restart by restart through the one-point path as the reference does."""
from copy import deepcopy

import numpy as np
import pytest
import scipy.optimize

from conftest import load_golden, oracle_state, scaled_err
from oracle import gp_oracle as orc
from test_gpu_gpr import make_gpr

pytestmark = pytest.mark.gpu
TOL = 1e-10


def grad_std_extended(st, x):
    """d std/dx_ from the oracle's own V_, kernel values and kernel gradient, with the
    contractions carried out in extended precision.  The reference's FP64 evaluation
    (gpr.py:1250-1252 forms V^T V explicitly) loses up to ~1e-5 of relative accuracy on the
    ill-conditioned golden cases (rbf_d2_n60: |V|_max = 190), so the device result is required
    to be as close to this value as the reference's own FP64 result is (or 1e-8)."""
    X_ = st.transform_X(x[None])
    Kt = orc.kernel_cross(st.kind, st.theta, X_, st.X_train_)[0].astype(np.longdouble)
    grad = orc.kernel_gradient_x(st.kind, st.theta, X_[0], st.X_train_).astype(np.longdouble)
    V = st.V_.astype(np.longdouble)
    w = V @ Kt
    var = np.longdouble(np.exp(st.theta[0])) - w @ w
    gs = -((V.T @ w) @ grad) / np.sqrt(var)
    if st.normalize_y:
        gs = gs * st.y_std * st.y_std
    return np.asarray(gs, dtype=float)


def check_grad_std(st, x, ours, ref):
    truth = grad_std_extended(st, x)
    scale = np.abs(truth).max()
    err_ref = scaled_err(ref, truth, scale)
    assert scaled_err(ours, truth, scale) < max(1e-8, 4 * err_ref)


def test_predict_grad_batch_like_oracle(golden):
    g = golden
    gpr = make_gpr(g)
    st = oracle_state(g)
    X = g["Xc"][:150]
    n0 = gpr.n_eval
    mean, std, gm, gs = gpr.predict_grad_batch(X)
    assert gpr.n_eval == n0 + len(X)
    sy = float(g["y_std"])
    m_ref, s_ref = gpr.predict(X, return_std=True)
    assert np.allclose(mean, m_ref, rtol=0, atol=1e-12 * max(1.0, np.abs(m_ref).max()))
    assert np.allclose(std, s_ref, rtol=0, atol=1e-10 * sy)
    for i in range(0, len(X), 7):
        mo, so, gmo, gso = orc.predict(st, X[i:i + 1], return_std=True, return_mean_grad=True,
                                       return_std_grad=True)
        assert scaled_err(gm[i], gmo, np.abs(gmo).max()) < TOL
        check_grad_std(st, X[i], gs[i], gso)
        # the one-point C entries give the same numbers
        _, _, gm1, gs1 = gpr.predict(X[i:i + 1], return_std=True, return_mean_grad=True,
                                     return_std_grad=True)
        assert scaled_err(gm[i], gm1, np.abs(gm1).max()) < 1e-12
        check_grad_std(st, X[i], gs1, gso)
    # golden vector of the reference itself (first candidate)
    assert scaled_err(gm[0], g["grad_mean"], np.abs(g["grad_mean"]).max()) < TOL
    check_grad_std(st, X[0], gs[0], g["grad_std"])
    # ragged sizes: one row, a tile boundary, more than one tile
    for M in (1, 127, 128, 129):
        Xr = np.resize(g["Xc"], (M, g["d"]))
        out = gpr.predict_grad_batch(Xr)
        k = min(M, len(X))
        assert np.array_equal(out[2][:k], gm[:k]) or scaled_err(out[2][:k], gm[:k], 1.0) < 1e-12
    _, _, gm_only = gpr.predict_grad_batch(X[:5], return_std_grad=False)
    assert scaled_err(gm_only, gm[:5], np.abs(gm).max()) < 1e-12


def test_predict_grad_batch_at_training_points():
    """At a training point r = 0: Matern gradients are defined as 0 there (kernels.py:363-432)
    and the std is tiny; nothing may turn into NaN."""
    for name in ("matern15_d5_n200", "matern25_d8_n300", "rbf_d8_n300"):
        g = load_golden(name)
        gpr = make_gpr(g)
        st = oracle_state(g)
        X = g["X_train"][:9]
        mean, std, gm, gs = gpr.predict_grad_batch(X)
        assert np.all(np.isfinite(gm)) and np.all(np.isfinite(gs))
        for i in (0, 4, 8):
            _, so, gmo, gso = orc.predict(st, X[i:i + 1], return_std=True, return_mean_grad=True,
                                          return_std_grad=True)
            assert scaled_err(gm[i], gmo, max(np.abs(gmo).max(), 1e-300)) < 1e-9
            if so[0] > 1e-6 * float(g["y_std"]):
                truth = grad_std_extended(st, X[i])
                scale = max(np.abs(truth).max(), 1e-300)
                assert scaled_err(gs[i], truth, scale) < max(1e-6, 4 * scaled_err(gso, truth, scale))


def test_logexp_batch_gradient(golden):
    from gpry_b200.acquisition_functions import LogExp
    g = golden
    gpr = make_gpr(g)
    acq = LogExp(zeta=g["zeta"])
    X = g["Xc"][:40]
    vals, grads = acq(X, gpr, eval_gradient=True)
    assert vals.shape == (40,) and grads.shape == (40, g["d"])
    plain = acq(X, gpr)
    both = np.isfinite(vals) & np.isfinite(plain)
    assert np.array_equal(np.isfinite(vals), np.isfinite(plain))
    assert scaled_err(vals[both], plain[both], 1.0) < 1e-9
    for i in range(0, 40, 9):
        v1, g1 = acq(X[i:i + 1], gpr, eval_gradient=True)
        if np.isfinite(v1[0]):
            assert abs(vals[i] - v1[0]) < 1e-9 * max(1.0, abs(v1[0]))
            assert scaled_err(grads[i], g1, np.abs(g1).max()) < 1e-5
        else:
            assert np.all(np.isinf(grads[i]))


def test_lockstep_matches_sequential_optimisation():
    """Every restart of the lock-step optimiser follows the iterates scipy produces when the
    same start is optimised alone through the one-point path (what the reference does,
    gp_acquisition.py:503-511)."""
    from gpry_b200.gp_acquisition import BatchOptimizer
    from gpry_b200.preprocessing import Normalize_bounds
    g = load_golden("rbf_d2_n60")
    gpr = make_gpr(g)
    opt = BatchOptimizer(g["bounds"], preprocessing_X=Normalize_bounds(g["bounds"]),
                         n_restarts_optimizer=6, verbose=0)
    rng = np.random.default_rng(3)
    opt.proposer.update(gpr)
    x0, value, optimise = opt._starting_points(gpr, list(range(6)), g["bounds"], rng)
    assert optimise.all() and np.all(np.isfinite(value[1:]))
    assert np.array_equal(x0[0], gpr.X_train[-1])
    # proposals are the best of n_repeats_propose + 1 finite draws -> better than a typical draw
    x0_opt = opt._to_opt(x0)
    ob = opt._opt_bounds(g["bounds"])
    together = opt._optimize_all(gpr, x0_opt, ob)
    for i in range(6):
        def one(x):
            val, grad = opt.acq_func(opt._from_opt(np.atleast_2d(x)), gpr, eval_gradient=True)
            return -val[0], -np.asarray(grad).reshape(-1)
        res = scipy.optimize.minimize(one, x0_opt[i], method="L-BFGS-B", jac=True, bounds=ob)
        assert together[i][1] <= -value[i] + 1e-9          # never worse than its start
        assert abs(together[i][1] - res.fun) < 1e-5 * max(1.0, abs(res.fun))
        assert np.allclose(together[i][0], res.x, atol=1e-4)


@pytest.mark.parametrize("name", ["rbf_d2_n60", "rbf_d8_n300"])
def test_batch_optimizer_multi_add(name):
    from gpry_b200.gp_acquisition import BatchOptimizer
    from gpry_b200.acquisition_functions import LogExp
    from gpry_b200.preprocessing import Normalize_bounds
    g = load_golden(name)
    gpr = make_gpr(g)
    n_train = gpr.n
    opt = BatchOptimizer(g["bounds"], preprocessing_X=Normalize_bounds(g["bounds"]),
                         n_restarts_optimizer="4d", verbose=0)
    assert opt.n_restarts_optimizer == 4 * g["d"]
    n_points = 3
    X, y_lies, acq_vals = opt.multi_add(gpr, n_points=n_points, rng=np.random.default_rng(11))
    assert X.shape == (n_points, g["d"]) and y_lies.shape == (n_points,)
    assert gpr.n == n_train                                 # the caller's regressor is untouched
    lo, hi = g["bounds"][:, 0], g["bounds"][:, 1]
    assert np.all(X >= lo - 1e-12) and np.all(X <= hi + 1e-12)
    assert np.all(np.isfinite(acq_vals))
    # first point: its value is the acquisition function of the un-augmented model there; it is
    # a local optimum reached from the best of several starts, so it is at least as good as the
    # start from the last training point and better than almost all of a random pool (the
    # surface is multi-modal: a local optimiser need not beat the best of 2000 random draws)
    acq = LogExp(dimension=g["d"])
    assert abs(acq(X[:1], gpr)[0] - acq_vals[0]) < 1e-8 * max(1.0, abs(acq_vals[0]))
    assert acq_vals[0] >= acq(gpr.X_train[-1:], gpr)[0] - 1e-9
    pool = lo + (hi - lo) * np.random.default_rng(5).random((2000, g["d"]))
    pool_vals = acq(pool, gpr)
    assert acq_vals[0] >= np.quantile(pool_vals[np.isfinite(pool_vals)], 0.9)
    assert np.allclose(y_lies[0], gpr.predict(X[:1])[0])
    # later points are conditioned on the lies: re-building that model reproduces the values
    g2 = deepcopy(gpr)
    g2.append_to_data(X[:1], y_lies[:1], fit_gpr=False, fit_classifier=False)   # as :488-491
    assert abs(acq(X[1:2], g2)[0] - acq_vals[1]) < 1e-7 * max(1.0, abs(acq_vals[1]))
    assert np.min(np.linalg.norm(X[1] - X[0])) > 0
    # one-restart entry of the reference's interface
    x, f = opt.optimize_acquisition_function(gpr, 0, rng=np.random.default_rng(1))
    assert x.shape == (g["d"],) and np.isfinite(f)
