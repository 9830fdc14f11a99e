"""GPU parity of the training-side path: kernel matrix + blocked FP64 Cholesky + L^-1 + alpha
(gpry_factorize) and the log marginal likelihood with its gradient (gpry_lml_batched), against
golden vectors from the real reference and against the CPU oracle."""
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR, oracle_state, scaled_err
from oracle import gp_oracle as orc

pytestmark = pytest.mark.gpu

TOL = 1e-10


@pytest.fixture(scope="module")
def dev():
    from gpry_b200 import DeviceGP
    d = DeviceGP(0)
    yield d
    d.close()


def test_golden_factorize(dev, golden):
    g = golden
    st = oracle_state(g)
    L, V, alpha_, logdet_half, info = dev.factorize(g["kind"], st.X_train_, st.noise2,
                                                    st.y_train_, g["theta"])
    assert info == 0
    N = g["N"]
    assert scaled_err(np.diag(L), g["L_diag"], 1.0) < TOL
    tol_c = max(TOL, 5e-17 * float(g["condK"]))     # reproducible to ~eps cond(K) only
    assert scaled_err(alpha_, g["alpha_"], np.abs(g["alpha_"]).max()) < tol_c
    assert scaled_err(V[[0, N // 2, N - 1]], g["V_rows"], np.abs(g["V_rows"]).max()) < tol_c
    assert abs(np.linalg.norm(V) - float(g["V_fro"])) < TOL * float(g["V_fro"])
    assert np.all(np.triu(L, 1) == 0) and np.all(np.triu(V, 1) == 0)
    assert abs(logdet_half - np.log(g["L_diag"]).sum()) < 1e-10 * max(1, abs(logdet_half))
    # full-matrix check against the oracle factor
    assert scaled_err(L, st.L_, np.abs(st.L_).max()) < TOL
    assert scaled_err(V, st.V_, np.abs(st.V_).max()) < TOL


def test_golden_lml(dev, golden):
    g = golden
    if "lml" not in g:
        pytest.skip("no LML in this fixture")
    st = oracle_state(g)
    lml, grad, info = dev.lml_batched(g["kind"], st.X_train_, st.noise2, st.y_train_,
                                      g["lml_thetas"])
    assert np.all(info == 0)
    assert scaled_err(lml, g["lml"], 1.0) < TOL
    for b in range(len(lml)):
        assert scaled_err(grad[b], g["lml_grad"][b], np.abs(g["lml_grad"][b]).max()) < TOL
    lml2, grad2, _ = dev.lml_batched(g["kind"], st.X_train_, st.noise2, st.y_train_,
                                     g["lml_thetas"], eval_gradient=False)
    assert grad2 is None and np.array_equal(lml2, lml)


def test_lml_nonpd(dev):
    z = np.load(os.path.join(GOLDEN_DIR, "lml_nonpd.npz"))
    lml, grad, info = dev.lml_batched("rbf", z["X_train_"], z["noise2"], z["y_train_"],
                                      z["theta"][None, :])
    assert info[0] > 0 and lml[0] == -np.inf and np.array_equal(grad[0], np.zeros(3))
    L, V, a, ld, info2 = dev.factorize("rbf", z["X_train_"], z["noise2"], z["y_train_"],
                                       z["theta"])
    assert info2 > 0


@pytest.mark.parametrize("kind,N,d", [("rbf", 1000, 8), ("matern25", 500, 6),
                                      ("matern15", 129, 2), ("rbf", 2000, 12), ("rbf", 203, 5),
                                      ("matern25", 61, 3)])
def test_oracle_lml(dev, kind, N, d):
    X, y, theta, bounds = orc.synthetic_problem(N, d)
    st = orc.GPState(kind, theta, X, y, bounds=bounds)
    rng = np.random.default_rng(7)
    thetas = np.array([theta, theta + 0.2 * rng.standard_normal(theta.shape)])
    lml, grad, info = dev.lml_batched(kind, st.X_train_, st.noise2, st.y_train_, thetas)
    assert np.all(info == 0)
    for b in range(2):
        lo, go = orc.log_marginal_likelihood(kind, thetas[b], st.X_train_, st.y_train_,
                                             st.noise2, eval_gradient=True)
        assert abs(lml[b] - lo) < TOL * abs(lo)
        assert scaled_err(grad[b], go, np.abs(go).max()) < TOL


def test_adopt_factorization(dev):
    N, d = 700, 7
    X, y, theta, bounds = orc.synthetic_problem(N, d)
    st = orc.GPState("rbf", theta, X, y, bounds=bounds)
    c, ell = orc.split_theta(theta)
    x_min, x_w = bounds[:, 0], bounds[:, 1] - bounds[:, 0]
    clip_hi = 1.1 * y.max() - 0.1 * y.min()
    Xc = np.random.default_rng(1).uniform(size=(2000, d))
    L, V, alpha_, _, info = dev.factorize("rbf", st.X_train_, st.noise2, st.y_train_, theta,
                                          keep_on_device=True)
    assert info == 0
    dev.adopt_factorization(c, ell, x_min, x_w, st.y_mean, st.y_std, clip_hi)
    m1, s1 = dev.predict(Xc, return_std=True)
    dev.upload("rbf", st.X_train_, alpha_, V, c, ell, x_min, x_w, st.y_mean, st.y_std, clip_hi)
    m2, s2 = dev.predict(Xc, return_std=True)
    assert np.array_equal(m1, m2) and np.array_equal(s1, s2)
    mo, so = orc.predict(st, Xc, return_std=True)
    assert scaled_err(m1, mo, st.y_std) < TOL
    assert scaled_err(s1 ** 2, so ** 2, st.y_std ** 2) < TOL
