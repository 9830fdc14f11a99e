"""Edge cases of the path on the GPU: non-finite candidates, variance clamp at (near-)
duplicate points, extreme shapes (N = 1, d = 128, N not a multiple of any tile), -inf
acquisition ordering, repeated uploads of different models into one state."""
import numpy as np
import pytest

from conftest import scaled_err
from oracle import gp_oracle as orc
from test_gpu_predict import upload_from_oracle

pytestmark = pytest.mark.gpu
TOL = 1e-10


@pytest.fixture(scope="module")
def dev():
    from gpry_b200 import DeviceGP
    d = DeviceGP(0)
    yield d
    d.close()


def test_nonfinite_candidates_rank_last(dev):
    X, y, theta, bounds = orc.synthetic_problem(200, 4)
    st = orc.GPState("rbf", theta, X, y, bounds=bounds)
    upload_from_oracle(dev, st)
    Xc = np.random.default_rng(0).uniform(size=(1000, 4))
    Xc[7, 2] = np.nan
    Xc[11, 0] = np.inf
    mean, std, acq = dev.predict_logexp(Xc, 0.3, st.noise_level, st.y_max)
    assert np.isnan(mean[7]) and np.isnan(acq[7])
    good = np.ones(1000, bool)
    good[[7, 11]] = False
    mo, so, ao = orc.predict_logexp(st, Xc[good], zeta=0.3)
    assert scaled_err(mean[good], mo, st.y_std) < TOL
    a, idx, m, s, Xo = dev.predict_logexp_topk(Xc, 0.3, st.noise_level, st.y_max, 1000)
    assert len(idx) == 1000 and sorted(idx.tolist()) == list(range(1000))
    nan_pos = [int(np.flatnonzero(idx == i)[0]) for i in np.flatnonzero(np.isnan(acq))]
    n_nan = int(np.isnan(acq).sum())
    assert sorted(nan_pos) == list(range(1000 - n_nan, 1000))          # NaN ranks last
    fin = ~np.isnan(a)
    assert np.all(np.diff(a[fin]) <= 0)                                  # descending
    ninf = np.flatnonzero(a == -np.inf)
    if len(ninf):                                                        # -inf after finite
        assert ninf[0] > np.flatnonzero(np.isfinite(a))[-1]


def test_variance_clamp_at_training_points(dev):
    X, y, theta, bounds = orc.synthetic_problem(300, 3)
    st = orc.GPState("rbf", theta, X, y, bounds=bounds, noise_level=1e-5)
    upload_from_oracle(dev, st)
    Xc = np.vstack([X[:100], X[:50] + 1e-9])
    mean, std = dev.predict(Xc, return_std=True)
    mo, so = orc.predict(st, Xc, return_std=True)
    assert np.all(std >= 0) and np.all(np.isfinite(std))
    assert scaled_err(mean, mo, st.y_std) < 1e-8          # cond(K) ~ 1e10 here: not a 1e-10 case
    assert np.max(np.abs(std ** 2 - so ** 2)) < 1e-6 * st.y_std ** 2
    acq = dev.predict_logexp(Xc, 0.4, 1e-3, st.y_max)[2]
    assert np.all((acq == -np.inf) | np.isfinite(acq))


@pytest.mark.parametrize("kind,N,d,M", [("rbf", 1, 1, 5), ("rbf", 2, 128, 300),
                                        ("matern25", 513, 64, 200), ("matern15", 1025, 2, 1000)])
def test_extreme_shapes(dev, kind, N, d, M):
    rng = np.random.default_rng(N + d)
    X = rng.uniform(size=(N, d))
    y = -0.5 * np.sum(((X - 0.5) / 0.3) ** 2, axis=1) if N > 1 else np.array([1.0])
    theta = np.log(np.concatenate([[1.5], np.full(d, 0.7 * np.sqrt(d))]))
    bounds = np.array([[0.0, 1.0]] * d)
    st = orc.GPState(kind, theta, X, y, bounds=bounds, normalize_y=N > 1)
    upload_from_oracle(dev, st)
    Xc = rng.uniform(size=(M, d))
    mean, std = dev.predict(Xc, return_std=True)
    mo, so = orc.predict(st, Xc, return_std=True)
    assert scaled_err(mean, mo, st.y_std) < TOL
    assert scaled_err(std ** 2, so ** 2, st.y_std ** 2) < TOL
    L, V, alpha_, _, info = dev.factorize(kind, st.X_train_, st.noise2, st.y_train_, theta)
    assert info == 0
    assert scaled_err(alpha_, st.alpha_, max(1e-300, np.abs(st.alpha_).max())) < 1e-9
    lml, grad, info = dev.lml_batched(kind, st.X_train_, st.noise2, st.y_train_, theta[None])
    lo, go = orc.log_marginal_likelihood(kind, theta, st.X_train_, st.y_train_, st.noise2,
                                         eval_gradient=True)
    assert abs(lml[0] - lo) < TOL * max(1.0, abs(lo))
    assert scaled_err(grad[0], go, max(1e-300, np.abs(go).max())) < 1e-9


def test_state_reuse_across_models(dev):
    """One state, several uploads of different (kind, N, d): buffers are re-used correctly."""
    for kind, N, d in [("rbf", 700, 9), ("matern25", 90, 3), ("rbf", 1300, 12), ("matern15", 90, 3)]:
        X, y, theta, bounds = orc.synthetic_problem(N, d, seed=N)
        st = orc.GPState(kind, theta, X, y, bounds=bounds)
        upload_from_oracle(dev, st)
        Xc = np.random.default_rng(d).uniform(size=(777, d))
        mean, std = dev.predict(Xc, return_std=True)
        mo, so = orc.predict(st, Xc, return_std=True)
        assert scaled_err(mean, mo, st.y_std) < TOL
        assert scaled_err(std ** 2, so ** 2, st.y_std ** 2) < TOL


def test_argument_errors(dev):
    from gpry_b200 import DeviceGP, GpryB200Error
    fresh = DeviceGP(0)
    with pytest.raises(GpryB200Error):
        fresh.predict(np.zeros((3, 2)))                  # no model uploaded
    fresh.close()
    X, y, theta, bounds = orc.synthetic_problem(50, 2)
    st = orc.GPState("rbf", theta, X, y, bounds=bounds)
    upload_from_oracle(dev, st)
    with pytest.raises(ValueError):
        dev.predict(np.zeros((3, 5)))                    # wrong dimensionality
    with pytest.raises(GpryB200Error):
        dev.predict_logexp_topk(np.zeros((3, 2)), 0.3, 0.01, 0.0, 0)      # K' < 1
