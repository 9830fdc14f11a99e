"""The oracle against the REAL reference, imported live (container only: skipped where
/root/reference does not exist, e.g. on the GPU box).  The golden fixtures pin the oracle to
reference outputs that were minted once; this test re-derives a few of them on the spot, at the
headline size, so that a drift of either side (or of the installed numpy / scipy /
scikit-learn under the reference) shows up immediately.

Measured when written: oracle == reference bit for bit for mean, std and LogExp at
N_train = 2000, d = 12 (same operations, same libraries, same order)."""
import warnings

import numpy as np
import pytest

from oracle import gp_oracle as orc
from oracle.ref_import import import_reference, reference_available

pytestmark = pytest.mark.skipif(not reference_available(),
                                reason="reference tree not present (GPU box)")


def _reference_gpr(gpry, kind, X, y, theta, bounds):
    from sklearn.base import clone
    from gpry.preprocessing import Normalize_bounds, Normalize_y
    kernel = {"rbf": "RBF", "matern25": {"Matern": {"nu": 2.5}}}[kind]
    gpr = gpry.gpr.GaussianProcessRegressor(
        kernel=kernel, bounds=bounds, noise_level=1e-2, preprocessing_X=Normalize_bounds(bounds),
        preprocessing_y=Normalize_y(), account_for_inf=None, verbose=0)
    gpr.kernel_ = clone(gpr.kernel)
    gpr.kernel_.theta = np.asarray(theta)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        gpr.append_to_data(X, y, fit_gpr=False)
    return gpr


@pytest.mark.parametrize("kind,N,d,M", [("rbf", 2000, 12, 4000), ("matern25", 500, 6, 2000)])
def test_predict_logexp_identical(kind, N, d, M):
    gpry = import_reference()
    X, y, theta, bounds = orc.synthetic_problem(N, d)
    gpr = _reference_gpr(gpry, kind, X, y, theta, bounds)
    st = orc.GPState(kind, theta, X, y, bounds=bounds)
    Xc = np.random.default_rng(4321).uniform(size=(M, d))
    mean, std = gpr.predict(Xc, return_std=True, validate=False)
    zeta = d ** (-0.85)
    with np.errstate(divide="ignore"):
        acq = gpry.acquisition_functions.LogExp.f(mean, std, gpr.y_max, gpr.noise_level, zeta)
    mo, so, ao = orc.predict_logexp(st, Xc, zeta=zeta)
    sy = st.y_std
    assert np.max(np.abs(mo - mean)) <= 1e-13 * sy
    assert np.max(np.abs(so ** 2 - std ** 2)) <= 1e-13 * sy ** 2
    fin = np.isfinite(acq)
    assert np.array_equal(np.isfinite(ao), fin)
    assert np.max(np.abs(ao[fin] - acq[fin])) <= 1e-9
    assert np.max(np.abs(st.alpha_ - gpr.alpha_)) <= 1e-12 * np.abs(gpr.alpha_).max()
    assert np.max(np.abs(st.V_ - gpr.V_)) <= 1e-12 * np.abs(gpr.V_).max()


def test_lml_identical():
    gpry = import_reference()
    N, d = 600, 8
    X, y, theta, bounds = orc.synthetic_problem(N, d)
    gpr = _reference_gpr(gpry, "rbf", X, y, theta, bounds)
    st = orc.GPState("rbf", theta, X, y, bounds=bounds)
    rng = np.random.default_rng(3)
    for _ in range(3):
        th = theta + 0.2 * rng.standard_normal(d + 1)
        v, g = gpr.log_marginal_likelihood(th, eval_gradient=True, clone_kernel=True)
        lml, grad = orc.log_marginal_likelihood("rbf", th, st.X_train_, st.y_train_, st.noise2,
                                                eval_gradient=True)
        assert abs(lml - v) <= 1e-12 * abs(v)
        assert np.max(np.abs(grad - g)) <= 1e-11 * np.abs(g).max()


def test_fit_start_points_and_optimum():
    """Same random state -> same restart points (``rng.uniform`` over the log-bounds in loop
    order, gpr.py:970-978) -> the oracle-side L-BFGS-B run ends at the reference's optimum."""
    gpry = import_reference()
    from gpry.preprocessing import Normalize_bounds, Normalize_y
    rng = np.random.default_rng(21)
    d, N = 2, 40
    bounds = np.array([[0.0, 1.0]] * d)
    X = rng.uniform(size=(N, d))
    y = -0.5 * np.sum(((X - 0.5) / 0.15) ** 2, axis=1)
    gpr = gpry.gpr.GaussianProcessRegressor(
        kernel="RBF", bounds=bounds, noise_level=1e-2, n_restarts_optimizer=4,
        preprocessing_X=Normalize_bounds(bounds), preprocessing_y=Normalize_y(),
        account_for_inf=None, random_state=7, verbose=0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        gpr.append_to_data(X, y, fit_gpr=True)
    z = np.load(__import__("os").path.join(__import__("conftest").GOLDEN_DIR, "fit_rbf_d2_n40.npz"))
    assert np.array_equal(gpr.kernel_.theta, z["theta_opt"])
    assert gpr.n_eval_loglike == int(z["n_eval_loglike"])
