"""Host logic of the regressor (model-update decisions, lazy L_/V_, masks, pickling) and of the
in-place GPry patch, run on the CPU against an oracle-backed stand-in for the device
(tests/fake_device.py).  The arithmetic of the real device is covered by the -m gpu tests; here
the point is that every code path of the Python layer executes and lands on the reference's
numbers."""
import pickle
from copy import deepcopy

import numpy as np
import pytest

from conftest import load_golden, scaled_err
from fake_device import FakeDeviceGP, install
from oracle.ref_import import import_reference, reference_available

TOL = 1e-10


def make_gpr(g, **kw):
    from gpry_b200.gpr import GaussianProcessRegressor
    from gpry_b200.preprocessing import Normalize_bounds, Normalize_y
    kernel = {"rbf": "RBF", "matern15": {"Matern": {"nu": 1.5}},
              "matern25": {"Matern": {"nu": 2.5}}}[g["kind"]]
    gpr = GaussianProcessRegressor(
        kernel=kernel, bounds=g["bounds"], noise_level=g["noise_level"],
        clip_factor=g["clip_factor"], preprocessing_X=Normalize_bounds(g["bounds"]),
        preprocessing_y=Normalize_y(), account_for_inf=None, verbose=0, **kw)
    gpr.kernel_ = deepcopy(gpr.kernel)
    gpr.kernel_.theta = g["theta"]
    gpr.append_to_data(g["X_train"], g["y_train"], fit_gpr=False)
    return gpr


def names(kind):
    return [c[0] for c in FakeDeviceGP.calls if c[0] == kind]


def test_mirror_host_logic(monkeypatch):
    install(monkeypatch)
    g = load_golden("rbf_d8_n300")
    gpr = make_gpr(g)
    sy = float(g["y_std"])
    # one factorisation kept on the device, adopted without any N^2 transfer
    assert names("factorize") == ["factorize"] and not names("upload") and not names("factor_download")
    mean, std = gpr.predict(g["Xc"], return_std=True)
    assert names("adopt") == ["adopt"]
    assert scaled_err(mean, g["mean"], sy) < TOL and scaled_err(std ** 2, g["std"] ** 2, sy ** 2) < TOL
    m1, s1, gm, gs = gpr.predict(g["Xc"][:1], return_std=True, return_mean_grad=True,
                                 return_std_grad=True)
    assert scaled_err(gm, g["grad_mean"], np.abs(g["grad_mean"]).max()) < TOL
    assert scaled_err(gs, g["grad_std"], np.abs(g["grad_std"]).max()) < 1e-8
    mb, sb, gmb, gsb = gpr.predict_grad_batch(g["Xc"][:5])
    assert scaled_err(gmb[0], gm, np.abs(gm).max()) < 1e-12 and mb.shape == (5,)
    # lazy host copies: fetched once, on first access
    assert not names("factor_download")
    assert gpr.V_.shape == (300, 300) and gpr.L_.shape == (300, 300)
    assert names("factor_download") == ["factor_download"]
    assert scaled_err(gpr.alpha_, g["alpha_"], np.abs(g["alpha_"]).max()) < TOL
    # bordered append for lies, refactorisation when the y pre-processor is refit
    lies = gpr.predict(g["Xc"][:3])
    gpr.append_to_data(g["Xc"][:3], lies, fit_gpr=False, fit_classifier=False)
    assert names("factor_append") == ["factor_append"] and len(names("factorize")) == 1
    assert gpr.n == 303 and gpr.V_.shape == (303, 303)
    gpr.append_to_data(g["Xc"][3:4], lies[:1], fit_gpr=False)          # pre-processors refit
    assert len(names("factorize")) == 2 and len(names("factor_append")) == 1
    # copies and pickles carry host factors and no device handle; they upload on first use
    n_dl = len(names("factor_download"))
    c = deepcopy(gpr)
    assert len(names("factor_download")) == n_dl + 1 and c._dev is None and not c._factor_resident
    p = pickle.loads(pickle.dumps(gpr))
    assert p._dev is None and np.array_equal(p.V_, gpr.V_)
    assert np.array_equal(c.predict(g["Xc"]), gpr.predict(g["Xc"]))
    assert names("upload") == ["upload"]
    # trust region and mask value go to the device per call
    gpr.trust_bounds = np.array(g["bounds"], dtype=float)
    gpr.trust_bounds[:, 1] = gpr.trust_bounds[:, 0] + 0.5 * (g["bounds"][:, 1] - g["bounds"][:, 0])
    out = ~np.all((g["Xc"] >= gpr.trust_bounds[:, 0]) & (g["Xc"] <= gpr.trust_bounds[:, 1]), axis=1)
    m = gpr.predict(g["Xc"])
    assert out.any() and np.all(m[out] == -np.inf) and np.all(np.isfinite(m[~out]))
    gpr.minus_inf_value = -1e300
    assert np.all(gpr.predict(g["Xc"])[out] == -1e300)
    assert np.all(np.isfinite(gpr.predict(g["Xc"], ignore_trust_region=True)))
    assert np.all(gpr.predict_logexp(g["Xc"], 0.5)[2][out] == -np.inf)
    # hyper-parameter fit in lock-step through the (fake) batched LML
    gpr2 = make_gpr(load_golden("rbf_d2_n60"), n_restarts_optimizer=3, random_state=3)
    gpr2.fit_gpr_hyperparameters()
    assert gpr2.fitted and np.isfinite(gpr2.log_marginal_likelihood_value_)
    assert gpr2.log_marginal_likelihood_value_ >= gpr2.log_marginal_likelihood(
        load_golden("rbf_d2_n60")["theta"]) - 1e-9


def test_failed_factorisation_raises_like_the_reference(monkeypatch):
    install(monkeypatch)
    g = load_golden("rbf_d2_n60")
    gpr = make_gpr(g)
    X_dup = np.vstack([g["X_train"][:2], g["X_train"][:2]])       # duplicates, no noise: singular
    gpr.noise_level = 0.0
    gpr.noise_level_ = np.zeros_like(gpr.noise_level_)
    with pytest.raises(np.linalg.LinAlgError, match="not returning a positive definite"):
        gpr.append_to_data(X_dup, np.zeros(4), noise_level=None, fit_gpr=False, fit_classifier=False)


@pytest.mark.skipif(not reference_available(), reason="reference tree not present")
def test_patched_reference_runs_and_matches_itself(monkeypatch):
    """gpry.gpr.GaussianProcessRegressor with the device-backed methods patched in gives the
    numbers of the unpatched class (here through the oracle-backed fake device)."""
    install(monkeypatch)
    gpry = import_reference()
    from gpry.preprocessing import Normalize_bounds, Normalize_y
    from sklearn.base import clone
    from gpry_b200 import integration
    g = load_golden("matern25_d8_n300")
    cls = gpry.gpr.GaussianProcessRegressor

    def build():
        r = cls(kernel={"Matern": {"nu": 2.5}}, bounds=g["bounds"], noise_level=g["noise_level"],
                account_for_inf=None, preprocessing_X=Normalize_bounds(g["bounds"]),
                preprocessing_y=Normalize_y(), verbose=0)
        r.kernel_ = clone(r.kernel)
        r.kernel_.theta = g["theta"]
        r.append_to_data(g["X_train"], g["y_train"], fit_gpr=False)
        return r
    ref = build()
    m_ref, s_ref = ref.predict(g["Xc"], return_std=True)
    lml_ref, grad_ref = ref.log_marginal_likelihood(g["theta"], eval_gradient=True)
    integration.patch_gpry(gpry)
    try:
        FakeDeviceGP.calls.clear()
        pat = build()
        assert names("factorize") == ["factorize"]
        m, s = pat.predict(g["Xc"], return_std=True)
        sy = float(g["y_std"])
        assert scaled_err(m, m_ref, sy) < TOL and scaled_err(s ** 2, s_ref ** 2, sy ** 2) < TOL
        assert scaled_err(pat.predict_std(g["Xc"]) ** 2, s_ref ** 2, sy ** 2) < TOL
        lml, grad = pat.log_marginal_likelihood(g["theta"], eval_gradient=True)
        assert abs(lml - lml_ref) < 1e-10 * abs(lml_ref)
        assert np.max(np.abs(grad - grad_ref)) < 1e-9 * np.max(np.abs(grad_ref))
        assert scaled_err(pat.V_, ref.V_, np.abs(ref.V_).max()) < TOL          # lazy property
        assert scaled_err(pat.alpha_, ref.alpha_, np.abs(ref.alpha_).max()) < TOL
        # GPry's own __deepcopy__ / dill pickling work on the patched class
        c = deepcopy(pat)
        assert c.__dict__.get("_dev") is None
        assert np.allclose(c.predict(g["Xc"]), m, rtol=0, atol=1e-12 * sy)
        import dill
        p = dill.loads(dill.dumps(pat))
        assert p.__dict__.get("_dev") is None and np.array_equal(p.V_, pat.V_)
        # lies are appended by bordering
        lies = pat.predict(g["Xc"][:2])
        pat.append_to_data(g["Xc"][:2], lies, fit_gpr=False, fit_classifier=False)
        assert names("factor_append") == ["factor_append"]
        ref.append_to_data(g["Xc"][:2], lies, fit_gpr=False, fit_classifier=False)
        assert scaled_err(pat.predict(g["Xc"]), ref.predict(g["Xc"]), sy) < TOL
    finally:
        integration.unpatch_gpry(gpry)
    assert not isinstance(cls.__dict__.get("V_"), property)
