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


def test_acquisition_engines_on_fake_device(monkeypatch):
    """NORA (ranked pool of the reference reproduced) and BatchOptimizer end to end."""
    install(monkeypatch)
    from conftest import golden_pool_candidates
    from gpry_b200.acquisition_functions import LogExp
    from gpry_b200.gp_acquisition import NORA, BatchOptimizer
    from gpry_b200.preprocessing import Normalize_bounds
    g = load_golden("rbf_d2_n60")
    gpr = make_gpr(g)
    Xp = golden_pool_candidates(g)
    nora = NORA(g["bounds"], acq_func=LogExp(zeta=g["zeta"]), kprime=64)
    X_pool, y_pool, acq_pool = nora.multi_add(gpr, n_points=int(g["pool_n_points"]), X_mc=Xp)
    assert np.array_equal(X_pool, Xp[g["pool_idx_single_sort_acq"]])
    assert np.array_equal(nora.last_pool_idx, g["pool_idx_single_sort_acq"])
    # the same sample again (same object): the rows proposed above are skipped by index, and
    # the result is the ranking of the sample without them (gp_acquisition.py:1037-1047)
    X2, _, _ = nora.multi_add(gpr, n_points=int(g["pool_n_points"]), X_mc=Xp)
    assert not nora.last_new_sample
    assert not set(map(bytes, X2)) & set(map(bytes, X_pool))
    fresh = NORA(g["bounds"], acq_func=LogExp(zeta=g["zeta"]), kprime=64)
    X2_ref, _, _ = fresh.multi_add(gpr, n_points=int(g["pool_n_points"]),
                                   X_mc=np.delete(Xp, g["pool_idx_single_sort_acq"], axis=0))
    assert np.array_equal(X2, X2_ref)
    assert len(nora._X_already_proposed) == 2 * int(g["pool_n_points"])
    # a sharded hand-over (what every rank of a multi-GPU run does) gives the same pool
    sharded = NORA(g["bounds"], acq_func=LogExp(zeta=g["zeta"]), kprime=64)
    X3, _, _ = sharded.multi_add(gpr, n_points=int(g["pool_n_points"]), X_shard=Xp)
    assert np.array_equal(X3, X_pool)
    opt = BatchOptimizer(g["bounds"], preprocessing_X=Normalize_bounds(g["bounds"]),
                         acq_func=LogExp(zeta=g["zeta"]), n_restarts_optimizer=4, verbose=0)
    n0 = gpr.n
    Xb, yb, ab = opt.multi_add(gpr, n_points=3, rng=np.random.default_rng(1))
    assert gpr.n == n0 and Xb.shape == (3, 2) and np.all(np.isfinite(ab))
    # the working copy factorises once (a deep copy carries host factors only), further lies
    # are appended by bordering
    assert len(names("factor_append")) == 1
    acq = LogExp(zeta=g["zeta"])
    assert abs(acq(Xb[:1], gpr)[0] - ab[0]) < 1e-8 * max(1.0, abs(ab[0]))


def test_device_classifier_plumbing(monkeypatch):
    """account_for_inf="SVM": trained on the host (scikit-learn), evaluated by the device
    entries (here the fake ones); the regressor binds it after every model / classifier change."""
    install(monkeypatch)
    from sklearn.svm import SVC
    from gpry_b200.gpr import GaussianProcessRegressor
    from gpry_b200.preprocessing import Normalize_bounds, Normalize_y
    from gpry_b200.svm import SVM
    rng = np.random.default_rng(0)
    bounds = np.array([[-1.0, 3.0]] * 3)
    X = bounds[:, 0] + 4.0 * rng.random((150, 3))
    y = -0.5 * np.sum(((X - 1.0) / 0.45) ** 2, axis=1)
    y[:5] = -np.inf
    gpr = GaussianProcessRegressor(kernel="RBF", bounds=bounds, noise_level=1e-2,
                                   preprocessing_X=Normalize_bounds(bounds),
                                   preprocessing_y=Normalize_y(), account_for_inf="SVM",
                                   inf_threshold=12.0, verbose=0, random_state=0)
    gpr.kernel_ = deepcopy(gpr.kernel)
    gpr.kernel_.theta = np.log([4.0, 0.35, 0.35, 0.35])
    gpr.append_to_data(X, y, fit_gpr=False)
    clf = gpr.infinities_classifier
    assert isinstance(clf, SVM) and 0 < clf.y_finite.sum() < len(y) and gpr.n == clf.y_finite.sum()
    Xc = bounds[:, 0] + 4.0 * rng.random((4000, 3))
    ref = SVC(C=1e7, kernel="rbf", gamma="scale").fit(clf.X_train, clf.y_finite)
    dec_ref = ref.decision_function(gpr.preprocessing_X.transform(Xc))
    clear = np.abs(dec_ref) > 1e-7
    m, s = gpr.predict(Xc, return_std=True)
    bad = (dec_ref <= 0) & clear
    assert bad.any() and np.all(m[bad] == -np.inf) and np.all(s[bad] == 0)
    assert np.all(np.isfinite(m[(dec_ref > 0) & clear]))
    assert np.array_equal(clf.predict(gpr.preprocessing_X.transform(Xc))[clear], (dec_ref > 0)[clear])
    assert np.all(gpr.predict_std(Xc)[bad] == 0)
    assert np.all(gpr.predict_logexp(Xc, 0.5)[2][bad] == -np.inf)
    # classifier-only refit (only infinite points appended) re-binds it on the device state
    v0 = clf.version
    gpr.append_to_data(bounds[:, 1][None] - 1e-3, np.array([-np.inf]), fit_gpr=False)
    assert clf.version > v0 and gpr._clf_bound == (id(clf), v0)      # stale until the next call
    assert gpr.predict(bounds[:, 1][None] - 1e-3)[0] == -np.inf
    assert gpr._clf_bound == (id(clf), clf.version)
