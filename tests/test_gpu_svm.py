"""the infinities classifier evaluated on the device.

Checker: scikit-learn's own ``SVC.predict`` / ``decision_function`` on the host (the reference's
SVM *is* that class, svm.py:20), and the plain regressor for the unmasked rows."""
import pickle
from copy import deepcopy

import numpy as np
import pytest
from sklearn.svm import SVC

pytestmark = pytest.mark.gpu


def problem(d=3, n=160, seed=0):
    rng = np.random.default_rng(seed)
    bounds = np.array([[-1.0, 3.0]] * d)
    X = bounds[:, 0] + (bounds[:, 1] - bounds[:, 0]) * rng.random((n, d))
    y = -0.5 * np.sum(((X - 1.0) / 0.45) ** 2, axis=1)       # spans ~ -60 .. 0
    y[rng.choice(n, 6, replace=False)] = -np.inf             # a few hard infinities
    return bounds, X, y


def make(bounds, X, y, account_for_inf, threshold=12.0, **kw):
    from gpry_b200.gpr import GaussianProcessRegressor
    from gpry_b200.preprocessing import Normalize_bounds, Normalize_y
    gpr = GaussianProcessRegressor(
        kernel="RBF", bounds=bounds, noise_level=1e-2,
        preprocessing_X=Normalize_bounds(bounds), preprocessing_y=Normalize_y(),
        account_for_inf=account_for_inf, inf_threshold=threshold, verbose=0, random_state=0, **kw)
    gpr.kernel_ = deepcopy(gpr.kernel)
    gpr.kernel_.theta = np.log([4.0] + [0.35] * len(bounds))
    gpr.append_to_data(X, y, fit_gpr=False)
    return gpr


def test_svm_matches_sklearn():
    from gpry_b200.svm import SVM
    rng = np.random.default_rng(1)
    X = rng.random((300, 4))
    y = -40.0 * np.sum((X - 0.5) ** 2, axis=1)
    svm = SVM(random_state=0)
    with pytest.raises(ValueError):
        svm.predict(X)
    finite = svm.fit(X, y, diff_threshold=12.0)
    assert np.array_equal(finite, y >= y.max() - 12.0) and 0 < finite.sum() < len(y)
    assert svm.n == 300 and svm.d == 4 and svm.abs_threshold == y.max() - 12.0
    assert np.array_equal(svm.is_finite(np.array([y.max() - 11.9, y.max() - 12.1, np.nan, -np.inf])),
                          [True, False, False, False])
    ref = SVC(C=1e7, kernel="rbf", gamma="scale").fit(X, finite)
    Z = rng.random((50000, 4))
    dec = svm.decision_function(Z)
    dec_ref = ref.decision_function(Z)
    assert np.max(np.abs(dec - dec_ref)) < 1e-9 * max(1.0, np.abs(dec_ref).max())
    clear = np.abs(dec_ref) > 1e-7
    assert np.array_equal(svm.predict(Z)[clear], ref.predict(Z)[clear])
    assert np.array_equal(svm.predict(X), finite)            # C = 1e7: training set separated
    # the corner cases of svm.py:246-264, 333-341
    assert np.all(SVM().fit(X, np.full(300, -np.inf), 5.0) == False)   # noqa: E712
    only_inf = SVM()
    only_inf.fit(X, np.full(300, -np.inf), 5.0)
    with pytest.warns(UserWarning):
        assert not only_inf.predict(Z[:10]).any()
    all_fin = SVM()
    all_fin.fit(X, y, diff_threshold=1e9)
    assert all_fin.all_finite and all_fin.predict(Z[:10]).all()
    # copies and pickles drop the device handle and still predict
    for clone in (deepcopy(svm), pickle.loads(pickle.dumps(svm))):
        assert clone._dev is None
        assert np.array_equal(clone.predict(Z[:1000]), svm.predict(Z[:1000]))


def test_regressor_with_device_classifier():
    from gpry_b200.acquisition_functions import LogExp
    bounds, X, y = problem()
    gpr = make(bounds, X, y, "SVM")
    clf = gpr.infinities_classifier
    finite_train = clf.y_finite
    assert 0 < finite_train.sum() < len(y) and gpr.n == finite_train.sum()
    plain = make(bounds, X[finite_train], y[finite_train], None)
    rng = np.random.default_rng(3)
    Xc = bounds[:, 0] + (bounds[:, 1] - bounds[:, 0]) * rng.random((30000, len(bounds)))
    X_ = gpr.preprocessing_X.transform(Xc)
    ref = SVC(C=1e7, kernel="rbf", gamma="scale").fit(clf.X_train, clf.y_finite)
    dec_ref = ref.decision_function(X_)
    clear = np.abs(dec_ref) > 1e-7
    fin = dec_ref > 0
    assert 0.02 < fin.mean() < 0.98
    # decision values through the regressor's own device state (un-transformed input)
    dec = gpr._device_state().classify(Xc)
    assert np.max(np.abs(dec - dec_ref)) < 1e-9 * max(1.0, np.abs(dec_ref).max())
    m0, s0 = plain.predict(Xc, return_std=True)
    m, s = gpr.predict(Xc, return_std=True)
    assert np.all(m[~fin & clear] == -np.inf) and np.all(s[~fin & clear] == 0)
    assert np.array_equal(m[fin & clear], m0[fin & clear])
    assert np.array_equal(s[fin & clear], s0[fin & clear])
    assert np.array_equal(gpr.predict(Xc), m)
    s_only = gpr.predict_std(Xc)
    assert np.array_equal(s_only[clear], s[clear])
    gpr.minus_inf_value = -1e300                             # read at call time
    assert np.all(gpr.predict(Xc)[~fin & clear] == -1e300)
    gpr.minus_inf_value = -np.inf
    # latency path and one-point gradient conventions (gpr.py:1153-1171)
    i_inf, i_fin = np.flatnonzero(~fin & clear)[0], np.flatnonzero(fin & clear)[0]
    out = gpr.predict(Xc[i_inf:i_inf + 1], return_std=True, return_mean_grad=True,
                      return_std_grad=True)
    assert out[0][0] == -np.inf and out[1][0] == 0 and np.all(out[2] == np.inf) \
        and np.all(out[3] == 0)
    out = gpr.predict(Xc[i_fin:i_fin + 1], return_std=True, return_mean_grad=True)
    ref_out = plain.predict(Xc[i_fin:i_fin + 1], return_std=True, return_mean_grad=True)
    assert all(np.array_equal(a, b) for a, b in zip(out, ref_out))
    mb, sb, gmb, gsb = gpr.predict_grad_batch(Xc[:300])
    sel = (~fin & clear)[:300]
    assert np.all(mb[sel] == -np.inf) and np.all(sb[sel] == 0) and np.all(gmb[sel] == np.inf) \
        and np.all(gsb[sel] == 0)
    # acquisition: masked rows are -inf, the ranking is the ranking of the masked values
    acq = LogExp(dimension=len(bounds))
    a = acq(Xc, gpr)
    a0 = acq(Xc, plain)
    assert np.all(a[~fin & clear] == -np.inf)
    assert np.array_equal(a[fin & clear], a0[fin & clear])
    Kp = 64
    ta, ti, tm, ts, tX = gpr.predict_logexp_topk(Xc, acq.zeta, Kp)
    expect = np.lexsort((np.arange(len(a)), -np.where(np.isnan(a), -np.inf, a)))[:Kp]
    assert np.array_equal(ti, expect) and np.array_equal(ta, a[expect])
    assert fin[ti].all() and np.array_equal(tX, Xc[ti])
    # copies keep working (the device binding is rebuilt lazily)
    g2 = deepcopy(gpr)
    assert g2._dev is None and np.array_equal(g2.predict(Xc[:2000]), m[:2000])
    # a refit of the classifier alone (only infinite points appended) re-binds it
    extra = bounds[:, 1] - 1e-3 * rng.random((5, len(bounds)))
    v0 = clf.version
    gpr.append_to_data(extra, np.full(5, -np.inf), fit_gpr=False)
    assert gpr.infinities_classifier.version > v0 and gpr.n == finite_train.sum()
    m2 = gpr.predict(extra)
    assert np.all(m2 == -np.inf)


def test_nora_and_sampler_with_device_classifier():
    from gpry_b200.gp_acquisition import NORA
    from gpry_b200.mc import ensemble_sample
    bounds, X, y = problem(d=2, n=120, seed=4)
    gpr = make(bounds, X, y, "SVM", threshold=8.0)
    clf = gpr.infinities_classifier
    res = ensemble_sample(gpr, n_walkers=6000, n_steps=80, seed=0, X_init="training")
    dec = gpr._device_state().classify(res.X)
    assert np.all(dec > 0) and np.all(np.isfinite(res.logp))
    nora = NORA(bounds, sampler="uniform", nsamples=40000, verbose=0)
    Xn, yn, an = nora.multi_add(gpr, n_points=3, rng=np.random.default_rng(2))
    assert np.all(np.isfinite(an)) and np.all(np.isfinite(yn))
    assert clf.predict(gpr.preprocessing_X.transform(Xn)).all()
