"""the batched-proposal ensemble sampler on the surrogate."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def gaussian_gpr(d=3, n=120, seed=0):
    """A surrogate of a Gaussian log-posterior (mean 0.5, sigma 0.08 per dim) in the unit cube."""
    from gpry_b200.gpr import GaussianProcessRegressor
    from gpry_b200.preprocessing import Normalize_bounds, Normalize_y
    rng = np.random.default_rng(seed)
    bounds = np.array([[0.0, 1.0]] * d)
    sigma = 0.08
    X = np.clip(0.5 + 2.5 * sigma * rng.standard_normal((n, d)), 0, 1)
    y = -0.5 * np.sum(((X - 0.5) / sigma) ** 2, axis=1)
    gpr = GaussianProcessRegressor(kernel="RBF", bounds=bounds, noise_level=1e-3,
                                   preprocessing_X=Normalize_bounds(bounds),
                                   preprocessing_y=Normalize_y(), n_restarts_optimizer=4,
                                   account_for_inf=None, verbose=0, random_state=1)
    gpr.append_to_data(X, y, fit_gpr=True)
    return gpr, bounds, sigma


def test_ensemble_sampler_recovers_gaussian():
    from gpry_b200.mc import ensemble_sample
    gpr, bounds, sigma = gaussian_gpr()
    d = gpr.d
    n0 = gpr.n_eval
    res = ensemble_sample(gpr, n_walkers=20000, n_steps=150, seed=5, X_init="training",
                          keep_every=50)
    assert res.X.shape == (20000, d) and res.logp.shape == (20000,)
    assert np.all(res.X >= bounds[:, 0]) and np.all(res.X <= bounds[:, 1])
    assert 0.2 < res.acceptance < 0.9
    assert res.n_eval == 20000 * (150 + 1) and gpr.n_eval == n0 + res.n_eval
    assert res.chain.shape == (3, 20000, d)
    # log-posterior values are the surrogate's mean at the final positions
    assert np.allclose(res.logp, gpr.predict(res.X), rtol=0, atol=1e-9)
    # moments of the target: the surrogate of a Gaussian is Gaussian to good accuracy
    assert np.all(np.abs(res.X.mean(axis=0) - 0.5) < 0.1 * sigma)
    assert np.all(np.abs(res.X.std(axis=0) / sigma - 1) < 0.08)
    # seeded: same seed, same sample; from a uniform start the sampler burns in to the same
    # distribution
    again = ensemble_sample(gpr, n_walkers=20000, n_steps=150, seed=5, X_init="training")
    assert np.array_equal(again.X, res.X)
    uni = ensemble_sample(gpr, n_walkers=20000, n_steps=400, seed=6)
    assert np.all(np.abs(uni.X.mean(axis=0) - 0.5) < 0.15 * sigma)
    assert np.all(np.abs(uni.X.std(axis=0) / sigma - 1) < 0.1)
    with pytest.raises(ValueError):
        ensemble_sample(gpr, n_walkers=7)


def test_ensemble_sampler_respects_trust_region_and_box():
    from gpry_b200.mc import ensemble_sample
    gpr, bounds, sigma = gaussian_gpr(d=2, n=80, seed=3)
    box = np.array([[0.5, 1.0], [0.0, 1.0]])            # half of the mode cut away
    res = ensemble_sample(gpr, bounds=box, n_walkers=8000, n_steps=200, seed=1)
    assert np.all(res.X[:, 0] >= 0.5) and np.all(res.X <= 1.0)
    # half-normal in x0: mean = 0.5 + sigma sqrt(2/pi)
    assert abs(res.X[:, 0].mean() - (0.5 + sigma * np.sqrt(2 / np.pi))) < 0.1 * sigma
    gpr.trust_bounds = np.array([[0.0, 0.5], [0.0, 1.0]])
    res = ensemble_sample(gpr, n_walkers=8000, n_steps=200, seed=2)
    assert np.all(res.X[:, 0] <= 0.5)
    res = ensemble_sample(gpr, n_walkers=8000, n_steps=200, seed=2, use_trust_region=False)
    assert (res.X[:, 0] > 0.5).mean() > 0.3


def test_nora_with_ensemble_sampler():
    from gpry_b200.gp_acquisition import NORA
    gpr, bounds, sigma = gaussian_gpr(d=3, n=60, seed=2)
    nora = NORA(bounds, sampler="ensemble", nsamples=20001, mc_steps=60, verbose=0)
    X, y, acq = nora.multi_add(gpr, n_points=3, rng=np.random.default_rng(0))
    assert X.shape == (3, 3) and np.all(np.isfinite(acq))
    assert nora._X_shard.shape == (20001, 3)
    # the pool is drawn where the surrogate posterior has mass, not uniformly in the box
    assert np.all(np.abs(nora._X_shard.mean(axis=0) - 0.5) < 0.05)
    assert np.all(nora._X_shard.std(axis=0) < 0.2)
    assert np.all(np.linalg.norm(X - 0.5, axis=1) < 0.6)


def test_nested_sampler_logp_closures():
    """The closures the reference gives its nested samplers (gp_acquisition.py:770, 784-793):
    UltraNest's is called with batches (ns_interfaces.py:448) and must not return -inf."""
    from conftest import load_golden
    from gpry_b200.gp_acquisition import NORA
    from test_gpu_gpr import make_gpr
    g = load_golden("rbf_d8_n300")
    gpr = make_gpr(g)
    lo, hi = g["bounds"][:, 0], g["bounds"][:, 1]
    gpr.trust_bounds = np.stack([lo + 0.2 * (hi - lo), hi - 0.2 * (hi - lo)], axis=1)
    X = g["Xc"]
    inside = np.all((X >= gpr.trust_bounds[:, 0]) & (X <= gpr.trust_bounds[:, 1]), axis=1)
    assert 0 < inside.sum() < len(X)
    logp = NORA.logp_function(gpr, "ultranest")
    n0 = gpr.n_eval
    batch = logp(X)
    assert gpr.n_eval == n0 + len(X) and gpr.minus_inf_value == -np.inf      # restored
    assert np.all(batch[~inside] == -1e-300) and np.all(np.isfinite(batch))
    ref = gpr.predict(X, ignore_trust_region=True)
    assert np.array_equal(batch[inside], ref[inside])
    one = NORA.logp_function(gpr, "polychord")
    i_in, i_out = np.flatnonzero(inside)[0], np.flatnonzero(~inside)[0]
    assert isinstance(one(X[i_in]), float) and abs(one(X[i_in]) - ref[i_in]) <= 1e-12 * abs(ref[i_in])
    assert one(X[i_out]) == -np.inf
    assert logp(X[i_in]).shape == (1,)
