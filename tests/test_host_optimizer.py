"""Host-side pieces of the acquisition optimiser: proposers and the lock-step L-BFGS-B."""
import numpy as np
import pytest
import scipy.optimize


class _FakeGpr:
    def __init__(self, X):
        self.X_train = X


def test_proposers_stay_in_bounds_and_are_seeded():
    from gpry_b200.proposal import UniformProposer, CentroidsProposer, PartialProposer
    d = 4
    bounds = np.array([[-1.0, 2.0]] * d)
    rng = np.random.default_rng(0)
    X_train = rng.uniform(-1, 2, size=(30, d))
    cen = CentroidsProposer(bounds)
    with pytest.raises(ValueError):
        cen.get(rng=rng)                                # update(gpr) not called yet
    cen.update(_FakeGpr(X_train))
    prop = PartialProposer(bounds, cen, random_proposal_fraction=0.25)
    for p in (UniformProposer(bounds), cen, prop):
        x = p.get(rng=np.random.default_rng(1))
        assert x.shape == (d,)
        B = p.get_batch(500, rng=np.random.default_rng(2))
        assert B.shape == (500, d)
        assert np.all(B >= bounds[:, 0]) and np.all(B <= bounds[:, 1])
        assert np.array_equal(B, p.get_batch(500, rng=np.random.default_rng(2)))
    # centroid proposals concentrate around the training cloud: closer on average to the
    # training mean than uniform draws
    far = np.linalg.norm(UniformProposer(bounds).get_batch(2000, rng=rng) - X_train.mean(0), axis=1)
    near = np.linalg.norm(cen.get_batch(2000, rng=rng) - X_train.mean(0), axis=1)
    assert np.median(near) < np.median(far)
    # shrinking the bounds keeps proposals inside the new box, falling back to all training
    # points when fewer than d + 1 lie inside (proposal.py:289-296)
    small = np.array([[0.0, 0.1]] * d)
    cen.update_bounds(small)
    B = cen.get_batch(100, rng=rng)
    assert np.all(B >= small[:, 0]) and np.all(B <= small[:, 1])
    with pytest.raises(ValueError):
        PartialProposer(bounds, cen, random_proposal_fraction=1.5)
    with pytest.raises(ValueError):
        PartialProposer(bounds, "not a proposer")


def test_lockstep_minimize_equals_independent_runs():
    from gpry_b200.lockstep import lockstep_minimize
    rng = np.random.default_rng(4)
    A = rng.normal(size=(5, 5))
    A = A @ A.T + 5 * np.eye(5)
    b = rng.normal(size=5)
    calls = []

    def batch(X):
        calls.append(len(X))
        vals = 0.5 * np.einsum("ni,ij,nj->n", X, A, X) - X @ b + np.sum(np.cos(X), axis=1)
        grads = X @ A - b - np.sin(X)
        return vals, grads

    x0s = rng.uniform(-2, 2, size=(7, 5))
    bounds = [(-2.0, 2.0)] * 5
    out = lockstep_minimize(batch, x0s, bounds)
    assert len(out) == 7 and max(calls) == 7             # evaluations were shared
    for i, (x, f) in enumerate(out):
        res = scipy.optimize.minimize(lambda t: tuple(v[0] for v in batch(t[None])), x0s[i],
                                      method="L-BFGS-B", jac=True, bounds=bounds)
        assert np.allclose(x, res.x, atol=1e-10) and abs(f - res.fun) < 1e-12
    # an exception in the batched objective reaches the caller instead of dead-locking
    def broken(X):
        raise FloatingPointError("boom")
    with pytest.raises(FloatingPointError):
        lockstep_minimize(broken, x0s, bounds)


def test_number_times_d():
    from gpry_b200.gp_acquisition import _number_times_d
    assert _number_times_d("5d", 4, "n") == 20
    assert _number_times_d("d", 4, "n") == 4
    assert _number_times_d(7, 4, "n") == 7
    with pytest.raises(ValueError):
        _number_times_d("many", 4, "n")
