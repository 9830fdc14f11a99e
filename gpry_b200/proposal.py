"""
Start-point proposers of the acquisition optimiser (reference: gpry/proposal.py).

Only what ``BatchOptimizer`` uses by default is mirrored: ``UniformProposer`` (proposal.py:
136-160), ``CentroidsProposer`` (:258-322) and their mixture ``PartialProposer`` (:163-215).
Besides the reference's one-point ``get`` every proposer has ``get_batch(n)``: the GPU scores
proposals by the thousand, so they are drawn by the thousand.
"""
import numpy as np

from .gpr import is_in_bounds


def _as_generator(rng):
    if isinstance(rng, np.random.Generator):
        return rng
    if rng is None:
        return np.random.default_rng()
    if isinstance(rng, np.random.RandomState):
        return rng
    return np.random.default_rng(rng)


class Proposer:
    """proposal.py:45-88."""

    def get(self, rng=None):
        raise NotImplementedError

    def get_batch(self, n, rng=None):
        rng = _as_generator(rng)
        return np.array([self.get(rng=rng) for _ in range(n)])

    def update_bounds(self, bounds):
        self.bounds = np.asarray(bounds, dtype=float)

    def update(self, gpr):
        pass


class UniformProposer(Proposer):
    """Uniform in the hypercube of the bounds (proposal.py:136-160)."""

    def __init__(self, bounds):
        self.update_bounds(bounds)

    def get(self, rng=None):
        rng = _as_generator(rng)
        return rng.uniform(self.bounds[:, 0], self.bounds[:, 1])

    def get_batch(self, n, rng=None):
        rng = _as_generator(rng)
        return rng.uniform(self.bounds[:, 0], self.bounds[:, 1], size=(n, len(self.bounds)))


class CentroidsProposer(Proposer):
    """Centroid of d+1 random training points, kicked per dimension towards one of them by an
    exponentially distributed factor (scale 1/lambd), clipped to the bounds
    (proposal.py:258-322)."""

    def __init__(self, bounds, lambd=1.0):
        self.training = None
        self.training_ = None
        self.lambd = lambd
        self.update_bounds(bounds)

    @property
    def d(self):
        return len(self.bounds)

    def _source(self, m):
        if self.training is None:
            raise ValueError("CentroidsProposer.update(gpr) has not been called")
        if self.training_ is not None and len(self.training_) >= m:
            return self.training_
        if len(self.training) < m:
            raise ValueError(f"need at least {m} training points, got {len(self.training)}")
        return self.training

    def get(self, rng=None):
        rng = _as_generator(rng)
        m = self.d + 1
        src = self._source(m)
        subset = src[rng.choice(len(src), size=m, replace=False)]
        centroid = np.average(subset, axis=0)
        towards = rng.choice(m, size=self.d, replace=False)
        kick = subset[towards, np.arange(self.d)] - centroid
        kick = kick * rng.exponential(scale=1 / self.lambd, size=self.d)
        return np.clip(centroid + kick, self.bounds[:, 0], self.bounds[:, 1])

    def get_batch(self, n, rng=None):
        rng = _as_generator(rng)
        d, m = self.d, self.d + 1
        src = self._source(m)
        # n subsets of m distinct rows: argsort of uniform keys = sampling without replacement
        pick = np.argsort(rng.random((n, len(src))), axis=1)[:, :m]
        subsets = src[pick]                                        # (n, m, d)
        centroid = subsets.mean(axis=1)
        towards = np.argsort(rng.random((n, m)), axis=1)[:, :d]    # (n, d), distinct per row
        kick = subsets[np.arange(n)[:, None], towards, np.arange(d)[None, :]] - centroid
        kick = kick * rng.exponential(scale=1 / self.lambd, size=(n, d))
        return np.clip(centroid + kick, self.bounds[:, 0], self.bounds[:, 1])

    def update(self, gpr):
        self.training = np.copy(gpr.X_train)
        self.update_bounds(self.bounds)

    def update_bounds(self, bounds):
        super().update_bounds(bounds)
        if self.training is not None:
            self.training_ = self.training[is_in_bounds(self.training, self.bounds)]


class PartialProposer(Proposer):
    """``true_proposer`` with a fraction of uniform draws mixed in (proposal.py:163-215)."""

    def __init__(self, bounds, true_proposer, random_proposal_fraction=0.25):
        if random_proposal_fraction > 1.0 or random_proposal_fraction < 0.0:
            raise ValueError("Cannot pass a fraction outside of [0,1]. You passed "
                             f"'random_proposal_fraction={random_proposal_fraction}'")
        if not isinstance(true_proposer, Proposer):
            raise ValueError("The true proposer needs to be a valid proposer.")
        self.rpf = random_proposal_fraction
        self.random_proposer = UniformProposer(bounds)
        self.true_proposer = true_proposer
        self.bounds = np.asarray(bounds, dtype=float)

    def get(self, rng=None):
        rng = _as_generator(rng)
        if rng.random() > self.rpf:
            return self.true_proposer.get(rng=rng)
        return self.random_proposer.get(rng=rng)

    def get_batch(self, n, rng=None):
        rng = _as_generator(rng)
        use_true = rng.random(n) > self.rpf
        out = self.random_proposer.get_batch(n, rng=rng)
        if use_true.any():
            out[use_true] = self.true_proposer.get_batch(int(use_true.sum()), rng=rng)
        return out

    def update(self, gpr):
        self.true_proposer.update(gpr)

    def update_bounds(self, bounds):
        self.bounds = np.asarray(bounds, dtype=float)
        self.random_proposer.update_bounds(bounds)
        self.true_proposer.update_bounds(bounds)
