"""
Batched-proposal surrogate sampler.

The reference samples the surrogate posterior with one-point ``gpr.predict`` calls from an
external MCMC / nested sampler (mc.py:96-100, 387-391; gp_acquisition.py:770-793): ~1 ms of
host work per proposal.  The device path scores 10^8 mean-only proposals per second, but only
if they arrive by the 10^5..10^7 -- so the front-end has to be batched.  This module is that
front-end: an affine-invariant ensemble sampler (Goodman & Weare 2010 "stretch move", the
algorithm of emcee) whose W walkers live on the GPU.  One half of the ensemble is updated at a
time from the other half, so each half-step is ONE ``gpry_predict`` call over W/2 rows (device
pointers in and out, nothing crosses PCIe), followed by an element-wise accept/reject.  torch
is used for the random numbers and the element-wise bookkeeping only.

The target is what the reference's samplers see: ``exp(gpr.predict(x))`` inside the prior box
(optionally the trust region), zero outside; the stretch move needs no proposal tuning.
"""
import numpy as np


class EnsembleResult:
    """Walker positions (W, d) and surrogate log-posterior (W,) after the last step, plus
    the acceptance rate and the number of surrogate evaluations."""

    def __init__(self, X, logp, acceptance, n_eval, chain=None, chain_logp=None):
        self.X = X
        self.logp = logp
        self.acceptance = acceptance
        self.n_eval = n_eval
        self.chain = chain
        self.chain_logp = chain_logp


def _predict_mean_device(gpr, dev, X):
    """Surrogate log-posterior for a CUDA tensor of points (mean only; rows the device-side
    infinities classifier rejects come back as ``minus_inf_value``)."""
    gpr.n_eval += int(X.shape[0])
    mean, _ = dev.predict(X, return_mean=True, return_std=False)
    return mean


def ensemble_sample(gpr, bounds=None, n_walkers=None, n_steps=200, stretch=2.0, seed=None,
                    X_init=None, keep_every=0, temperature=1.0, use_trust_region=True):
    """Samples ``exp(gpr.predict(x) / temperature)`` inside ``bounds`` (default: the
    regressor's prior bounds, intersected with its trust region if it has one and
    ``use_trust_region``).

    n_walkers : even, default ``max(1000 d, 4096)``; n_steps : full ensemble updates;
    X_init : optional (>= n_walkers, d) starting positions; "training" = training points drawn
    with weights exp(y - y_max) plus a jitter of 1e-3 box widths (short burn-in when the mode
    is much smaller than the box); default: uniform in the box;
    keep_every : if > 0, every ``keep_every``-th ensemble is kept (host numpy) in
    ``result.chain`` / ``result.chain_logp``.
    Returns an ``EnsembleResult`` with numpy arrays.
    """
    import torch
    if gpr.infinities_classifier is not None and not gpr._classifier_on_device():
        raise NotImplementedError("the batched sampler keeps its walkers on the device; use "
                                  "account_for_inf='SVM' (gpry_b200.svm, evaluated on the "
                                  "device) instead of a host-side classifier")
    d = gpr.d
    box = np.array(gpr.bounds if bounds is None else bounds, dtype=float)
    if use_trust_region and gpr.trust_bounds is not None:
        box[:, 0] = np.maximum(box[:, 0], gpr.trust_bounds[:, 0])
        box[:, 1] = np.minimum(box[:, 1], gpr.trust_bounds[:, 1])
    W = max(1000 * d, 4096) if n_walkers is None else int(n_walkers)
    if W % 2 or W < 2 * (d + 1):
        raise ValueError("n_walkers must be even and at least 2 (d + 1)")
    if stretch <= 1:
        raise ValueError("stretch must be > 1")
    dev = gpr._device_state()
    gpr._set_masks(dev, trust=False)    # the box below already contains the trust region
    device = torch.device("cuda", gpr.device)
    gen = torch.Generator(device=device)
    gen.manual_seed(int(np.random.SeedSequence(seed).generate_state(1)[0]))
    lo = torch.as_tensor(box[:, 0], device=device)
    hi = torch.as_tensor(box[:, 1], device=device)
    if X_init is None:
        X = lo + (hi - lo) * torch.rand((W, d), dtype=torch.float64, device=device, generator=gen)
    elif isinstance(X_init, str):
        if X_init != "training":
            raise ValueError(f"unknown X_init {X_init!r}")
        rng = np.random.default_rng(seed)
        Xt = np.asarray(gpr.X_train, dtype=float)
        ok = np.all((Xt >= box[:, 0]) & (Xt <= box[:, 1]), axis=1)
        if not ok.any():
            raise ValueError("no training point inside the sampling box")
        w = np.exp(np.asarray(gpr.y_train)[ok] - np.max(np.asarray(gpr.y_train)[ok]))
        pick = rng.choice(np.flatnonzero(ok), size=W, p=w / w.sum())
        X0 = Xt[pick] + 1e-3 * (box[:, 1] - box[:, 0]) * rng.standard_normal((W, d))
        X = torch.as_tensor(np.clip(X0, box[:, 0], box[:, 1]), device=device)
    else:
        X_init = np.asarray(X_init, dtype=float)
        if X_init.ndim != 2 or X_init.shape[1] != d or len(X_init) < W:
            raise ValueError(f"X_init must be (>= {W}, {d})")
        X = torch.as_tensor(np.ascontiguousarray(X_init[:W]), device=device)
    X = X.contiguous()
    logp = _predict_mean_device(gpr, dev, X) / temperature
    inside = ((X >= lo) & (X <= hi)).all(dim=1)
    logp = torch.where(inside, logp, torch.full_like(logp, -float("inf")))
    n_eval = W
    half = W // 2
    accepted = torch.zeros((), dtype=torch.float64, device=device)
    chain, chain_logp = [], []
    a = float(stretch)
    for step in range(n_steps):
        for first in (0, half):
            sl = slice(first, first + half)
            other = slice(half - first, W - first)          # the complementary half
            Xa, Xb = X[sl], X[other]
            partner = torch.randint(0, half, (half,), device=device, generator=gen)
            u = torch.rand(half, dtype=torch.float64, device=device, generator=gen)
            z = ((a - 1.0) * u + 1.0) ** 2 / a                # g(z) ~ 1/sqrt(z) on [1/a, a]
            Xp = Xb[partner]
            Y = (Xp + z[:, None] * (Xa - Xp)).contiguous()
            ok = ((Y >= lo) & (Y <= hi)).all(dim=1)
            lp_new = _predict_mean_device(gpr, dev, Y) / temperature
            n_eval += half
            lp_new = torch.where(ok, lp_new, torch.full_like(lp_new, -float("inf")))
            log_ratio = (d - 1) * torch.log(z) + lp_new - logp[sl]
            log_ratio = torch.where(torch.isnan(log_ratio),
                                    torch.full_like(log_ratio, -float("inf")), log_ratio)
            lu = torch.log(torch.rand(half, dtype=torch.float64, device=device, generator=gen))
            acc = lu < log_ratio
            X[sl] = torch.where(acc[:, None], Y, Xa)
            logp[sl] = torch.where(acc, lp_new, logp[sl])
            accepted += acc.sum()
        if keep_every and (step + 1) % keep_every == 0:
            chain.append(X.cpu().numpy())
            chain_logp.append((logp * temperature).cpu().numpy())
    acceptance = float(accepted.item()) / max(1, n_steps * W)
    return EnsembleResult(X.cpu().numpy(), (logp * temperature).cpu().numpy(), acceptance, n_eval,
                          np.array(chain) if chain else None,
                          np.array(chain_logp) if chain_logp else None)
