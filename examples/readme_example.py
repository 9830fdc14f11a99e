"""
The reference's README example (examples/readme_example.py: a curved 2-D log-likelihood, NORA
acquisition) with the surrogate numerics on a B200.  The `Runner` orchestration, the nested
sampler and the final MC sample are outside this repository's scope, so
the loop is written out: NORA with its uniform test sampler -> evaluate the truth -> append
and refit the hyper-parameters every few iterations.

    python examples/readme_example.py            # needs a B200 and the built library
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpry_b200.acquisition_functions import LogExp           # noqa: E402
from gpry_b200.gp_acquisition import NORA                     # noqa: E402
from gpry_b200.gpr import GaussianProcessRegressor            # noqa: E402
from gpry_b200.preprocessing import Normalize_bounds, Normalize_y   # noqa: E402


def loglike(X):
    X = np.atleast_2d(X)
    return -0.5 * (X[:, 0] ** 2 / 1.5 + (X[:, 1] - 0.5 * X[:, 0] ** 2) ** 2 / 0.5)


def main(n_iter=12, n_points=2, seed=1):
    bounds = np.array([[-4.0, 4.0], [-2.0, 6.0]])
    d = bounds.shape[0]
    rng = np.random.default_rng(seed)
    gpr = GaussianProcessRegressor(kernel="RBF", bounds=bounds, noise_level=1e-2,
                                   n_restarts_optimizer=10 + 2 * d, random_state=seed,
                                   preprocessing_X=Normalize_bounds(bounds),
                                   preprocessing_y=Normalize_y(), account_for_inf=None, verbose=0)
    X0 = rng.uniform(bounds[:, 0], bounds[:, 1], size=(4 * d, d))
    gpr.append_to_data(X0, loglike(X0), fit_gpr=True)
    nora = NORA(bounds, acq_func=LogExp(dimension=d), nsamples=20000, kprime=128)
    for it in range(n_iter):
        X_new, y_lie, acq = nora.multi_add(gpr, n_points=n_points, rng=rng, force_resample=True)
        y_new = loglike(X_new)
        gpr.append_to_data(X_new, y_new, fit_gpr=(it % 3 == 2))
        print(f"iter {it:2d}: N={gpr.n:3d}  max|y_lie - y_true|={np.max(np.abs(y_lie - y_new)):.3g} "
              f"n_eval={gpr.n_eval} n_eval_loglike={gpr.n_eval_loglike}")
    # quality of the surrogate where the posterior mass is
    Xt = rng.normal(size=(4000, d)) * [np.sqrt(1.5), 1.0]
    Xt[:, 1] = 0.5 * Xt[:, 0] ** 2 + np.sqrt(0.5) * rng.normal(size=4000)
    Xt = Xt[np.all((Xt >= bounds[:, 0]) & (Xt <= bounds[:, 1]), axis=1)]
    err = np.abs(gpr.predict(Xt) - loglike(Xt))
    print(f"surrogate error on {len(Xt)} posterior draws: median {np.median(err):.3g}, "
          f"95% {np.quantile(err, 0.95):.3g}")
    return gpr, err


if __name__ == "__main__":
    main()
