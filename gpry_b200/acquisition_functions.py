"""
``LogExp`` acquisition function with the interface of
``gpry.acquisition_functions.LogExp`` (reference acquisition_functions.py:868-1074).

``LogExp.f`` (the static (mu, std) -> value map NORA uses, :1068-1074) is a few flops per
candidate; on pools it is evaluated inside the fused CUDA pass
(``GaussianProcessRegressor.predict_logexp[_topk]``).  ``LogExp.__call__(X, gp)`` (the
BaseLogExp value branch, :936-992) routes through that fused pass as well and then applies the
reference's validity mask.
"""
import numpy as np


class LogExp:
    """A(X) = 2 zeta (mu(X) - y_max) + log sqrt(max(sigma(X)^2 - sigma_n^2, 0))."""

    def __init__(self, zeta=None, sigma_n=None, dimension=None, zeta_scaling=0.85, fixed=False):
        self.sigma_n = sigma_n
        self.fixed = fixed
        self.zeta_scaling = zeta_scaling
        if zeta is None:
            if dimension is None:
                raise ValueError("Pass either 'zeta' or 'dimension' (for auto-scaling).")
            zeta = self.auto_zeta(dimension, zeta_scaling)
        self.zeta = zeta

    @staticmethod
    def auto_zeta(dimension, scaling=0.85):
        """acquisition_functions.py:933-934."""
        return dimension ** (-scaling)

    @staticmethod
    def f(mu, std, baseline, noise_level, zeta):
        """acquisition_functions.py:1068-1074 (host, elementwise; used on the few survivors
        the GPU returns and on conditioned sigmas inside the ranked pool)."""
        with np.errstate(divide="ignore", invalid="ignore"):
            return (2 * zeta * (mu - baseline) +
                    np.log(np.sqrt(np.clip(std ** 2. - noise_level ** 2., 0., None))))

    def noise_var(self, gp):
        """acquisition_functions.py:974-981."""
        if self.sigma_n is None:
            sigma_n = gp.noise_level
            return float(np.mean(sigma_n)) if np.iterable(sigma_n) else sigma_n
        return self.sigma_n

    def __call__(self, X, gp, eval_gradient=False):
        """acquisition_functions.py:936-992: value; ``-inf`` where sigma^2 - sigma_n^2 <= 0 or
        the mean is not finite (classifier / trust-region rows)."""
        X = np.atleast_2d(np.asarray(X, dtype=float))
        noise_var = self.noise_var(gp)
        if eval_gradient:   # acquisition_functions.py:966-969, 993-1007 (one point)
            if X.shape[0] > 1:    # many points in one device pass (the ndim > 1 branch below)
                mu, std, mu_grad, std_grad = gp.predict_grad_batch(X)
            else:
                mu, std, mu_grad, std_grad = gp.predict(
                    X, return_std=True, return_mean_grad=True, return_std_grad=True)
            var = std ** 2 - noise_var ** 2.
            mask = (var > 0) & np.isfinite(mu)
            values = np.where(mask, self.f(mu, std, gp.y_max, noise_var, self.zeta), -np.inf)
            if np.array(std_grad).ndim > 1:
                # Row by row the rule of the one-point branch (:1002-1007): a finite gradient
                # wherever std > sigma_n -- also for a row whose MEAN is masked (outside the
                # trust region: value -inf, gradient still defined), so that a restart sees the
                # same numbers whether it is evaluated alone or in a batch.  (The reference's
                # own many-point branch, :996-1001, is unreachable there and lacks this
                # broadcast over the dimensions.)
                std_grad, mu_grad = np.array(std_grad), np.array(mu_grad)
                fin = std > noise_var
                grad = np.full_like(std_grad, np.inf)
                if np.any(fin):
                    grad[fin] = std_grad[fin] / (std[fin] - noise_var)[:, None] \
                        + 2 * self.zeta * mu_grad[fin]
            elif std[0] > noise_var:
                grad = std_grad / (std[0] - noise_var) + 2 * self.zeta * mu_grad
            else:
                grad = np.ones_like(std_grad) * np.inf
            return values, grad
        # one fused device pass; classifier / trust-region rows come back with a non-finite mean
        mu, std, values = gp.predict_logexp(X, self.zeta, noise_var)
        var = std ** 2 - noise_var ** 2.
        mask = (var > 0) & np.isfinite(mu)
        values = np.where(mask, values, -np.inf)
        return values

    def __repr__(self):
        return str(self.__class__) + "with zeta={0:.3f}".format(self.zeta)
