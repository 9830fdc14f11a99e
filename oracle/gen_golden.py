"""
Golden-vector generator: runs the REAL reference (GPry 3.0.0 imported from /root/reference
through ``oracle/ref_import.py``) and writes small ``tests/golden/*.npz`` fixtures.

Run in the build container only (``python oracle/gen_golden.py``); the fixtures are committed
and travel to the GPU box, the reference does not.  Each fixture stores the inputs needed to
rebuild the case (or the seed that regenerates them with ``numpy.random.default_rng``) and
the reference's outputs at the drop-in boundary:

  predict(return_std)            gpr.py:1022-1273      -> mean, std
  predict(return_mean_grad, ..)  gpr.py:1236-1266      -> grad_mean, grad_std (1 point)
  predict_std                    gpr.py:1275-1352
  LogExp.f / LogExp.__call__     acquisition_functions.py:936-1009,1068-1074
  log_marginal_likelihood        gpr.py:876-881 -> sklearn _gpr.py:541-656   (value, gradient)
  L_, V_, alpha_                 gpr.py:1453-1465      (checksums + a few rows)
  RankedPool.add                 gp_acquisition.py:1290-1670 (final pool indices, X, y, acq)
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle.ref_import import import_reference  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def target(X):
    return -0.5 * np.sum(((X - 0.5) / 0.15) ** 2, axis=1)


def make_reference_gpr(gpry, kind, X, y, theta, bounds, noise_level=1e-2, normalize=True,
                       clip_factor=1.1):
    """Fixed-theta reference GPR."""
    from sklearn.base import clone
    from gpry.preprocessing import Normalize_bounds, Normalize_y
    kernel = {"rbf": "RBF", "matern15": {"Matern": {"nu": 1.5}},
              "matern25": {"Matern": {"nu": 2.5}}}[kind]
    gpr = gpry.gpr.GaussianProcessRegressor(
        kernel=kernel, bounds=bounds, noise_level=noise_level, clip_factor=clip_factor,
        preprocessing_X=Normalize_bounds(bounds) if normalize else None,
        preprocessing_y=Normalize_y() if normalize else None,
        account_for_inf=None, verbose=0)
    gpr.kernel_ = clone(gpr.kernel)
    gpr.kernel_.theta = np.asarray(theta)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        gpr.append_to_data(X, y, fit_gpr=False)
    return gpr


def case(gpry, name, kind, N, d, M, seed, ell, c=1.0, bounds=None, normalize=True,
         with_lml=True, pool=None, noise_level=1e-2, store_train=True, clip_factor=1.1):
    rng = np.random.default_rng(seed)
    if bounds is None:
        bounds = np.array([[0.0, 1.0]] * d)
    lo, hi = bounds[:, 0], bounds[:, 1]
    U = rng.uniform(size=(N, d))
    X = lo + U * (hi - lo)
    y = target(U)
    Xc = lo + rng.uniform(size=(M, d)) * (hi - lo)
    theta = np.log(np.concatenate([[c], np.full(d, ell) * (1 + 0.1 * np.arange(d) / d)]))
    gpr = make_reference_gpr(gpry, kind, X, y, theta, bounds, noise_level, normalize,
                             clip_factor=clip_factor)
    mean, std = gpr.predict(Xc, return_std=True, validate=False)
    std_only = gpr.predict_std(Xc, validate=False)
    mean_only = gpr.predict(Xc, validate=False)
    zeta = d ** (-0.85)
    LogExp = gpry.acquisition_functions.LogExp
    with np.errstate(divide="ignore"):
        acq_f = LogExp.f(mean, std, gpr.y_max, gpr.noise_level, zeta)
    acq_call = LogExp(zeta=zeta)(Xc, gpr)
    m1, s1, gm, gs = gpr.predict(Xc[:1], return_std=True, return_mean_grad=True,
                                 return_std_grad=True, validate=False)
    out = dict(
        kind=kind, N=N, d=d, M=M, seed=seed, theta=theta, bounds=bounds,
        noise_level=noise_level, normalize=normalize, zeta=zeta, clip_factor=clip_factor,
        n_clipped=int(np.sum(mean >= clip_factor * max(y) - (clip_factor - 1) * min(y))),
        Xc=Xc, mean=mean, std=std, std_only=std_only, mean_only=mean_only,
        acq_f=acq_f, acq_call=acq_call, grad_mean=gm, grad_std=gs,
        y_mean=getattr(gpr.preprocessing_y, "mean_", 0.0),
        y_std=getattr(gpr.preprocessing_y, "std_", 1.0),
        y_max=gpr.y_max, alpha_=gpr.alpha_, V_rows=gpr.V_[[0, N // 2, N - 1]],
        L_diag=np.diag(gpr.L_), V_fro=np.linalg.norm(gpr.V_),
        condK=np.linalg.cond(gpr.L_) ** 2,
    )
    if store_train:
        out.update(X_train=X, y_train=y)
    if with_lml:
        thetas = [theta, theta + 0.3 * rng.standard_normal(theta.shape)]
        lml, grad = [], []
        for th in thetas:
            v, g = gpr.log_marginal_likelihood(th, eval_gradient=True, clone_kernel=True)
            lml.append(v), grad.append(g)
        out.update(lml_thetas=np.array(thetas), lml=np.array(lml), lml_grad=np.array(grad))
    if pool is not None:
        from functools import partial
        Mp, n_points = pool
        Xp = lo + np.random.default_rng(seed + 1).uniform(size=(Mp, d)) * (hi - lo)
        yp, sp = gpr.predict(Xp, return_std=True, validate=False)
        acq_func = partial(LogExp.f, baseline=gpr.y_max, noise_level=gpr.noise_level,
                           zeta=zeta)
        with np.errstate(divide="ignore"):
            ap = acq_func(yp, sp)
            for method in ("single sort acq", "bulk"):
                rp = gpry.gp_acquisition.RankedPool(n_points, gpr=gpr, acq_func=acq_func,
                                                    verbose=0)
                rp.add(Xp, yp, sp, ap, method=method)
                rp = rp.copy(drop_empty=True)
                Xsel = rp.X[:n_points]
                idx = np.array([int(np.flatnonzero(np.all(Xp == x, axis=1))[0])
                                for x in Xsel])
                tag = method.replace(" ", "_")
                out[f"pool_idx_{tag}"] = idx
                out[f"pool_y_{tag}"] = rp.y[:n_points]
                out[f"pool_sigma_{tag}"] = rp.sigma[:n_points]
                out[f"pool_acq_cond_{tag}"] = rp.acq_cond[:n_points]
        out.update(pool_M=Mp, pool_n_points=n_points, pool_seed=seed + 1,
                   pool_acq_top=np.sort(ap)[::-1][:64],
                   pool_argsort_top=np.argsort(ap)[::-1][:64])
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(f"{name}: N={N} d={d} M={M} kind={kind} cond(K)={out['condK']:.3g} "
          f"std range [{std.min():.3g},{std.max():.3g}] finite acq "
          f"{np.isfinite(acq_f).mean():.3f}")


def nonpd_case(gpry):
    """LML with a non-PD kernel matrix must return (-inf, 0)  (sklearn:_gpr.py:590-593)."""
    rng = np.random.default_rng(5)
    d, N = 2, 40
    X = rng.uniform(size=(N, d))
    X[1] = X[0]  # duplicate point, zero noise -> singular
    y = target(X)
    theta = np.log([1.0, 5.0, 5.0])
    bounds = np.array([[0.0, 1.0]] * d)
    gpr = make_reference_gpr(gpry, "rbf", X[2:], y[2:], theta, bounds, noise_level=1e-2)
    gpr.X_train_ = X.copy()
    gpr.y_train_ = (y - y.mean()) / y.std()
    gpr.alpha = np.zeros(N)
    v, g = gpr.log_marginal_likelihood(theta, eval_gradient=True, clone_kernel=True)
    np.savez_compressed(os.path.join(OUT, "lml_nonpd.npz"), X_train_=gpr.X_train_,
                        y_train_=gpr.y_train_, noise2=gpr.alpha, theta=theta, lml=v, grad=g)
    print("lml_nonpd:", v, g)


def fit_case(gpry):
    """Reference hyper-parameter fit (gpr.py:883-994): 4 restarts from a seeded RandomState."""
    from gpry.preprocessing import Normalize_bounds, Normalize_y
    rng = np.random.default_rng(21)
    d, N = 2, 40
    bounds = np.array([[0.0, 1.0]] * d)
    X = rng.uniform(size=(N, d))
    y = target(X)
    import warnings
    gpr = gpry.gpr.GaussianProcessRegressor(
        kernel="RBF", bounds=bounds, noise_level=1e-2, n_restarts_optimizer=4,
        preprocessing_X=Normalize_bounds(bounds), preprocessing_y=Normalize_y(),
        account_for_inf=None, random_state=7, verbose=0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        gpr.append_to_data(X, y, fit_gpr=True)
    Xc = rng.uniform(size=(64, d))
    mean, std = gpr.predict(Xc, return_std=True, validate=False)
    np.savez_compressed(os.path.join(OUT, "fit_rbf_d2_n40.npz"), X_train=X, y_train=y,
                        bounds=bounds, theta_opt=gpr.kernel_.theta,
                        lml_opt=gpr.log_marginal_likelihood_value_,
                        n_eval_loglike=gpr.n_eval_loglike, Xc=Xc, mean=mean, std=std,
                        kernel_bounds=gpr.kernel_.bounds, theta_init=gpr.kernel.theta)
    print("fit_rbf_d2_n40: theta_opt", gpr.kernel_.theta, "lml", gpr.log_marginal_likelihood_value_,
          "n_eval_loglike", gpr.n_eval_loglike)


def fit_case_d8(gpry):
    """A second reference fit (gpr.py:883-994): d = 8, N = 300, 6 restarts (the first from the
    kernel's initial theta is not available to an unfitted regressor: all six are prior draws)."""
    from gpry.preprocessing import Normalize_bounds, Normalize_y
    import warnings
    rng = np.random.default_rng(22)
    d, N, n_restarts = 8, 300, 6
    bounds = np.array([[0.0, 1.0]] * d)
    X = rng.uniform(size=(N, d))
    y = target(X)
    gpr = gpry.gpr.GaussianProcessRegressor(
        kernel="RBF", bounds=bounds, noise_level=1e-2, n_restarts_optimizer=n_restarts,
        preprocessing_X=Normalize_bounds(bounds), preprocessing_y=Normalize_y(),
        account_for_inf=None, random_state=7, verbose=0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        gpr.append_to_data(X, y, fit_gpr=True)
    Xc = rng.uniform(size=(64, d))
    mean, std = gpr.predict(Xc, return_std=True, validate=False)
    np.savez_compressed(os.path.join(OUT, "fit_rbf_d8_n300.npz"), X_train=X, y_train=y,
                        bounds=bounds, theta_opt=gpr.kernel_.theta, n_restarts=n_restarts,
                        lml_opt=gpr.log_marginal_likelihood_value_,
                        n_eval_loglike=gpr.n_eval_loglike, Xc=Xc, mean=mean, std=std,
                        kernel_bounds=gpr.kernel_.bounds, theta_init=gpr.kernel.theta)
    print("fit_rbf_d8_n300: theta_opt", gpr.kernel_.theta, "lml",
          gpr.log_marginal_likelihood_value_, "n_eval_loglike", gpr.n_eval_loglike)


def loop_case(gpry):
    """BASELINE config #1 in miniature (examples/readme_example.py): an active-learning loop
    on a 2-D curved log-likelihood with the reference's own pieces -- predict, LogExp.f,
    RankedPool(method="single sort acq"), append_to_data -- at fixed hyper-parameters, with the
    MC sample replaced by seeded uniform draws (NORA's test sampler).  Stores the sequence of
    acquired points: given identical MC samples the acquisitions must be identical."""
    from functools import partial
    import warnings
    d, n_points, n_iter, n_mc = 2, 2, 5, 3000
    bounds = np.array([[-4.0, 4.0], [-2.0, 6.0]])

    def loglike(X):   # curved ("banana") Gaussian
        return -0.5 * (X[:, 0] ** 2 / 1.5 + (X[:, 1] - 0.5 * X[:, 0] ** 2) ** 2 / 0.5)

    rng = np.random.default_rng(2024)
    X0 = rng.uniform(bounds[:, 0], bounds[:, 1], size=(12, d))
    theta = np.log([25.0, 0.25, 0.2])
    gpr = make_reference_gpr(gpry, "rbf", X0, loglike(X0), theta, bounds, noise_level=1e-2)
    LogExp = gpry.acquisition_functions.LogExp
    zeta = d ** (-0.85)
    acquired, y_lies, acqs = [], [], []
    for it in range(n_iter):
        X_mc = rng.uniform(bounds[:, 0], bounds[:, 1], size=(n_mc, d))
        y_mc, s_mc = gpr.predict(X_mc, return_std=True, validate=False)
        acq_func = partial(LogExp.f, baseline=gpr.y_max, noise_level=gpr.noise_level, zeta=zeta)
        with np.errstate(divide="ignore"):
            a_mc = acq_func(y_mc, s_mc)
            pool = gpry.gp_acquisition.RankedPool(n_points, gpr=gpr, acq_func=acq_func, verbose=0)
            pool.add(X_mc, y_mc, s_mc, a_mc, method="single sort acq")
            pool = pool.copy(drop_empty=True)
            X_new, y_lie = pool.X[:n_points].copy(), pool.y[:n_points].copy()
            acq_new = acq_func(y_lie, pool.sigma[:n_points])
        acquired.append(X_new), y_lies.append(y_lie), acqs.append(acq_new)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            gpr.append_to_data(X_new, loglike(X_new), fit_gpr=False)
    Xc = rng.uniform(bounds[:, 0], bounds[:, 1], size=(64, d))
    mean, std = gpr.predict(Xc, return_std=True, validate=False)
    np.savez_compressed(os.path.join(OUT, "loop_banana_d2.npz"), bounds=bounds, X0=X0, theta=theta,
                        seed=2024, n_points=n_points, n_iter=n_iter, n_mc=n_mc, zeta=zeta,
                        acquired=np.array(acquired), y_lies=np.array(y_lies),
                        acqs=np.array(acqs), Xc=Xc, mean=mean, std=std, n_train=gpr.n)
    print("loop_banana_d2: acquired", np.array(acquired).shape, "final N", gpr.n)


def bench_problem(N, d, seed=1234):
    """The synthetic workload of bench.py / SURVEY.md 8(d) (same as oracle.synthetic_problem)."""
    rng = np.random.default_rng(seed)
    X = rng.uniform(size=(N, d))
    y = target(X)
    ell = 0.5 if d <= 8 else (1.0 if d <= 16 else 1.5)
    theta = np.log(np.concatenate([[1.0], np.full(d, ell)]))
    return X, y, theta, np.array([[0.0, 1.0]] * d)


def config_cases(gpry, which=("C", "D", "E")):
    """Reference outputs AT THE SIZES of BASELINE.json configs[2], [3], [4] on bench.py's own
    synthetic workload.  Only seeds + outputs are stored (inputs are regenerated)."""
    from functools import partial
    LogExp = gpry.acquisition_functions.LogExp
    if "C" in which:   # N_train = 2000, d = 12, RBF: scores of 20000 candidates + ranked pool
        N, d, M, n_points = 2000, 12, 20000, 12
        X, y, theta, bounds = bench_problem(N, d)
        gpr = make_reference_gpr(gpry, "rbf", X, y, theta, bounds)
        Xc = np.random.default_rng(4321).uniform(size=(M, d))
        mean, std = gpr.predict(Xc, return_std=True, validate=False)
        zeta = d ** (-0.85)
        acq_func = partial(LogExp.f, baseline=gpr.y_max, noise_level=gpr.noise_level, zeta=zeta)
        with np.errstate(divide="ignore"):
            acq = acq_func(mean, std)
            rp = gpry.gp_acquisition.RankedPool(n_points, gpr=gpr, acq_func=acq_func, verbose=0)
            rp.add(Xc, mean, std, acq, method="single sort acq")
            rp = rp.copy(drop_empty=True)
        idx = np.array([int(np.flatnonzero(np.all(Xc == x, axis=1))[0]) for x in rp.X[:n_points]])
        np.savez_compressed(
            os.path.join(OUT, "config_c_n2000_d12.npz"), N=N, d=d, M=M, seed=1234, cand_seed=4321,
            theta=theta, zeta=zeta, noise_level=1e-2, mean=mean, std=std, acq=acq,
            y_mean=gpr.preprocessing_y.mean_, y_std=gpr.preprocessing_y.std_, y_max=gpr.y_max,
            pool_n_points=n_points, pool_idx=idx, pool_y=rp.y[:n_points],
            pool_sigma=rp.sigma[:n_points], pool_acq_cond=rp.acq_cond[:n_points],
            alpha_head=gpr.alpha_[:16], L_diag=np.diag(gpr.L_), condK=np.linalg.cond(gpr.L_) ** 2)
        print("config_c: cond(K)=%.3g std [%.3g, %.3g] pool idx %s" % (
            np.linalg.cond(gpr.L_) ** 2, std.min(), std.max(), idx))
    if "D" in which:   # N_train = 4000, d = 20: LML + gradient at restart points
        N, d = 4000, 20
        X, y, theta, bounds = bench_problem(N, d)
        rng = np.random.default_rng(7)
        out = dict(N=N, d=d, seed=1234, noise_level=1e-2)
        for kind in ("rbf", "matern25"):
            gpr = make_reference_gpr(gpry, kind, X[:64], y[:64], theta, bounds)
            # the LML only needs the training set in the transformed space and alpha
            from gpry.preprocessing import Normalize_y
            py = Normalize_y()
            py.fit(X, y)
            gpr.X_train_ = X.copy()          # bounds [0, 1]^d: Normalize_bounds is the identity
            gpr.y_train_ = py.transform(y)
            gpr.alpha = np.full(N, (1e-2 / py.std_) ** 2)
            lo, hi = gpr.kernel_.bounds[:, 0], gpr.kernel_.bounds[:, 1]
            thetas = [theta, theta + 0.1 * rng.standard_normal(d + 1)]
            if kind == "rbf":     # one start drawn like fit_gpr_hyperparameters does (gpr.py:976)
                thetas.append(rng.uniform(lo, hi))
            lml, grad = [], []
            for th in thetas:
                v, g = gpr.log_marginal_likelihood(th, eval_gradient=True, clone_kernel=True)
                lml.append(v), grad.append(g)
                print(f"config_d {kind}: lml {v:.12g}  |grad| {np.abs(g).max():.4g}")
            out.update({f"thetas_{kind}": np.array(thetas), f"lml_{kind}": np.array(lml),
                        f"grad_{kind}": np.array(grad)})
        np.savez_compressed(os.path.join(OUT, "config_d_n4000_d20.npz"), **out)
    if "E" in which:   # N_train = 2000, d = 16: mean only (surrogate-MCMC proposals)
        N, d, M = 2000, 16, 20000
        X, y, theta, bounds = bench_problem(N, d)
        gpr = make_reference_gpr(gpry, "rbf", X, y, theta, bounds)
        Xc = np.random.default_rng(4321).uniform(size=(M, d))
        mean = gpr.predict(Xc, validate=False)
        np.savez_compressed(os.path.join(OUT, "config_e_n2000_d16.npz"), N=N, d=d, M=M, seed=1234,
                            cand_seed=4321, theta=theta, noise_level=1e-2, mean=mean,
                            y_std=gpr.preprocessing_y.std_, condK=np.linalg.cond(gpr.L_) ** 2)
        print("config_e: mean range [%.4g, %.4g]" % (mean.min(), mean.max()))


def main():
    os.makedirs(OUT, exist_ok=True)
    gpry = import_reference()
    if len(sys.argv) > 1 and sys.argv[1] == "fit_d8":
        fit_case_d8(gpry)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "configs":
        config_cases(gpry, tuple(sys.argv[2:]) or ("C", "D", "E"))
        return
    case(gpry, "rbf_d2_n60", "rbf", 60, 2, 128, 11, 0.3, pool=(4000, 4))
    case(gpry, "rbf_d8_n300", "rbf", 300, 8, 256, 12, 0.5, pool=(20000, 8))
    case(gpry, "matern25_d8_n300", "matern25", 300, 8, 256, 13, 0.8)
    case(gpry, "matern15_d5_n200", "matern15", 200, 5, 256, 14, 0.7,
         bounds=np.array([[-2.0, 3.0], [0.0, 10.0], [1.0, 1.5], [-5.0, -1.0], [0.0, 1.0]]))
    case(gpry, "rbf_d12_n500_c4", "rbf", 500, 12, 256, 15, 1.0, c=4.0)
    case(gpry, "rbf_d3_n100_raw", "rbf", 100, 3, 128, 16, 0.4, normalize=False)
    case(gpry, "rbf_d8_n1000", "rbf", 1000, 8, 512, 1234, 0.5, with_lml=False,
         store_train=False)
    # upper clipping of the mean active (clip_factor = 1: nothing may exceed max(y_train))
    case(gpry, "rbf_d4_n150_clip1", "rbf", 150, 4, 2048, 18, 0.35, with_lml=False,
         clip_factor=1.0)
    # higher dimensions, and sizes whose pools take the INT8 tensor-core contraction
    # (N_pad >= 512, more than 64 candidates)
    extra_cases(gpry)
    nonpd_case(gpry)
    fit_case(gpry)
    fit_case_d8(gpry)
    loop_case(gpry)
    config_cases(gpry)


def extra_cases(gpry):
    case(gpry, "rbf_d16_n400", "rbf", 400, 16, 256, 19, 1.0)
    case(gpry, "matern15_d20_n600", "matern15", 600, 20, 700, 20, 1.5, c=2.0)
    case(gpry, "matern25_d6_n640", "matern25", 640, 6, 700, 21, 0.6, with_lml=False)
    # ranked pool of the reference on a model whose pool scoring takes the INT8 contraction
    case(gpry, "rbf_d8_n700_pool", "rbf", 700, 8, 128, 22, 0.5, with_lml=False,
         pool=(30000, 8))


if __name__ == "__main__":
    main()
