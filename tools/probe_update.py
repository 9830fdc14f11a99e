"""Cost of one model update (``append_to_data(fit_gpr=False)`` = ``_update_model``: K build,
Cholesky, L^-1, alpha_ on the device) followed by a small predict, through the Python API:
(a) as shipped -- factor consumed on the device, host L_/V_ fetched lazily (never here);
(b) the same with the host copies forced (what an eager implementation pays every update);
plus acquisition-optimiser figures: one batched gradient call and one ``BatchOptimizer.multi_add``."""
import sys, time, json
from copy import deepcopy
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from gpry_b200.gpr import GaussianProcessRegressor
from gpry_b200.preprocessing import Normalize_bounds, Normalize_y
from gpry_b200.gp_acquisition import BatchOptimizer
from bench import synthetic_problem

for N, d in [(1000, 8), (2000, 12), (4000, 20)]:
    X, y, theta, bounds = synthetic_problem(N + 8, d)
    gpr = GaussianProcessRegressor(kernel="RBF", bounds=bounds, noise_level=1e-2,
                                   preprocessing_X=Normalize_bounds(bounds),
                                   preprocessing_y=Normalize_y(), account_for_inf=None, verbose=0)
    gpr.kernel_ = deepcopy(gpr.kernel)
    gpr.kernel_.theta = theta
    gpr.append_to_data(X[:N], y[:N], fit_gpr=False)
    gpr.predict(X[:4], return_std=True)
    out = {"N": N, "d": d}
    for mode in ("lazy", "eager"):
        ts = []
        for i in range(4):
            t0 = time.perf_counter()
            gpr.append_to_data(X[N + 2 * i + (mode == "eager"):][:1], y[N + 2 * i + (mode == "eager"):][:1],
                               fit_gpr=False, fit_classifier=False)
            if mode == "eager":
                gpr.V_, gpr.L_
            gpr.predict(X[:4], return_std=True)
            ts.append(time.perf_counter() - t0)
        out[f"update_plus_predict_ms_{mode}"] = round(min(ts) * 1e3, 2)
    Xg = np.random.default_rng(0).uniform(size=(5 * d, d))
    gpr.predict_grad_batch(Xg)
    t0 = time.perf_counter()
    for _ in range(5):
        gpr.predict_grad_batch(Xg)
    out["grad_batch_5d_points_ms"] = round((time.perf_counter() - t0) / 5 * 1e3, 3)
    t0 = time.perf_counter()
    for i in range(len(Xg)):
        gpr.predict(Xg[i:i + 1], return_std=True, return_mean_grad=True, return_std_grad=True)
    out["one_point_calls_5d_points_ms"] = round((time.perf_counter() - t0) * 1e3, 3)
    if N <= 2000:
        opt = BatchOptimizer(bounds, preprocessing_X=Normalize_bounds(bounds), verbose=0)
        opt.multi_add(gpr, n_points=1, rng=np.random.default_rng(1))
        e0 = gpr.n_eval
        t0 = time.perf_counter()
        opt.multi_add(gpr, n_points=2, rng=np.random.default_rng(2))
        out["batch_optimizer_2_points_s"] = round(time.perf_counter() - t0, 3)
        out["batch_optimizer_gp_evals"] = int(gpr.n_eval - e0)
    print(json.dumps(out), flush=True)
