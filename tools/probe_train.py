"""Timing probe of the training side: LML+grad per evaluation and factorize, vs the CPU oracle."""
import sys, time, json
import numpy as np
sys.path.insert(0, ".")
from oracle import gp_oracle as orc
from gpry_b200 import DeviceGP
dev = DeviceGP(0)
cfgs = [(1000, 8, 8), (2000, 12, 8), (4000, 20, 8)]
if len(sys.argv) > 1:
    cfgs = [(int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]))]
for N, d, B in cfgs:
    X, y, theta, bounds = orc.synthetic_problem(N, d)
    y_mean, y_std = y.mean(), y.std()
    X_ = X; y_ = (y - y_mean) / y_std
    noise2 = np.full(N, (1e-2 / y_std) ** 2)
    rng = np.random.default_rng(7)
    thetas = theta + 0.1 * rng.standard_normal((B, d + 1))
    dev.lml_batched("rbf", X_, noise2, y_, thetas)
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter(); lml, grad, info = dev.lml_batched("rbf", X_, noise2, y_, thetas); t1 = time.perf_counter()
        best = min(best, t1 - t0)
    t0, t1 = 0.0, best
    t2 = time.perf_counter(); dev.factorize("rbf", X_, noise2, y_, theta, want_L=False, want_V=False); t3 = time.perf_counter()
    t4 = time.perf_counter(); dev.factorize("rbf", X_, noise2, y_, theta); t5 = time.perf_counter()
    out = dict(N=N, d=d, lml_grad_ms_per_eval=(t1 - t0) / B * 1e3, factorize_device_ms=(t3 - t2) * 1e3,
               factorize_with_LV_to_host_ms=(t5 - t4) * 1e3, info=info.tolist())
    if N <= 2000:
        t6 = time.perf_counter(); lo, go = orc.log_marginal_likelihood("rbf", thetas[0], X_, y_, noise2, eval_gradient=True); t7 = time.perf_counter()
        out["cpu_oracle_ms"] = (t7 - t6) * 1e3
        out["lml_relerr"] = abs(lml[0] - lo) / abs(lo); out["grad_err"] = float(np.max(np.abs(grad[0] - go)) / np.max(np.abs(go)))
    print(json.dumps(out))
