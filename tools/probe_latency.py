"""Latency of small predict calls through the Python API (numpy in / numpy out), as issued by
nested samplers / MCMC (1 point) and by the ranked pool (a few rows)."""
import sys, time, json
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from oracle import gp_oracle as orc
from gpry_b200 import DeviceGP
from test_gpu_predict import upload_from_oracle
dev = DeviceGP(0)
for N, d in [(200, 4), (2000, 12)]:
    X, y, theta, bounds = orc.synthetic_problem(N, d)
    st = orc.GPState("rbf", theta, X, y, bounds=bounds)
    upload_from_oracle(dev, st)
    out = {"N": N, "d": d}
    for M in (1, 16, 128, 1024, 16384):
        Xc = np.random.default_rng(M).uniform(size=(M, d))
        for what in ("mean", "mean_std"):
            f = (lambda: dev.predict(Xc)) if what == "mean" else (lambda: dev.predict(Xc, return_std=True))
            for _ in range(5): f()
            t0 = time.perf_counter()
            reps = 50 if M <= 1024 else 10
            for _ in range(reps): f()
            out[f"{what}_M{M}_us"] = round((time.perf_counter() - t0) / reps * 1e6, 1)
    t0 = time.perf_counter(); orc.predict(st, Xc[:1], return_std=True); 
    t0 = time.perf_counter()
    for _ in range(20): orc.predict(st, Xc[:1], return_std=True)
    out["cpu_oracle_mean_std_M1_us"] = round((time.perf_counter() - t0) / 20 * 1e6, 1)
    print(json.dumps(out))
