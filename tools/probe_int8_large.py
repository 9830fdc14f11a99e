"""INT8 vs FP64 contraction at the upper end of the supported size range (int32 digit sums must not
overflow: |S_g| <= 7 N 2^14 < 2^31 up to N = 16384).  No oracle: the two device paths are compared."""
import sys, time, json
import numpy as np
sys.path.insert(0, ".")
import torch
from gpry_b200 import DeviceGP

dev = DeviceGP(0)
for N, d, M in [(8192, 10, 60000), (16384, 8, 30000)]:
    rng = np.random.default_rng(N)
    X = rng.uniform(size=(N, d))
    y = -0.5 * np.sum(((X - 0.5) / 0.15) ** 2, axis=1)
    y_ = (y - y.mean()) / y.std()
    theta = np.log(np.concatenate([[1.0], np.full(d, 0.6)]))
    noise2 = np.full(N, (1e-2 / y.std()) ** 2)
    t0 = time.perf_counter()
    _, _, alpha_, _, info = dev.factorize("rbf", X, noise2, y_, theta, want_L=False, want_V=False,
                                          keep_on_device=True)
    t_fact = time.perf_counter() - t0
    assert info == 0
    dev.adopt_factorization(1.0, np.full(d, 0.6), None, None, y.mean(), y.std(), np.inf)
    Xd = torch.rand((M, d), dtype=torch.float64, device="cuda")
    res, tms = {}, {}
    for mode in ("fp64", "int8_1pass", "int8"):
        dev.set_contract_mode(mode)
        dev.predict(Xd, return_std=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        m, s = dev.predict(Xd, return_std=True)
        torch.cuda.synchronize()
        tms[mode] = time.perf_counter() - t0
        res[mode] = (m.cpu().numpy(), s.cpu().numpy())
    sy = y.std()
    out = {"N": N, "d": d, "M": M, "factorize_s": round(t_fact, 3),
           "ms": {k: round(v * 1e3, 2) for k, v in tms.items()},
           "var_diff_int8_vs_fp64": float(np.max(np.abs(res["int8"][1] ** 2 - res["fp64"][1] ** 2)) / sy ** 2),
           "var_diff_1pass_vs_fp64": float(np.max(np.abs(res["int8_1pass"][1] ** 2 - res["fp64"][1] ** 2)) / sy ** 2),
           "std_range": [float(res["fp64"][1].min()), float(res["fp64"][1].max())],
           "mean_equal": bool(np.array_equal(res["int8"][0], res["fp64"][0]))}
    print(json.dumps(out), flush=True)
dev.close()
