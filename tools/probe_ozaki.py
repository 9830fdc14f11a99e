"""INT8-split contraction vs the FP64 DMMA contraction: agreement and speed."""
import sys, time, json
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch
from oracle import gp_oracle as orc
from gpry_b200 import DeviceGP
from test_gpu_predict import upload_from_oracle

dev = DeviceGP(0)
for kind, N, d, M in [("rbf", 600, 5, 5000), ("rbf", 2000, 12, 2_000_000), ("matern25", 1000, 8, 400_000)]:
    X, y, theta, bounds = orc.synthetic_problem(N, d)
    st = orc.GPState(kind, theta, X, y, bounds=bounds)
    upload_from_oracle(dev, st)
    gen = torch.Generator(device="cuda"); gen.manual_seed(1)
    Xd = torch.rand((M, d), dtype=torch.float64, device="cuda", generator=gen)
    out = {"kind": kind, "N": N, "d": d, "M": M}
    res = {}
    for mode in ("fp64", "int8_1pass", "int8"):
        dev.set_contract_mode(mode)
        m, s = dev.predict(Xd, return_std=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        m, s = dev.predict(Xd, return_std=True)
        torch.cuda.synchronize()
        out[f"{mode}_ms"] = round((time.perf_counter() - t0) * 1e3, 2)
        out[f"{mode}_Mcand_s"] = round(M / (time.perf_counter() - t0) * 1e-6, 2)
        res[mode] = (m.cpu().numpy(), s.cpu().numpy())
    sy = st.y_std
    out["mean_equal"] = bool(np.array_equal(res["fp64"][0], res["int8"][0]))
    out["var_diff_int8_vs_fp64"] = float(np.max(np.abs(res["fp64"][1] ** 2 - res["int8"][1] ** 2)) / sy ** 2)
    n = min(M, 3000)
    mo, so = orc.predict(st, Xd[:n].cpu().numpy(), return_std=True)
    out["var_diff_1pass_vs_2pass"] = float(np.max(np.abs(res["int8_1pass"][1] ** 2 - res["int8"][1] ** 2)) / sy ** 2)
    for mode in ("fp64", "int8_1pass", "int8"):
        out[f"var_err_{mode}_vs_oracle"] = float(np.max(np.abs(res[mode][1][:n] ** 2 - so ** 2)) / sy ** 2)
    print(json.dumps(out), flush=True)
dev.close()
