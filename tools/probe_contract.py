"""Timing probe of the candidate-side pipeline (not the bench): stage times from the library's
own CUDA events, at N=2000,d=12 and N=1000,d=8."""
import sys, time, json
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch
from oracle import gp_oracle as orc
from gpry_b200 import DeviceGP
from test_gpu_predict import upload_from_oracle

M = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
cfgs = [(2000, 12), (1000, 8)] if len(sys.argv) < 3 else [(int(sys.argv[2]), int(sys.argv[3]))]
import os
dev = DeviceGP(0)
if os.environ.get("GPRY_B200_OZ_DBG"):        # timing experiments: wrong results on purpose
    dev.set_contract_mode("int8", guard=False)
for N, d in cfgs:
    X, y, theta, bounds = orc.synthetic_problem(N, d)
    st = orc.GPState("rbf", theta, X, y, bounds=bounds)
    upload_from_oracle(dev, st)
    Xd = torch.rand((M, d), dtype=torch.float64, device="cuda")
    s = torch.cuda.current_stream()
    zeta = orc.auto_zeta(d)
    for _ in range(2):
        dev.predict_logexp_topk(Xd, zeta, st.noise_level, st.y_max, 1024, stream=s, device_out=True, want_X=False)
    torch.cuda.synchronize()
    dev.set_profiling(True); dev.timings(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    reps = 3
    for _ in range(reps):
        dev.predict_logexp_topk(Xd, zeta, st.noise_level, st.y_max, 1024, stream=s, device_out=True, want_X=False)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    tm = dev.timings(reset=True); dev.set_profiling(False)
    flop = (N * (N + 1) + 2 * N) * M
    out = dict(N=N, d=d, M=M, ms_per_pass=ms, cand_per_s=M / ms * 1e3,
               contract_tflops=flop / (tm["contract_ms"] / reps) * 1e-9,
               stage_ms={k: v / reps for k, v in tm.items()})
    print(json.dumps(out))
