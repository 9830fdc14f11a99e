#!/bin/bash
# first end-to-end GPU check: sanitizer on a tiny case, the gpu test-suite, a timing probe
mkdir -p gpurun_out
python - > gpurun_out/sanity.txt 2>&1 <<'PY'
import numpy as np, sys, time
sys.path.insert(0, ".")
from oracle import gp_oracle as orc
from gpry_b200 import DeviceGP
sys.path.insert(0, "tests")
from test_gpu_predict import upload_from_oracle
dev = DeviceGP(0)
X, y, theta, bounds = orc.synthetic_problem(300, 5)
st = orc.GPState("rbf", theta, X, y, bounds=bounds)
upload_from_oracle(dev, st)
Xc = np.random.default_rng(0).uniform(size=(1000, 5))
m, s = dev.predict(Xc, return_std=True)
mo, so = orc.predict(st, Xc, return_std=True)
print("max mean err", np.abs(m - mo).max(), "max var err", np.abs(s**2 - so**2).max())
PY
cat gpurun_out/sanity.txt
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.txt
