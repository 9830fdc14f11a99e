"""Small end-to-end pass over every C-ABI entry point, for compute-sanitizer
(memcheck / racecheck / initcheck):  compute-sanitizer --tool memcheck python tools/sanitize_smoke.py"""
import sys
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from oracle import gp_oracle as orc
from gpry_b200 import DeviceGP
from test_gpu_predict import upload_from_oracle

dev = DeviceGP(0)
for kind, N, d, M in [("rbf", 300, 5, 700), ("matern25", 140, 33, 130), ("rbf", 700, 6, 900)]:   # the last one takes the INT8 contraction
    X, y, theta, bounds = orc.synthetic_problem(N, d)
    st = orc.GPState(kind, theta, X, y, bounds=bounds)
    upload_from_oracle(dev, st)
    Xc = np.random.default_rng(0).uniform(size=(M, d))
    m, s = dev.predict(Xc, return_std=True)
    m2, _ = dev.predict(Xc)
    a, idx, mm, ss, Xo = dev.predict_logexp_topk(Xc, 0.3, st.noise_level, st.y_max, 64)
    g = dev.mean_grad(Xc[0]); gs, sd = dev.std_grad(Xc[0])
    S = dev.posterior_cov(Xc[:200])
    K = dev.kernel_cross(kind, theta, st.X_train_[:40], st.X_train_[:50])
    G = dev.kernel_gradient_x(st.X_train_[3])
    L, V, al, ld, info = dev.factorize(kind, st.X_train_, st.noise2, st.y_train_, theta, keep_on_device=True)
    c, ell = orc.split_theta(theta)
    dev.adopt_factorization(c, ell, bounds[:, 0], bounds[:, 1] - bounds[:, 0], st.y_mean, st.y_std, np.inf)
    m3, s3 = dev.predict(Xc, return_std=True)
    L2, V2 = dev.factor_download()
    assert np.array_equal(L2, L) and np.array_equal(V2, V)
    # small-batch path, batched gradients, device-side masks
    for Ms in (1, 7, 64):
        dev.predict(Xc[:Ms], return_std=True)
    pm, ps, gm, gsd = dev.predict_grad(Xc[:131])
    dev.set_trust_region(np.stack([bounds[:, 0] + 0.2, bounds[:, 1] - 0.2], axis=1), -np.inf)
    dev.set_mask_value(-1e300)
    rng = np.random.default_rng(1)
    dev.set_classifier((rng.uniform(size=(37, d)), rng.standard_normal(37), 0.1, 0.7))
    dec = dev.classify(Xc)
    mt, st_ = dev.predict(Xc, return_std=True)
    at, it, _, _, _ = dev.predict_logexp_topk(Xc, 0.3, st.noise_level, st.y_max, 64)
    assert np.all(mt[dec <= 0] == -1e300) and np.all(st_[dec <= 0] == 0)
    dev.set_classifier(None)
    dev.set_trust_region(None)
    lml, grad, inf = dev.lml_batched(kind, st.X_train_, st.noise2, st.y_train_, np.array([theta, theta + 0.1, theta - 0.1]))
    mo, so = orc.predict(st, Xc, return_std=True)
    print(kind, N, d, "mean err", np.abs(m - mo).max(), "var err", np.abs(s**2 - so**2).max(), "lml", lml, "info", info, inf)
dev.close()
print("done")
