"""Full-size (BASELINE config) GPU checks through size-independent properties: the oracle
cannot score 10^6..10^7 candidates at N_train = 2000 in seconds, so the large runs are pinned
by invariances (tile / chunk / order independence, exact ranking, interpolation at training
points) plus an oracle comparison on a sub-sample drawn from every part of the pool."""
import numpy as np
import pytest

from oracle import gp_oracle as orc
from test_gpu_predict import upload_from_oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def setup():
    import torch
    from gpry_b200 import DeviceGP
    N, d = 2000, 12
    X, y, theta, bounds = orc.synthetic_problem(N, d)
    st = orc.GPState("rbf", theta, X, y, bounds=bounds)
    dev = DeviceGP(0)
    upload_from_oracle(dev, st)
    M = 1_500_000
    gen = torch.Generator(device="cuda")
    gen.manual_seed(4321)
    Xd = torch.rand((M, d), dtype=torch.float64, device="cuda", generator=gen)
    yield dev, st, Xd
    dev.close()


def test_fullsize_invariances(setup):
    import torch
    dev, st, Xd = setup
    M, d = Xd.shape
    zeta = orc.auto_zeta(d)
    s = torch.cuda.current_stream()
    mean, std, acq = dev.predict_logexp(Xd, zeta, st.noise_level, st.y_max, stream=s)
    torch.cuda.synchronize()
    # order independence: a permuted pool gives the permuted result, bit for bit
    perm = torch.randperm(M, device="cuda")
    m2, s2, a2 = dev.predict_logexp(Xd[perm].contiguous(), zeta, st.noise_level, st.y_max, stream=s)
    assert torch.equal(m2, mean[perm]) and torch.equal(s2, std[perm]) and torch.equal(a2, acq[perm])
    # chunk independence: ragged splits of the pool give the same bits
    cut = 777_777
    ma, sa, _ = dev.predict_logexp(Xd[:cut].contiguous(), zeta, st.noise_level, st.y_max, stream=s)
    mb, sb, _ = dev.predict_logexp(Xd[cut:].contiguous(), zeta, st.noise_level, st.y_max, stream=s)
    assert torch.equal(torch.cat([ma, mb]), mean) and torch.equal(torch.cat([sa, sb]), std)
    # bounds: 0 <= var <= c (in original units: std <= sqrt(c) * y_std), clip on the mean
    c = float(np.exp(st.theta[0]))
    assert float(std.min()) >= 0.0 and float(std.max()) <= np.sqrt(c) * st.y_std * (1 + 1e-12)
    clip_hi = st.clip_factor * max(st.y_train) - (st.clip_factor - 1) * min(st.y_train)
    assert float(mean.max()) <= clip_hi
    # exact ranking of the whole pool
    Kp = 1024
    a, idx, m, sd, _ = dev.predict_logexp_topk(Xd, zeta, st.noise_level, st.y_max, Kp, stream=s,
                                               device_out=True, want_X=False)
    ref_vals, ref_idx = torch.sort(acq, descending=True, stable=True)
    assert torch.equal(idx, ref_idx[:Kp]) and torch.equal(a, ref_vals[:Kp])
    assert torch.equal(m, mean[idx]) and torch.equal(sd, std[idx])
    # oracle parity on a sub-sample spread over the whole pool (first / middle / last tiles)
    pick = torch.cat([torch.arange(0, 700), torch.arange(M // 2, M // 2 + 700),
                      torch.arange(M - 700, M)]).cuda()
    Xs = Xd[pick].cpu().numpy()
    mo, so, ao = orc.predict_logexp(st, Xs)
    assert np.max(np.abs(mean[pick].cpu().numpy() - mo)) < 1e-10 * st.y_std
    assert np.max(np.abs(std[pick].cpu().numpy() ** 2 - so ** 2)) < 1e-10 * st.y_std ** 2


def test_interpolates_training_points(setup):
    dev, st, _ = setup
    mean, std = dev.predict(st.X_train[:512], return_std=True)
    # at a training point the posterior std is of the order of the noise level and the mean
    # reproduces y within a few noise sigmas (reference test_io.py:60-61 pins the same thing)
    assert np.all(std < 20 * st.noise_level)
    assert np.all(np.abs(mean - st.y_train[:512]) < 5 * np.maximum(std, st.noise_level) + 1e-8)
    mo, so = orc.predict(st, st.X_train[:512], return_std=True)
    assert np.max(np.abs(mean - mo)) < 1e-10 * st.y_std
    assert np.max(np.abs(std ** 2 - so ** 2)) < 1e-10 * st.y_std ** 2
