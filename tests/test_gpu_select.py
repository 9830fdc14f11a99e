"""GPU tests of the fused finish + streaming selection behind ``gpry_predict_logexp_topk``
(csrc/predict.cu ``finish_select_kernel``, csrc/topk.cu): the K' records that leave the GPU must
be exactly the first K' of ``np.lexsort((index, -acq))`` over the scores that
``gpry_predict_logexp`` returns for the same pool -- the order ``RankedPool.add(method="single
sort acq")`` visits them in (gp_acquisition.py:1326-1333) -- whatever the pool size (one tile,
ragged tiles, several compaction periods, several host staging blocks), with the device masks
(trust region / classifier), NaN rows, ties, and the skip list of already proposed rows
(gp_acquisition.py:1037-1047)."""
import numpy as np
import pytest

from oracle import gp_oracle as orc
from test_gpu_predict import upload_from_oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    from gpry_b200 import DeviceGP
    d = DeviceGP(0)
    yield d
    d.close()


def reference_order(acq, Kp, exclude=None):
    key = np.where(np.isnan(acq), -np.inf, acq)
    nan_last = np.isnan(acq)
    order = np.lexsort((np.arange(len(acq)), nan_last, -key))
    if exclude is not None and len(exclude):
        order = order[~np.isin(order, exclude)]
    return order[:Kp]


def model(dev, N=300, d=5, kind="rbf"):
    X, y, theta, bounds = orc.synthetic_problem(N, d)
    st = orc.GPState(kind, theta, X, y, bounds=bounds)
    upload_from_oracle(dev, st)
    return st


@pytest.mark.parametrize("M", [1, 63, 65, 128, 129, 5000, 40000, 700001])
@pytest.mark.parametrize("Kp", [1, 16, 1024, 2048])
def test_selection_equals_full_sort(dev, M, Kp):
    st = model(dev)
    zeta = orc.auto_zeta(5)
    Xc = np.random.default_rng(M + Kp).uniform(size=(M, 5))
    mean, std, acq = dev.predict_logexp(Xc, zeta, st.noise_level, st.y_max)
    a, idx, m, s, Xo = dev.predict_logexp_topk(Xc, zeta, st.noise_level, st.y_max, Kp,
                                              idx_offset=7)
    order = reference_order(acq, Kp)
    assert len(a) == min(Kp, M)
    assert np.array_equal(idx - 7, order)
    # (values: the scoring call may take the latency path (M <= 64) or the contraction's fused
    # epilogue where the selection runs the tiled kernels: same numbers to round-off)
    assert np.allclose(a, acq[order], rtol=1e-10, atol=1e-10)
    assert np.allclose(m, mean[order], rtol=1e-11, atol=1e-11 * st.y_std)
    assert np.allclose(s, std[order], rtol=1e-10, atol=1e-11 * st.y_std)
    assert np.array_equal(Xo, Xc[order])


def test_int8_model_selection_and_device_io(dev):
    import torch
    st = model(dev, N=600, d=8)         # N_pad >= 512: INT8 contraction, row blocks split over CTAs
    zeta = orc.auto_zeta(8)
    M, Kp = 120000, 512
    Xc = np.random.default_rng(5).uniform(size=(M, 8))
    mean, std, acq = dev.predict_logexp(Xc, zeta, st.noise_level, st.y_max)
    order = reference_order(acq, Kp)
    Xd = torch.from_numpy(Xc).cuda()
    s = torch.cuda.current_stream()
    a, idx, m, sd, Xo = dev.predict_logexp_topk(Xd, zeta, st.noise_level, st.y_max, Kp,
                                                idx_offset=10 ** 10, stream=s, device_out=True)
    torch.cuda.synchronize()
    assert np.array_equal(idx.cpu().numpy() - 10 ** 10, order)
    assert np.array_equal(a.cpu().numpy(), acq[order])
    assert np.array_equal(m.cpu().numpy(), mean[order])
    assert np.array_equal(sd.cpu().numpy(), std[order])
    assert np.array_equal(Xo.cpu().numpy(), Xc[order])


def test_skip_list(dev):
    st = model(dev)
    zeta = orc.auto_zeta(5)
    M, Kp = 90000, 256
    Xc = np.random.default_rng(11).uniform(size=(M, 5))
    _, _, acq = dev.predict_logexp(Xc, zeta, st.noise_level, st.y_max)
    best = reference_order(acq, 40)
    skip = np.sort(np.concatenate([best[::2], [0, M - 1, 40000]]))
    a, idx, *_ = dev.predict_logexp_topk(Xc, zeta, st.noise_level, st.y_max, Kp, exclude=skip)
    assert np.array_equal(idx, reference_order(acq, Kp, exclude=skip))
    # the list is per call: cleared by the next call without one
    a2, idx2, *_ = dev.predict_logexp_topk(Xc, zeta, st.noise_level, st.y_max, Kp)
    assert np.array_equal(idx2, reference_order(acq, Kp))
    # everything but 3 rows skipped: 3 records come back
    few = np.random.default_rng(0).uniform(size=(300, 5))
    keep = np.array([5, 17, 299])
    a3, idx3, *_ = dev.predict_logexp_topk(few, zeta, st.noise_level, st.y_max, 64,
                                           exclude=np.setdiff1d(np.arange(300), keep))
    assert sorted(idx3.tolist()) == keep.tolist() and len(a3) == 3


def test_masks_ties_and_nan(dev):
    st = model(dev)
    zeta = orc.auto_zeta(5)
    M, Kp = 50000, 2048
    Xc = np.random.default_rng(3).uniform(size=(M, 5))
    Xc[100] = np.nan
    Xc[40001, 2] = np.inf
    # a trust region that leaves ~3 % of the pool: most of the K' records are -inf ties, which
    # must come out in ascending index order after the finite ones
    tb = np.array([[0.0, 0.5]] * 5)
    dev.set_trust_region(tb, -np.inf)
    try:
        mean, std, acq = dev.predict_logexp(Xc, zeta, st.noise_level, st.y_max)
        n_finite = int(np.sum(np.isfinite(acq)))
        assert 0 < n_finite < Kp
        a, idx, m, s, _ = dev.predict_logexp_topk(Xc, zeta, st.noise_level, st.y_max, Kp)
        order = reference_order(acq, Kp)
        assert np.array_equal(idx, order)
        assert np.allclose(a, acq[order], rtol=1e-10, atol=1e-10, equal_nan=True)
        assert np.array_equal(np.isfinite(a), np.isfinite(acq[order]))
        assert np.allclose(m, mean[order], rtol=1e-11, atol=1e-11 * st.y_std, equal_nan=True)
        assert np.allclose(s, std[order], rtol=1e-10, atol=1e-11 * st.y_std, equal_nan=True)
    finally:
        dev.set_trust_region(None)


def test_host_staging_blocks(dev):
    """Host candidates are staged in blocks of 56 x 2 x n_sm x 128 rows; indices must stay
    global across block boundaries and the selection state must carry over."""
    st = model(dev, N=130, d=3)
    zeta = orc.auto_zeta(3)
    M, Kp = 2_300_000, 1024
    Xc = np.random.default_rng(8).uniform(size=(M, 3))
    _, _, acq = dev.predict_logexp(Xc, zeta, st.noise_level, st.y_max)
    a, idx, _, _, Xo = dev.predict_logexp_topk(Xc, zeta, st.noise_level, st.y_max, Kp)
    order = reference_order(acq, Kp)
    assert np.array_equal(idx, order) and np.array_equal(Xo, Xc[order])
    assert idx.max() > 56 * 2 * 132 * 128      # survivors beyond the first block
