"""The INT8 split of the variance contraction (ozaki.cu) against the oracle, the committed
golden vectors and the FP64 DMMA contraction, on models large enough to take that path
(N_pad >= 512, d <= 32, more than 64 candidates per call)."""
import numpy as np
import pytest

from conftest import load_golden, oracle_state, scaled_err
from oracle import gp_oracle as orc
from test_gpu_predict import upload_from_oracle

pytestmark = pytest.mark.gpu
TOL = 1e-10


@pytest.fixture(scope="module")
def dev():
    from gpry_b200 import DeviceGP
    d = DeviceGP(0)
    yield d
    d.close()


@pytest.mark.parametrize("name", ["rbf_d12_n500_c4", "rbf_d8_n1000"])
def test_int8_matches_golden_and_fp64(dev, name):
    g = load_golden(name)
    st = oracle_state(g)
    upload_from_oracle(dev, st)
    sy = float(g["y_std"])
    Xc = np.resize(g["Xc"], (max(len(g["Xc"]), 700), g["d"]))        # > 64 rows: tiled path
    n = len(g["Xc"])
    out = {}
    for mode in ("fp64", "int8_1pass", "int8"):
        dev.set_contract_mode(mode)
        out[mode] = dev.predict_logexp(Xc, float(g["zeta"]), float(g["noise_level"]), st.y_max)
    dev.set_contract_mode("int8")
    # the one-pass and the two-pass integer kernels form the same exact digit sums
    assert scaled_err(out["int8"][1] ** 2, out["int8_1pass"][1] ** 2, sy ** 2) < 1e-14
    for mode in ("fp64", "int8_1pass", "int8"):
        mean, std, acq = out[mode]
        assert scaled_err(mean[:n], g["mean"], sy) < TOL
        assert scaled_err(std[:n] ** 2, g["std"] ** 2, sy ** 2) < TOL
    assert np.array_equal(out["fp64"][0], out["int8"][0])              # the mean is not touched
    assert scaled_err(out["int8"][1] ** 2, out["fp64"][1] ** 2, sy ** 2) < 1e-12
    # ranking through the fused call: same survivors, same order
    a8, i8, *_ = dev.predict_logexp_topk(Xc[:n], float(g["zeta"]), float(g["noise_level"]),
                                         st.y_max, 64)
    dev.set_contract_mode("fp64")
    a64, i64, *_ = dev.predict_logexp_topk(Xc[:n], float(g["zeta"]), float(g["noise_level"]),
                                           st.y_max, 64)
    dev.set_contract_mode("int8")
    resolved = np.abs(np.diff(a64)) > 1e-9            # ties closer than the tolerance may swap
    assert np.array_equal(i8[:-1][resolved], i64[:-1][resolved])


@pytest.mark.parametrize("kind,N,d", [("rbf", 513, 3), ("matern15", 1100, 7), ("matern25", 2000, 12),
                                      ("rbf", 3000, 32)])
def test_int8_ragged_sizes_vs_oracle(dev, kind, N, d):
    X, y, theta, bounds = orc.synthetic_problem(N, d)
    st = orc.GPState(kind, theta, X, y, bounds=bounds)
    upload_from_oracle(dev, st)
    dev.set_contract_mode("int8")
    rng = np.random.default_rng(N)
    for M in (65, 129, 1000, 40000):
        Xc = rng.uniform(size=(M, d))
        mean, std = dev.predict(Xc, return_std=True)
        pick = np.unique(np.concatenate([np.arange(min(M, 40)), np.arange(max(M - 40, 0), M)]))
        mo, so = orc.predict(st, Xc[pick], return_std=True)
        assert scaled_err(mean[pick], mo, st.y_std) < TOL
        assert scaled_err(std[pick] ** 2, so ** 2, st.y_std ** 2) < TOL
        # chunk / order independence of the integer path: bit-identical under permutation
        perm = rng.permutation(M)
        m2, s2 = dev.predict(np.ascontiguousarray(Xc[perm]), return_std=True)
        assert np.array_equal(m2, mean[perm]) and np.array_equal(s2, std[perm])
    # training points: tiny variances, clamp at zero, never NaN
    Xt = X[:200]
    mean, std = dev.predict(Xt, return_std=True)
    mo, so = orc.predict(st, Xt, return_std=True)
    assert np.all(np.isfinite(std)) and np.all(std >= 0)
    assert scaled_err(std ** 2, so ** 2, st.y_std ** 2) < TOL


def test_int8_nonfinite_and_extreme_scales(dev):
    # wide dynamic range in V (small noise -> entries of L^-1 up to ~1e4) and a large constant c
    N, d = 640, 2
    X, y, theta, bounds = orc.synthetic_problem(N, d)
    theta = theta.copy()
    theta[0] = np.log(1e4)
    st = orc.GPState("rbf", theta, X, y, bounds=bounds, noise_level=1e-4)
    upload_from_oracle(dev, st)
    rng = np.random.default_rng(5)
    Xc = rng.uniform(size=(3000, d))
    Xc[7, 1] = np.nan
    Xc[11, 0] = np.inf
    res = {}
    for mode in ("fp64", "int8"):
        dev.set_contract_mode(mode, guard=False)      # the raw integer path, whatever the guard says
        res[mode] = dev.predict_logexp(Xc, 0.3, st.noise_level, st.y_max)
    dev.set_contract_mode("int8")
    m8, s8, a8 = res["int8"]
    m64, s64, a64 = res["fp64"]
    assert np.isnan(m8[7]) and np.isnan(s8[7]) and np.isnan(a8[7])
    assert np.isfinite(s8[11]) and s8[11] == s64[11]            # k* = 0 everywhere: var = c
    good = np.ones(len(Xc), bool)
    good[[7, 11]] = False
    mo, so = orc.predict(st, Xc[good][:600], return_std=True)
    # ill-conditioned (cond(K) >> 1e10): compare both device paths with the oracle on equal
    # footing -- the integer path must be as close to it as the FP64 path is (or within 1e-10)
    c = float(np.exp(theta[0]))
    err64 = np.max(np.abs(s64[good][:600] ** 2 - so ** 2)) / (c * st.y_std ** 2)
    err8 = np.max(np.abs(s8[good][:600] ** 2 - so ** 2)) / (c * st.y_std ** 2)
    assert err8 < max(1e-10, 4 * err64)


def test_guard_headline_model_stays_int8(dev):
    """N_train = 2000, d = 12, c = 1 (the benchmark's model): estimate and probe are far below
    the tolerance, the INT8 kernel is used."""
    X, y, theta, bounds = orc.synthetic_problem(2000, 12)
    st = orc.GPState("rbf", theta, X, y, bounds=bounds)
    upload_from_oracle(dev, st)
    dev.set_contract_mode("int8")
    info = dev.contract_info()
    assert info["requested"] == "int8" and info["in_use"] == "int8" and info["guard"]
    assert info["estimate_sigma"] < info["tolerance"]
    assert info["estimate_sigma"] < info["bound_worst_case"]
    assert info["probe_diff"] is not None and info["probe_diff"] * 16 <= info["tolerance"]


@pytest.mark.parametrize("c,ell,noise", [(1e6, 8.0, 1e-2), (1.0, 1e-3, 1e-2), (100.0, 1.0, 1e-2),
                                         (1e4, 1.0, 1e-4)])
def test_guard_falls_back_when_the_split_is_not_safe(dev, c, ell, noise):
    """theta where GPry's own fits drift (c = 1e6, l ~ 8: SURVEY section 7), a vanishing length
    scale, and large output scales: whatever the guard decides, the result must be as close to
    the reference as the FP64 kernel's (or within 1e-10); where the estimate exceeds the
    tolerance the FP64 kernel must be the one in use."""
    N, d = 900, 6
    X, y, theta, bounds = orc.synthetic_problem(N, d)
    theta = np.log(np.concatenate([[c], np.full(d, ell)]))
    try:
        st = orc.GPState("rbf", theta, X, y, bounds=bounds, noise_level=noise)
    except np.linalg.LinAlgError:
        pytest.skip("not positive definite on the host either")
    upload_from_oracle(dev, st)
    Xc = np.concatenate([np.random.default_rng(5).uniform(size=(1500, d)), X[:300]])
    mo, so = orc.predict(st, Xc, return_std=True)
    dev.set_contract_mode("fp64")
    _, s64 = dev.predict(Xc, return_std=True)
    dev.set_contract_mode("int8")
    info = dev.contract_info()
    _, s8 = dev.predict(Xc, return_std=True)
    scale = np.maximum(so ** 2, st.y_std ** 2)
    err64 = np.max(np.abs(s64 ** 2 - so ** 2) / scale)
    err8 = np.max(np.abs(s8 ** 2 - so ** 2) / scale)
    assert err8 <= max(1e-10, 2 * err64), (info, err8, err64)
    if info["estimate_sigma"] > info["tolerance"]:
        assert info["in_use"] == "fp64" and np.array_equal(s8, s64)
    if c >= 1e4:
        assert info["in_use"] == "fp64"


def test_int8_used_only_where_supported(dev):
    """Small models and small batches silently take the FP64 path: same numbers whatever the
    mode."""
    for N, d, M in [(300, 4, 500), (600, 4, 64)]:
        X, y, theta, bounds = orc.synthetic_problem(N, d)
        st = orc.GPState("rbf", theta, X, y, bounds=bounds)
        upload_from_oracle(dev, st)
        Xc = np.random.default_rng(1).uniform(size=(M, d))
        dev.set_contract_mode("fp64")
        a = dev.predict(Xc, return_std=True)
        dev.set_contract_mode("int8")
        b = dev.predict(Xc, return_std=True)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_regressor_contraction_switch():
    """`GaussianProcessRegressor(contraction=...)` selects the kernel; results agree to 1e-12."""
    from copy import deepcopy
    from test_gpu_gpr import make_gpr
    g = load_golden("rbf_d8_n1000")
    sy = float(g["y_std"])
    Xc = np.resize(g["Xc"], (800, g["d"]))
    out = {}
    for mode in (None, "fp64", "int8_1pass"):
        gpr = make_gpr(g, contraction=mode)
        out[mode] = gpr.predict(Xc, return_std=True)
        assert deepcopy(gpr).contraction == mode
    assert np.array_equal(out[None][0], out["fp64"][0])
    assert scaled_err(out[None][1] ** 2, out["fp64"][1] ** 2, sy ** 2) < 1e-12
    assert scaled_err(out[None][1] ** 2, out["int8_1pass"][1] ** 2, sy ** 2) < 1e-14
    assert not np.array_equal(out[None][1], out["fp64"][1])      # really two different kernels


def test_int8_wide_models(dev):
    """d > 32 (candidate coordinates in shared memory instead of registers) takes the INT8
    contraction too."""
    N, d, M = 600, 40, 900
    X, y, theta, bounds = orc.synthetic_problem(N, d)
    st = orc.GPState("matern25", theta, X, y, bounds=bounds)
    upload_from_oracle(dev, st)
    Xc = np.random.default_rng(2).uniform(size=(M, d))
    dev.set_contract_mode("fp64")
    m64, s64 = dev.predict(Xc, return_std=True)
    dev.set_contract_mode("int8")
    m8, s8 = dev.predict(Xc, return_std=True)
    mo, so = orc.predict(st, Xc[:200], return_std=True)
    assert np.array_equal(m8, m64) and not np.array_equal(s8, s64)
    assert scaled_err(s8 ** 2, s64 ** 2, st.y_std ** 2) < 1e-12
    assert scaled_err(m8[:200], mo, st.y_std) < TOL
    assert scaled_err(s8[:200] ** 2, so ** 2, st.y_std ** 2) < TOL
