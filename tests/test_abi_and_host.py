"""CPU-only checks: the C-ABI library loads and exports every symbol the header declares;
host-side logic (kernel descriptors, pre-processors, ranked pool control flow, NORA
pre-selection bound) behaves like the reference.  No compute call touches a GPU here."""
import os
import re
from copy import deepcopy
from functools import partial

import numpy as np
import pytest

from conftest import ROOT, golden_pool_candidates, load_golden, oracle_state
from oracle import gp_oracle as orc


def header_functions():
    src = open(os.path.join(ROOT, "include", "gpry_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gpry_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    import ctypes
    from gpry_b200 import _lib
    lib = _lib.load_library()
    names = header_functions()
    assert len(names) >= 18
    for name in names:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes prototype"
    assert set(_lib.SIGNATURES) == set(names)
    assert lib.gpry_abi_version() == 1


def test_no_gpu_fails_loudly():
    """Without a CUDA device the product path must raise, not fall back."""
    from gpry_b200 import _lib
    lib = _lib.load_library()
    if lib.gpry_device_count() > 0:
        pytest.skip("a GPU is visible")
    from gpry_b200 import DeviceGP, GpryB200Error
    with pytest.raises(GpryB200Error):
        DeviceGP(0)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "gpry_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in re.sub(r'""".*?"""', "", src, flags=re.S), fn


def test_kernel_descriptors():
    from gpry_b200.kernels import ConstantKernel as C, RBF, Matern
    d = 3
    k = C(10.0, [1e-4, 1e6]) * RBF([0.1] * d, [1e-3, 1e1], prior_bounds=np.array([[0, 1.]] * d))
    assert np.allclose(k.theta, np.log([10.0, 0.1, 0.1, 0.1]))
    assert np.allclose(k.bounds, np.log([[1e-4, 1e6]] + [[1e-3, 10.0]] * d))
    k.theta = np.log([2.0, 0.5, 0.6, 0.7])
    assert k.k1.constant_value == pytest.approx(2.0) and np.allclose(k.k2.length_scale,
                                                                      [0.5, 0.6, 0.7])
    assert k.device_spec(d)[0] == "rbf" and k.theta_is_standard(d)
    k2 = k.clone_with_theta(np.log([1.0, 1.0, 1.0, 1.0]))
    assert k.k1.constant_value == pytest.approx(2.0) and k2.k1.constant_value == 1.0
    m = C(1.0) * Matern([1.0] * d, nu=2.5)
    assert m.device_spec(d)[0] == "matern25"
    with pytest.raises(ValueError):
        Matern(nu=0.5)
    dyn = RBF([0.2, 0.3], "dynamic", prior_bounds=np.array([[0, 2.], [0, 4.]]))
    assert np.allclose(dyn.bounds, np.log([[2e-3, 200.], [4e-3, 400.]]))
    assert len(k.hyperparameters) == 2 and k.hyperparameters[1].n_elements == d


def test_preprocessing():
    from gpry_b200.preprocessing import Normalize_bounds, Normalize_y
    b = np.array([[-1.0, 3.0], [2.0, 2.5]])
    nb = Normalize_bounds(b)
    X = np.array([[0.0, 2.25], [3.0, 2.0]])
    assert np.array_equal(nb.transform(X), orc.normalize_bounds_transform(X, b))
    assert np.allclose(nb.inverse_transform(nb.transform(X)), X)
    ny = Normalize_y()
    y = np.array([1.0, -2.0, np.inf, 5.0])
    ny.fit(None, y)
    assert (ny.mean_, ny.std_) == orc.normalize_y_fit(y)
    with pytest.raises(TypeError):
        Normalize_y().transform(y)


class OracleBackedGPR:
    """A stand-in regressor (CPU oracle) with the few methods RankedPool's 'refit'
    conditioning uses, to test the pool's control flow without a GPU."""

    def __init__(self, st):
        self.st = st
        self.d = st.d
        self.noise_level = st.noise_level
        self.y_max = st.y_max

    def predict_std(self, X, validate=True):
        return orc.predict_std(self.st, X)

    def append_to_data(self, X, y, fit_gpr=False, fit_classifier=False):
        self.st = self.st.appended(X, y)

    def __deepcopy__(self, memo):
        return OracleBackedGPR(self.st)


@pytest.mark.parametrize("name", ["rbf_d2_n60", "rbf_d8_n300"])
@pytest.mark.parametrize("method", ["single sort acq", "bulk"])
def test_ranked_pool_control_flow(name, method):
    from gpry_b200.acquisition_functions import LogExp
    from gpry_b200.gp_acquisition import RankedPool
    g = load_golden(name)
    st = oracle_state(g)
    Xp = golden_pool_candidates(g)
    y, sigma, acq = orc.predict_logexp(st, Xp, zeta=g["zeta"])
    keep = np.argsort(acq)[::-1][:300]
    acq_func = partial(LogExp.f, baseline=st.y_max, noise_level=st.noise_level, zeta=g["zeta"])
    n_points = int(g["pool_n_points"])
    pool = RankedPool(n_points, gpr=OracleBackedGPR(st), acq_func=acq_func, verbose=0)
    with np.errstate(divide="ignore"):
        pool.add(Xp[keep], y[keep], sigma[keep], acq[keep], method=method)
    tag = method.replace(" ", "_")
    assert np.array_equal(keep[pool.idx[:n_points]], g[f"pool_idx_{tag}"])
    assert acq[keep][-1] <= pool.min_acq
    c = pool.copy(drop_empty=True)
    assert len(c.y) == n_points + 1 and not hasattr(c, "_gpr")
    # LogExp.f is the reference's static formula
    mu, s = np.array([1.0, 2.0]), np.array([0.5, 0.005])
    with np.errstate(divide="ignore"):
        v = LogExp.f(mu, s, 3.0, 0.01, 0.3)
    assert v[0] == pytest.approx(2 * 0.3 * (1 - 3) + 0.5 * np.log(0.25 - 1e-4)) and v[1] == -np.inf


def test_gpr_host_logic_without_gpu():
    """Constructor, kernel auto-construction and argument checks need no device."""
    from gpry_b200.gpr import GaussianProcessRegressor
    from gpry_b200.preprocessing import Normalize_bounds, Normalize_y
    bounds = np.array([[0.0, 1.0]] * 4)
    gpr = GaussianProcessRegressor(kernel={"Matern": {"nu": 2.5}}, bounds=bounds,
                                   preprocessing_X=Normalize_bounds(bounds),
                                   preprocessing_y=Normalize_y(), verbose=0)
    assert gpr.d == 4 and gpr.n == 0 and not gpr.fitted
    assert repr(gpr.kernel).startswith("3.16**2 * Matern(")
    assert np.allclose(gpr.kernel.theta, np.log([10.0] + [0.1] * 4))       # gpr.py:351-352
    assert np.allclose(gpr.kernel.bounds, np.log([[1e-4, 1e6]] + [[1e-3, 10.0]] * 4))
    with pytest.raises(ValueError):
        GaussianProcessRegressor(kernel="Foo", bounds=bounds)
    with pytest.raises(ValueError):
        GaussianProcessRegressor(bounds=bounds, clip_factor=0.5)
    import pickle
    g2 = pickle.loads(pickle.dumps(gpr))
    assert g2._dev is None and g2.d == 4
    assert deepcopy(gpr).kernel == gpr.kernel


def test_unpickled_regressor_uses_the_local_gpu(monkeypatch):
    """A regressor broadcast from another rank (the reference's mpi.bcast of the GPR,
    gp_acquisition.py:453) must run on the receiving process's GPU: device ordinals are per
    process, one process per GPU (LOCAL_RANK)."""
    import pickle
    from gpry_b200.gpr import GaussianProcessRegressor
    from gpry_b200.svm import SVM
    bounds = np.array([[0.0, 1.0]] * 3)
    gpr = GaussianProcessRegressor(bounds=bounds, verbose=0, device=0, account_for_inf="SVM")
    blob = pickle.dumps(gpr)
    assert pickle.loads(blob).device == 0                 # single process: kept
    monkeypatch.setenv("LOCAL_RANK", "1")
    g2 = pickle.loads(blob)
    assert g2.device == 1 and isinstance(g2.infinities_classifier, SVM)
    assert g2.infinities_classifier.device is None
    assert deepcopy(g2).device == 1


def test_fit_start_points_like_reference(monkeypatch):
    """The restart points of ``fit_gpr_hyperparameters`` (gpr.py:970-978): the first from the
    current theta when the regressor has been fitted before, the others ``rng.uniform`` over the
    LOG-bounds in loop order from ``check_random_state(random_state)`` -- drawn here exactly as
    the reference (and sklearn) would, so that a fit is reproducible against it."""
    from gpry_b200.gpr import GaussianProcessRegressor
    bounds = np.array([[0.0, 1.0]] * 3)
    seen = {}

    def fake_lockstep(self, theta_initials, hyper_bounds):
        seen["starts"] = [np.array(t) for t in theta_initials]
        return [(np.array(t), float(i)) for i, t in enumerate(theta_initials)]

    monkeypatch.setattr(GaussianProcessRegressor, "_lockstep_optimization", fake_lockstep)
    monkeypatch.setattr(GaussianProcessRegressor, "_update_model", lambda self: self)
    g = GaussianProcessRegressor(kernel="RBF", bounds=bounds, verbose=0, random_state=7,
                                 account_for_inf=None)
    g.fit_gpr_hyperparameters(n_restarts=5)                  # never fitted: all five are draws
    lo, hi = g.kernel.bounds[:, 0], g.kernel.bounds[:, 1]
    rs = np.random.RandomState(7)
    expect = [rs.uniform(lo, hi) for _ in range(5)]
    assert all(np.array_equal(a, b) for a, b in zip(seen["starts"], expect))
    assert np.allclose(g.kernel_.theta, expect[0], rtol=1e-14) and g.fitted   # argmin of the fakes
    # fitted before: restart 0 = current theta, then four draws from a FRESH generator of the seed
    g.fit_gpr_hyperparameters(n_restarts=5)
    rs = np.random.RandomState(7)
    assert np.allclose(seen["starts"][0], expect[0], rtol=1e-14)
    assert all(np.array_equal(a, rs.uniform(lo, hi)) for a in seen["starts"][1:])
    # a Generator as random_state is used as is (what the restart-parallel fit hands each rank)
    g.random_state = np.random.default_rng(3)
    g.fit_gpr_hyperparameters(n_restarts=3, start_from_current=False)
    gen = np.random.default_rng(3)
    assert all(np.array_equal(a, gen.uniform(lo, hi)) for a in seen["starts"])


def test_lockstep_restart_driver_without_gpu():
    """The lock-step multi-restart driver (one batched objective call per round) reaches the
    optima a serial L-BFGS-B reaches, and propagates a failing evaluation instead of hanging."""
    import scipy.optimize
    from gpry_b200.gpr import GaussianProcessRegressor
    bounds = np.array([[0.0, 1.0]] * 2)
    g = GaussianProcessRegressor(kernel="RBF", bounds=bounds, verbose=0)
    target, w = np.array([1.0, -1.0, 0.5]), np.array([1.0, 2.0, 3.0])
    batch_sizes = []

    def fake_batch(thetas, eval_gradient=True):
        thetas = np.atleast_2d(thetas)
        batch_sizes.append(len(thetas))
        return -np.sum(w * (thetas - target) ** 2, axis=1), -2 * w * (thetas - target)

    g.log_marginal_likelihood_batch = fake_batch
    box = np.array([[-5.0, 5.0]] * 3)
    starts = [np.zeros(3), np.full(3, 3.0), np.array([-4.0, 2.0, 1.0]), np.array([4.0, -4.0, 4.0])]
    optima = g._lockstep_optimization(starts, box)
    for th0, (x, f) in zip(starts, optima):
        ref = scipy.optimize.minimize(lambda t: (np.sum(w * (t - target) ** 2),
                                                 2 * w * (t - target)),
                                      th0, method="L-BFGS-B", jac=True, bounds=box)
        assert np.allclose(x, ref.x, atol=1e-12) and f == pytest.approx(ref.fun, abs=1e-15)
    assert max(batch_sizes) == 4 and len(batch_sizes) < 4 * 12

    def boom(thetas, eval_gradient=True):
        raise ValueError("device failure")

    g.log_marginal_likelihood_batch = boom
    with pytest.raises(ValueError):
        g._lockstep_optimization(starts, box)


def test_bench_reference_arm_runs_on_cpu():
    """`bench.py --impl reference` (the CPU arm) prints exactly one JSON line with the contract's
    keys; it needs no GPU."""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                          "--steps", "1", "--warmup", "1", "--cpu-chunk", "300",
                          "--ntrain", "200", "--dim", "4"],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step",
                "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config",
                "cpu_baseline", "e2e"):
        assert key in d
    assert d["impl"] == "reference" and d["value"] > 0 and d["cpu_baseline"]["kind"] == "port"
    assert d["e2e"]["h2d_bytes_per_step"] == 0


def test_build_entry_point():
    """__graft_entry__.build() compiles the library (nvcc cross-compiles without a GPU)."""
    import importlib
    ge = importlib.import_module("__graft_entry__")
    ge.build()
    assert os.path.exists(os.path.join(ROOT, "gpry_b200", "libgpry_b200.so"))


def test_bench_watchdog_prints_what_was_measured():
    """A multi-GPU bench run that stops making progress must not hang its caller: the watchdog
    dumps the stacks, rank 0 prints the line with what has been measured so far (marked
    "incomplete"), and the process exits; once the line is out, only exit."""
    import json
    import subprocess
    import sys
    code = (
        "import sys, os\n"
        "sys.argv = ['bench.py']\n"
        "import bench\n"
        "bench._REAL_STDOUT_FD = 1\n"
        "bench.PARTIAL['line'] = {'metric': bench.METRIC, 'value': 2.0, 'n_gpus': 4}\n"
        "bench.PARTIAL['section'] = 'secondary figures'\n"
        "bench._watchdog_fire()\n")
    out = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True)
    assert out.returncode == 0
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["value"] == 2.0 and "secondary figures" in line["incomplete"]
    code2 = code.replace("bench._watchdog_fire()", "bench.emit_line({'a': 1}); bench._watchdog_fire()")
    out = subprocess.run([sys.executable, "-c", code2], cwd=ROOT, capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.strip() == '{"a": 1}'
