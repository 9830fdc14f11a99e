"""world_size-2 tests (gloo, CPU) of the multi-process host logic: strided sharding and merge
(mpi.py:105-131), restart split (mpi.py:80-102), survivor all-gather, best-fit selection and
the sharded NORA ranking (result on 2 ranks == result on 1 rank), and the restart-split
BatchOptimizer (every rank returns the same batch), and the restart-parallel hyper-parameter
fit with the real regressor class on an oracle-backed stand-in for the device."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, golden_pool_candidates, load_golden, oracle_state
from oracle import gp_oracle as orc


class FakeDeviceGPR:
    """CPU stand-in (oracle arithmetic) exposing what NORA / parallel helpers call."""

    class _PY:
        def __init__(self, std_):
            self.std_ = std_

    def __init__(self, st):
        self.st = st
        self.d = st.d
        self.noise_level = st.noise_level
        self.y_max = st.y_max
        self.preprocessing_y = self._PY(st.y_std)
        self.n_eval = 0

    def predict(self, X, return_std=False, validate=True, return_mean_grad=False,
                return_std_grad=False):
        self.n_eval += len(X)
        return orc.predict(self.st, X, return_std=return_std, return_mean_grad=return_mean_grad,
                           return_std_grad=return_std_grad)

    def predict_std(self, X, validate=True):
        return orc.predict_std(self.st, X)

    def predict_logexp_topk(self, X, zeta, Kp, exclude=None, **kw):
        m, s, a = orc.predict_logexp(self.st, X, zeta=zeta)
        order = np.lexsort((np.arange(len(a)), -a))
        if exclude is not None and len(exclude):
            order = order[~np.isin(order, exclude)]
        order = order[:Kp]
        return a[order], order.astype(np.int64), m[order], s[order], X[order]

    def _device_state(self):
        return self

    # --- what BatchOptimizer / LogExp.__call__ use (oracle arithmetic, one point at a time)
    infinities_classifier = None
    trust_bounds = None

    @property
    def X_train(self):
        return self.st.X_train

    def predict_logexp(self, X, zeta, noise_level=None):
        self.n_eval += len(X)
        return orc.predict_logexp(self.st, X, zeta=zeta)

    def predict_grad_batch(self, X):
        self.n_eval += len(X)
        rows = [orc.predict(self.st, x[None], return_std=True, return_mean_grad=True,
                            return_std_grad=True) for x in X]
        return (np.array([r[0][0] for r in rows]), np.array([r[1][0] for r in rows]),
                np.array([r[2] for r in rows]), np.array([r[3] for r in rows]))

    def append_to_data(self, X, y, noise_level=None, fit_gpr=False, fit_classifier=False):
        self.st = self.st.appended(X, y)
        self.y_max = self.st.y_max

    def posterior_cov(self, X):
        X_ = self.st.transform_X(X)
        Ks = orc.kernel_cross(self.st.kind, self.st.theta, X_, self.st.X_train_)
        U = self.st.V_ @ Ks.T
        return orc.kernel_cross(self.st.kind, self.st.theta, X_, X_) - U.T @ U


def orc_ranked_rows(gpr, X, n_points, zeta):
    """Single-process reference ranking of a whole sample (oracle arithmetic)."""
    m, s, a = orc.predict_logexp(gpr.st, X, zeta=zeta)
    return orc.ranked_pool_select(gpr.st, X, m, s, a, n_points, zeta=zeta)[1]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gpry_b200 import parallel
    from gpry_b200.acquisition_functions import LogExp
    from gpry_b200.gp_acquisition import NORA
    res = {}
    assert parallel.size() == world and parallel.rank() == rank
    # strided split + merge round trip
    vals = np.arange(11, dtype=float) if rank == 0 else None
    mine = parallel.step_split(vals)
    assert np.array_equal(mine, np.arange(11, dtype=float)[rank::world])
    merged = parallel.merge_step_split(mine * 2)
    if rank == 0:
        assert np.array_equal(merged, 2 * np.arange(11, dtype=float))
    else:
        assert merged is None
    assert list(parallel.split_number_for_parallel_processes(5, 3)) == [2, 2, 1]
    assert list(parallel.split_number_for_parallel_processes(64, 8)) == [8] * 8
    assert parallel.bcast("x" if rank == 0 else None) == "x"
    assert parallel.allgather(rank) == list(range(world))
    assert parallel.max_scalar(float(rank)) == world - 1
    # ragged survivor all-gather
    n = 3 + rank
    a = np.arange(n, dtype=float) + 10 * rank
    rec = parallel.allgather_survivors(a, np.arange(n) + 100 * rank, a + 0.5, a + 0.25,
                                       np.outer(a, np.ones(2)))
    assert len(rec[0]) == sum(3 + r for r in range(world))
    assert np.array_equal(rec[1][:3], [0, 1, 2]) and rec[4].shape == (len(rec[0]), 2)
    lml, theta, best = parallel.best_fit_across_processes(float(rank), np.full(3, rank))
    assert best == world - 1 and lml == world - 1 and np.array_equal(theta, np.full(3, world - 1.))
    # compute_y_parallel + sharded NORA on the golden ranked-pool case
    g = load_golden("rbf_d2_n60")
    gpr = FakeDeviceGPR(oracle_state(g))
    Xp = golden_pool_candidates(g)
    y, s = parallel.compute_y_parallel(gpr, Xp if rank == 0 else None, None, None,
                                       ensure_sigma_y=True)
    if rank == 0:
        yo, so = orc.predict(gpr.st, Xp, return_std=True)
        # sharded vs full-batch BLAS calls differ by round-off only
        assert np.allclose(y, yo, rtol=1e-11, atol=1e-11) and np.allclose(s, so, rtol=1e-9, atol=1e-12)
        assert len(y) == len(Xp)
    nora = NORA(g["bounds"], acq_func=LogExp(zeta=g["zeta"]), kprime=64)
    n_points = int(g["pool_n_points"])
    X_pool, y_pool, acq_pool = nora.multi_add(gpr, n_points=n_points, X_mc=Xp)
    assert np.array_equal(X_pool, Xp[g["pool_idx_single_sort_acq"]])
    np.save(os.path.join(out_dir, f"pool_{rank}.npy"), X_pool)
    # sharded hand-over: every rank gives only ITS rows (no broadcast of the sample); second call
    # on the same shards skips the rows proposed by the first (by global index, on every rank)
    nora_s = NORA(g["bounds"], acq_func=LogExp(zeta=g["zeta"]), kprime=64)
    shard = np.ascontiguousarray(Xp[rank::world])
    Xs1, _, _ = nora_s.multi_add(gpr, n_points=n_points, X_shard=shard)
    assert np.array_equal(Xs1, X_pool)
    assert np.array_equal(nora_s.last_pool_idx, g["pool_idx_single_sort_acq"])
    Xs2, _, _ = nora_s.multi_add(gpr, n_points=n_points, X_shard=shard)
    assert not nora_s.last_new_sample
    assert not set(map(bytes, Xs2)) & set(map(bytes, Xs1))
    one = NORA(g["bounds"], acq_func=LogExp(zeta=g["zeta"]), kprime=64)
    ref2 = orc_ranked_rows(gpr, np.delete(Xp, g["pool_idx_single_sort_acq"], axis=0), n_points,
                           g["zeta"])
    assert np.array_equal(Xs2, ref2)
    # X_mc held by rank 0 only: tensor broadcast, no pickle
    nora_b = NORA(g["bounds"], acq_func=LogExp(zeta=g["zeta"]), kprime=64)
    Xb1, _, _ = nora_b.multi_add(gpr, n_points=n_points, X_mc=Xp if rank == 0 else None)
    assert np.array_equal(Xb1, X_pool)
    # K' small and many ranks' worth of survivors: only the best K' of the union are ranked
    # (the posterior covariance stays K' x K'), the exactness bound still holds
    nora_k = NORA(g["bounds"], acq_func=LogExp(zeta=g["zeta"]), kprime=16)
    Xk, _, _ = nora_k.multi_add(gpr, n_points=n_points, X_shard=shard)
    assert np.array_equal(Xk, X_pool)
    # per-rank generators: children of one SeedSequence (mpi.py:31-50)
    r1 = parallel.get_random_generator(123).uniform()
    both = parallel.allgather(r1)
    assert both[0] != both[1]
    assert both[rank] == np.random.default_rng(np.random.SeedSequence(123).spawn(world)[rank]).uniform()
    # BatchOptimizer: restarts split over the ranks (mpi.split_number_for_parallel_processes,
    # gp_acquisition.py:454-458), results gathered, every rank picks the same optimum and
    # appends the same lie
    from gpry_b200.gp_acquisition import BatchOptimizer
    from gpry_b200.preprocessing import Normalize_bounds
    opt = BatchOptimizer(g["bounds"], preprocessing_X=Normalize_bounds(g["bounds"]),
                         acq_func=LogExp(zeta=g["zeta"]), n_restarts_optimizer=5, verbose=0)
    gpr3 = FakeDeviceGPR(oracle_state(g))
    n0 = len(gpr3.X_train)
    Xb, yb, ab = opt.multi_add(gpr3, n_points=2, rng=np.random.default_rng(40 + rank))
    assert len(gpr3.X_train) == n0                      # the caller's regressor is untouched
    assert Xb.shape == (2, g["d"]) and np.all(np.isfinite(ab)) and gpr3.n_eval > 0
    lo, hi = g["bounds"][:, 0], g["bounds"][:, 1]
    assert np.all(Xb >= lo - 1e-12) and np.all(Xb <= hi + 1e-12)
    acq = LogExp(zeta=g["zeta"])
    assert abs(acq(Xb[:1], gpr3)[0] - ab[0]) < 1e-8 * max(1.0, abs(ab[0]))
    assert ab[0] >= acq(gpr3.X_train[-1:], gpr3)[0] - 1e-9      # restart 0 starts there (rank 0)
    np.save(os.path.join(out_dir, f"bopt_{rank}.npy"), np.concatenate([Xb.ravel(), yb, ab]))
    # restart-parallel hyper-parameter fit (Runner._fit_gpr_parallel, run.py:1238-1301) with the
    # real regressor class on an oracle-backed stand-in for the device: restarts split over
    # the ranks, (lml, theta) all-gathered, every rank ends with the winner's model
    import gpry_b200.device as dev_mod
    import gpry_b200.gpr as gpr_mod
    from fake_device import FakeDeviceGP
    ws = {}
    gpr_mod.DeviceGP = FakeDeviceGP
    gpr_mod.workspace = dev_mod.workspace = lambda device=0: ws.setdefault(device, FakeDeviceGP(device))
    from gpry_b200.preprocessing import Normalize_y
    z = np.load(os.path.join(ROOT, "tests", "golden", "fit_rbf_d2_n40.npz"))
    fit = gpr_mod.GaussianProcessRegressor(
        kernel="RBF", bounds=z["bounds"], noise_level=1e-2, n_restarts_optimizer=4,
        preprocessing_X=Normalize_bounds(z["bounds"]), preprocessing_y=Normalize_y(),
        account_for_inf=None, random_state=100 + rank, verbose=0)
    best_rank = parallel.fit_gpr_parallel(fit, z["X_train"], z["y_train"])
    thetas = parallel.allgather(np.array(fit.kernel_.theta))
    assert all(np.array_equal(t, thetas[0]) for t in thetas) and 0 <= best_rank < world
    lmls = parallel.allgather(float(fit.log_marginal_likelihood_value_))
    assert lmls[0] == lmls[1] and np.isfinite(lmls[0]) and fit.fitted
    preds = parallel.allgather(fit.predict(z["Xc"]))
    assert np.array_equal(preds[0], preds[1])
    # the reference's own fit of this case (golden): the parallel fit is at least as good
    assert lmls[0] >= float(z["lml_opt"]) - 1e-6 * abs(float(z["lml_opt"]))
    # same integer seed on every rank: the ranks must still draw DIFFERENT restart points
    # (children of one SeedSequence, mpi.py:31-50), and the seed is restored afterwards
    fit2 = gpr_mod.GaussianProcessRegressor(
        kernel="RBF", bounds=z["bounds"], noise_level=1e-2, n_restarts_optimizer=4,
        preprocessing_X=Normalize_bounds(z["bounds"]), preprocessing_y=Normalize_y(),
        account_for_inf=None, random_state=7, verbose=0)
    seen = {}
    orig = fit2._lockstep_optimization
    fit2._lockstep_optimization = lambda starts, b: (seen.setdefault("starts", np.array(starts)),
                                                     orig(starts, b))[1]
    parallel.fit_gpr_parallel(fit2, z["X_train"], z["y_train"])
    starts = parallel.allgather(seen["starts"])
    assert not np.array_equal(starts[0], starts[1]) and fit2.random_state == 7
    # a rank whose own fit ends at a non-positive-definite matrix still takes part in the
    # exchange and adopts the winner (instead of leaving the others in the all-gather)
    fit3 = gpr_mod.GaussianProcessRegressor(
        kernel="RBF", bounds=z["bounds"], noise_level=1e-2, n_restarts_optimizer=4,
        preprocessing_X=Normalize_bounds(z["bounds"]), preprocessing_y=Normalize_y(),
        account_for_inf=None, random_state=11, verbose=0)
    if rank == 1:
        real_fit = fit3.fit_gpr_hyperparameters

        def failing_fit(**kw):
            real_fit(**kw)
            raise np.linalg.LinAlgError("not positive definite (simulated)")
        fit3.fit_gpr_hyperparameters = failing_fit
    winner = parallel.fit_gpr_parallel(fit3, z["X_train"], z["y_train"])
    assert winner == 0
    th3 = parallel.allgather(np.array(fit3.kernel_.theta))
    assert np.array_equal(th3[0], th3[1]) and fit3.fitted
    p3 = parallel.allgather(fit3.predict(z["Xc"]))
    assert np.array_equal(p3[0], p3[1])
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a = np.load(tmp_path / "pool_0.npy")
    b = np.load(tmp_path / "pool_1.npy")
    assert np.array_equal(a, b)      # every rank ends with the same pool (as after the bcast)
    assert np.array_equal(np.load(tmp_path / "bopt_0.npy"), np.load(tmp_path / "bopt_1.npy"))


def test_serial_fallbacks():
    from gpry_b200 import parallel
    assert parallel.size() == 1 and parallel.rank() == 0 and parallel.is_main_process()
    v = np.arange(5.0)
    assert parallel.step_split(v) is v and parallel.merge_step_split(v) is v
    assert parallel.bcast(3) == 3 and parallel.allgather(3) == [3]
    assert parallel.max_scalar(2.5) == 2.5
