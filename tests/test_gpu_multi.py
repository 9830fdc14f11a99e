"""2-GPU test (NCCL): restart-parallel fit leaves every rank with the same, best model, and the
sharded NORA ranking equals the single-process ranking.  Skipped with fewer than 2 GPUs."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    from conftest import golden_pool_candidates, load_golden
    from gpry_b200 import parallel
    from gpry_b200.acquisition_functions import LogExp
    from gpry_b200.gp_acquisition import NORA
    from gpry_b200.gpr import GaussianProcessRegressor
    from gpry_b200.preprocessing import Normalize_bounds, Normalize_y
    from test_gpu_gpr import make_gpr
    # --- restart-parallel fit
    z = np.load(os.path.join(root, "tests", "golden", "fit_rbf_d2_n40.npz"))
    gpr = GaussianProcessRegressor(
        kernel="RBF", bounds=z["bounds"], noise_level=1e-2, n_restarts_optimizer=6,
        preprocessing_X=Normalize_bounds(z["bounds"]), preprocessing_y=Normalize_y(),
        account_for_inf=None, random_state=100 + rank, verbose=0)
    best_rank = parallel.fit_gpr_parallel(gpr, z["X_train"], z["y_train"])
    thetas = parallel.allgather(gpr.kernel_.theta)
    assert all(np.array_equal(t, thetas[0]) for t in thetas)
    assert gpr.fitted and np.isfinite(gpr.log_marginal_likelihood_value_)
    m = gpr.predict(z["Xc"])
    ms = parallel.allgather(m)
    assert all(np.array_equal(x, ms[0]) for x in ms)
    # --- sharded NORA == golden single-process ranking
    g = load_golden("rbf_d8_n300")
    gpr2 = make_gpr(g)
    Xp = golden_pool_candidates(g)
    nora = NORA(g["bounds"], acq_func=LogExp(zeta=g["zeta"]), kprime=128)
    X_pool, y_pool, acq_pool = nora.multi_add(gpr2, n_points=int(g["pool_n_points"]), X_mc=Xp)
    assert np.array_equal(X_pool, Xp[g["pool_idx_single_sort_acq"]])
    # every rank hands in only its rows; the exchange is gpry_allgather_topk (NCCL inside the
    # library); second call on the same shards skips what the first proposed
    comm = parallel.device_comm()
    assert comm is not None and comm.comm_info()["size"] == world
    shard = np.ascontiguousarray(Xp[rank::world])
    nora2 = NORA(g["bounds"], acq_func=LogExp(zeta=g["zeta"]), kprime=128)
    Xs1, _, _ = nora2.multi_add(gpr2, n_points=int(g["pool_n_points"]), X_shard=shard)
    assert np.array_equal(nora2.last_pool_idx, g["pool_idx_single_sort_acq"])
    Xs2, _, _ = nora2.multi_add(gpr2, n_points=int(g["pool_n_points"]), X_shard=shard)
    assert not set(map(bytes, Xs2)) & set(map(bytes, Xs1))
    # --- gpry_allgather_topk against the tensor-collective + numpy merge
    rng = np.random.default_rng(7 + rank)
    n_loc, Kq, dd = 50 + 13 * rank, 64, 3
    a = np.sort(rng.normal(size=n_loc))[::-1].copy()
    a[-3:] = -np.inf
    i = (np.arange(n_loc) * world + rank).astype(np.int64)
    m, s_, Xr = rng.normal(size=n_loc), rng.uniform(size=n_loc), rng.uniform(size=(n_loc, dd))
    got = comm.allgather_topk(a, i, m, s_, Xr, Kq)
    ua, ui, um, us, uX = parallel.allgather_survivors(a, i, m, s_, Xr)
    order = np.lexsort((ui, -ua))
    assert np.array_equal(got[1], ui[order[:Kq]]) and np.array_equal(got[0], ua[order[:Kq]])
    assert np.array_equal(got[2], um[order[:Kq]]) and np.array_equal(got[3], us[order[:Kq]])
    assert np.array_equal(got[4], uX[order[:Kq]]) and got[5] == ua[order[Kq]]
    # --- gpry_bcast_state: rank 0's model -> a fresh state on the other GPU, same predictions
    from gpry_b200 import DeviceGP
    probe = DeviceGP(rank)
    probe.comm_share(comm)
    src = gpr2._device_state()
    if rank == 0:
        src.comm_share(comm)
        src.bcast_state(0)
    else:
        probe.bcast_state(0)
    Xq = g["Xc"]
    mine = (src if rank == 0 else probe).predict(Xq, return_std=True)
    both = parallel.allgather(mine)
    assert np.array_equal(both[0][0], both[1][0]) and np.array_equal(both[0][1], both[1][1])
    probe.close()
    # --- BatchOptimizer: restarts split over the ranks, every rank returns the same batch
    from gpry_b200.gp_acquisition import BatchOptimizer
    opt = BatchOptimizer(g["bounds"], preprocessing_X=Normalize_bounds(g["bounds"]),
                         n_restarts_optimizer=6, verbose=0)
    Xb, yb, ab = opt.multi_add(gpr2, n_points=2, rng=np.random.default_rng(50 + rank))
    both = parallel.allgather((Xb, yb, ab))
    assert all(np.array_equal(b[0], both[0][0]) and np.array_equal(b[2], both[0][2])
               for b in both)
    assert Xb.shape == (2, g["d"]) and np.all(np.isfinite(ab))
    np.save(os.path.join(out_dir, f"ok_{rank}.npy"), np.array([best_rank]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_gpus(tmp_path):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok_0.npy") and os.path.exists(tmp_path / "ok_1.npy")
