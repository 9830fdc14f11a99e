"""Two PROCESSES sharing ONE B200 (gloo for the collectives, both ranks compute on cuda:0):
the multi-rank acquisition step and the restart-parallel fit with the real library on a box
that has a single GPU.  NCCL refuses two ranks on one device, so the exchange goes over gloo;
the device work is the same code the NCCL runs use (bench.py at N > 1, test_gpu_multi.py)."""
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
    import torch.distributed as dist
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), LOCAL_RANK="0")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from conftest import golden_pool_candidates, load_golden
    from gpry_b200 import parallel
    from gpry_b200.acquisition_functions import LogExp
    from gpry_b200.gp_acquisition import NORA
    from gpry_b200.gpr import GaussianProcessRegressor
    from gpry_b200.preprocessing import Normalize_bounds, Normalize_y
    from test_gpu_gpr import make_gpr
    # --- sharded acquisition step: each rank hands in ITS rows only; the merged pool must be
    # the reference's single-process ranking of the whole sample (golden), on every rank
    for name in ("rbf_d8_n300", "rbf_d8_n700_pool"):
        g = load_golden(name)
        if "pool_M" not in g:
            continue
        gpr = make_gpr(g)
        Xp = golden_pool_candidates(g)
        n_points = int(g["pool_n_points"])
        nora = NORA(g["bounds"], acq_func=LogExp(zeta=g["zeta"]), kprime=128)
        shard = np.ascontiguousarray(Xp[rank::world])
        X_pool, y_pool, acq_pool = nora.multi_add(gpr, n_points=n_points, X_shard=shard)
        assert np.array_equal(nora.last_pool_idx, g["pool_idx_single_sort_acq"]), name
        assert np.array_equal(X_pool, Xp[g["pool_idx_single_sort_acq"]])
        pools = parallel.allgather((X_pool, y_pool, acq_pool))
        assert all(np.array_equal(p[0], pools[0][0]) and np.array_equal(p[1], pools[0][1])
                   and np.array_equal(p[2], pools[0][2]) for p in pools)
        # same shards again: the rows just proposed are skipped on the device, by global index
        X2, _, _ = nora.multi_add(gpr, n_points=n_points, X_shard=shard)
        assert not nora.last_new_sample
        assert not set(map(bytes, X2)) & set(map(bytes, X_pool))
        # single-process ranking of the reduced sample, computed by rank 0 alone on its GPU
        if rank == 0:
            from gpry_b200.gp_acquisition import ranked_pool_from_scores
            from functools import partial
            Xr = np.delete(Xp, g["pool_idx_single_sort_acq"], axis=0)
            a, i, m, s, Xs = gpr.predict_logexp_topk(Xr, float(g["zeta"]), 512)
            f = partial(LogExp.f, baseline=gpr.y_max, noise_level=gpr.noise_level,
                        zeta=float(g["zeta"]))
            pool = ranked_pool_from_scores(gpr, Xs, m, s, a, n_points, f)
            ref2 = pool.copy(drop_empty=True).X[:n_points]
        else:
            ref2 = None
        ref2 = parallel.bcast(ref2)
        assert np.array_equal(X2, ref2), name
    # --- restart-parallel fit: distinct starting points per rank, one winner, same state
    z = np.load(os.path.join(root, "tests", "golden", "fit_rbf_d2_n40.npz"))
    fit = GaussianProcessRegressor(
        kernel="RBF", bounds=z["bounds"], noise_level=1e-2, n_restarts_optimizer=6,
        preprocessing_X=Normalize_bounds(z["bounds"]), preprocessing_y=Normalize_y(),
        account_for_inf=None, random_state=7, verbose=0)
    best_rank = parallel.fit_gpr_parallel(fit, z["X_train"], z["y_train"])
    thetas = parallel.allgather(np.array(fit.kernel_.theta))
    assert all(np.array_equal(t, thetas[0]) for t in thetas) and 0 <= best_rank < world
    alphas = parallel.allgather(np.array(fit.alpha_))
    assert np.array_equal(alphas[0], alphas[1])           # bit-identical factorisation
    preds = parallel.allgather(fit.predict(z["Xc"], return_std=True))
    assert np.array_equal(preds[0][0], preds[1][0]) and np.array_equal(preds[0][1], preds[1][1])
    assert fit.log_marginal_likelihood_value_ >= float(z["lml_opt"]) - 1e-6 * abs(float(z["lml_opt"]))
    np.save(os.path.join(out_dir, f"ok_{rank}.npy"), np.array([best_rank]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_one_gpu(tmp_path):
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok_0.npy") and os.path.exists(tmp_path / "ok_1.npy")
