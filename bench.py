#!/usr/bin/env python
"""
bench.py -- GP predict + LogExp + ranked-pool pre-selection throughput (candidates / s).

Workload (BASELINE.json configs[2], the one the metric is quoted on): N_train = 2000, d = 12,
ConstantKernel x RBF at fixed theta, 12.5e6 synthetic candidates PER GPU (weak scaling:
8 GPUs -> 10^8), K' = 1024 survivors per GPU merged through an NCCL all-gather.

One step = one pass of the hot path over the rank's candidate pool:
    K* build -> variance contraction (exact INT8 split on tcgen05, or FP64 DMMA with
    --contract fp64) -> finish/LogExp -> top-K' -> all-gather+merge.

  python bench.py [--gpus N] [--steps K] [--warmup W]          our arm (CUDA, sm_100a)
  python bench.py --impl reference ...                         reference arm: the CPU port of
                                                               GPry's own path (oracle/), host cores
Prints ONE JSON line (see README / DESIGN.md "Measurement").
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "gp_predict_logexp_candidates_per_sec"
UNIT = "candidates/s"
FP64_DGEMM_FALLBACK_TFLOPS = 35.4   # cublasDgemm 8192^3 on this pool (profiles/r01_fp64_peaks.txt)


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=5)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--ntrain", type=int, default=2000)
    p.add_argument("--dim", type=int, default=12)
    p.add_argument("--pool", type=int, default=12_500_000, help="candidates per GPU")
    p.add_argument("--kp", type=int, default=1024, help="survivors per GPU (K')")
    p.add_argument("--e2e-steps", type=int, default=2)
    p.add_argument("--cpu-seconds", type=float, default=12.0)
    p.add_argument("--cpu-chunk", type=int, default=20000)
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--no-secondary", action="store_true")
    p.add_argument("--contract", choices=("fp64", "int8"), default="int8",
                   help="variance contraction: FP64 DMMA or the exact INT8 split (tcgen05)")
    return p.parse_args()


# ------------------------------------------------------------------------------------------
# synthetic workload; identical for both arms
# ------------------------------------------------------------------------------------------
def synthetic_problem(N, d, seed=1234):
    rng = np.random.default_rng(seed)
    X = rng.uniform(size=(N, d))
    y = -0.5 * np.sum(((X - 0.5) / 0.15) ** 2, axis=1)
    ell = 0.5 if d <= 8 else (1.0 if d <= 16 else 1.5)
    theta = np.log(np.concatenate([[1.0], np.full(d, ell)]))
    bounds = np.array([[0.0, 1.0]] * d)
    return X, y, theta, bounds


def candidates_host(M, d, seed):
    return np.random.default_rng(seed).uniform(size=(M, d))


def use_all_host_cores():
    """torchrun exports OMP_NUM_THREADS=1; the CPU arm should use every host core it can."""
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=os.cpu_count())
    except Exception:
        pass


def cpu_thread_info():
    try:
        from threadpoolctl import threadpool_info
        blas = [i for i in threadpool_info() if i.get("user_api") == "blas"]
        n = max([i.get("num_threads", 1) for i in blas] or [os.cpu_count()])
        return int(n), ",".join(sorted({str(i.get("internal_api")) for i in blas}))
    except Exception:
        return os.cpu_count(), "unknown"


def time_cpu_port(N, d, chunk, budget_s, max_chunks=16):
    """Times the CPU port of the reference path (oracle/gp_oracle.py: cdist + exp + dtrmm +
    einsum, all host cores through BLAS) on a bounded sample of the same workload."""
    from oracle import gp_oracle as orc
    use_all_host_cores()
    X, y, theta, bounds = synthetic_problem(N, d)
    st = orc.GPState("rbf", theta, X, y, bounds=bounds)
    Xc = candidates_host(chunk, d, 4321)
    orc.predict_logexp(st, Xc[:2000])   # warm-up
    t_used, n_done, best = 0.0, 0, None
    while (t_used < budget_s and n_done < max_chunks) or n_done == 0:
        t0 = time.perf_counter()
        out = orc.predict_logexp(st, Xc)
        dt = time.perf_counter() - t0
        t_used += dt
        n_done += 1
        best = dt if best is None else min(best, dt)
    return chunk * n_done / t_used, chunk / best, n_done, st, Xc, out


# ------------------------------------------------------------------------------------------
# reference arm
# ------------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import gp_oracle as orc
    use_all_host_cores()
    N, d = args.ntrain, args.dim
    X, y, theta, bounds = synthetic_problem(N, d)
    st = orc.GPState("rbf", theta, X, y, bounds=bounds)
    sample = 2 * args.cpu_chunk
    Xc = candidates_host(sample, d, 4321)

    def step():
        for i in range(0, sample, args.cpu_chunk):
            orc.predict_logexp(st, Xc[i:i + args.cpu_chunk])

    for _ in range(max(1, min(args.warmup, 2))):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = sample * args.steps / dt
    cores, blas = cpu_thread_info()
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"GP predict+LogExp N_train={N} d={d} RBF, CPU port of the "
                               "reference path (cdist+exp+dtrmm+einsum)",
                   "sample_candidates_per_step": sample, "chunk": args.cpu_chunk},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{sample} candidates/step in chunks of {args.cpu_chunk}, "
                                   f"BLAS={blas}, os.cpu_count={os.cpu_count()}"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "200", "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.f.read().splitlines():
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])), mx.append(float(parts[1])), pw.append(float(parts[2]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.f.name)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)),
                       power_w_max=float(max(pw)), reasons=sorted(reasons), samples=len(sm))
        return out


# ------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------
def fit_state_for_bench(dev, N, d, kind="rbf"):
    """Training state for the synthetic problem: Normalize_bounds / Normalize_y scalars on the
    host (O(N)), kernel matrix + Cholesky + L^-1 + alpha on the GPU (gpry_factorize)."""
    X, y, theta, bounds = synthetic_problem(N, d)
    y_mean, y_std = float(np.mean(y)), float(np.std(y))
    noise_level = 1e-2
    X_ = (X - bounds[:, 0]) / (bounds[:, 1] - bounds[:, 0])
    y_ = (y - y_mean) / y_std
    noise2 = np.full(N, (noise_level / y_std) ** 2)
    L, V, alpha_, _, info = dev.factorize(kind, X_, noise2, y_, theta, want_L=False)
    assert info == 0, "synthetic kernel matrix not positive definite"
    clip_hi = 1.1 * y.max() - 0.1 * y.min()
    model = dict(kind=kind, X_=X_, alpha_=alpha_, V=V, c=float(np.exp(theta[0])),
                 ell=np.exp(theta[1:]), x_min=bounds[:, 0], x_width=bounds[:, 1] - bounds[:, 0],
                 y_mean=y_mean, y_std=y_std, clip_hi=clip_hi, y_max=float(y.max()),
                 noise_level=noise_level, zeta=float(d) ** -0.85)
    return model


def secondary_figures(dev, dev_t, world=1, dist=None):
    """BASELINE.json configs[3] and configs[4]: every GPU works on its own share (restarts /
    proposals split across ranks, no exchange); figures are whole-job (max time over ranks)."""
    import torch
    out = {}

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev_t)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # config D: LML + gradient, N_train = 4000, d = 20, 8 restarts' worth of theta per call
    N, d, B = 4000, 20, 8
    X, y, theta, bounds = synthetic_problem(N, d)
    y_ = (y - y.mean()) / y.std()
    noise2 = np.full(N, (1e-2 / y.std()) ** 2)
    thetas = theta + 0.1 * np.random.default_rng(7).standard_normal((B, d + 1))
    dev.lml_batched("rbf", X, noise2, y_, thetas)       # warm-up: buffers sized for this batch
    dt = 1e9
    for _ in range(3):
        t0 = time.perf_counter()
        lml, grad, info = dev.lml_batched("rbf", X, noise2, y_, thetas)
        dt = min(dt, time.perf_counter() - t0)
    dt = max_over_ranks(dt)
    flop = N ** 3 + (3 * d + 4 + 2 * (d + 1)) * N ** 2 / 2
    out["lml_grad"] = {"n_train": N, "dim": d, "restarts_per_gpu": B, "restarts_total": B * world,
                       "evals_per_s": B * world / dt, "ms_per_eval_per_gpu": dt / B * 1e3,
                       "tflops_algorithmic_per_gpu": flop * B / dt * 1e-12,
                       "all_pd": bool(np.all(info == 0))}
    # config E: mean-only proposals (surrogate MCMC), N_train = 2000, d = 16, 10^7 per step
    N, d, M = 2000, 16, 10_000_000
    model = fit_state_for_bench(dev, N, d)
    dev.upload(model["kind"], model["X_"], model["alpha_"], model["V"], model["c"], model["ell"],
               model["x_min"], model["x_width"], model["y_mean"], model["y_std"],
               model["clip_hi"])
    Xd = torch.rand((M, d), dtype=torch.float64, device=dev_t)
    s = torch.cuda.current_stream()
    dev.predict(Xd, return_std=False, stream=s)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(3):
        dev.predict(Xd, return_std=False, stream=s)
    e1.record()
    torch.cuda.synchronize()
    ms = max_over_ranks(e0.elapsed_time(e1) / 3)
    del Xd
    # same model: batched gradients for the acquisition optimiser's restarts (host in / out,
    # as the lock-step L-BFGS-B drivers call it) and the one-point latency path
    Xg = np.random.default_rng(5).uniform(size=(256, d))
    dev.predict_grad(Xg)
    dtg = 1e9
    for _ in range(5):
        t0 = time.perf_counter()
        dev.predict_grad(Xg)
        dtg = min(dtg, time.perf_counter() - t0)
    dev.predict(Xg[:1], return_std=True)
    t0 = time.perf_counter()
    for _ in range(50):
        dev.predict(Xg[:1], return_std=True)
    dt1 = (time.perf_counter() - t0) / 50
    out["grad_batch"] = {"n_train": N, "dim": d, "points_per_call": 256,
                         "ms_per_call": max_over_ranks(dtg) * 1e3,
                         "what": "mean, std, d mean/dx, d std/dx per point, host to host"}
    out["latency"] = {"n_train": N, "dim": d, "what": "predict(1 point, return_std) host to host",
                      "us_per_call": max_over_ranks(dt1) * 1e6}
    # config B (and its Matern-5/2 repeat): N_train = 1000, d = 8, 10^6 candidates, mean+std+acq
    for kind in ("rbf", "matern25"):
        Nb, db, Mb = 1000, 8, 1_000_000
        mb = fit_state_for_bench(dev, Nb, db, kind)
        dev.upload(mb["kind"], mb["X_"], mb["alpha_"], mb["V"], mb["c"], mb["ell"], mb["x_min"],
                   mb["x_width"], mb["y_mean"], mb["y_std"], mb["clip_hi"])
        Xb = torch.rand((Mb, db), dtype=torch.float64, device=dev_t)
        dev.predict_logexp(Xb, mb["zeta"], mb["noise_level"], mb["y_max"], stream=s)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(3):
            dev.predict_logexp(Xb, mb["zeta"], mb["noise_level"], mb["y_max"], stream=s)
        e1.record()
        torch.cuda.synchronize()
        msb = max_over_ranks(e0.elapsed_time(e1) / 3)
        out["config_b_" + kind] = {"n_train": Nb, "dim": db, "candidates_per_gpu": Mb,
                                   "candidates_per_s": Mb * world / msb * 1e3, "ms_per_step": msb}
        del Xb
    out["mean_only"] = {"n_train": N, "dim": d, "proposals_per_step_per_gpu": M,
                        "proposals_per_s": M * world / ms * 1e3, "ms_per_step": ms}
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist
    from gpry_b200 import DeviceGP

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev_t = torch.device("cuda", local)
    N, d, M, Kp = args.ntrain, args.dim, args.pool, args.kp
    dev = DeviceGP(local)
    dev.set_contract_mode(args.contract)

    # ---- model: fitted on rank 0, broadcast once (per refit), uploaded on every GPU ----
    if rank == 0:
        model = fit_state_for_bench(dev, N, d)
    t_bcast_ms = 0.0
    if world > 1:
        meta = [model if rank == 0 else None]
        big = {}
        if rank == 0:
            big = {k: torch.from_numpy(np.ascontiguousarray(model[k])).to(dev_t)
                   for k in ("X_", "alpha_", "V")}
            meta = [{k: v for k, v in model.items() if k not in big}]
        dist.broadcast_object_list(meta, src=0)
        if rank != 0:
            model = meta[0]
            big = {"X_": torch.empty((N, d), dtype=torch.float64, device=dev_t),
                   "alpha_": torch.empty(N, dtype=torch.float64, device=dev_t),
                   "V": torch.empty((N, N), dtype=torch.float64, device=dev_t)}
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in ("X_", "alpha_", "V"):
            dist.broadcast(big[k], src=0)
        e1.record()
        torch.cuda.synchronize()
        t_bcast_ms = e0.elapsed_time(e1)
        if rank != 0:
            for k in big:
                model[k] = big[k].cpu().numpy()
    dev.upload(model["kind"], model["X_"], model["alpha_"], model["V"], model["c"], model["ell"],
               model["x_min"], model["x_width"], model["y_mean"], model["y_std"],
               model["clip_hi"])
    zeta, sig, ymax = model["zeta"], model["noise_level"], model["y_max"]

    # ---- candidates: this rank's shard, generated on the device (Philox), resident in HBM ----
    gen = torch.Generator(device=dev_t)
    gen.manual_seed(4321 + rank)
    Xd = torch.rand((M, d), dtype=torch.float64, device=dev_t, generator=gen)
    idx_offset = rank * M
    stream = torch.cuda.current_stream()

    def merge(acq, idx):
        """All-gather of the per-GPU survivor lists + final top-K' (every rank ends with the
        same merged list, as after gp_acquisition.py:1190 bcast)."""
        if world == 1:
            return acq, idx
        ga = torch.empty(world * Kp, dtype=torch.float64, device=dev_t)
        gi = torch.empty(world * Kp, dtype=torch.int64, device=dev_t)
        dist.all_gather_into_tensor(ga, acq.contiguous())
        dist.all_gather_into_tensor(gi, idx.contiguous())
        vals, pos = dev.topk(ga, Kp, stream=stream)
        return vals, gi[pos]

    def step():
        acq, idx, mean, std, _ = dev.predict_logexp_topk(
            Xd, zeta, sig, ymax, Kp, idx_offset=idx_offset, stream=stream, device_out=True,
            want_X=False)
        return merge(acq, idx)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    dev.set_profiling(True)
    dev.timings(reset=True)
    barrier()
    clocks = ClockSampler(local) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        top_acq, top_idx = step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clk = clocks.stop() if clocks else {}
    tm = dev.timings(reset=True)
    dev.set_profiling(False)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev_t)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * M * args.steps / (ms * 1e-3)

    # ---- end-to-end through the public API with HOST buffers (pinned) ----
    e2e = None
    if not args.no_e2e:
        Xh = torch.empty((M, d), dtype=torch.float64, pin_memory=True)
        Xh.copy_(Xd)

        def step_e2e():
            acq, idx, mean, std, Xo = dev.predict_logexp_topk(
                Xh, zeta, sig, ymax, Kp, idx_offset=idx_offset, stream=stream)
            if world > 1:
                a, i = merge(torch.from_numpy(acq).to(dev_t), torch.from_numpy(idx).to(dev_t))
                return a.cpu().numpy(), i.cpu().numpy()
            return acq, idx

        step_e2e()
        barrier()
        e0.record()
        for _ in range(args.e2e_steps):
            step_e2e()
        e1.record()
        barrier()
        ms2 = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms2], dtype=torch.float64, device=dev_t)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms2 = float(t.item())
        e2e = {"value": world * M * args.e2e_steps / (ms2 * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": world * M * d * 8,
               "d2h_bytes_per_step": world * Kp * (4 + d) * 8,
               "ms_per_step": ms2 / args.e2e_steps}
        del Xh

    # ---- roofline of the dominant kernel (variance contraction, FP64 tensor pipe) ----
    n_launch = max(tm["contract_launches"], 1.0)
    cands_per_launch = M * args.steps / n_launch
    flop_per_cand = N * (N + 1) + 2 * N          # V k* (lower tri.) + sum of squares
    contract_ms_per_launch = tm["contract_ms"] / n_launch
    achieved = flop_per_cand * cands_per_launch / (contract_ms_per_launch * 1e-3) * 1e-12
    peak, peak_src = FP64_DGEMM_FALLBACK_TFLOPS, "cublasDgemm 8192^3, profiles/r01_fp64_peaks.txt"
    try:   # measure the FP64 tensor peak live (MEASURED_PEAKS.json has no FP64 entry)
        a = torch.randn(8192, 8192, dtype=torch.float64, device=dev_t)
        b = torch.randn(8192, 8192, dtype=torch.float64, device=dev_t)
        (a @ b)
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(3):
            e0.record()
            (a @ b)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        peak, peak_src = 2 * 8192 ** 3 / best * 1e-9, "cublasDgemm 8192^3 measured in this run"
        del a, b
    except Exception:
        pass
    int8 = args.contract == "int8" and N > 384                  # the library's own criterion
    prof_name = "r01_oz_contract_ncu.json" if int8 else "r01_contract_ncu.json"
    traffic, traffic_src = None, None
    try:   # dram bytes per launch of the same kernel from the committed ncu --set full capture
        with open(os.path.join(ROOT, "profiles", prof_name)) as f:
            prof = json.load(f)
        if N == 2000 and d == 12:
            traffic = prof["traffic_bytes_per_launch"] * (cands_per_launch / (296 * 128))
            traffic_src = prof["source"]
    except Exception:
        pass
    stage_ms = {k: tm[k] / args.steps for k in ("build_ms", "contract_ms", "finish_ms", "topk_ms")}
    if int8:
        # The contraction runs on the INT8 tensor pipe: each FP64 multiply-add of the algorithm
        # costs 28 int8 multiply-adds (7 x 7 digit products with p + q <= 6), so the pipe's
        # measured rate / 28 is the ceiling in algorithmic FP64 flop/s.
        int8_peak = dev.int8_peak_tops()
        pairs = 28
        n_rb = -(-N // 128)          # 128-row blocks of V, k chunks of 32 up to the block's last row
        executed = pairs * 2.0 * 128 * 128 * 32 * sum(4 * (rb + 1) for rb in range(n_rb)) / 128
        roofline = {"bound": "tensor", "achieved": achieved, "peak": int8_peak / pairs,
                    "unit": "TFLOP/s", "frac": achieved / (int8_peak / pairs), "traffic": traffic,
                    "traffic_source": traffic_src,
                    "peak_source": "tcgen05.mma kind::i8 128x256x32 issue rate measured in this run "
                                   f"({int8_peak:.0f} TOPS) / 28 int8 products per FP64 product",
                    "kernel": "oz2_contract_kernel (exact INT8 split, two passes of 128x128x32 tcgen05.mma kind::i8, TMEM)",
                    "flop_per_candidate": flop_per_cand,
                    "int8_tops_executed": executed * cands_per_launch
                    / (contract_ms_per_launch * 1e-3) * 1e-12,
                    "int8_tops_peak": int8_peak,
                    "fp64_dgemm_tflops": peak, "fp64_dgemm_source": peak_src,
                    "vs_fp64_tensor_peak": achieved / peak,
                    "ms_per_launch": contract_ms_per_launch, "stage_ms_per_step": stage_ms}
    else:
        roofline = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                    "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                    "peak_source": peak_src,
                    "kernel": "var_contract_kernel (FP64 DMMA.8x8x4)",
                    "flop_per_candidate": flop_per_cand,
                    "ms_per_launch": contract_ms_per_launch, "stage_ms_per_step": stage_ms}

    # ---- every rank checks a sample of ITS OWN shard against the oracle (first / middle / last
    # tiles), max error reduced over ranks ----
    shard_err = None
    if not args.no_cpu_baseline:
        from oracle import gp_oracle as orc
        Xp, yp, thp, bp = synthetic_problem(N, d)
        st_o = orc.GPState("rbf", thp, Xp, yp, bounds=bp)
        pick = torch.cat([torch.arange(0, 500), torch.arange(M // 2, M // 2 + 500),
                          torch.arange(M - 500, M)]).to(dev_t)
        Xs = Xd[pick].contiguous()
        mg, sg, ag = dev.predict_logexp(Xs, zeta, sig, ymax, stream=stream)
        mo, so, ao = orc.predict_logexp(st_o, Xs.cpu().numpy())
        errs = torch.tensor([np.max(np.abs(mg.cpu().numpy() - mo)) / st_o.y_std,
                             np.max(np.abs(sg.cpu().numpy() ** 2 - so ** 2)) / st_o.y_std ** 2],
                            dtype=torch.float64, device=dev_t)
        if world > 1:
            dist.all_reduce(errs, op=dist.ReduceOp.MAX)
        shard_err = {"n_per_rank": int(len(pick)), "mean_err": float(errs[0]),
                     "var_err": float(errs[1])}

    # ---- agreement check + CPU baseline (rank 0, N = 1 run only for the baseline) ----
    agreement, cpu_baseline = None, None
    if rank == 0 and not args.no_cpu_baseline:
        # the CPU baseline is timed at N = 1 only; larger runs just verify agreement
        thr, thr_best, n_chunks, st, Xc, (mo, so, ao) = time_cpu_port(
            N, d, args.cpu_chunk, args.cpu_seconds if world == 1 else 0.0)
        mg, sg, ag = dev.predict_logexp(Xc, zeta, sig, ymax)
        ok = np.isfinite(ao) & (so ** 2 - sig ** 2 > 1e-6 * st.y_std ** 2)
        agreement = {
            "n": int(len(Xc)),
            "mean_err": float(np.max(np.abs(mg - mo)) / st.y_std),
            "var_err": float(np.max(np.abs(sg ** 2 - so ** 2)) / st.y_std ** 2),
            "acq_err": float(np.max(np.abs(ag[ok] - ao[ok]))),
            "tolerance": 1e-10,
        }
        # ranked list vs an independent sort of the same device scores
        acq_full = dev.predict_logexp(Xd, zeta, sig, ymax, stream=stream)[2]
        ref_vals, ref_idx = torch.sort(acq_full, descending=True, stable=True)
        a1, i1, _, _, _ = dev.predict_logexp_topk(Xd, zeta, sig, ymax, Kp, stream=stream,
                                                  device_out=True, want_X=False)
        agreement["topk_identical"] = bool(torch.equal(i1, ref_idx[:Kp]))
        agreement["all_shards"] = shard_err
        agreement["verified"] = bool(agreement["mean_err"] < 1e-10
                                     and agreement["var_err"] < 1e-10
                                     and agreement["topk_identical"]
                                     and shard_err["mean_err"] < 1e-10
                                     and shard_err["var_err"] < 1e-10)
        cores, blas = cpu_thread_info()
        if world == 1:
            cpu_baseline = {"value": thr, "unit": UNIT, "cores": cores, "kind": "port",
                            "sample": f"{n_chunks} x {args.cpu_chunk} candidates of the same "
                                      f"workload (N_train={N}, d={d}); best chunk {thr_best:.0f} "
                                      f"cand/s; BLAS={blas}; os.cpu_count={os.cpu_count()}"}

    # ---- secondary figures of the same path (not the headline metric): LML+gradient
    # evaluations/s at config D and mean-only proposals/s at config E, this GPU only ----
    secondary = None
    if not args.no_secondary:     # every rank runs them; rank 0 reports the whole-job figures
        secondary = secondary_figures(dev, dev_t, world, dist if world > 1 else None)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"NORA ranked-pool scoring: predict mean+std + LogExp + "
                                   f"top-{Kp}, N_train={N}, d={d}, RBF, {M} candidates/GPU",
                       "candidates_per_gpu": M, "candidates_total": world * M, "n_train": N,
                       "dim": d, "kprime": Kp, "parallelism": f"candidate-sharded x{world}",
                       "contraction": args.contract,
                       "contraction_arithmetic": (
                           "f64 operands split exactly into 7 int8 digits each; int8 x int8 -> "
                           "int32 on the tensor cores (exact), recombined and squared in f64"
                           if args.contract == "int8" else "f64 tensor cores (DMMA)"),
                       "l2": "inputs (1.2 GB/GPU) and K* scratch (>500 MB) exceed the 126 MB L2",
                       "state_bcast_ms": t_bcast_ms},
            "e2e": e2e, "gpu_launches": int(tm["launches"]), "roofline": roofline,
            "cpu_baseline": cpu_baseline, "agreement": agreement, "clocks": clk,
            "secondary": secondary,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    dev.close()


def main():
    args = parse_args()
    # Only the JSON line may reach stdout: libraries (NCCL prints its version banner there) are
    # pointed at stderr for the duration of the run.
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    try:
        import io
        buf = io.StringIO()
        real_stdout, sys.stdout = sys.stdout, buf
        try:
            if args.impl == "reference":
                run_reference(args)
            else:
                run_ours(args)
        finally:
            sys.stdout = real_stdout
    finally:
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        os.close(saved_stdout)
    out = buf.getvalue()
    if out:
        sys.stdout.write(out)
        sys.stdout.flush()


if __name__ == "__main__":
    main()
