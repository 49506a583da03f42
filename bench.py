#!/usr/bin/env python
"""Benchmark of the spatial_gradient hot path (BASELINE.json metric: posterior-sample x grid-point / s).

    python bench.py --gpus N --steps K --warmup W            # one process per GPU (torchrun for N > 1)
    python bench.py --impl reference --steps K --warmup W    # the restated reference algorithm on host cores

A "step" is one pass of the hot path over a FIXED batch of synthetic posterior samples (strong scaling: the
same S_total samples per step at every N, split into contiguous shards over the ranks as the reference's mclapply
does): covariance assembly -> batched Cholesky -> triangular inverse -> cross-covariances -> fused TRMM + Gram +
5x5 Cholesky + draw for every grid point, then the single NCCL gather of the draws to rank 0 and the batch-median /
HPD summary there (both inside the timed region).
`value` is measured with inputs resident in HBM (one process per GPU, torch.distributed NCCL gather); `e2e` is the
same step through the C ABI R binds -- `spw_gradient_run` with HOST buffers and `ngpu = N` driven from rank 0
(in-process sharding, grouped NCCL send/recv gather, summary, draws and estimates copied back to the host).
Rank 0 also recomputes one gathered sample alone and requires bit-identical draws (`parity_checked`).
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "spatial_gradient posterior-sample x grid-point / s"
UNIT = "sample*gridpoint/s"

WORKLOADS = {
    # BASELINE.json configs[4] (the configuration the metric's 2/4/8-GPU figures are quoted on; it fits one GPU)
    "c5": dict(name="C5: synthetic N=5000 uniform coords, matern2, 100x100 grid", n=5000, cov="matern2", gside=100,
               phi=(100.0, 200.0), samples=64, nbatch=8, g_cpu=16),
    # BASELINE.json configs[3]
    "c4": dict(name="C4: synthetic N=2000 uniform coords, gaussian, 60x60 grid", n=2000, cov="gaussian", gside=60,
               phi=(3000.0, 6000.0), samples=512, nbatch=8, g_cpu=32),
    # BASELINE.json configs[1] / configs[2] at their sizes (synthetic coordinates; the real datasets are test fixtures)
    "c2": dict(name="C2-size: N=155 (meuse), matern1, 50x50 grid", n=155, cov="matern1", gside=50,
               phi=(2.0, 8.0), samples=4000, nbatch=100, g_cpu=2500),
    "c3": dict(name="C3-size: N=506 (boston.c), matern2, 134 grid points (12x12 here)", n=506, cov="matern2", gside=12,
               phi=(30.0, 60.0), samples=4000, nbatch=100, g_cpu=144),
    # BASELINE.json configs[0] (README synthetic)
    "c1": dict(name="C1: README synthetic N=100, matern2, 19x19 grid", n=100, cov="matern2", gside=19,
               phi=(10.0, 30.0), samples=4000, nbatch=100, g_cpu=361),
}


def ncomp(cov):
    return 2 if cov == "matern1" else 5


def make_inputs(w, S, seed0=0):
    """Synthetic inputs of SURVEY section 8d (NumPy Philox; shared by the GPU path, the oracle and the CPU baseline)."""
    P = lambda s: np.random.Generator(np.random.Philox(seed0 + s))   # noqa: E731
    n, c = w["n"], ncomp(w["cov"])
    coords = P(1).random((n, 2))
    gs = np.linspace(0.05, 0.95, w["gside"])
    grid = np.array([[a, b] for b in gs for a in gs])
    r3 = P(3)
    sig2 = r3.uniform(0.5, 2.0, S)
    phi = r3.uniform(w["phi"][0], w["phi"][1], S)
    f = 10 * (np.sin(3 * np.pi * coords[:, 0]) + np.cos(3 * np.pi * coords[:, 1]))
    z = (f - f.mean())[None, :] + 0.3 * P(4).standard_normal((S, n))
    eps = P(5).standard_normal((S, grid.shape[0], c))
    return coords, grid, phi, sig2, z, eps


def flops_per_unit(w):
    """Algorithmic flops per (sample, grid point) -- SURVEY section 8d: c N^2 (TRMM) + (N^3/3 potrf + N^3/3 trtri
    + N^2) / G + O(cN)."""
    n, c, G = w["n"], ncomp(w["cov"]), w["gside"] ** 2
    return c * n * (n + 1.0) + (c * (c + 1) + 2 * c) * n + (2.0 * n ** 3 / 3.0 + n * n) / G


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                pw.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        busy = [s for s, p in zip(sm, pw) if p > 0.5 * max(pw)] if pw else sm
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def config_of(w, world, S_total):
    """The `config` object -- identical for our arm and the reference arm so that the driver compares like with like."""
    G = w["gside"] ** 2
    return {"workload": w["name"], "cov_type": w["cov"], "n": w["n"], "grid_points": G,
            "samples_per_step": S_total, "units_per_step": S_total * G, "nbatch": w["nbatch"],
            "flops_per_unit": flops_per_unit(w), "parallelism": "sample-sharded x%d" % world,
            "l2": "inputs larger than L2: every step streams %.1f GB of factors and cross-covariances per %d samples "
                  "(L2 is 126 MB); three distinct sample sets cycle" %
                  ((2 * w["n"] ** 2 * 8 + ncomp(w["cov"]) * G * w["n"] * 8) * min(S_total, 8) / 1e9, min(S_total, 8))}


def cpu_baseline_leg(w, workers=None):
    """Bounded CPU sample of the workload: `workers` forked single-threaded workers (the reference's mclapply shape,
    R/spatial_gradient.R:29-32), one posterior sample each, g_cpu grid points, extrapolated to the whole grid."""
    from oracle import baseline
    cores = os.cpu_count() or 1
    workers = max(1, cores - 1) if workers is None else workers      # detectCores() - 1
    coords, grid, phi, sig2, z, eps = make_inputs(w, workers)         # one sample per worker, whatever the core count
    G = grid.shape[0]
    g_cpu = min(w["g_cpu"], G)
    r = baseline.gradient_reference_step(coords, grid, w["cov"], phi, sig2, z, eps, g_cpu, workers)
    r["sample"] = ("%d forked workers x 1 BLAS thread (the reference's mclapply shape), 1 sample each, %d of %d grid "
                   "points, extrapolated per worker as G/(t_factor + G*t_grid) with t_factor=%.2fs, t_grid=%.4fs, summed "
                   "over workers; restated reference algorithm (NumPy+OpenBLAS), not R"
                   % (workers, g_cpu, G, r["t_factor"], r["t_grid"]))
    r["workers"] = workers
    r["host"] = baseline.host_info()
    return r


# ---- reference arm: the restated reference algorithm on the host cores --------------------------------------
def run_reference(args, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    S_total = args.samples or w["samples"]
    vals, walls = [], []
    for it in range(args.warmup + args.steps):
        r = cpu_baseline_leg(w)
        if it >= args.warmup:
            vals.append(r["units_per_s"])
            walls.append(r["wall"])
    value = float(np.median(vals))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(np.mean(walls)) * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_of(w, world, S_total),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": r["workers"], "kind": "port", "sample": r["sample"],
                             "host": r["host"]},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---- our arm -------------------------------------------------------------------------------------------------
def run_ours(args, w):
    import torch
    import torch.distributed as dist
    import spwombling_b200 as spw
    from spwombling_b200 import device as dev
    import ctypes

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)
        # host-side rendezvous for the e2e leg: ranks > 0 must not spin in an NCCL barrier kernel on the GPUs that
        # rank 0's in-process spw_gradient_run(ngpu = N) is using
        cpu_group = dist.new_group(backend="gloo")
    lib = spw.load()
    lib.spw_set_device(local)
    n, cov, c = w["n"], w["cov"], ncomp(w["cov"])
    S_total = args.samples or w["samples"]
    nbatch_sum = min(w["nbatch"], max(1, S_total // 2))
    lo, hi = dev.shard_bounds(S_total, world, rank)
    npool = 3                                            # distinct sample sets cycled through the steps
    pools = [make_inputs(w, S_total, seed0=1000 * i) for i in range(npool)]
    coords, grid = pools[0][0], pools[0][1]
    G = grid.shape[0]
    coords_d = dev.coords_to_device(coords, device)
    grid_d = dev.coords_to_device(grid, device)
    up = lambda a: torch.from_numpy(np.ascontiguousarray(a[lo:hi])).to(device)   # noqa: E731
    devb = [dict(phi=up(p[2]), sig2=up(p[3]), z=up(p[4]), eps=up(p[5])) for p in pools]
    draws_buf = torch.empty((hi - lo, c, G), dtype=torch.float64, device=device)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def barrier_host():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier(group=cpu_group)

    def step_resident(i):
        # one process per GPU: this rank's contiguous shard, ONE gather to rank 0, summary on rank 0
        d = devb[i % npool]
        mine = dev.gradient_shard(coords_d, grid_d, cov, d["phi"], d["sig2"], d["z"], d["eps"], out=draws_buf)
        full = dev.gather_shards(mine, S_total) if world > 1 else mine
        if rank == 0:
            return full, dev.hpd_summary_dev(full, nbatch_sum)
        return None, None

    def step_e2e(i):
        # the C ABI R binds (spw_gradient_run): host buffers in, estimates + draws out, ngpu = N driven from rank 0
        if rank != 0:
            return None
        p = pools[i % npool]
        return spw.spatial_gradient(coords, {"phi": p[2], "sig2": p[3], "z": p[4]}, cov, grid, nbatch=nbatch_sum,
                                    return_mcmc=True, eps=p[5], ngpu=world, rounded=False)

    def timed(fn, steps, warmup, barrier=barrier):
        for i in range(warmup):
            fn(i)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(warmup + i)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    units_per_step = S_total * G
    # ---- value: HBM-resident inputs, phase events and clocks sampled during the timed region ----
    for i in range(args.warmup):
        step_resident(i)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    lib.spw_profile_enable(1)
    lib.spw_profile_read(None, None, None, 1)
    spw.launch_count(reset=True)
    ms_total = timed(step_resident, args.steps, 0)
    launches = spw.launch_count()
    ph_ms = (ctypes.c_double * 6)()
    ph_calls = (ctypes.c_longlong * 6)()
    main_flops = ctypes.c_double(0.0)
    lib.spw_profile_read(ctypes.cast(ph_ms, ctypes.c_void_p), ctypes.cast(ph_calls, ctypes.c_void_p),
                         ctypes.cast(ctypes.byref(main_flops), ctypes.c_void_p), 1)
    lib.spw_profile_enable(0)
    clocks = sampler.stop() if rank == 0 else None
    value = units_per_step * args.steps / (ms_total * 1e-3)
    # ---- multi-GPU parity, visible to the driver: the last sample of the gathered block (it lives on the last rank)
    #      recomputed alone on rank 0 must be bit-identical (sharding and batching change no arithmetic) ----
    last_i = args.warmup + args.steps - 1
    full, summ_dev = step_resident(last_i)
    parity = None
    if rank == 0:
        p = pools[last_i % npool]
        one = lambda a: torch.from_numpy(np.ascontiguousarray(a[S_total - 1:S_total])).to(device)   # noqa: E731
        alone = dev.gradient_shard(coords_d, grid_d, cov, one(p[2]), one(p[3]), one(p[4]), one(p[5]))
        parity = bool(torch.equal(alone[0], full[S_total - 1]))
        summ_value_leg = summ_dev.cpu().numpy()
        draws_value_leg = full[S_total - 1].cpu().numpy()
    barrier()
    # ---- e2e: host buffers -> spw_gradient_run(ngpu = N) -> estimates + draws on the host ----
    ms_e2e = timed(step_e2e, args.steps, max(1, args.warmup // 2), barrier_host)
    e2e_value = units_per_step * args.steps / (ms_e2e * 1e-3)
    h2d = (S_total * (2 + n + G * c) + coords.size + grid.size) * 8
    d2h = (c * 3 * G + S_total * c * G) * 8
    out_e2e = step_e2e(last_i)
    # ---- optional: the WHOLE job of the workload (all S posterior samples, nbatch = 200) through spw_gradient_run, wall
    #      clock including input marshalling, every H2D / D2H copy, the gather and the summary ----
    full_job = None
    if args.full_job:
        if rank == 0:
            Sf = args.full_job
            fin = make_inputs(w, Sf, seed0=7000)
            t0 = time.perf_counter()
            fo = spw.spatial_gradient(fin[0], {"phi": fin[2], "sig2": fin[3], "z": fin[4]}, cov, fin[1], nbatch=200,
                                      return_mcmc=True, eps=fin[5], ngpu=world, rounded=False)
            dt = time.perf_counter() - t0
            from oracle.summary import summarise_draws
            sub = np.linspace(0, G - 1, 7).astype(int)
            ok = all(np.array_equal(summarise_draws(np.asarray(fo["grad.%s.mcmc" % nm])[:, sub], 200), fo[en][sub])
                     for nm, en in zip(("s1", "s2", "s11", "s12", "s22")[:c],
                                       ("grad1.est", "grad2.est", "grad11.est", "grad12.est", "grad22.est")[:c]))
            full_job = {"samples": Sf, "grid_points": G, "nbatch": 200, "ngpu": world, "wall_s": dt,
                        "units_per_s_wall": Sf * G / dt, "h2d_gb": (fin[5].nbytes + fin[4].nbytes) / 1e9,
                        "d2h_gb": Sf * G * c * 8 / 1e9, "summary_matches_oracle_on_returned_draws": bool(ok)}
            del fin, fo
        barrier_host()
    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    # the C-ABI path (in-process sharding, its own gather) against the torch.distributed path: same bytes
    names = ("s1", "s2", "s11", "s12", "s22")[:c]
    est_names = ("grad1.est", "grad2.est", "grad11.est", "grad12.est", "grad22.est")[:c]
    e2e_draws = np.stack([np.asarray(out_e2e["grad.%s.mcmc" % nm])[S_total - 1] for nm in names])
    e2e_summ = np.stack([np.asarray(out_e2e[k]).T for k in est_names])
    parity_e2e = bool(np.array_equal(e2e_draws, draws_value_leg) and np.array_equal(e2e_summ, summ_value_leg))
    if not (parity and parity_e2e):
        raise SystemExit("bench.py: multi-GPU parity check failed (gathered sample vs single-rank recompute: %s; "
                         "spw_gradient_run vs torch.distributed path: %s)" % (parity, parity_e2e))
    # ---- roofline of the dominant kernel (fused TRMM + Gram + draw) ----
    peaks = {}
    for f in ("profiles/FP64_PEAK.json", "MEASURED_PEAKS.json"):
        pth = os.path.join(ROOT, f)
        if os.path.exists(pth):
            peaks.update(json.load(open(pth)))
    peak_tf = float(peaks.get("fp64_tensor_tflops", 37.0))
    main_ms, main_calls = ph_ms[5], max(1, ph_calls[5])
    achieved_tf = main_flops.value / (main_ms * 1e-3) / 1e12 if main_ms > 0 else None
    traffic = None
    tp = os.path.join(ROOT, "profiles", "main_kernel_traffic.json")
    if os.path.exists(tp):
        rec = json.load(open(tp)).get(args.workload)
        if rec:     # ncu capture of one launch, scaled to the average number of samples per launch of this run
            traffic = rec["bytes_per_sample"] * (hi - lo) * args.steps / main_calls
    names_ph = ["cov_assemble", "potrf", "trtri", "trmv", "cross_cov", "gradient_main"]
    phases = {nm: {"ms_per_step": ph_ms[i] / args.steps, "launch_groups": int(ph_calls[i])} for i, nm in enumerate(names_ph)}
    roofline = {"bound": "tensor", "kernel": "gradient_tma_kernel (TMA-fed fused TRMM + Gram + 5x5 chol + draw)",
                "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                "frac": (achieved_tf / peak_tf) if achieved_tf else None, "traffic": traffic,
                "peak_source": "measured FP64 tensor pipe (DMMA.8x8x4 micro-benchmark, profiles/FP64_PEAK.json; "
                               "MEASURED_PEAKS.json has no FP64 entry)",
                "ms_per_launch": main_ms / main_calls, "flops_per_launch": main_flops.value / main_calls,
                "share_of_step": main_ms / ms_total,
                "note": "rank 0's launches; traffic = DRAM bytes per launch from one ncu --set full capture "
                        "(profiles/main_kernel_traffic.json), scaled to this run's samples per launch"}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_of(w, world, S_total),
            "algorithmic_tflops": value * flops_per_unit(w) / 1e12,
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / args.steps,
                    "path": "spw_gradient_run (C ABI, host buffers, ngpu=%d in-process from rank 0%s)" %
                            (world, ", gather via NCCL" if lib.spw_last_gather_used_nccl() == 1 else
                             (", gather via peer copies" if world > 1 else ""))},
            "parity_checked": True,
            "parity": {"gathered_sample_vs_single_rank_recompute": "bit-identical",
                       "c_abi_path_vs_torch_distributed_path": "bit-identical (draws of the last sample and all estimates)"},
            "gpu_launches": int(launches), "roofline": roofline, "phases": phases}
    if full_job:
        line["full_job"] = full_job
    # ---- CPU baseline (rank 0, N = 1 only): bounded sample of the same workload ----
    if world == 1 and not args.no_cpu:
        r = cpu_baseline_leg(w)
        line["cpu_baseline"] = {"value": r["units_per_s"], "unit": UNIT, "cores": r["workers"], "kind": "port",
                                "sample": r["sample"], "host": r["host"]}
    # ---- secondary metric: hlmBayes_sp MCMC iterations / s at 1 GPU ----
    if world == 1 and not args.no_hlm:
        try:                                  # a secondary metric must never take the headline line down with it
            line["hlm"] = bench_hlm(spw, args)
        except Exception as exc:              # noqa: BLE001 -- reported, not swallowed
            line["hlm"] = {"error": "%s: %s" % (type(exc).__name__, exc)}
        line["spsample"] = bench_spsample(spw, args)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    print(json.dumps(line), flush=True)


def bench_hlm(spw, args):
    """hlmBayes_sp iterations/s (BASELINE.json's second metric), N = 5000 matern2 and the README's N = 100, through the
    C ABI with host buffers (wall clock of spw_hlm_run: transfers, plan and graph set-up included)."""
    peak_tf = 37.0
    pk = os.path.join(ROOT, "profiles", "FP64_PEAK.json")
    if os.path.exists(pk):
        peak_tf = float(json.load(open(pk)).get("fp64_tensor_tflops", peak_tf))
    out = {}
    for n, niter in ((5000, 96), (100, 2000)):     # long enough that the per-call graph set-up (~20 ms) is < 3 %
        rng = np.random.Generator(np.random.Philox(2))
        coords = rng.random((n, 2))
        y = 10 * (np.sin(3 * np.pi * coords[:, 0]) + np.cos(3 * np.pi * coords[:, 1])) + rng.standard_normal(n)
        r6 = np.random.Generator(np.random.Philox(6))
        norm, unif = r6.standard_normal(3 * niter), r6.random(3 * niter)
        kw = dict(y=y, coords=coords, nburn=0, report=max(1, niter // 2), cov_type="matern2", verbose=False)
        spw.hlmBayes_sp(niter=max(2, niter // 8), norm=norm, unif=unif, **kw)    # warm-up (plans, clocks)
        dt = 1e30
        for _ in range(3):                                                       # best of 3 (wall clock, host jitter)
            t0 = time.perf_counter()
            spw.hlmBayes_sp(niter=niter, norm=norm, unif=unif, **kw)
            dt = min(dt, time.perf_counter() - t0)
        its = niter / dt
        tf = its * 3 * (n + 1) ** 3 / 3.0 / 1e12
        out["n%d" % n] = {"iters_per_s": its, "niter": niter, "ms_per_factorisation": dt / niter / 3 * 1e3,
                          "cholesky_tflops": tf, "cholesky_frac": tf / peak_tf,
                          "schedule": ("single persistent CTA" if n <= 200 else
                                       "shadow schedule (SPW_LA=0)" if os.environ.get("SPW_LA", "")[:1] == "0" or n < 1024
                                       else "look-ahead schedule with a resident diagonal chain (potrf_la.cu)"),
                          "note": "wall clock of spw_hlm_run through the C ABI (H2D/D2H, plan and CUDA-graph set-up "
                                  "included); 3 Cholesky factorisations of order n+1 per iteration ((n+1)^3/3 flop "
                                  "each); cholesky_frac = cholesky_tflops / %.1f "
                                  "(measured FP64 tensor peak)" % peak_tf}
    if not args.no_cpu:
        from oracle import baseline
        n = 5000
        rng = np.random.Generator(np.random.Philox(2))
        coords = rng.random((n, 2))
        y = rng.standard_normal(n)
        dt, _ = baseline.hlm_reference_iteration(coords, y, "matern2")
        out["n5000"]["cpu_baseline_iters_per_s"] = 1.0 / dt
        out["n5000"]["cpu_baseline_note"] = ("restated reference iteration (3 x dpotrf + dpotri + quadratic form, "
                                             "NumPy+OpenBLAS all threads), not R")
    return out


def bench_spsample(spw, args):
    """spsample latent-process draws per second at N = 2000, through the C ABI with host buffers."""
    n, S = 2000, 48
    rng = np.random.Generator(np.random.Philox(8))
    coords = rng.random((n, 2))
    y = 10 * (np.sin(3 * np.pi * coords[:, 0]) + np.cos(3 * np.pi * coords[:, 1])) + rng.standard_normal(n)
    X = np.ones((n, 1))
    phi, sig2, tau2 = rng.uniform(20, 40, S), rng.uniform(0.5, 2, S), rng.uniform(0.5, 2, S)
    eb, ez = rng.standard_normal((S, 1)), rng.standard_normal((S, n))
    kw = dict(y=y, X=X, coords=coords, cov_type="matern2")
    spw.spsample(phi=phi[:8], sig2=sig2[:8], tau2=tau2[:8], eps_beta=eb[:8], eps_z=ez[:8], **kw)
    dt = 1e30
    for _ in range(3):
        t0 = time.perf_counter()
        spw.spsample(phi=phi, sig2=sig2, tau2=tau2, eps_beta=eb, eps_z=ez, **kw)
        dt = min(dt, time.perf_counter() - t0)
    out = {"n": n, "samples": S, "samples_per_s": S / dt,
           "algorithmic_tflops": S * (8.0 / 3.0) * n ** 3 / dt / 1e12,
           "note": "wall clock of spw_spsample_run (H2D/D2H included); per sample 3 Cholesky factorisations, 2 triangular "
                   "inverses and 2 Gram products (8/3 n^3 flop) replace the reference's 3 dpotrf + 2 dpotri"}
    if not args.no_cpu:
        from oracle.spsample import spsample as oracle_spsample
        t0 = time.perf_counter()
        oracle_spsample(y, X, coords, phi[:2], sig2[:2], tau2[:2], "matern2", eb[:2], ez[:2])
        out["cpu_baseline_samples_per_s"] = 2 / (time.perf_counter() - t0)
        out["cpu_baseline_note"] = "restated reference loop (NumPy+OpenBLAS, all BLAS threads, 2 samples), not R"
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c5", choices=sorted(WORKLOADS))
    ap.add_argument("--samples", type=int, default=0, help="posterior samples per step (whole job, split over the ranks)")
    ap.add_argument("--full-job", type=int, default=0, metavar="S",
                    help="also run the workload's whole job with S posterior samples through spw_gradient_run (wall clock)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    ap.add_argument("--no-hlm", action="store_true", help="skip the hlmBayes_sp secondary metric")
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, w)
    else:
        args.warmup = max(args.warmup, 3)
        run_ours(args, w)


if __name__ == "__main__":
    main()
