#!/usr/bin/env python
"""bench.py — EDMD fit throughput (lift + Gram + solve) on BASELINE config 5.

One step = one full bilinear fit of this rank's snapshot shard:
    kf_accumulate_dev (fused lift -> L2-resident panel -> FP64 DMMA Gram/cross-covariance)
    [N>1: one NCCL all-reduce of the packed partial Grams]
    kf_solve_dev      (assemble G, C; pivoted-Cholesky basic LS solve; K to the host)

Workload: synthetic snapshot pairs, n = nzeta = 12, m = 3, dictionary
{'poly','gaussian'},[3,569] -> N = 1024, bilinear -> P = 4096.  Weak scaling: every rank owns
`--snapshots-per-gpu` pairs (default 1.25e7 = 10^8 / 8, so N = 8 is exactly the named 10^8 job).

    python bench.py --gpus 1 --steps 3 --warmup 3
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference      # CPU arm: the NumPy/SciPy oracle on host cores
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

if "reference" in sys.argv:
    # The CPU arm uses every host core whatever the launcher exported: torchrun sets OMP_NUM_THREADS=1 for its workers, which
    # would time the reference on ONE core (and made the N > 1 reference runs of round 1 time out).  Must precede `import numpy`.
    for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_v] = str(os.cpu_count() or 1)

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "edmd_fit_snapshots_per_sec"
UNIT = "snapshots/s"
NZETA, M_IN, N_LIFT, P_REG = 12, 3, 1024, 4096
OBS_TYPE, OBS_DEGREE = ["poly", "gaussian"], [3, 569]
FLOPS_ALGO_PER_PAIR = P_REG * (P_REG + 1) + 2 * P_REG * P_REG      # P(P+1) + 2 P Pc


_REAL_STDOUT = None


def emit(line):
    """The ONE JSON line of the contract, on the real stdout."""
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def workload_constants():
    """A0 (spectral radius 0.95), B0_i (||.||_2 = 0.1) from seed 1; gaussian centres from seed 2."""
    rng = np.random.default_rng(1)
    A0 = rng.standard_normal((NZETA, NZETA))
    A0 *= 0.95 / np.max(np.abs(np.linalg.eigvals(A0)))
    B0 = rng.standard_normal((M_IN, NZETA, NZETA))
    B0 = np.stack([0.1 * b / np.linalg.norm(b, 2) for b in B0])
    centres = 2 * np.random.default_rng(2).random((NZETA, OBS_DEGREE[1])) - 1
    return A0, B0, centres


def gen_numpy(M, seed):
    """CPU generator of the same distribution (for the CPU arm)."""
    A0, B0, _ = workload_constants()
    rng = np.random.default_rng(seed)
    zeta = 2 * rng.random((M, NZETA)) - 1
    u = 2 * rng.random((M, M_IN)) - 1
    beta = zeta @ A0.T + sum(u[:, i:i + 1] * (zeta @ B0[i].T) for i in range(M_IN)) + 0.01 * rng.standard_normal((M, NZETA))
    return zeta, np.clip(beta, -1, 1), u


def gen_torch(M, device, seed):
    """Device generator (torch's Philox): feature-major tensors (nzeta, M) = column-major M x nzeta."""
    import torch
    A0, B0, _ = workload_constants()
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    tA = torch.tensor(A0, device=device)
    tB = torch.tensor(B0, device=device)
    alpha = torch.empty((NZETA, M), dtype=torch.float64, device=device)
    beta = torch.empty((NZETA, M), dtype=torch.float64, device=device)
    u = torch.empty((M_IN, M), dtype=torch.float64, device=device)
    step = 1 << 21
    for lo in range(0, M, step):
        hi = min(M, lo + step)
        z = torch.rand((NZETA, hi - lo), dtype=torch.float64, device=device, generator=g) * 2 - 1
        uu = torch.rand((M_IN, hi - lo), dtype=torch.float64, device=device, generator=g) * 2 - 1
        b = tA @ z
        for i in range(M_IN):
            b += uu[i:i + 1] * (tB[i] @ z)
        b += 0.01 * torch.randn((NZETA, hi - lo), dtype=torch.float64, device=device, generator=g)
        alpha[:, lo:hi], beta[:, lo:hi], u[:, lo:hi] = z, b.clamp_(-1, 1), uu
    return alpha, beta, u


def fp64_peak_committed():
    """The builder's earlier measurement (profiles/r01_fp64_dgemm_peak.json): kept as a cross-check only."""
    try:
        return float(json.load(open(os.path.join(ROOT, "profiles", "r01_fp64_dgemm_peak.json")))["dgemm_tflops_sustained"])
    except Exception:
        return None


def fp64_peak_measured(dev, n=8192, reps=12):
    """Roofline denominator measured IN THIS RUN on the bench box: cuBLAS DGEMM n^3 in FP64 (torch.matmul), CUDA events, after
    warm-up; `burst` = best single GEMM, `sustained` = all reps back to back.  MEASURED_PEAKS.json carries no FP64 figure."""
    import torch
    a = torch.randn((n, n), dtype=torch.float64, device=dev)
    b = torch.randn((n, n), dtype=torch.float64, device=dev)
    c = torch.empty((n, n), dtype=torch.float64, device=dev)
    for _ in range(3):
        torch.matmul(a, b, out=c)
    torch.cuda.synchronize(dev)
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    evs[0].record()
    for i in range(reps):
        torch.matmul(a, b, out=c)
        evs[i + 1].record()
    torch.cuda.synchronize(dev)
    per = [evs[i].elapsed_time(evs[i + 1]) for i in range(reps)]
    flop = 2.0 * n ** 3
    del a, b, c
    return {"burst": flop / (min(per) * 1e-3) / 1e12, "sustained": flop * reps / (evs[0].elapsed_time(evs[reps]) * 1e-3) / 1e12}


def gram_traffic_from_profile(kernel="gram_tma"):
    """dram bytes per launch of the dominant kernel from the committed ncu capture (NOT measured in this run)."""
    for name in (f"r02_{kernel}_traffic.json", f"r01_{kernel}_traffic.json"):
        try:
            d = json.load(open(os.path.join(ROOT, "profiles", name)))
            return float(d["dram_bytes_per_launch"]), f"profiles/{name} ({d.get('source', 'ncu --set full')})"
        except Exception:
            continue
    return None, None


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "200", "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [ln.strip().split(", ") for ln in open(self.f.name) if ln.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for nm, v in zip(names, r[5:9]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        busy = [s for s in sm if s > 0.5 * max(mx or [1])] or sm
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


_CPU_SOLVE_S = None


def cpu_fit_rates(sample, threads):
    """Time the oracle (kind 'port': NumPy/SciPy restatement, OpenBLAS) on a bounded sample of the workload.
    Returns lift+Gram rate (snapshots/s), solve seconds, and what was timed.  The per-snapshot part (lift + Gram) is timed
    on every call; the P = 4096 dgeqp3 solve — paid once per fit, independent of the snapshot count — is timed ONCE per
    process and reused (it is ~15 s on 16 cores; timing it in each of 25 steps is what overran the driver's limit in round 1)."""
    global _CPU_SOLVE_S
    import oracle as O
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=threads)
    except Exception:
        pass
    _, _, centres = workload_constants()
    prog = O.build_program(OBS_TYPE, OBS_DEGREE, NZETA, centres)
    alpha, beta, u = gen_numpy(sample, seed=1234)
    t0 = time.perf_counter()
    Px, Py = O.build_regressors("bilinear", prog, alpha, beta, u)
    G, C = O.gram(Px, Py)
    t_lg = time.perf_counter() - t0
    assert np.all(np.isfinite(G)) and np.all(np.isfinite(C))
    if _CPU_SOLVE_S is None:
        t0 = time.perf_counter()
        K = O.mldivide(Px, Py)          # dgeqp3 — the routine behind MATLAB's `\` (Ksysid.m:1069)
        _CPU_SOLVE_S = time.perf_counter() - t0
        assert K.shape == (P_REG, P_REG)
    return sample / t_lg, _CPU_SOLVE_S, t_lg


def run_reference(args, rank):
    """--impl reference: the reference's CPU implementation of the path.  MATLAB/Octave are absent
    (BASELINE.md §4), so this is the oracle port on the host cores, vectorised (which flatters the CPU
    against the reference's interpreted per-snapshot loop, Ksysid.m:1030-1065)."""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    sample = args.cpu_sample
    vals, times = [], []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        rate_lg, t_solve, t_lg = cpu_fit_rates(sample, threads)
        dt = time.perf_counter() - t0
        if i >= args.warmup:
            # whole-job CPU estimate for the same per-rank shard: per-snapshot work scales with M, the solve is paid once
            Mjob = args.snapshots_per_gpu * args.gpus
            vals.append(Mjob / (Mjob / rate_lg + t_solve))
            times.append(dt)
    v = float(np.median(vals))
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(times)), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(args),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "cpu_sample_snapshots": sample,
                             "sample": f"{sample} snapshots of the same workload: NumPy lift + OpenBLAS Gram timed per snapshot in every step, "
                                       f"dgeqp3 solve (P=4096) timed once per run; extrapolated linearly in M to the "
                                       f"{args.snapshots_per_gpu * args.gpus}-snapshot job"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def config_dict(args, extra=None):
    d = {"workload": "BASELINE config 5: synthetic snapshots, n=12, m=3, dictionary {poly,gaussian},[3,569] -> N=1024, "
                     "bilinear -> P=4096 lifted regressor, lasso=Inf (least squares), snapshot-sharded",
         "snapshots_per_gpu": args.snapshots_per_gpu, "total_snapshots": args.snapshots_per_gpu * args.gpus,
         "P": P_REG, "N": N_LIFT, "nzeta": NZETA, "m": M_IN, "model_type": "bilinear",
         "parallelism": f"snapshot-sharded x{args.gpus}, one NCCL all-reduce of the packed partial Grams",
         "l2_policy": "inputs (>= 216 B/snapshot x shard) exceed L2 for shards >= 1M snapshots; every step re-reads them from HBM"}
    if extra:
        d.update(extra)
    return d


def lasso3a_sweep(fit, rank, world, dev, reps=2):
    """BASELINE config 3a as an extra (`"lasso3a"` in the JSON line): snake data, bilinear, fourier degree 4 (P = 1464), the
    lasso vector logspace(-2, 2, 64) swept by the exact active-set solver.  One kf_fit call per rank on its shard of the
    snapshot pairs (host buffers): all-reduce of the partial Grams, then the sweep is split by COLUMNS of K over the ranks inside
    the library (Ksysid.m:1370-1387 re-fits per lasso value).  Time = whole call, max over ranks; certified by the Frank-Wolfe gaps."""
    import io
    import contextlib
    import torch
    import torch.distributed as dist
    from koopfit.ksysid import Ksysid
    from koopfit.sharding import shard_bounds
    z = np.load(os.path.join(ROOT, "tests", "golden", "snake_data.npz"))
    data = {}
    for split in ("train", "val"):
        data[split] = [{k: z[f"{split}{i}_{k}"] for k in ("t", "y", "u")} for i in range(int(z[f"{split}_n"]))]
    lassos = np.logspace(-2, 2, 64)
    with contextlib.redirect_stdout(io.StringIO()):
        ks = Ksysid(data, model_type="bilinear", obs_type=["fourier"], obs_degree=[4], lasso=lassos, dim_red=False, fitter=fit)
    N, sp = ks.params["N"], ks.snapshotPairs
    lo, hi = shard_bounds(sp["alpha"].shape[0], rank, world)
    al, be, uu = (np.asfortranarray(sp[k][lo:hi]) for k in ("alpha", "beta", "u"))
    secs, res = [], None
    for _ in range(reps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        res = fit.fit(ks.basis, "bilinear", al, be, uu, least_squares=False, t=lassos * N, psd_shift="never")
        tt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        secs.append(float(tt.item()))
    rel = res["qp_gap"] / np.abs(res["objective"])
    return {"workload": "BASELINE config 3a: snake-data, bilinear, fourier 4 (P = 1464), 64 budgets logspace(-2,2,64)*N, exact active set",
            "n_gpus": world, "seconds": min(secs), "seconds_all": secs, "solve_ms": res["info"]["t_solve_ms"],
            "lift_gram_ms": res["info"]["t_lift_gram_ms"], "budgets": int(lassos.size), "total_steps": int(res["qp_iters"].sum()),
            "unconverged": int(res["info"]["qp_capped"]), "worst_rel_gap": float(rel.max()),
            "split": "columns of K over the ranks, step scalars reduced on the device (NCCL, in stream order)" if world > 1 else "single GPU",
            "api": "kf_fit (host shard per rank; library-owned NCCL communicator)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--snapshots-per-gpu", type=int, default=12_500_000)
    ap.add_argument("--cpu-sample", type=int, default=8192)
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-fast-mode", action="store_true", help="skip the extra pc_cols = N measurement")
    ap.add_argument("--no-lasso3a", action="store_true", help="skip the extra config-3a lasso sweep")
    ap.add_argument("--workload", default="config5", choices=["config5", "lasso3a"],
                    help="lasso3a: only the config-3a lasso sweep (prints its own JSON line, metric lasso_sweep_seconds)")
    ap.add_argument("--option", action="append", default=[], help="kf_set_option name=value")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    # stdout carries exactly ONE JSON line: anything libraries print (e.g. "NCCL version ...") goes to stderr
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    import koopfit

    if world != args.gpus and world > 1:
        args.gpus = world
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    fit = koopfit.Fitter(device=local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        # the data-path collective lives INSIDE libkoopfit.so: rank 0 draws the NCCL id, torch.distributed only carries it over
        ids = [koopfit.Fitter.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        fit.comm_init(world, rank, ids[0])
    for kv in args.option:
        k, v = kv.split("=")
        fit.set_option(k, float(v))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        fit.sync()

    if args.workload == "lasso3a":
        out = lasso3a_sweep(fit, rank, world, dev)
        if rank == 0:
            emit({"metric": "lasso_sweep_seconds", "value": out["seconds"], "unit": "s", "n_gpus": world, "steps": 1, "warmup": 1,
                  "ms_per_step": 1e3 * out["seconds"], "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
                  "data": "reference fixture (snake-data.mat)", "config": {"workload": out["workload"]}, "lasso3a": out})
        fit.close()
        if world > 1:
            dist.destroy_process_group()
        return

    fit.set_option("profile", 1)
    M = args.snapshots_per_gpu
    _, _, centres = workload_constants()
    basis = koopfit.Basis(OBS_TYPE, OBS_DEGREE, NZETA, centres)
    alpha, beta, u = gen_torch(M, dev, seed=1000 + rank)
    torch.cuda.synchronize()
    kstream = torch.cuda.ExternalStream(fit.stream, device=dev)

    def step_resident(pc_cols=0):
        # kf_fit_dev: lift + Gram of this rank's device-resident shard, [N > 1: ONE ncclAllReduce of the packed partial Grams,
        # inside the library], assemble, pivoted-Cholesky solve (refinement passes if the conditioning asks for them), K to the host
        return fit.fit_dev(basis, "bilinear", M, NZETA, M_IN, alpha.data_ptr(), beta.data_ptr(), u.data_ptr(), ls_method="gram",
                           pc_cols=pc_cols)

    # ---------------- device-resident timing (value) ----------------
    for _ in range(args.warmup):
        res = step_resident()
    barrier()
    fit.counters(reset=True)
    fit.engine_info(reset=True)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    gram_ms, liftgram_ms, solve_ms = [], [], []
    e0.record(kstream)
    for _ in range(args.steps):
        res = step_resident()
        t = fit.last_times()
        gram_ms.append(t["gram_kernel_ms"]); liftgram_ms.append(t["lift_gram_ms"]); solve_ms.append(t["solve_ms"])
    e1.record(kstream)
    barrier()
    clocks = sampler.stop() if sampler else None
    ms_total = e0.elapsed_time(e1)
    flops_issued, launches = fit.counters()
    eng = fit.engine_info()
    tmax = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_step = float(tmax.item()) / args.steps
    value = M * world / (ms_step * 1e-3)
    assert res["rank"] == P_REG, f"unexpected rank {res['rank']}"
    assert res["info"]["passes"] == 1, "the benchmarked workload is well conditioned: one data pass"
    assert np.all(np.isfinite(res["K"]))

    # ---------------- end-to-end through the public host API (kf_fit, host buffers) at every N ----------------
    h_alpha = torch.empty((NZETA, M), dtype=torch.float64).pin_memory()
    h_beta = torch.empty((NZETA, M), dtype=torch.float64).pin_memory()
    h_u = torch.empty((M_IN, M), dtype=torch.float64).pin_memory()
    h_alpha.copy_(alpha); h_beta.copy_(beta); h_u.copy_(u)
    torch.cuda.synchronize()
    na, nb, nu = h_alpha.numpy().T, h_beta.numpy().T, h_u.numpy().T     # (M x nzeta) column-major views
    h2d = (2 * NZETA + M_IN) * 8 * M
    d2h = P_REG * P_REG * 8

    def step_e2e():
        # this rank's HOST shard: block-wise H2D overlapped with the lift + Gram, in-library all-reduce, solve, K to the host
        return fit.fit(basis, "bilinear", na, nb, nu, ls_method="gram")
    step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        r2 = step_e2e()
    barrier()
    tt = torch.tensor([(time.perf_counter() - t0) * 1e3 / args.e2e_steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    e2e_val = M * world / (float(tt.item()) * 1e-3)
    assert r2["rank"] == P_REG

    # ---------------- opt-in fast mode (not the headline): only K(:,1:N), the columns A and B are cut from ----------------
    fast = None
    if not args.no_fast_mode:
        step_resident(pc_cols=N_LIFT)
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record(kstream)
        rf = step_resident(pc_cols=N_LIFT)
        f1.record(kstream)
        barrier()
        tf = torch.tensor([f0.elapsed_time(f1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tf, op=dist.ReduceOp.MAX)
        assert rf["rank"] == P_REG and rf["K"].shape == (P_REG, N_LIFT)
        fast = {"pc_cols": N_LIFT, "value": M * world / (float(tf.item()) * 1e-3), "unit": UNIT, "ms_per_step": float(tf.item()), "steps": 1,
                "note": "kf_problem.pc_cols = N: 616 accumulator tiles instead of 1000 and N right-hand sides; K(:,1:N) identical to the full solve"}

    # ---------------- K1 alone (lift-only mode): HBM-bound, 8 (2 nzeta + m) + 16 P bytes per pair ----------------
    lift_only = None
    if world == 1 and rank == 0:
        try:
            Ml = min(M, 65536)
            out_buf = torch.empty((2 * P_REG, Ml), dtype=torch.float64, device=dev)
            for _ in range(2):
                fit.regressors_dev(basis, "bilinear", Ml, NZETA, M_IN, alpha.data_ptr(), beta.data_ptr(), u.data_ptr(), out_buf.data_ptr(), ld=Ml)
            barrier()
            l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            l0.record(kstream)
            for _ in range(5):
                fit.regressors_dev(basis, "bilinear", Ml, NZETA, M_IN, alpha.data_ptr(), beta.data_ptr(), u.data_ptr(), out_buf.data_ptr(), ld=Ml)
            l1.record(kstream)
            barrier()
            lms = l0.elapsed_time(l1) / 5
            lbytes = Ml * (8.0 * (2 * NZETA + M_IN) + 16.0 * P_REG)
            hbm = 6554.6
            try:
                hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
            except Exception:
                pass
            lift_only = {"bound": "hbm", "kernel": "kf_lift_stream_kernel: materialising lift ([Px | Py] written to HBM; not on the fit path)", "pairs": Ml,
                         "achieved": lbytes / lms / 1e6, "peak": hbm, "unit": "GB/s", "frac": lbytes / lms / 1e6 / hbm,
                         "algorithmic_bytes_per_pair": 8.0 * (2 * NZETA + M_IN) + 16.0 * P_REG, "ms": lms}
            del out_buf
        except Exception as exc:      # an extra, never the reason for a failed bench
            lift_only = {"error": str(exc)[:200]}

    # ---------------- FP64 tensor peak measured in this run (rank 0), inputs freed first ----------------
    del alpha, beta, u
    torch.cuda.empty_cache()
    peak_now = fp64_peak_measured(dev) if rank == 0 else None

    # ---------------- extra: config 3a lasso sweep at this N ----------------
    lasso3a = None
    if not args.no_lasso3a:
        try:
            fit.set_option("profile", 0)
            lasso3a = lasso3a_sweep(fit, rank, world, dev)
        except Exception as exc:
            lasso3a = {"error": str(exc)[:300]}

    if rank == 0:
        committed = fp64_peak_committed()
        tiles_flops = flops_issued / args.steps                      # DMMA flops issued per step (incl. solver GEMMs)
        gk = float(np.mean(gram_ms)) if gram_ms and np.mean(gram_ms) > 0 else float(np.mean(liftgram_ms))
        lg = float(np.mean(liftgram_ms))
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        if eng["engine"] == 2:
            # INT8 tensor-core engine (Ozaki scheme II, FP64-exact): the dominant kernel is oz_gemm_kernel (tcgen05.mma kind::i8).
            # achieved = INT8 operations issued per contraction launch / mean duration of the sampled launches (CUDA events on
            # the launching stream, launch isolated from the other pipeline); peak = 2 x the MEASURED dense bf16 rate of this
            # pool's B200 (INT8 runs at twice the bf16 rate on these tensor cores): burst figure, the kernel is timed alone.
            launches_per_step = max(eng["gram_launches"], 1)
            k_ms = gk / launches_per_step
            ach = eng["i8_ops_per_launch"] / (k_ms * 1e-3) / 1e12
            bf16 = float(peaks.get("bf16_tflops_sustained", 1377.3))
            bf16_burst = float(peaks.get("bf16_tflops", 1654.5))
            peak_i8 = 2.0 * bf16
            traffic, traffic_src = gram_traffic_from_profile("oz_gemm")
            roof = {"bound": "tensor", "kernel": "oz::oz_gemm_kernel (tcgen05.mma kind::i8, INT32 accumulators in TMEM, TMA operands; "
                                                   "15 residue GEMMs per panel = one FP64-exact Gram / cross product, Ozaki scheme II)",
                    "achieved": ach, "peak": peak_i8, "unit": "TFLOP/s", "frac": ach / peak_i8,
                    "unit_note": "INT8 tensor operations (TOP/s); 2 ops per multiply-accumulate",
                    "peak_source": f"2 x MEASURED_PEAKS.json bf16_tflops_sustained ({bf16}: cuBLAS bf16 8192^3 back to back for 4 s) - the INT8 rate "
                                   "of the tcgen05 tensor cores is twice the bf16 rate, and the sampled launches sit INSIDE a long step that runs "
                                   "at the 1 kW power cap (see clocks: sw_power_cap, SM clock well below max), so the sustained figure applies; "
                                   "nominal INT8 dense peak 4500",
                    "frac_of_2x_bf16_burst": ach / (2.0 * bf16_burst), "frac_of_nominal_4500": ach / 4500.0,
                    "traffic": traffic, "traffic_unit": "bytes per launch (one 4096-snapshot panel, all 15 moduli)",
                    "traffic_source": f"NOT measured in this run: {traffic_src}" if traffic_src else None,
                    "kernel_ms_per_launch": k_ms, "launches_per_step": launches_per_step, "sampled_launches_per_step": eng["gram_sampled"],
                    "time_basis": "mean of the isolated, sampled contraction launches (CUDA events on the launching stream)",
                    "int8_ops_per_launch": eng["i8_ops_per_launch"],
                    "algorithmic_flops_per_pair": FLOPS_ALGO_PER_PAIR,
                    "fp64_equivalent_achieved": FLOPS_ALGO_PER_PAIR * M / (lg * 1e-3) / 1e12,
                    "fp64_equivalent_note": "algorithmic FP64 flops P(P+1) + 2 P Pc per pair / measured lift + Gram phase: what an FP64 "
                                            "tensor pipe would have to sustain; this box's measured cuBLAS DGEMM peak is below",
                    "fp64_dgemm_peak_this_run": peak_now["sustained"], "fp64_dgemm_peak_committed_crosscheck": committed,
                    "lift_gram_ms_per_step": lg, "solve_ms_per_step": float(np.mean(solve_ms)),
                    "engine": "int8 (auto: P >= 1024 and >= 4 panels; kf_set_option gram_engine=1 selects the FP64 DMMA kernel)"}
        else:
            peak = peak_now["sustained"]
            gram_flops = 1000 * 2.0 * 128 * 128 * M                       # 10 Kronecker blocks x (36 G + 64 C) tiles
            achieved = gram_flops / (lg * 1e-3) / 1e12
            traffic, traffic_src = gram_traffic_from_profile("gram_tma")
            roof = {"bound": "tensor", "kernel": "kf_gram_tma_kernel<true> (FP64 DMMA.8x8x4 Gram/cross-covariance, tensor-map TMA operands)",
                    "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                    "traffic": traffic, "traffic_unit": "bytes per launch (one 4096-snapshot panel)",
                    "traffic_source": f"NOT measured in this run: {traffic_src}" if traffic_src else None,
                    "algorithmic_flops_per_launch": FLOPS_ALGO_PER_PAIR * 4096, "issued_flops_per_launch": 1000 * 2.0 * 128 * 128 * 4096,
                    "peak_source": "cuBLAS DGEMM 8192^3 fp64 (torch.matmul) measured in THIS run on this box, 12 GEMMs back to back "
                                   "(sustained); MEASURED_PEAKS.json has no FP64 entry",
                    "peak_burst": peak_now["burst"], "peak_committed_crosscheck": committed,
                    "time_basis": "lift_gram_ms_per_step: CUDA events on the library's stream around the whole lift + Gram phase",
                    "flops_basis": "DMMA flops actually issued per step by the Gram kernel (Kronecker blocks: 1000 tiles x 2x128x128 per snapshot = 3.28e7/pair)",
                    "achieved_algorithmic": FLOPS_ALGO_PER_PAIR * M / (lg * 1e-3) / 1e12,
                    "algorithmic_flops_per_pair": FLOPS_ALGO_PER_PAIR, "issued_flops_per_pair": 1000 * 2.0 * 128 * 128,
                    "gram_kernel_ms_per_step_estimate": gk,
                    "gram_kernel_ms_note": "ESTIMATE: mean of the isolated every-61st-launch samples x launches per step (not a measured total)",
                    "achieved_kernel_only_estimate": gram_flops / (gk * 1e-3) / 1e12,
                    "lift_gram_ms_per_step": lg,
                    "solve_ms_per_step": float(np.mean(solve_ms)), "dmma_flops_issued_per_step_all_kernels": tiles_flops,
                    "engine": "fp64 dmma"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": config_dict(args),
            "clocks": clocks,
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "kf_fit: this rank's pinned HOST shard, block-wise H2D overlapped with lift + Gram, "
                           + ("ncclAllReduce inside the library, " if world > 1 else "") + "solve, K to the host"},
            "gpu_launches": int(launches),
            "fast_mode": fast,
            "lift_only": lift_only,
            "lasso3a": lasso3a,
            "roofline": roof,
        }
        if not args.no_cpu_baseline:
            rate_lg, t_solve, t_lg = cpu_fit_rates(args.cpu_sample, os.cpu_count())
            Mjob = M * world
            line["cpu_baseline"] = {"value": Mjob / (Mjob / rate_lg + t_solve), "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                                    "cpu_sample_snapshots": args.cpu_sample,
                                    "sample": f"{args.cpu_sample} snapshots of the same workload (oracle: NumPy lift + OpenBLAS Gram {t_lg:.1f} s, "
                                              f"dgeqp3 solve {t_solve:.1f} s), per-snapshot part extrapolated linearly to {Mjob} snapshots"}
        emit(line)
    fit.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
