"""BASELINE config 3a swept across GPUs (torchrun, one process per GPU): snake-data, bilinear, fourier degree 4
(P = 1464), lasso vector logspace(-2, 2, 64) * N.  Snapshots are sharded (one NCCL all-reduce of the partial Grams),
then the exact active-set solver is split by COLUMNS of K: every rank factors its own column block for all budgets
and only a few scalars per step cross NVLink.  Rank 0 prints the timing and the certified gaps."""
import json, os, sys, time
import numpy as np
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import koopfit
from koopfit.ksysid import Ksysid
from koopfit.sharding import fit_sharded, shard_bounds
from conftest import unpack, GOLDEN

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
snake = unpack(np.load(os.path.join(GOLDEN, "snake_data.npz")))
fit = koopfit.Fitter(local)
nb = int(os.environ.get("KF_SWEEP_N", "64"))
lassos = np.logspace(-2, 2, 64)[:nb]
import io, contextlib
with contextlib.redirect_stdout(io.StringIO()):
    ks = Ksysid(snake, model_type="bilinear", obs_type=["fourier"], obs_degree=[4], lasso=lassos, dim_red=False, fitter=fit)
N, sp = ks.params["N"], ks.snapshotPairs
P = 2 * N
lo, hi = shard_bounds(sp["alpha"].shape[0], rank, world)
a = torch.from_numpy(np.ascontiguousarray(sp["alpha"][lo:hi].T)).to(dev)
b = torch.from_numpy(np.ascontiguousarray(sp["beta"][lo:hi].T)).to(dev)
u = torch.from_numpy(np.ascontiguousarray(sp["u"][lo:hi].T)).to(dev)
split = os.environ.get("KF_SPLIT", "columns")
times = []
for rep in range(int(os.environ.get("KF_REPS", "2"))):
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    t0 = time.time()
    res = fit_sharded(fit, ks.basis, "bilinear", a, b, u, P, budgets=lassos * N, split=split, psd_shift="never")
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    times.append(time.time() - t0)
f = res["objective"]
rel = res["qp_gap"] / np.abs(f)
l1 = np.array([np.abs(res["K_all"][:, :, i]).sum() for i in range(nb)])
out = dict(n_gpus=world, split=split, P=int(P), budgets=nb, seconds=times, best_seconds=min(times), solve_ms=res["info"]["t_solve_ms"],
           capped=int(res["info"]["qp_capped"]), worst_rel_gap=float(rel.max()), total_steps=int(res["qp_iters"].sum()),
           max_l1_excess=float((l1 / (lassos * N) - 1).max()), objective_first_last=[float(f[0]), float(f[-1])],
           nnz_last=int(np.count_nonzero(res["K_all"][:, :, -1])))
if rank == 0:
    print(json.dumps(out), flush=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"config3_sweep_n{world}_{split}.json"), "w"), indent=1)
if world > 1:
    dist.destroy_process_group()
