"""BASELINE config 3a at full size on one GPU: snake-data, bilinear, fourier degree 4 (N = 732, P = 1464, cond(G) ~ 2e13),
the whole lasso vector logspace(-2, 2, 64) * N in ONE kf_fit call (exact active-set solver), every budget certified by
its Frank-Wolfe gap (kf_result.qp_gap).  KF_SWEEP_N limits the number of budgets (the smallest ones)."""
import sys, time, os, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import koopfit
from koopfit.ksysid import Ksysid
from conftest import unpack, GOLDEN

snake = unpack(np.load(os.path.join(GOLDEN, "snake_data.npz")))
fit = koopfit.Fitter(0)
if os.environ.get("KF_AS_FRAC"):
    fit.set_option("as_frac", float(os.environ["KF_AS_FRAC"]))
for kv in filter(None, os.environ.get("KF_OPTS", "").split(",")):
    fit.set_option(kv.split("=")[0], float(kv.split("=")[1]))
if os.environ.get("KF_AS_DIAG"):
    fit.set_option("as_diag", 1)
nb = int(os.environ.get("KF_SWEEP_N", "64"))
lassos = np.logspace(-2, 2, 64)[:nb]
ks = Ksysid(snake, model_type="bilinear", obs_type=["fourier"], obs_degree=[4], lasso=lassos, dim_red=False, fitter=fit)
N = ks.params["N"]
sp = ks.snapshotPairs
basis = ks.basis
t0 = time.time()
res = fit.fit(basis, "bilinear", sp["alpha"], sp["beta"], sp["u"], least_squares=False, t=lassos * N, psd_shift="never",
              qp_max_iter=int(os.environ.get("KF_QP_ITERS", "0")))
dt = time.time() - t0
f = res["objective"]
rel = res["qp_gap"] / np.abs(f)
rows = []
for i in range(nb):
    K = res["K_all"][:, :, i]
    rows.append(dict(budget=float(lassos[i] * N), objective=float(f[i]), l1=float(res["l1norm"][i]), rel_gap=float(rel[i]),
                     steps=int(res["qp_iters"][i]), nnz=int(np.count_nonzero(K))))
    print(rows[-1], flush=True)
out = dict(P=int(res["P"]), N=int(N), budgets=nb, seconds=dt, solve_ms=res["info"]["t_solve_ms"], lift_gram_ms=res["info"]["t_lift_gram_ms"],
           capped=int(res["info"]["qp_capped"]), worst_rel_gap=float(rel.max()), total_steps=int(res["qp_iters"].sum()), rows=rows)
print({k: v for k, v in out.items() if k != "rows"}, flush=True)
json.dump(out, open("gpurun_out/config3_sweep.json", "w"), indent=1)
if os.environ.get("KF_SAVE_K"):
    for i in [int(x) for x in os.environ["KF_SAVE_K"].split(",")]:
        np.save(f"gpurun_out/c3a_K_{i}.npy", res["K_all"][:, :, i])
