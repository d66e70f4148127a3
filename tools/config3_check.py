"""Config 3a at full size on the GPU: snake-data, bilinear, fourier degree 4 (N = 732, P = 1464):
LS through the QRCP route vs the oracle's dgeqp3, and the L1-ball QP for a few budgets (timing, KKT check)."""
import sys, time, os, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import koopfit, oracle as O
from conftest import unpack, GOLDEN

snake = unpack(np.load(os.path.join(GOLDEN, "snake_data.npz")))
k = O.KsysidOracle(snake, model_type="bilinear", obs_type=["fourier"], obs_degree=[4])
print("N", k.N, "pairs", k.pairs["alpha"].shape, flush=True)
fit = koopfit.Fitter(0)
basis = koopfit.Basis(["fourier"], [4], 3)
out = {}
t0 = time.time(); res = fit.fit(basis, "bilinear", k.pairs["alpha"], k.pairs["beta"], k.pairs["u"], want_gram=True); t_gpu = time.time() - t0
t0 = time.time(); res = fit.fit(basis, "bilinear", k.pairs["alpha"], k.pairs["beta"], k.pairs["u"], want_gram=True); t_gpu = time.time() - t0
print("GPU LS (QR) s", round(t_gpu, 3), "rank", res["rank"], res["info"], flush=True)
t0 = time.time()
Px, Py = O.build_regressors("bilinear", k.prog, k.pairs["alpha"], k.pairs["beta"], k.pairs["u"])
Ko, info = O.mldivide(Px, Py, return_info=True); t_cpu = time.time() - t0
G, C = O.gram(Px, Py)
relK = np.linalg.norm(res["K"] - Ko) / np.linalg.norm(Ko)
N = k.N
relA = np.linalg.norm(res["K"].T[:N, :N] - Ko.T[:N, :N]) / np.linalg.norm(Ko.T[:N, :N])
relB = np.linalg.norm(res["K"].T[:N, N:] - Ko.T[:N, N:]) / np.linalg.norm(Ko.T[:N, N:])
print("CPU oracle s", round(t_cpu, 2), "rank", info["rank"], "relK", relK, "relA", relA, "relB", relB,
      "relG", np.linalg.norm(res["G"] - G) / np.linalg.norm(G), flush=True)
out.update(ls=dict(gpu_s=t_gpu, cpu_s=t_cpu, rank_gpu=int(res["rank"]), rank_cpu=int(info["rank"]), relK=relK, relA=relA, relB=relB))
# QP: budgets of the 64-value sweep logspace(-2, 2, 64) * N, one call per budget so that each is timed;
# inner sweeps bounded (qp_max_iter) so an ill-conditioned Gram cannot run away
all_l = np.logspace(-2, 2, 64)
sel = [int(x) for x in os.environ.get("KF_QP_BUDGETS", "0,21").split(",")]
out["qp"] = []
for idx in sel:
    lam = all_l[idx]
    t0 = time.time()
    rq = fit.fit(basis, "bilinear", k.pairs["alpha"], k.pairs["beta"], k.pairs["u"], least_squares=False, t=[lam * N],
                 psd_shift="never", qp_max_iter=int(os.environ.get("KF_QP_SWEEPS", "300")))
    t_qp = time.time() - t0
    K = rq["K"]
    grad = G @ K - C
    nz = K != 0
    lam_est = np.median(np.abs(grad[nz])) if nz.any() else 0.0
    viol = np.max(np.abs(grad[~nz])) - lam_est if (~nz).any() else 0.0
    spread = np.max(np.abs(np.abs(grad[nz]) - lam_est)) if nz.any() else 0.0
    rec = dict(budget=float(lam * N), seconds=t_qp, evals=int(rq["qp_iters"][0]), capped=int(rq["info"]["qp_capped"]), l1=float(rq["l1norm"][0]),
               nnz=int(nz.sum()), multiplier=float(lam_est), kkt_inactive_viol=float(viol), kkt_active_spread=float(spread),
               objective=float(O.qp_objective(G, C, K)))
    print(rec, flush=True)
    out["qp"].append(rec)
json.dump(out, open("gpurun_out/config3_check.json", "w"), indent=1)
