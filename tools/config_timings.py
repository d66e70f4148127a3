"""GPU fit vs CPU oracle for the BASELINE configs whose data ships with the repo (parity-test cases, not bench
lines): wall time of one kf_fit through the C ABI (second call, host buffers) beside the NumPy/SciPy oracle on the
box's host cores, with the parity figure.  Writes gpurun_out/config_timings.json."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import koopfit, oracle as O
from conftest import unpack, GOLDEN

arm = unpack(np.load(os.path.join(GOLDEN, "arm_data.npz")))
snake = unpack(np.load(os.path.join(GOLDEN, "snake_data.npz")))
fit = koopfit.Fitter(0)
rows = []


def relF(a, b):
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


def ls_case(name, data, model, types, degs, delays=0, centres=None):
    k = O.KsysidOracle(data, model_type=model, obs_type=types, obs_degree=degs, delays=delays, centres=centres)
    nv = k.nzeta + (k.m if model == "nonlinear" else 0)
    basis = koopfit.Basis(types, degs, nv, centres=centres)
    for _ in range(2):
        t0 = time.perf_counter()
        res = fit.fit(basis, model, k.pairs["alpha"], k.pairs["beta"], k.pairs["u"])
        t_gpu = time.perf_counter() - t0
    t0 = time.perf_counter()
    koop = O.get_koopman(model, k.prog, k.pairs, lasso=1e6, N=k.N, n=k.n, nd=delays)
    t_cpu = time.perf_counter() - t0
    rec = dict(config=name, M=int(k.pairs["alpha"].shape[0]), P=int(koop["Px"].shape[1]), rank_gpu=int(res["rank"]),
               rank_cpu=int(koop["info"]["rank"]), method="qr" if res["info"]["ls_method_used"] == 2 else "gram",
               gpu_s=t_gpu, cpu_oracle_s=t_cpu, speedup=t_cpu / t_gpu, relF_K=relF(res["K"], koop["K"]),
               gpu_lift_gram_ms=res["info"]["t_lift_gram_ms"], gpu_solve_ms=res["info"]["t_solve_ms"])
    print(rec, flush=True)
    rows.append(rec)


ls_case("C1 arm bilinear poly2", arm, "bilinear", ["poly"], [2])
ls_case("C1' arm bilinear poly3", arm, "bilinear", ["poly"], [3])
ls_case("C2a arm linear poly3 delays=1", arm, "linear", ["poly"], [3], delays=1)
ls_case("C2b arm nonlinear poly3 delays=1", arm, "nonlinear", ["poly"], [3], delays=1)
ls_case("C3a snake bilinear fourier4 (LS)", snake, "bilinear", ["fourier"], [4])
cen = 2 * np.random.default_rng(0).random((3, 4)) - 1
ls_case("C3b snake bilinear gaussian4 (LS)", snake, "bilinear", ["gaussian"], [4], centres=cen)
# C3b lasso sweep: 64 budgets
k = O.KsysidOracle(snake, model_type="bilinear", obs_type=["gaussian"], obs_degree=[4], centres=cen)
basis = koopfit.Basis(["gaussian"], [4], 3, centres=cen)
lassos = np.logspace(-2, 2, 64)
for _ in range(2):
    t0 = time.perf_counter()
    rq = fit.fit(basis, "bilinear", k.pairs["alpha"], k.pairs["beta"], k.pairs["u"], least_squares=False, t=lassos * k.N, psd_shift="never")
    t_gpu = time.perf_counter() - t0
Px, Py = O.build_regressors("bilinear", k.prog, k.pairs["alpha"], k.pairs["beta"], k.pairs["u"])
t0 = time.perf_counter()
G, C = O.gram(Px, Py)
worst = 0.0
for i in range(0, 64, 8):
    Ko, _ = O.solve_l1ball_qp(G, C, lassos[i] * k.N)
    fo, fg = O.qp_objective(G, C, Ko), O.qp_objective(G, C, rq["K_all"][:, :, i])
    worst = max(worst, abs(fg - fo) / abs(fo))
t_cpu = (time.perf_counter() - t0) * 8
rec = dict(config="C3b snake bilinear gaussian4, 64-budget lasso sweep", M=int(Px.shape[0]), P=16, gpu_s=t_gpu,
           cpu_oracle_s_est=t_cpu, worst_rel_objective_gap_on_8_budgets=worst, qp_capped=int(rq["info"]["qp_capped"]))
print(rec, flush=True)
rows.append(rec)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "config_timings.json"), "w"), indent=1)
