"""Small-P lasso sweeps (BASELINE config 3b: snake, bilinear, gaussian 4 -> P = 16; and gaussian 30 -> P = 68): the two QP solvers
(qp_method 1 = coordinate descent in lockstep, 2 = exact active set) — wall time of a 64-budget kf_fit and the objective agreement."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import koopfit, oracle as O
from conftest import unpack, GOLDEN

snake = unpack(np.load(os.path.join(GOLDEN, "snake_data.npz")))
fit = koopfit.Fitter(0)
out = []
for ng, methods in ((4, (1, 2)), (30, (2,)), (100, (2,))):       # coordinate descent on the larger ones: minutes (measured), skipped
    cen = 2 * np.random.default_rng(0).random((3, ng)) - 1
    k = O.KsysidOracle(snake, model_type="bilinear", obs_type=["gaussian"], obs_degree=[ng], centres=cen)
    basis = koopfit.Basis(["gaussian"], [ng], 3, centres=cen)
    lassos = np.logspace(-2, 2, 64)
    rec = {"gaussians": ng}
    objs = {}
    for method in methods:
        fit.set_option("qp_method", method)
        for _ in range(2):
            t0 = time.perf_counter()
            rq = fit.fit(basis, "bilinear", k.pairs["alpha"], k.pairs["beta"], k.pairs["u"], least_squares=False, t=lassos * k.N, psd_shift="never")
            dt = time.perf_counter() - t0
        objs[method] = rq["objective"].copy()
        rec["P"] = int(rq["P"])
        rec[f"method{method}"] = {"seconds": round(dt, 4), "capped": int(rq["info"]["qp_capped"]), "worst_rel_gap": float((rq["qp_gap"] / np.abs(rq["objective"])).max()),
                                  "steps": int(rq["qp_iters"].sum())}
    if len(objs) == 2:
        rec["max_rel_objective_diff"] = float(np.max(np.abs(objs[1] - objs[2]) / np.abs(objs[1])))
    print(rec, flush=True)
    out.append(rec)
fit.set_option("qp_method", 0)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "qp_small_timing.json"), "w"), indent=1)
