"""Config 4 timing on the shipped subset: the 23 fits per random system of evaluate_rand_models.m:47-143 (linear deg 1-13,
bilinear 1-6 by least squares; nonlinear 1-4 with lasso = 4) solved by kf_fit_batch (one CTA per problem, concurrent; QP fits
with an inactive budget included) vs one kf_fit call per problem, replicated to the size of the full experiment
(312 systems -> 7176 fits) by repeating the subset."""
import os, sys, time, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import koopfit, oracle as O
from conftest import unpack, GOLDEN
z = np.load(os.path.join(GOLDEN, "rsys_subset.npz"))
systems = [unpack(z, prefix=f"s{i}_") for i in range(int(z["nsys"]))]
fit = koopfit.Fitter(0)
problems = []
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 104          # 3 systems x 104 = 312
pairs = []
for d in systems:
    base = O.KsysidOracle(d, model_type="linear", obs_type=["poly"], obs_degree=[1])
    pairs.append(tuple(np.asfortranarray(base.pairs[k]) for k in ("alpha", "beta", "u")))
for r in range(reps):
    for (a, b, u) in pairs:
        a2, b2, u2 = (a.copy(order="F"), b.copy(order="F"), u.copy(order="F")) if r else (a, b, u)   # distinct uploads, as for distinct systems
        for model, degs in (("linear", range(1, 14)), ("bilinear", range(1, 7)), ("nonlinear", range(1, 5))):
            for deg in degs:
                pb = dict(basis=koopfit.Basis(["poly"], [deg], 2 if model == "nonlinear" else 1), model_type=model, alpha=a2, beta=b2, u=u2)
                if model == "nonlinear":
                    N = fit.dims(pb["basis"], model, 1)[1]
                    pb.update(least_squares=False, t=[4.0 * N], psd_shift="as_reference")      # lasso = 4 (evaluate_rand_models.m:121)
                problems.append(pb)
print("problems", len(problems), flush=True)
fit.fit_batch(problems[:69])
t0 = time.perf_counter(); res = fit.fit_batch(problems); t_batch = time.perf_counter() - t0
sub = problems[:69 * 4]
t0 = time.perf_counter()
seq = [fit.fit(p["basis"], p["model_type"], p["alpha"], p["beta"], p["u"],
               **{k: v for k, v in p.items() if k not in ("basis", "model_type", "alpha", "beta", "u")}) for p in sub]
t_seq = (time.perf_counter() - t0) / len(sub) * len(problems)
err = max(np.linalg.norm(a["K"] - b["K"]) / np.linalg.norm(b["K"]) for a, b in zip(res[:len(sub)], seq))
nqp = sum(1 for p in problems if p["model_type"] == "nonlinear")
nqp_batched = sum(1 for p, r in zip(problems, res) if p["model_type"] == "nonlinear" and r["info"]["ls_method_used"] == 2 and r["qp_iters"][0] == 0)
out = dict(problems=len(problems), qp_problems=nqp, qp_answered_by_the_concurrent_kernel=nqp_batched, batch_s=t_batch, kernel_ms=fit.last_times()["solve_ms"], sequential_s_extrapolated=t_seq,
           speedup=t_seq / t_batch, max_rel_diff_vs_sequential=float(err))
print(out, flush=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "batch_timing.json"), "w"), indent=1)
