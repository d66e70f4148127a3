"""Config 4 timing on the shipped subset: the 19 least-squares fits per random system (linear deg 1-13, bilinear 1-6)
solved by kf_fit_batch (one CTA per problem, concurrent) vs one kf_fit call per problem, replicated to the size of the
full experiment (312 systems -> 5928 LS fits) by repeating the subset."""
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
        for model, degs in (("linear", range(1, 14)), ("bilinear", range(1, 7))):
            for deg in degs:
                problems.append(dict(basis=koopfit.Basis(["poly"], [deg], 1), model_type=model, alpha=a2, beta=b2, u=u2))
print("problems", len(problems), flush=True)
fit.fit_batch(problems[:57])
t0 = time.perf_counter(); res = fit.fit_batch(problems); t_batch = time.perf_counter() - t0
sub = problems[:57 * 4]
t0 = time.perf_counter()
seq = [fit.fit(p["basis"], p["model_type"], p["alpha"], p["beta"], p["u"]) for p in sub]
t_seq = (time.perf_counter() - t0) / len(sub) * len(problems)
err = max(np.linalg.norm(a["K"] - b["K"]) / np.linalg.norm(b["K"]) for a, b in zip(res[:len(sub)], seq))
out = dict(problems=len(problems), batch_s=t_batch, kernel_ms=fit.last_times()["solve_ms"], sequential_s_extrapolated=t_seq,
           speedup=t_seq / t_batch, max_rel_diff_vs_sequential=float(err))
print(out, flush=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "batch_timing.json"), "w"), indent=1)
