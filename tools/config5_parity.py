"""Config 5 parity on a subsample: the bench workload's first M snapshots through kf_fit (Gram route and
QR route) against the oracle's dgeqp3 solution; prints rel-F errors of G, C, K and A, B_i."""
import os, sys, time, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, koopfit, oracle as O
M = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
_, _, centres = bench.workload_constants()
alpha, beta, u = bench.gen_numpy(M, seed=7)
prog = O.build_program(bench.OBS_TYPE, bench.OBS_DEGREE, bench.NZETA, centres)
basis = koopfit.Basis(bench.OBS_TYPE, bench.OBS_DEGREE, bench.NZETA, centres)
fit = koopfit.Fitter(0)
t0 = time.time(); Px, Py = O.build_regressors("bilinear", prog, alpha, beta, u); G, C = O.gram(Px, Py)
Ko, info = O.mldivide(Px, Py, return_info=True); t_cpu = time.time() - t0
rel = lambda a, b: float(np.linalg.norm(a - b) / np.linalg.norm(b))
out = dict(M=M, P=int(Px.shape[1]), rank_cpu=int(info["rank"]), cpu_s=t_cpu, cond_Px=float(info["diagR"][0] / info["diagR"][info["rank"] - 1]))
for method in ("gram", "qr"):
    t0 = time.time(); r = fit.fit(basis, "bilinear", alpha, beta, u, want_gram=True, ls_method=method); dt = time.time() - t0
    N = 1024
    out[method] = dict(seconds=dt, rank=int(r["rank"]), relG=rel(r["G"], G), relC=rel(r["C"], C), relK=rel(r["K"], Ko),
                       relA=rel(r["K"].T[:N, :N], Ko.T[:N, :N]), relB=rel(r["K"].T[:N, N:], Ko.T[:N, N:]))
print(json.dumps(out, indent=1))
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "config5_parity.json"), "w"), indent=1)
