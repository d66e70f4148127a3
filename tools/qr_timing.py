"""QRCP route timings: kf_mldivide-shaped problems on device data through kf_fit(ls_method=qr) is awkward to
isolate, so this times kf_mldivide's factorisation via the library counters: host copy excluded by timing a second call pair."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import koopfit

fit = koopfit.Fitter(0)
rng = np.random.default_rng(0)
out = []
for M, P, Pc in ((16384, 4096, 4096), (20000, 1464, 1464), (8192, 1024, 1024)):
    A = np.asfortranarray(rng.standard_normal((M, P)))
    B = np.asfortranarray(rng.standard_normal((M, Pc)))
    rec = {"M": M, "P": P, "Pc": Pc}
    Xs = {}
    for blocked, nb in ((1, 32), (1, 64), (1, 128), (0, 32)):
        fit.set_option("qr_blocked", blocked)
        fit.set_option("qr_nb", nb)
        fit.mldivide(A[:256, :64], B[:256, :8])
        ts = []
        for _ in range(2):
            t0 = time.perf_counter()
            X, r, perm = fit.mldivide(A, B)
            ts.append(time.perf_counter() - t0)
        Xs[blocked] = X
        rec[f"blocked_nb{nb}" if blocked else "unblocked"] = {"wall_s_incl_copies": round(min(ts), 4), "rank": int(r)}
    rec["rel_diff"] = float(np.linalg.norm(Xs[1] - Xs[0]) / np.linalg.norm(Xs[0]))
    # copies alone (H2D of [A | B], D2H of X): a rank-0 problem skips the factorisation loop quickly
    t0 = time.perf_counter(); fit.mldivide(np.zeros_like(A), B); rec["zero_matrix_wall_s"] = round(time.perf_counter() - t0, 4)
    print(rec, flush=True)
    out.append(rec)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "qr_timing.json"), "w"), indent=1)
