"""Config 2b (nonlinear, poly 3, delays 1; rank 1216 of 1330, cond ~2e7): error of the blocked and the column-by-column QRCP route
and of the LAPACK oracle against the x87 extended-precision basic solution on the same basic set."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import koopfit
import oracle as O
from conftest import unpack, GOLDEN            # noqa

relF = lambda a, b: float(np.linalg.norm(a - b) / np.linalg.norm(b))
data = unpack(np.load(os.path.join(GOLDEN, 'arm_data.npz')))
out = {}
for name, model, rank in (("2a", "linear", 723), ("2b", "nonlinear", 1216)):
    k = O.KsysidOracle(data, model_type=model, obs_type=["poly"], obs_degree=[3], delays=1)
    koop = O.get_koopman(model, k.prog, k.pairs, lasso=1e6, N=k.N, n=k.n, nd=1)
    nv = 15 + (3 if model == "nonlinear" else 0)
    basis = koopfit.Basis(["poly"], [3], nv)
    fit = koopfit.Fitter(0)
    Ks = {}
    for blocked in (1, 0):
        fit.set_option("qr_blocked", blocked)
        res = fit.fit(basis, model, k.pairs["alpha"], k.pairs["beta"], k.pairs["u"], ls_method="qr")
        assert res["rank"] == rank and set(res["perm"][:rank].tolist()) == set(koop["info"]["perm"][:rank].tolist())
        Ks[blocked] = res["K"]
    fit.close()
    nc = 15
    basic = np.sort(koop["info"]["perm"][:rank])
    t0 = time.time()
    truth = O.basic_solution_extended(koop["Px"], koop["Py"][:, :nc], basic).astype(np.float64)
    rec = {"M": int(koop["Px"].shape[0]), "P": int(koop["Px"].shape[1]), "rank": rank, "columns_checked": nc,
           "err_vs_truth": {"lapack_dgeqp3_oracle": relF(koop["K"][:, :nc], truth), "gpu_blocked": relF(Ks[1][:, :nc], truth),
                            "gpu_column_by_column": relF(Ks[0][:, :nc], truth)},
           "diff_vs_lapack": {"gpu_blocked": relF(Ks[1][:, :nc], koop["K"][:, :nc]), "gpu_column_by_column": relF(Ks[0][:, :nc], koop["K"][:, :nc])},
           "truth_seconds": round(time.time() - t0, 1)}
    print(name, rec, flush=True)
    out[name] = rec
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "qr_truth_check.json"), "w"), indent=1)
