"""Short INT8-engine fit for an ncu launch list (per-kernel durations of one chunk pipeline)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench, koopfit
M = int(os.environ.get("KF_M", "65536"))
dev = torch.device("cuda:0")
_, _, centres = bench.workload_constants()
basis = koopfit.Basis(bench.OBS_TYPE, bench.OBS_DEGREE, bench.NZETA, centres)
fit = koopfit.Fitter(0)
fit.set_option("gram_engine", float(os.environ.get("KF_ENGINE", "2")))
fit.set_option("overlap", float(os.environ.get("KF_OVERLAP", "1")))
a, b, u = bench.gen_torch(M, dev, seed=1)
torch.cuda.synchronize()
for _ in range(2):
    r = fit.fit_dev(basis, "bilinear", M, bench.NZETA, bench.M_IN, a.data_ptr(), b.data_ptr(), u.data_ptr(), ls_method="gram")
print(r["info"]["t_lift_gram_ms"], r["info"]["t_solve_ms"], r["rank"])
