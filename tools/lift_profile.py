"""One launch of the materialising lift per (dictionary, variant) for ncu: config-5 dictionary and poly 4 bilinear, poly 3 linear; 65 536 pairs.
argv[1]: comma list of lift_smem_kb values (wide kernel)."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import koopfit

dev = torch.device("cuda", 0)
fit = koopfit.Fitter(0)
rng = np.random.default_rng(2)
M, nz, m = 65536, 12, 3
kbs = [float(x) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["100"])]
for types, degs, cen, model in ((["poly", "gaussian"], [3, 569], 2 * rng.random((12, 569)) - 1, "bilinear"), (["poly"], [4], None, "bilinear"),
                                (["poly"], [3], None, "linear")):
    basis = koopfit.Basis(types, degs, nz, centres=cen)
    _, N, P = fit.dims(basis, model, m)
    a = torch.rand((nz, M), dtype=torch.float64, device=dev) * 2 - 1
    b = torch.rand((nz, M), dtype=torch.float64, device=dev) * 2 - 1
    u = torch.rand((m, M), dtype=torch.float64, device=dev) * 2 - 1
    o = torch.empty((2 * P, M), dtype=torch.float64, device=dev)
    for kb in kbs:
        fit.set_option("lift_smem_kb", kb)
        fit.regressors_dev(basis, model, M, nz, m, a.data_ptr(), b.data_ptr(), u.data_ptr(), o.data_ptr())
        fit.sync()
    del o
