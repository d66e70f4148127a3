"""Diagnostic: kf_mldivide on config 2a's materialised [Px | Py(:, :15)] — blocked / no-WY / column-by-column vs the extended truth."""
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
k = O.KsysidOracle(data, model_type="linear", obs_type=["poly"], obs_degree=[3], delays=1)
koop = O.get_koopman("linear", k.prog, k.pairs, lasso=1e6, N=k.N, n=k.n, nd=1)
Px, Py = koop["Px"], koop["Py"][:, :15]
rank = 723
basic = np.sort(koop["info"]["perm"][:rank])
truth = O.basic_solution_extended(Px, Py, basic).astype(np.float64)
fit = koopfit.Fitter(0)
print("lapack", relF(koop["K"][:, :15], truth))
for mode, nb, ks in ((1, 32, 64), (1, 64, 64), (1, 128, 64), (1, 128, 16), (1, 128, 1), (2, 32, 1), (0, 32, 1)):
    fit.set_option("qr_blocked", mode)
    fit.set_option("qr_nb", nb)
    fit.set_option("qr_ksplit", ks)
    X, r, perm = fit.mldivide(Px, Py)
    print("mode", mode, "nb", nb, "ksplit", ks, "rank", r, "same set", set(perm[:r].tolist()) == set(basic.tolist()), "err vs truth", relF(X, truth), flush=True)
# random permutation of the columns: is the unpivoted first stage sensitive to the column order?
rng = np.random.default_rng(0)
pc = rng.permutation(Px.shape[1])
fit.set_option("qr_ksplit", 64)
fit.set_option("qr_nb", 128)
for mode in (1, 0):
    fit.set_option("qr_blocked", mode)
    X, r, perm = fit.mldivide(Px[:, pc], Py)
    Xo = np.zeros_like(X); Xo[pc] = X
    print("permuted columns, mode", mode, "rank", r, "err vs truth", relF(Xo, truth), flush=True)
