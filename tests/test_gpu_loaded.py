"""`loaded` models (Ksysid.m:539-626 def_observables_loaded, 1006-1057 regressors, 1192-1203 / 1251-1259 /
1320-1327 model slicing, 1657-1670 / 1751-1764 validation): the lifted state is [1; w] (x) psi.  GPU path through the C ABI
(kf_problem.w / nw) against the NumPy oracle on the same inputs."""
import numpy as np
import pytest

import koopfit
import oracle as O
from koopfit.ksysid import Ksysid

pytestmark = pytest.mark.gpu


def relF(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def loaded_pairs(M, n, m, nw, seed=0):
    rng = np.random.default_rng(seed)
    alpha = 2 * rng.random((M, n)) - 1
    u = 2 * rng.random((M, m)) - 1
    w = 2 * rng.random((M, nw)) - 1
    A0 = 0.9 * np.linalg.qr(rng.standard_normal((n, n)))[0]
    beta = np.clip(alpha @ A0.T + 0.2 * u @ rng.standard_normal((m, n)) + 0.3 * w[:, :1] * alpha + 0.01 * rng.standard_normal((M, n)), -1, 1)
    return alpha, beta, u, w


@pytest.mark.parametrize("model", ["linear", "bilinear", "nonlinear"])
@pytest.mark.parametrize("nw", [1, 2])
def test_loaded_fit_matches_oracle(fitter, model, nw):
    M, n, m = 4321, 3, 2
    alpha, beta, u, w = loaded_pairs(M, n, m, nw, seed=nw)
    nv = n + (m if model == "nonlinear" else 0)
    basis = koopfit.Basis(["poly"], [2], nv)
    prog = O.build_program(["poly"], [2], nv)
    Px, Py = O.build_regressors(model, prog, alpha, beta, u, w)
    N = prog.N
    assert Px.shape[1] == O.regressor_width(model, N, m, nw)
    Ko = O.mldivide(Px, Py)
    for route in ("gram", "qr"):
        res = fitter.fit(basis, model, alpha, beta, u, w=w, want_gram=True, want_regressors=True, ls_method=route)
        assert res["P"] == Px.shape[1]
        assert np.array_equal(res["Px"], Px) and np.array_equal(res["Py"], Py)       # products only: bit-exact, same association
        assert relF(res["G"], Px.T @ Px) < 1e-13 and relF(res["C"], Px.T @ Py) < 1e-13
        assert res["rank"] == Px.shape[1] and relF(res["K"], Ko) < 1e-9, route


def test_loaded_constant_zero_load_is_rank_deficient_like_mldivide(fitter):
    """The shipped arm file carries w = 0 (no load): with loaded = true the w psi columns vanish and mldivide's basic
    solution leaves their rows of K at zero."""
    alpha, beta, u, _ = loaded_pairs(3000, 3, 2, 1, seed=5)
    w = np.zeros((3000, 2))
    basis = koopfit.Basis(["poly"], [2], 3)
    prog = O.build_program(["poly"], [2], 3)
    Px, Py = O.build_regressors("bilinear", prog, alpha, beta, u, w)
    Ko, info = O.mldivide(Px, Py, return_info=True)
    res = fitter.fit(basis, "bilinear", alpha, beta, u, w=w, ls_method="gram")
    assert res["rank"] == info["rank"] == prog.N * 3
    assert set(res["perm"][:res["rank"]].tolist()) == set(info["perm"][:info["rank"]].tolist())
    assert relF(res["K"], Ko) < 1e-9


def _loaded_data(seed=0, ntrials=4, T=400, n=2, m=1, nw=1):
    rng = np.random.default_rng(seed)
    trials = []
    for k in range(ntrials + 2):
        wk = rng.uniform(0.0, 2.0, size=nw)                     # one load per trial
        y = np.zeros((T, n))
        u = np.cumsum(rng.standard_normal((T, m)), axis=0) * 0.05
        y[0] = rng.uniform(-0.5, 0.5, n)
        for i in range(T - 1):
            y[i + 1] = 0.95 * y[i] + 0.1 * np.tanh(u[i].sum()) - 0.02 * wk.sum() * y[i] ** 3 + 0.01 * wk.sum()
        trials.append({"t": 0.01 * np.arange(T), "y": y, "u": u, "w": np.tile(wk, (T, 1))})
    return {"train": trials[:ntrials], "val": trials[ntrials:]}


@pytest.mark.parametrize("model", ["linear", "bilinear", "nonlinear"])
def test_ksysid_loaded_option_end_to_end(fitter, model):
    """Ksysid(..., 'loaded', true): snapshot pairs carry w, the model matrices have N (nw+1) lifted states, and the loaded
    validation rollout (re-expansion with the actual load each step) matches the oracle's."""
    data = _loaded_data()
    ks = Ksysid(data, model_type=model, obs_type=["poly"], obs_degree=[2], loaded=True, dim_red=False, fitter=fitter).train_models()
    n, m, nw, N = ks.params["n"], ks.params["m"], ks.params["nw"], ks.params["N"]
    assert nw == 1 and "w" in ks.snapshotPairs
    # oracle on the same pre-processing
    merged = O.merge_trials(data["train"])
    tr, sc = O.get_scale(merged)
    pairs = O.get_snapshot_pairs(tr, 0)
    assert np.allclose(pairs["w"], ks.snapshotPairs["w"]) and np.allclose(pairs["alpha"], ks.snapshotPairs["alpha"])
    nv = n + (m if model == "nonlinear" else 0)
    prog = O.build_program(["poly"], [2], nv)
    koop = O.get_koopman(model, prog, pairs, lasso=1e6, N=prog.N, n=n, loaded=True)
    assert relF(ks.koopData["K"], koop["K"]) < 1e-9
    NL = N * (nw + 1)
    if model == "linear":
        mo = O.get_model(koop, n)
        assert ks.model["A"].shape == (NL, NL) and ks.model["B"].shape == (NL, m) and ks.model["C"].shape == (n, NL)
        assert relF(ks.model["A"], mo["A"]) < 1e-8 and relF(ks.model["B"], mo["B"]) < 1e-8
        val = O.scale_data(data["val"][0], sc)
        want = O.val_model(mo, prog, val, 0, n, loaded=True)["error"]["rmse"]
        got = ks.val_model(ks.model, ks.valdata[0])["error"]["rmse"]
        assert np.abs(got - want).max() < 1e-6
    elif model == "bilinear":
        mo = O.get_BLmodel(koop, n)
        assert ks.model["A"].shape == (NL, NL) and ks.model["B"].shape == (NL, NL * m)
        assert relF(ks.model["A"], mo["A"]) < 1e-9 and relF(ks.model["B"], mo["B"]) < 1e-9
        r = ks.val_BLmodel(ks.model, ks.valdata[0])
        assert np.all(np.isfinite(r["error"]["rmse"])) and r["sim"]["z"].shape[1] == NL
    else:
        assert ks.model["F_sym"].shape == (n, NL)
        zeta, uu, ww = pairs["alpha"][7], pairs["u"][7], pairs["w"][7]
        want = koop["K"][:, :n].T @ O.load_lift(O.lift(prog, np.concatenate([zeta, uu])[None, :]), ww[None, :])[0]
        assert np.allclose(ks.model["F_func"](zeta, uu, ww), want, rtol=1e-9, atol=1e-12)


def test_loaded_needs_the_load_field(fitter):
    data = _loaded_data()
    for tr in data["train"]:
        tr.pop("w")
    with pytest.raises(ValueError, match="required load field"):
        Ksysid(data, model_type="linear", obs_type=["poly"], obs_degree=[2], loaded=True, dim_red=False, fitter=fitter)


@pytest.mark.parametrize("N,m,h,nz", [(34, 3, 10, "h"), (34, 3, 10, 1), (7, 1, 1, 1), (100, 2, 25, "h")])
def test_mpc_costB_bilinear_matches_kmpc(fitter, N, m, h, nz):
    """Kmpc.get_costB_bilinear (Kmpc.m:569-596) on the GPU against the NumPy restatement, single problem and a batch."""
    rng = np.random.default_rng(N + h)
    Amat = 0.9 * np.linalg.qr(rng.standard_normal((N, N)))[0]
    Bmat = 0.1 * rng.standard_normal((N, N * m))
    rows = h if nz == "h" else 1
    z = rng.standard_normal((rows, N))
    got = fitter.mpc_costB_bilinear(Amat, Bmat, z, h)
    want = O.mpc_costB_bilinear(Amat, Bmat, z, h)
    assert got.shape == want.shape == (N * (h + 1), m * h)
    assert np.abs(got - want).max() <= 1e-12 * max(1.0, np.abs(want).max())
    zb = rng.standard_normal((5, rows, N))
    gb = fitter.mpc_costB_bilinear(Amat, Bmat, zb, h)
    for b in range(5):
        assert np.abs(gb[b] - O.mpc_costB_bilinear(Amat, Bmat, zb[b], h)).max() <= 1e-12 * max(1.0, np.abs(want).max())
