"""GPU edge cases: ragged / tiny / degenerate inputs through the C ABI, compared with the oracle."""
import numpy as np
import pytest

import koopfit
import oracle as O

pytestmark = pytest.mark.gpu


def relF(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def data(M, n, m, seed=0):
    rng = np.random.default_rng(seed)
    alpha = 2 * rng.random((M, n)) - 1
    u = 2 * rng.random((M, m)) - 1
    beta = np.clip(0.8 * alpha + (0.1 * u @ rng.standard_normal((m, n)) if m else 0) + 0.01 * rng.standard_normal((M, n)), -1, 1)
    return alpha, beta, u


@pytest.mark.parametrize("M", [1, 5, 17, 255, 257])
@pytest.mark.parametrize("model", ["linear", "bilinear"])
def test_fewer_snapshots_than_regressors(fitter, M, model):
    """M < P: G, C exact; `\\` returns a basic solution of rank <= M (QR route, as MATLAB does)."""
    alpha, beta, u = data(M, 3, 2, seed=M)
    basis = koopfit.Basis(["poly"], [2], 3)
    prog = O.build_program(["poly"], [2], 3)
    Px, Py = O.build_regressors(model, prog, alpha, beta, u)
    res = fitter.fit(basis, model, alpha, beta, u, want_gram=True, ls_method="qr")
    G, C = O.gram(Px, Py)
    assert relF(res["G"], G) < 1e-13 and relF(res["C"], C) < 1e-13
    Ko, info = O.mldivide(Px, Py, return_info=True)
    assert res["rank"] == info["rank"] <= min(M, Px.shape[1])
    # any basic solution reproduces the data to rounding when M <= rank; compare the fitted values
    assert np.linalg.norm(Px @ res["K"] - Px @ Ko) <= 1e-8 * max(1.0, np.linalg.norm(Px @ Ko))


def test_no_inputs(fitter):
    """m = 0: P = N for every model type."""
    alpha, beta, u = data(900, 3, 0)
    basis = koopfit.Basis(["poly"], [3], 3)
    prog = O.build_program(["poly"], [3], 3)
    for model in ("linear", "bilinear", "nonlinear"):
        Px, Py = O.build_regressors(model, prog, alpha, beta, u)
        res = fitter.fit(basis, model, alpha, beta, u, want_gram=True)
        assert res["P"] == prog.N == Px.shape[1]
        assert relF(res["G"], Px.T @ Px) < 1e-13
        assert relF(res["K"], O.mldivide(Px, Py)) < 1e-9


def test_context_reuse_across_shapes(fitter):
    """One context, alternating dictionaries / models / sizes: buffers and layouts are rebuilt correctly."""
    for i, (types, degs, n, m, model, M) in enumerate([(["poly"], [2], 4, 2, "bilinear", 3000), (["fourier_sparser"], [2], 2, 1, "linear", 700),
                                                        (["poly"], [4], 2, 1, "nonlinear", 5000), (["poly"], [2], 4, 2, "bilinear", 129)]):
        alpha, beta, u = data(M, n, m, seed=i)
        nv = n + (m if model == "nonlinear" else 0)
        basis = koopfit.Basis(types, degs, nv)
        prog = O.build_program(types, degs, nv)
        Px, Py = O.build_regressors(model, prog, alpha, beta, u)
        res = fitter.fit(basis, model, alpha, beta, u, want_gram=True, ls_method="gram")
        assert relF(res["G"], Px.T @ Px) < 1e-13 and relF(res["C"], Px.T @ Py) < 1e-13


def test_unknown_observable_type_is_ignored(fitter):
    # the reference's if/elseif chain silently skips unknown names (Ksysid.m:486-501)
    b1 = koopfit.Basis(["poly", "armshape"], [2, 5], 3)
    b2 = koopfit.Basis(["poly"], [2], 3)
    assert fitter.dims(b1, "linear", 1) == fitter.dims(b2, "linear", 1)


def test_dim_red_regressors_and_fit(fitter, arm_data):
    """dim_red: regressors use [zeta; pcs' psi_full; 1] (Ksysid.m:1594-1618) for all three layouts."""
    for model in ("linear", "bilinear", "nonlinear"):
        k = O.KsysidOracle(arm_data, model_type=model, obs_type=["poly"], obs_degree=[2], dim_red=True)
        nv = 6 + (3 if model == "nonlinear" else 0)
        basis = koopfit.Basis(["poly"], [2], nv, pcs=k.prog.pcs)
        assert fitter.dims(basis, model, 3)[1] == k.N
        sel = slice(0, 4000)
        a, b, u = k.pairs["alpha"][sel], k.pairs["beta"][sel], k.pairs["u"][sel]
        Px, Py = O.build_regressors(model, k.prog, a, b, u)
        res = fitter.fit(basis, model, a, b, u, want_gram=True, want_regressors=True)
        assert np.abs(res["Px"] - Px).max() < 1e-12 and np.abs(res["Py"] - Py).max() < 1e-12
        assert relF(res["G"], Px.T @ Px) < 1e-12
        Ko, info = O.mldivide(Px, Py, return_info=True)
        assert res["rank"] == info["rank"]
        assert relF(res["K"], Ko) < 1e-8


def test_bad_arguments(fitter):
    alpha, beta, u = data(50, 3, 2)
    basis = koopfit.Basis(["poly"], [2], 3)
    with pytest.raises(KeyError):
        fitter.fit(basis, "quadratic", alpha, beta, u)                       # Invalid model_type
    with pytest.raises(koopfit.KoopfitError):
        fitter.fit(basis, "nonlinear", alpha, beta, u)                       # nv must be nzeta + m
    with pytest.raises(koopfit.KoopfitError):
        fitter.fit(basis, "linear", alpha, beta, u, least_squares=False, t=[])   # QP needs budgets
