"""The INT8 tensor-core Gram engine (csrc/ozaki.cu + oz_gemm.cuh: Ozaki scheme II — fixed-point residues, 15 INT8 GEMMs on
tcgen05 with INT32 accumulators in TMEM, CRT reconstruction) must be FP64-exact: G = Px'Px (Ksysid.m:1114) and C = Px'Py (1125)
to 1e-13 of the oracle like the DMMA engine (measured: closer than plain FP64 accumulation), the same K, for all three regressor
layouts, ragged sizes, the pc_cols fast mode and sharded accumulation."""
import numpy as np
import pytest

import koopfit
import oracle as O

pytestmark = pytest.mark.gpu


def relF(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


@pytest.fixture
def i8(fitter):
    fitter.set_option("gram_engine", 2)
    yield fitter
    fitter.set_option("gram_engine", 0)


def synth(M, n, m, seed=0):
    rng = np.random.default_rng(seed)
    alpha = 2 * rng.random((M, n)) - 1
    u = 2 * rng.random((M, m)) - 1
    A0 = 0.9 * np.linalg.qr(rng.standard_normal((n, n)))[0]
    beta = np.clip(alpha @ A0.T + 0.2 * u @ rng.standard_normal((m, n)) + 0.1 * alpha * u[:, :1] + 0.01 * rng.standard_normal((M, n)), -1, 1)
    return alpha, beta, u


@pytest.mark.parametrize("model", ["linear", "bilinear", "nonlinear"])
@pytest.mark.parametrize("types,degs", [(["poly"], [3]), (["poly", "gaussian"], [2, 37]), (["fourier_sparser", "hermite"], [2, 2])])
def test_int8_engine_matches_oracle(i8, model, types, degs):
    M, n, m = 9001, 4, 2                                   # ragged: not a multiple of the chunk
    alpha, beta, u = synth(M, n, m, seed=len(types))
    nv = n + (m if model == "nonlinear" else 0)
    ng = sum(d for t, d in zip(types, degs) if t == "gaussian")
    cen = 2 * np.random.default_rng(1).random((nv, ng)) - 1 if ng else None
    basis = koopfit.Basis(types, degs, nv, cen)
    prog = O.build_program(types, degs, nv, cen)
    Px, Py = O.build_regressors(model, prog, alpha, beta, u)
    G, C = Px.T @ Px, Px.T @ Py
    res = i8.fit(basis, model, alpha, beta, u, want_gram=True, ls_method="gram")
    assert relF(res["G"], G) < 1e-13 and relF(res["C"], C) < 1e-13, (relF(res["G"], G), relF(res["C"], C))
    assert np.array_equal(res["G"], res["G"].T)
    if all(t in ("poly", "gaussian") for t in types):      # hermite duplicates columns: rank-deficient by construction
        Ko = O.mldivide(Px, Py)
        assert relF(res["K"], Ko) < 1e-9


def test_int8_engine_agrees_with_dmma_and_is_closer_to_the_exact_gram(fitter):
    """Same panel through both engines: they agree to rounding, and against an extended-precision Gram the INT8 engine (exact
    integer accumulation) is at least as accurate as FP64 DMMA accumulation."""
    alpha, beta, u = synth(20000, 6, 2, seed=3)
    cen = 2 * np.random.default_rng(2).random((6, 60)) - 1
    basis = koopfit.Basis(["poly", "gaussian"], [2, 60], 6, cen)
    prog = O.build_program(["poly", "gaussian"], [2, 60], 6, cen)
    Px, _ = O.build_regressors("bilinear", prog, alpha, beta, u)
    out = {}
    for eng in (1, 2):
        fitter.set_option("gram_engine", eng)
        try:
            out[eng] = fitter.fit(basis, "bilinear", alpha, beta, u, want_gram=True, ls_method="gram")
        finally:
            fitter.set_option("gram_engine", 0)
    assert relF(out[2]["G"], out[1]["G"]) < 1e-14 and relF(out[2]["C"], out[1]["C"]) < 1e-14
    assert relF(out[2]["K"], out[1]["K"]) < 1e-10
    P = Px.shape[1]
    sub = slice(0, min(P, 96))
    Gx = (Px[:, sub].astype(np.longdouble).T @ Px[:, sub].astype(np.longdouble)).astype(np.float64)
    e_dmma, e_i8 = relF(out[1]["G"][sub, sub], Gx), relF(out[2]["G"][sub, sub], Gx)
    assert e_i8 < 1e-15 and e_i8 <= 2 * e_dmma + 1e-16, (e_i8, e_dmma)


def test_int8_engine_pc_cols_and_shards(i8):
    import torch
    alpha, beta, u = synth(12000, 5, 2, seed=8)
    basis = koopfit.Basis(["poly"], [3], 5)
    prog = O.build_program(["poly"], [3], 5)
    Px, Py = O.build_regressors("bilinear", prog, alpha, beta, u)
    Ko = O.mldivide(Px, Py)
    N, P = prog.N, Px.shape[1]
    fast = i8.fit(basis, "bilinear", alpha, beta, u, ls_method="gram", pc_cols=N)
    assert fast["K"].shape == (P, N) and relF(fast["K"], Ko[:, :N]) < 1e-9
    full = i8.fit(basis, "bilinear", alpha, beta, u, want_gram=True, ls_method="gram")
    dev = torch.device("cuda:0")
    first, keep = True, []
    for lo, hi in ((0, 5000), (5000, 12000)):
        ts = [torch.tensor(np.ascontiguousarray(x[lo:hi].T), device=dev) for x in (alpha, beta, u)]
        torch.cuda.synchronize()
        keep += ts
        i8.accumulate_dev(basis, "bilinear", hi - lo, 5, 2, ts[0].data_ptr(), ts[1].data_ptr(), ts[2].data_ptr(), reset=first)
        first = False
    i8.accum_buffer()
    i8.sync()
    res = i8.solve_dev(want_gram=True, ls_method="gram")
    assert relF(res["G"], full["G"]) < 1e-15 and relF(res["C"], full["C"]) < 1e-15 and relF(res["K"], Ko) < 1e-9


def test_int8_engine_scales_rows_of_very_different_magnitude(i8):
    """Rows spanning 12 orders of magnitude (per-row exponents): every entry of G keeps FP64-level relative accuracy against
    the product of its row norms."""
    rng = np.random.default_rng(5)
    M, n = 8192, 3
    alpha = (2 * rng.random((M, n)) - 1) * np.array([1.0, 1e-6, 1e-12])
    beta = 0.5 * alpha
    u = np.zeros((M, 0))
    basis = koopfit.Basis(["poly"], [1], n)
    prog = O.build_program(["poly"], [1], n)
    Px, Py = O.build_regressors("nonlinear", prog, alpha, beta, u)
    res = i8.fit(basis, "nonlinear", alpha, beta, u, want_gram=True, ls_method="gram")
    G = (Px.astype(np.longdouble).T @ Px.astype(np.longdouble)).astype(np.float64)
    nrm = np.sqrt(np.diag(G))
    assert np.max(np.abs(res["G"] - G) / np.outer(nrm, nrm)) < 1e-15


@pytest.mark.parametrize("pc", [0, 256])
def test_int8_engine_kronecker_block_symmetry(fitter, pc):
    """Bilinear regressor with N a multiple of 256: only the lower half of every G block and the C blocks (a <= b) are
    contracted, the rest is mirrored (G_(a,b) is symmetric, C_(a,b) = C_(b,a)); same G, C, K as without the shortcut and as the
    oracle.  pc_cols = N restricts Py to its first block: the C symmetry is then not used, the G symmetry still is."""
    n, m, M = 4, 2, 9000
    alpha, beta, u = synth(M, n, m, seed=21)
    cen = 2 * np.random.default_rng(4).random((n, 241)) - 1
    basis = koopfit.Basis(["poly", "gaussian"], [2, 241], n, cen)          # N = 4 + 10 + 241 + 1 = 256
    prog = O.build_program(["poly", "gaussian"], [2, 241], n, cen)
    assert prog.N == 256
    Px, Py = O.build_regressors("bilinear", prog, alpha, beta, u)
    out = {}
    for sym in (1, 0):
        fitter.set_option("gram_engine", 2)
        fitter.set_option("oz_sym", sym)
        try:
            out[sym] = fitter.fit(basis, "bilinear", alpha, beta, u, want_gram=True, ls_method="gram", pc_cols=pc)
        finally:
            fitter.set_option("gram_engine", 0)
            fitter.set_option("oz_sym", 1)
    G, C = Px.T @ Px, Px.T @ Py
    assert relF(out[1]["G"], G) < 1e-13 and np.array_equal(out[1]["G"], out[1]["G"].T)
    cols = slice(0, pc if pc else C.shape[1])
    assert relF(out[1]["C"][:, cols], C[:, cols]) < 1e-13
    assert np.array_equal(out[1]["G"], out[0]["G"])                      # exact integer arithmetic: the mirrored entries are identical
    assert np.array_equal(out[1]["C"][:, cols], out[0]["C"][:, cols])
    assert relF(out[1]["K"], out[0]["K"]) < 1e-12
