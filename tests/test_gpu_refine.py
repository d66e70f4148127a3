"""GPU parity tests of the Gram route on ILL-CONDITIONED regressors.

`K = Px \\ Py` (Ksysid.m:1069) is a QR solve in the reference and never squares the condition number; the Gram route
— the only one that shards and scales — does, so it re-orthogonalises with extra data passes (multi-level pivoted
Cholesky-QR, api.cu: gram_ls_step / refine_pass).  These tests run it where it is weakest: BASELINE config 2a
(cond(Px(:,basic)) = 1.9e7, rank 723 of 819), 2b (rank 1216 of 1330) and 3a's least-squares fit (cond 4e6), through
kf_fit and through the staged 2-shard kf_accumulate_dev -> kf_solve_dev path (KF_EAGAIN loop).
Tolerance (north_star): relative Frobenius error of K (A, B, F) <= 1e-9 against the dgeqp3 oracle, and closer than 1e-9 to
the extended-precision basic solution where that is affordable.
"""
import numpy as np
import pytest

import koopfit
import oracle as O

pytestmark = pytest.mark.gpu


def relF(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def _oracle_case(data, model, types, degs, nd):
    k = O.KsysidOracle(data, model_type=model, obs_type=types, obs_degree=degs, delays=nd)
    koop = O.get_koopman(model, k.prog, k.pairs, lasso=1e6, N=k.N, n=k.n, nd=nd)
    nv = k.nzeta + (k.m if model == "nonlinear" else 0)
    return k, koop, koopfit.Basis(types, degs, nv)


@pytest.fixture(scope="module")
def case2a(arm_data):
    return _oracle_case(arm_data, "linear", ["poly"], [3], 1)


@pytest.fixture(scope="module")
def case2b(arm_data):
    return _oracle_case(arm_data, "nonlinear", ["poly"], [3], 1)


@pytest.fixture(scope="module")
def case3a(snake_data):
    return _oracle_case(snake_data, "bilinear", ["fourier"], [4], 0)


def _check(res, koop, rank, tol=1e-9):
    assert res["info"]["ls_method_used"] == 1                       # KF_LS_GRAM
    assert res["rank"] == rank == koop["info"]["rank"]
    assert set(res["perm"][:rank].tolist()) == set(koop["info"]["perm"][:rank].tolist())   # mldivide's basic set
    assert np.all(res["K"][np.asarray(res["perm"][rank:], dtype=int)] == 0)                  # basic solution: exact zeros
    err = relF(res["K"], koop["K"])
    assert err < tol, err
    return err


def test_config2a_gram_route_refined(fitter, case2a):
    """Config 2a through ls_method='gram': cond 1.9e7 -> plain Gram-Cholesky is off by ~1e-3; the refined
    route must reproduce the basic set and K to 1e-9, with the refinement reported in kf_info."""
    k, koop, basis = case2a
    res = fitter.fit(basis, "linear", k.pairs["alpha"], k.pairs["beta"], k.pairs["u"], ls_method="gram")
    info = res["info"]
    assert info["cond_est"] > 1e5 and info["refine_passes"] >= 1 and info["refine_capped"] == 0
    assert info["passes"] == 1 + info["refine_passes"]
    _check(res, koop, 723)
    N = k.N
    K, Ko = res["K"], koop["K"]
    assert relF(K.T[:N, :N], Ko.T[:N, :N]) < 1e-9 and relF(K.T[:N, N:], Ko.T[:N, N:]) < 1e-9     # A, B
    # against the extended-precision basic solution on the same set: as close as LAPACK's own QR
    basic = np.sort(koop["info"]["perm"][:723])
    truth = O.basic_solution_extended(koop["Px"], koop["Py"][:, :N], basic).astype(np.float64)
    e_gpu, e_lapack = relF(K[:, :N], truth), relF(Ko[:, :N], truth)
    assert e_gpu < 1e-9 and e_gpu < 20 * max(e_lapack, 1e-12), (e_gpu, e_lapack)


def test_config2a_refinement_off_shows_the_gap(fitter, case2a):
    """With refine = 0 the Gram route is cond^2 * eps: this is the failure the refinement removes (and the reason
    the option exists at all is to measure it)."""
    k, koop, basis = case2a
    fitter.set_option("refine", 0)
    try:
        res = fitter.fit(basis, "linear", k.pairs["alpha"], k.pairs["beta"], k.pairs["u"], ls_method="gram")
    finally:
        fitter.set_option("refine", 1)
    assert res["info"]["refine_passes"] == 0 and res["info"]["passes"] == 1
    assert relF(res["K"], koop["K"]) > 1e-7


def test_config2b_gram_route_refined(fitter, case2b):
    k, koop, basis = case2b
    res = fitter.fit(basis, "nonlinear", k.pairs["alpha"], k.pairs["beta"], k.pairs["u"], ls_method="gram")
    assert res["info"]["refine_passes"] >= 1 and res["info"]["refine_capped"] == 0
    _check(res, koop, 1216, tol=2e-9)                                # whole K; the consumed columns below to 1e-9
    assert relF(res["K"][:, :15], koop["K"][:, :15]) < 1e-9          # F = K(:,1:nzeta)' (Ksysid.m:1329)


def test_config3a_ls_gram_route_refined(fitter, case3a):
    """Snake, bilinear, fourier 4 (P = 1464, full rank, cond 4e6): least squares through the Gram route."""
    k, koop, basis = case3a
    res = fitter.fit(basis, "bilinear", k.pairs["alpha"], k.pairs["beta"], k.pairs["u"], ls_method="gram")
    assert res["info"]["refine_passes"] >= 1
    _check(res, koop, 1464)
    N = k.N
    assert relF(res["K"].T[:N, :N], koop["K"].T[:N, :N]) < 1e-9 and relF(res["K"].T[:N, N:], koop["K"].T[:N, N:]) < 1e-9


@pytest.mark.parametrize("which", ["2a", "3a"])
def test_two_shard_staged_path_refined(fitter, case2a, case3a, which):
    """The sharded path: two shards accumulated one after the other (what two ranks + all-reduce produce), then
    kf_solve_dev; it answers KF_EAGAIN (solve_dev -> None) until the refinement passes over BOTH shards are in."""
    import torch
    k, koop, basis = case2a if which == "2a" else case3a
    model, rank = ("linear", 723) if which == "2a" else ("bilinear", 1464)
    al, be, uu = k.pairs["alpha"], k.pairs["beta"], k.pairs["u"]
    M, nz, m = al.shape[0], al.shape[1], uu.shape[1]
    dev = torch.device("cuda:0")
    cut = M // 2 + 77
    shards = []
    for lo, hi in ((0, cut), (cut, M)):
        shards.append(tuple(torch.tensor(np.ascontiguousarray(x[lo:hi].T), device=dev) for x in (al, be, uu)) + (hi - lo,))
    torch.cuda.synchronize()

    def data_pass(reset):
        for i, (ta, tb, tu, n) in enumerate(shards):
            fitter.accumulate_dev(basis, model, n, nz, m, ta.data_ptr(), tb.data_ptr(), tu.data_ptr(), reset=reset and i == 0)
        fitter.accum_buffer()          # what a rank hands to the all-reduce (also writes the snapshot-count trailer)
        fitter.sync()

    data_pass(True)
    passes = 1
    while True:
        res = fitter.solve_dev(ls_method="gram")
        if res is not None:
            break
        data_pass(False)
        passes += 1
        assert passes <= 5
    assert passes >= 2 and res["info"]["passes"] == passes
    _check(res, koop, rank)


def test_well_conditioned_fit_stays_single_pass(fitter):
    """Config-5-like data (cond ~1e2): no refinement, one data pass — the benchmarked path is unchanged."""
    rng = np.random.default_rng(0)
    M, n, m = 20000, 6, 2
    alpha = 2 * rng.random((M, n)) - 1
    u = 2 * rng.random((M, m)) - 1
    beta = np.clip(alpha @ (0.9 * np.linalg.qr(rng.standard_normal((n, n)))[0]).T + 0.01 * rng.standard_normal((M, n)), -1, 1)
    cen = 2 * np.random.default_rng(2).random((n, 40)) - 1
    basis = koopfit.Basis(["poly", "gaussian"], [2, 40], n, cen)
    res = fitter.fit(basis, "bilinear", alpha, beta, u, ls_method="gram")
    assert res["info"]["passes"] == 1 and res["info"]["refine_passes"] == 0
    assert res["info"]["cond_est"] < 1e3
    prog = O.build_program(["poly", "gaussian"], [2, 40], n, cen)
    Px, Py = O.build_regressors("bilinear", prog, alpha, beta, u)
    assert relF(res["K"], O.mldivide(Px, Py)) < 1e-9


def test_refine_always_on_well_conditioned_data(fitter):
    """refine = 2 forces one verification pass even on well-conditioned data: same answer, two passes."""
    rng = np.random.default_rng(1)
    M, n, m = 5000, 3, 2
    alpha = 2 * rng.random((M, n)) - 1
    u = 2 * rng.random((M, m)) - 1
    beta = np.clip(alpha @ (0.9 * np.linalg.qr(rng.standard_normal((n, n)))[0]).T + 0.01 * rng.standard_normal((M, n)), -1, 1)
    basis = koopfit.Basis(["poly"], [3], n)
    prog = O.build_program(["poly"], [3], n)
    for model in ("linear", "bilinear", "nonlinear"):
        nv = n + (m if model == "nonlinear" else 0)
        b = koopfit.Basis(["poly"], [3], nv)
        pr = O.build_program(["poly"], [3], nv)
        Px, Py = O.build_regressors(model, pr, alpha, beta, u)
        Ko = O.mldivide(Px, Py)
        fitter.set_option("refine", 2)
        try:
            res = fitter.fit(b, model, alpha, beta, u, ls_method="gram")
        finally:
            fitter.set_option("refine", 1)
        assert res["info"]["passes"] == 2 and res["rank"] == Px.shape[1]
        assert relF(res["K"], Ko) < 1e-11
    # pc_cols fast mode through the refined route: the first Pc columns only
    Px, Py = O.build_regressors("bilinear", prog, alpha, beta, u)
    Ko = O.mldivide(Px, Py)
    fitter.set_option("refine", 2)
    try:
        res = fitter.fit(basis, "bilinear", alpha, beta, u, ls_method="gram", pc_cols=prog.N)
    finally:
        fitter.set_option("refine", 1)
    assert res["K"].shape == (Px.shape[1], prog.N) and relF(res["K"], Ko[:, :prog.N]) < 1e-11
