"""GPU parity tests (run with -m gpu on the B200 box): the CUDA path through the C ABI
against the CPU oracle on the same inputs."""
import numpy as np
import pytest

import koopfit
import oracle as O

pytestmark = pytest.mark.gpu

EPS = np.finfo(float).eps


def relF(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def centres_for(types, degs, nv, seed=0):
    ng = sum(d for t, d in zip(types, degs) if t == "gaussian")
    return 2 * np.random.default_rng(seed).random((nv, ng)) - 1 if ng else None


LIFT_CASES = [
    (["poly"], [2], 6), (["poly"], [3], 15), (["poly"], [13], 1), (["hermite"], [3], 4),
    (["fourier"], [4], 3), (["fourier_sparser"], [3], 3), (["gaussian"], [8], 3),
    (["poly", "gaussian"], [3, 569], 12),
    (["poly", "fourier", "hermite", "gaussian", "fourier_sparser"], [2, 1, 2, 5, 2], 3),
]


@pytest.mark.parametrize("types,degs,nv", LIFT_CASES)
def test_lift_matches_oracle(fitter, types, degs, nv):
    cen = centres_for(types, degs, nv)
    basis = koopfit.Basis(types, degs, nv, cen)
    prog = O.build_program(types, degs, nv, cen)
    V = 2 * np.random.default_rng(3).random((1000, nv)) - 1
    got = fitter.lift(basis, V)
    want = O.lift(prog, V)
    assert got.shape == want.shape
    if all(t in ("poly", "hermite") for t in types):
        assert np.array_equal(got, want)                    # products only: bit-exact
    # sin/cos/exp: CUDA libm vs numpy, a few ulp of the result scale
    assert np.abs(got - want).max() <= 8 * EPS


def synth(M, n, m, seed=0):
    rng = np.random.default_rng(seed)
    alpha = 2 * rng.random((M, n)) - 1
    u = 2 * rng.random((M, m)) - 1
    A0 = 0.9 * np.linalg.qr(rng.standard_normal((n, n)))[0]
    B0 = 0.2 * rng.standard_normal((n, m))
    beta = np.clip(alpha @ A0.T + u @ B0.T + 0.1 * alpha * u[:, :1] + 0.01 * rng.standard_normal((M, n)), -1, 1)
    return alpha, beta, u


@pytest.mark.parametrize("model", ["linear", "bilinear", "nonlinear"])
@pytest.mark.parametrize("types,degs", [(["poly"], [2]), (["poly", "gaussian"], [2, 7]), (["fourier_sparser"], [2])])
@pytest.mark.parametrize("M", [777, 5000])
def test_gram_and_ls_synthetic(fitter, model, types, degs, M):
    """G = Px'Px, C = Px'Py (Ksysid.m:1114,1125) to 1e-13 and K = Px\\Py to 1e-9 on well-conditioned data,
    including ragged M (not a multiple of any tile) and P below one tile."""
    n, m = 3, 2
    alpha, beta, u = synth(M, n, m)
    nv = n + (m if model == "nonlinear" else 0)
    cen = centres_for(types, degs, nv)
    basis = koopfit.Basis(types, degs, nv, cen)
    prog = O.build_program(types, degs, nv, cen)
    Px, Py = O.build_regressors(model, prog, alpha, beta, u)
    res = fitter.fit(basis, model, alpha, beta, u, want_gram=True, ls_method="gram")
    G, C = O.gram(Px, Py)
    assert res["G"].shape == G.shape
    assert relF(res["G"], G) < 1e-13 and relF(res["C"], C) < 1e-13
    assert np.array_equal(res["G"], res["G"].T)             # exactly symmetric by construction
    Kref, info = O.mldivide(Px, Py, return_info=True)
    if info["rank"] == Px.shape[1]:
        assert res["rank"] == info["rank"]
        assert relF(res["K"], Kref) < 1e-9


def test_regressors_materialised(fitter):
    """koopData.Px / Py (Ksysid.m:1085-1086): the three layouts of Ksysid.m:1039-1063."""
    alpha, beta, u = synth(1234, 3, 2)
    for model in ("linear", "bilinear", "nonlinear"):
        nv = 3 + (2 if model == "nonlinear" else 0)
        basis = koopfit.Basis(["poly"], [3], nv)
        prog = O.build_program(["poly"], [3], nv)
        Px, Py = O.build_regressors(model, prog, alpha, beta, u)
        res = fitter.fit(basis, model, alpha, beta, u, want_regressors=True, ls_method="gram")
        assert np.array_equal(res["Px"], Px) and np.array_equal(res["Py"], Py)


def test_config1_arm_bilinear_poly2(fitter, arm_data):
    """BASELINE config 1: bilinear, poly 2, lasso=Inf on the arm data: P=112, rank 100 (exactly
    rank-deficient) -> mldivide's pivoted basic solution; A, B to 1e-9, val RMSE to 1e-6."""
    k = O.KsysidOracle(arm_data, model_type="bilinear", obs_type=["poly"], obs_degree=[2]).train_models()
    koop = k.koopData[0]
    basis = koopfit.Basis(["poly"], [2], 6)
    res = fitter.fit(basis, "bilinear", k.pairs["alpha"], k.pairs["beta"], k.pairs["u"], want_gram=True, ls_method="gram")
    G, C = O.gram(koop["Px"], koop["Py"])
    assert relF(res["G"], G) < 1e-13 and relF(res["C"], C) < 1e-13
    assert res["rank"] == koop["info"]["rank"] == 100
    assert set(res["perm"][:100].tolist()) == set(koop["info"]["perm"][:100].tolist())
    K = res["K"]
    N = 28
    A, B = K.T[:N, :N], K.T[:N, N:]
    assert relF(A, k.model["A"]) < 1e-9 and relF(B, k.model["B"]) < 1e-9
    assert relF(K, koop["K"]) < 1e-9
    for trial in range(5):
        want = k.validate(trial=trial)["error"]["rmse"]
        got = O.val_BLmodel({"A": A, "B": B, "C": k.model["C"]}, k.prog, k.valdata[trial], 0, 6)["error"]["rmse"]
        assert np.abs(got - want).max() < 1e-6


def test_sharded_accumulation_equals_single_pass(fitter):
    """Snapshot sharding: accumulating the shards one after the other into the same
    accumulator gives the single-pass G, C up to summation order."""
    import torch
    alpha, beta, u = synth(6000, 3, 2, seed=5)
    basis = koopfit.Basis(["poly"], [3], 3)
    P = fitter.dims(basis, "bilinear", 2)[2]
    full = fitter.fit(basis, "bilinear", alpha, beta, u, want_gram=True, ls_method="gram")
    dev = torch.device("cuda:0")
    first = True
    keep = []
    for lo, hi in ((0, 2500), (2500, 6000)):
        ta = torch.tensor(np.ascontiguousarray(alpha[lo:hi].T), device=dev)
        tb = torch.tensor(np.ascontiguousarray(beta[lo:hi].T), device=dev)
        tu = torch.tensor(np.ascontiguousarray(u[lo:hi].T), device=dev)
        torch.cuda.synchronize()
        keep += [ta, tb, tu]
        fitter.accumulate_dev(basis, "bilinear", hi - lo, 3, 2, ta.data_ptr(), tb.data_ptr(), tu.data_ptr(), reset=first)
        first = False
    fitter.sync()
    res = fitter.solve_dev(P, want_gram=True, ls_method="gram")
    assert relF(res["G"], full["G"]) < 1e-14 and relF(res["C"], full["C"]) < 1e-14
    assert relF(res["K"], full["K"]) < 1e-10


def test_errors_are_reported(fitter):
    alpha, beta, u = synth(100, 3, 2)
    basis = koopfit.Basis(["poly"], [2], 4)                  # wrong nv for a linear model with nzeta=3
    with pytest.raises(koopfit.KoopfitError):
        fitter.fit(basis, "linear", alpha, beta, u)


def test_mldivide_qr_rank_deficient(fitter):
    """kf_mldivide = MATLAB A\\B: QRCP, rank by max(size)*eps(|R11|), basic solution (Ksysid.m:1069,1216)."""
    rng = np.random.default_rng(7)
    A = rng.standard_normal((3000, 40))
    A[:, 7] = A[:, 3] - 2 * A[:, 11]            # exact dependencies -> rank 37
    A[:, 20] = A[:, 0]
    A[:, 33] = 0.5 * A[:, 1] + A[:, 2]
    B = rng.standard_normal((3000, 9))
    X, rank, perm = fitter.mldivide(A, B)
    Xo, info = O.mldivide(A, B, return_info=True)
    assert rank == info["rank"] == 37
    assert set(perm[:rank].tolist()) == set(info["perm"][:rank].tolist())
    assert np.all(X[perm[rank:]] == 0)
    assert relF(X, Xo) < 1e-11
    # full rank, ragged sizes
    A = rng.standard_normal((1001, 131)); B = rng.standard_normal((1001, 3))
    X, rank, _ = fitter.mldivide(A, B)
    assert rank == 131 and relF(X, O.mldivide(A, B)) < 1e-11


def test_config1_qr_route(fitter, arm_data):
    k = O.KsysidOracle(arm_data, model_type="bilinear", obs_type=["poly"], obs_degree=[2]).train_models()
    basis = koopfit.Basis(["poly"], [2], 6)
    res = fitter.fit(basis, "bilinear", k.pairs["alpha"], k.pairs["beta"], k.pairs["u"], ls_method="qr")
    assert res["info"]["ls_method_used"] == 2 and res["rank"] == 100
    assert relF(res["K"], k.koopData[0]["K"]) < 1e-10


def test_config1_gpu_against_extended_precision_truth(fitter, arm_data):
    """with no reference output to pin K, the x87 extended-precision basic solution (same basic set) is the
    ground truth; the GPU's QRCP route must be as close to it as the float64 LAPACK oracle is, the Gram route within
    cond^2 eps (still inside the 1e-9 tolerance)."""
    k = O.KsysidOracle(arm_data, model_type="bilinear", obs_type=["poly"], obs_degree=[2]).train_models()
    koop = k.koopData[0]
    basic = koop["info"]["perm"][:100]
    truth = O.basic_solution_extended(koop["Px"], koop["Py"], basic).astype(np.float64)
    e_oracle = relF(koop["K"], truth)
    basis = koopfit.Basis(["poly"], [2], 6)
    qr = fitter.fit(basis, "bilinear", k.pairs["alpha"], k.pairs["beta"], k.pairs["u"], ls_method="qr")
    gr = fitter.fit(basis, "bilinear", k.pairs["alpha"], k.pairs["beta"], k.pairs["u"], ls_method="gram")
    assert set(qr["perm"][:100].tolist()) == set(basic.tolist()) == set(gr["perm"][:100].tolist())
    e_qr, e_gram = relF(qr["K"], truth), relF(gr["K"], truth)
    assert e_qr < max(20 * e_oracle, 1e-13), (e_qr, e_oracle)
    assert e_gram < 1e-9, e_gram


@pytest.mark.parametrize("model,P,rank", [("linear", 819, 723), ("nonlinear", 1330, 1216)])
def test_config2_poly3_delay1(fitter, arm_data, model, P, rank):
    """BASELINE config 2: poly 3, delays=1 on the arm data (cond ~2e7, exactly rank-deficient): the QRCP
    route must reproduce mldivide's basic solution; A, B (or F) to 1e-9."""
    k = O.KsysidOracle(arm_data, model_type=model, obs_type=["poly"], obs_degree=[3], delays=1)
    koop = O.get_koopman(model, k.prog, k.pairs, lasso=1e6, N=k.N, n=k.n, nd=1)
    assert koop["Px"].shape[1] == P and koop["info"]["rank"] == rank
    nv = 15 + (3 if model == "nonlinear" else 0)
    basis = koopfit.Basis(["poly"], [3], nv)
    res = fitter.fit(basis, model, k.pairs["alpha"], k.pairs["beta"], k.pairs["u"])      # AUTO -> QR
    assert res["info"]["ls_method_used"] == 2
    assert res["rank"] == rank
    assert set(res["perm"][:rank].tolist()) == set(koop["info"]["perm"][:rank].tolist())
    K, Ko = res["K"], koop["K"]
    N = k.N
    if model == "linear":
        assert relF(K.T[:N, :N], Ko.T[:N, :N]) < 1e-9 and relF(K.T[:N, N:], Ko.T[:N, N:]) < 1e-9
    else:
        assert relF(K[:, :15], Ko[:, :15]) < 1e-9


def test_fit_sharded_single_rank_lasso_vector(fitter):
    """sharding.fit_sharded with one rank: device-resident shard -> accumulate -> (no-op all-reduce) -> budgets."""
    import torch
    from koopfit.sharding import fit_sharded
    alpha, beta, u = synth(3000, 3, 2, seed=9)
    basis = koopfit.Basis(["poly"], [2], 3)
    prog = O.build_program(["poly"], [2], 3)
    Px, Py = O.build_regressors("linear", prog, alpha, beta, u)
    G, C = O.gram(Px, Py)
    P = Px.shape[1]
    dev = torch.device("cuda:0")
    ta, tb, tu = (torch.tensor(np.ascontiguousarray(x.T), device=dev) for x in (alpha, beta, u))
    l1 = np.abs(np.linalg.solve(G, C)).sum()
    budgets = np.array([0.2, 0.6]) * l1
    res = fit_sharded(fitter, basis, "linear", ta, tb, tu, P, budgets=budgets, psd_shift="never")
    for i, t in enumerate(budgets):
        Ko, _ = O.solve_l1ball_qp(G, C, t)
        fo, fg = O.qp_objective(G, C, Ko), O.qp_objective(G, C, res["K_all"][:, :, i])
        assert abs(fg - fo) <= 1e-8 * abs(fo)
    res = fit_sharded(fitter, basis, "linear", ta, tb, tu, P, ls_method="gram")
    assert relF(res["K"], O.mldivide(Px, Py)) < 1e-9


def test_config5_parity_on_subsample(fitter):
    """BASELINE config 5 (n=12, m=3, {'poly','gaussian'},[3,569] -> N=1024, bilinear P=4096) on a subsample of the
    bench workload: the Gram route the benchmark times (Kronecker-block DMMA accumulation + pivoted Cholesky) against
    the oracle's dgeqp3 solution: A, B_i to 1e-9 (parity for C5 on a subsample)."""
    import bench
    M = 6144
    _, _, centres = bench.workload_constants()
    alpha, beta, u = bench.gen_numpy(M, seed=7)
    prog = O.build_program(bench.OBS_TYPE, bench.OBS_DEGREE, bench.NZETA, centres)
    basis = koopfit.Basis(bench.OBS_TYPE, bench.OBS_DEGREE, bench.NZETA, centres)
    Px, Py = O.build_regressors("bilinear", prog, alpha, beta, u)
    assert Px.shape == (M, 4096)
    res = fitter.fit(basis, "bilinear", alpha, beta, u, want_gram=True, ls_method="gram")
    G, C = O.gram(Px, Py)
    assert relF(res["G"], G) < 1e-13 and relF(res["C"], C) < 1e-13
    Ko, info = O.mldivide(Px, Py, return_info=True)
    assert res["rank"] == info["rank"] == 4096
    N = 1024
    assert relF(res["K"].T[:N, :N], Ko.T[:N, :N]) < 1e-9
    for i in range(3):
        assert relF(res["K"].T[:N, N * (i + 1):N * (i + 2)], Ko.T[:N, N * (i + 1):N * (i + 2)]) < 1e-9


@pytest.mark.parametrize("model,ls_method", [("linear", "gram"), ("bilinear", "gram"), ("nonlinear", "gram"), ("bilinear", "qr"), ("linear", "qr")])
def test_pc_cols_fast_mode(fitter, model, ls_method):
    """Opt-in fast mode (SURVEY §8a note): only the K columns consumed downstream — K(:,1:N) for A, B
    (Ksysid.m:1199-1200, 1258-1259) or K(:,1:nzeta) for F (1329) — are computed; they must equal the same columns
    of the full solve."""
    n, m = 6, 2
    alpha, beta, u = synth(5000, n, m, seed=4)
    nv = n + (m if model == "nonlinear" else 0)
    basis = koopfit.Basis(["poly"], [4 if model != "nonlinear" else 3], nv)     # N = 210 / 165 > 128: several tile columns
    _, N, P = fitter.dims(basis, model, m)
    pc = n if model == "nonlinear" else N
    full = fitter.fit(basis, model, alpha, beta, u, ls_method=ls_method)
    fast = fitter.fit(basis, model, alpha, beta, u, ls_method=ls_method, pc_cols=pc)
    assert fast["K"].shape == (P, pc) and full["K"].shape == (P, P)
    assert fast["rank"] == full["rank"]
    assert relF(fast["K"], full["K"][:, :pc]) < 1e-9      # same G; only the split-K summation order may differ
    with pytest.raises(koopfit.KoopfitError):
        fitter.fit(basis, model, alpha, beta, u, pc_cols=pc, least_squares=False, t=[1.0])


@pytest.mark.parametrize("model,types,degs,nz,m,M", [
    ("bilinear", ["poly"], [2], 4, 2, 4096),                 # even M: 128-bit stores, 16-snapshot tiles
    ("linear", ["poly"], [3], 5, 2, 4097),                   # odd M: scalar stores, tail tile
    ("nonlinear", ["poly", "fourier_sparser"], [2, 2], 3, 2, 1000),
    ("bilinear", ["poly"], [4], 8, 1, 530),                  # 495 features: 8-snapshot tiles
    ("linear", ["hermite", "gaussian"], [2, 7], 3, 1, 777),
])
def test_device_lift_only_mode(fitter, model, types, degs, nz, m, M):
    """kf_regressors_dev / kf_lift_dev (lift-only mode on device buffers, shared-memory tile kernel): [Px | Py] and Psi
    equal the oracle's regressors (Ksysid.m:1019-1065) — polynomial / Hermite columns bit for bit — and the level-by-level
    kernel (option lift_tile = 0) gives the identical bytes."""
    import torch
    rng = np.random.default_rng(M)
    alpha, beta, u = 2 * rng.random((M, nz)) - 1, 2 * rng.random((M, nz)) - 1, 2 * rng.random((M, m)) - 1
    nv = nz + (m if model == "nonlinear" else 0)
    cen = centres_for(types, degs, nv)
    basis = koopfit.Basis(types, degs, nv, cen)
    prog = O.build_program(types, degs, nv, cen)
    Px, Py = O.build_regressors(model, prog, alpha, beta, u)
    P = Px.shape[1]
    dev = torch.device("cuda", 0)
    ta, tb, tu = (torch.from_numpy(np.ascontiguousarray(x.T)).to(dev) for x in (alpha, beta, u))
    outs = []
    for tile in (1, 0):
        fitter.set_option("lift_tile", tile)
        out = torch.full((2 * P, M), float("nan"), dtype=torch.float64, device=dev)
        fitter.regressors_dev(basis, model, M, nz, m, ta.data_ptr(), tb.data_ptr(), tu.data_ptr(), out.data_ptr())
        fitter.sync()
        outs.append(out.cpu().numpy().T)
    fitter.set_option("lift_tile", 1)
    got = outs[0]
    assert np.array_equal(outs[0], outs[1])
    exact = all(t in ("poly", "hermite") for t in types)
    if exact:
        assert np.array_equal(got[:, :P], Px) and np.array_equal(got[:, P:], Py)
    else:
        assert np.abs(got[:, :P] - Px).max() < 1e-14 and np.abs(got[:, P:] - Py).max() < 1e-14
    # lift.econ_full on device points
    V = np.concatenate([alpha, u], axis=1)[:, :nv] if model == "nonlinear" else alpha
    tv = torch.from_numpy(np.ascontiguousarray(V.T)).to(dev)
    psi = torch.empty((prog.N, M), dtype=torch.float64, device=dev)
    fitter.lift_dev(basis, M, tv.data_ptr(), psi.data_ptr())
    fitter.sync()
    want = O.lift(prog, V)
    assert np.abs(psi.cpu().numpy().T - want).max() < (1e-300 if exact else 1e-14)


@pytest.mark.parametrize("model,types,degs,nz,m,M", [
    ("bilinear", ["poly"], [4], 8, 2, 9001),                       # 495 features in several groups, odd row count -> narrow kernel only
    ("bilinear", ["poly"], [4], 8, 2, 9002),                       # even, rows not 128-byte aligned: 64-snapshot tiles, ragged last tile
    ("bilinear", ["poly", "gaussian"], [2, 300], 6, 5, 4096),      # m = 5: general block loop; aligned rows: 32-snapshot tiles
    ("linear", ["poly", "gaussian"], [3, 40], 10, 3, 8192),        # u rows appended
    ("nonlinear", ["poly", "fourier_sparser", "hermite"], [3, 2, 3], 9, 3, 12288 + 6),   # variables 10..12 come from u
])
def test_materialised_regressors_wide_tiles(fitter, model, types, degs, nz, m, M):
    """kf_regressors_dev through the wide-tile kernel (32 / 64 snapshots per tile, contiguous slot ranges, quad-wise gaussians):
    identical bytes to the narrow tile kernel (lift_wide = 0) for both tile widths, and the oracle's [Px | Py]
    (Ksysid.m:1019-1065)."""
    import torch
    rng = np.random.default_rng(M)
    alpha, beta, u = 2 * rng.random((M, nz)) - 1, 2 * rng.random((M, nz)) - 1, 2 * rng.random((M, m)) - 1
    nv = nz + (m if model == "nonlinear" else 0)
    cen = centres_for(types, degs, nv)
    basis = koopfit.Basis(types, degs, nv, cen)
    prog = O.build_program(types, degs, nv, cen)
    Px, Py = O.build_regressors(model, prog, alpha, beta, u)
    P = Px.shape[1]
    dev = torch.device("cuda", 0)
    ta, tb, tu = (torch.from_numpy(np.ascontiguousarray(x.T)).to(dev) for x in (alpha, beta, u))
    outs = []
    try:
        for wide, ls, kb in ((0, 0, 64), (1, 0, 64), (1, 32, 24), (1, 64, 48)):
            fitter.set_option("lift_wide", wide)
            fitter.set_option("lift_ls", ls)
            fitter.set_option("lift_smem_kb", kb)
            out = torch.full((2 * P, M), float("nan"), dtype=torch.float64, device=dev)
            fitter.regressors_dev(basis, model, M, nz, m, ta.data_ptr(), tb.data_ptr(), tu.data_ptr(), out.data_ptr())
            fitter.sync()
            outs.append(out.cpu().numpy().T)
    finally:
        fitter.set_option("lift_wide", 1)
        fitter.set_option("lift_ls", 0)
        fitter.set_option("lift_smem_kb", 64)
    for o in outs[1:]:
        assert np.array_equal(outs[0], o)
    if all(t in ("poly", "hermite") for t in types):
        assert np.array_equal(outs[0][:, :P], Px) and np.array_equal(outs[0][:, P:], Py)
    else:
        assert np.abs(outs[0][:, :P] - Px).max() < 1e-14 and np.abs(outs[0][:, P:] - Py).max() < 1e-14


@pytest.mark.parametrize("M,P,Pc,dep", [(5000, 600, 37, 5), (3001, 256, 256, 0), (9000, 1100, 300, 40)])
def test_blocked_qrcp_matches_unblocked(fitter, M, P, Pc, dep):
    """Blocked route of kf_mldivide (stage 1: 128-column Householder panels + compact-WY updates on the DMMA GEMMs; stage 2:
    column-pivoted QR of R1) against the column-by-column QRCP (option qr_blocked = 0) and the LAPACK oracle: same rank, same
    basic set, the basic solution of MATLAB's A \\ B (Ksysid.m:1069, 1216)."""
    rng = np.random.default_rng(M + P)
    A = rng.standard_normal((M, P)) * np.logspace(0, -5, P)[None, :]        # graded columns: cond ~ 1e5+
    idx = rng.choice(P, 3 * dep, replace=False).reshape(dep, 3)               # disjoint triples: no tied residuals between dependent columns
    for i, j, k in idx:                                                       # exact dependencies -> rank P - dep
        A[:, i] = 0.5 * A[:, j] - 2.0 * A[:, k]
    B = rng.standard_normal((M, Pc))
    Xo, info = O.mldivide(A, B, return_info=True)
    res = {}
    try:
        for blocked in (1, 0):
            fitter.set_option("qr_blocked", blocked)
            res[blocked] = fitter.mldivide(A, B)
    finally:
        fitter.set_option("qr_blocked", 1)
    (X1, r1, p1), (X0, r0, p0) = res[1], res[0]
    assert r1 == r0 == info["rank"] == P - dep
    assert set(p1[:r1].tolist()) == set(p0[:r0].tolist()) == set(info["perm"][:r1].tolist())
    assert np.all(X1[p1[r1:]] == 0)
    assert relF(X1, X0) < 1e-10 and relF(X1, Xo) < 1e-9
