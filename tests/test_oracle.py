"""CPU tests: the oracle against the reference's own golden vectors and known answers."""
import numpy as np
import pytest

import oracle as O


def test_partitions_order_matches_reference_doc():
    # partitions.m:206-219 — last variable ascending in the outermost loop
    assert O.partitions_ones(1, 3).tolist() == [[1, 0, 0], [0, 1, 0], [0, 0, 1]]
    assert O.partitions_ones(2, 3).tolist() == [[2, 0, 0], [1, 1, 0], [0, 2, 0], [1, 0, 1], [0, 1, 1], [0, 0, 2]]
    assert O.partitions_ones(2, 2).tolist() == [[2, 0], [1, 1], [0, 2]]
    assert O.partitions_ones(3, 1).tolist() == [[3]]


@pytest.mark.parametrize("nv,d", [(6, 2), (6, 3), (15, 3), (1, 13), (3, 4)])
def test_poly_dimension(nv, d):
    # N = C(nzeta+d, d) (Ksysid.m:641)
    from math import comb
    assert O.build_program(["poly"], [d], nv).N == comb(nv + d, d)


def test_fourier_dimension():
    # N = nzeta + (1+2d)^nzeta (Ksysid.m:705); snake fourier-4: 732
    assert O.build_program(["fourier"], [4], 3).N == 3 + 9 ** 3


def test_golden_lifted_states(arm_data, golden_Z):
    """res_lin.Z / res_bilin.Z shipped with the reference = lift.econ_full(scaledown.y(Y)) for
    poly-3 + dim_red models: pins get_scale, the snapshot set, the monomial set, pca and the
    [zeta; pcs' psi; 1] layout."""
    k = O.KsysidOracle(arm_data, model_type="linear", obs_type=["poly"], obs_degree=[3], dim_red=True)
    assert k.N == 34 and k.pairs["alpha"].shape == (11999, 6)
    for key in ("lin", "bil"):
        Y, Zg = golden_Z[f"{key}_Y"], golden_Z[f"{key}_Z"]
        Ysc = (Y - k.scale["y_offset"]) / k.scale["y_factor"]
        Z = O.lift(k.prog, Ysc[:Zg.shape[0]])
        assert Z.shape == Zg.shape
        assert np.abs(Z - Zg).max() < 1e-12
        assert np.all(Zg[:, -1] == 1.0)


def test_nonlinear_dim_red_dimension(arm_data):
    # shipped nonlinear model: N = 88
    k = O.KsysidOracle(arm_data, model_type="nonlinear", obs_type=["poly"], obs_degree=[3], dim_red=True)
    assert k.N == 88


def test_pair_counts(arm_data, snake_data, rsys_data):
    assert O.KsysidOracle(arm_data, delays=1).pairs["alpha"].shape == (11998, 15)
    assert O.KsysidOracle(snake_data).pairs["alpha"].shape == (19999, 3)
    assert O.KsysidOracle(rsys_data[0]).pairs["alpha"].shape[1] == 1


def test_config1_rank_and_basic_solution(arm_data):
    k = O.KsysidOracle(arm_data, model_type="bilinear", obs_type=["poly"], obs_degree=[2]).train_models()
    koop = k.koopData[0]
    assert koop["Px"].shape == (11999, 112)
    assert koop["info"]["rank"] == 100
    K = koop["K"]
    nonbasic = koop["info"]["perm"][100:]
    assert np.all(K[nonbasic] == 0)                          # basic solution: zeros off the pivot set
    # normal equations hold on the basic set
    G, Cc = O.gram(koop["Px"], koop["Py"])
    S = koop["info"]["perm"][:100]
    assert np.abs(G[np.ix_(S, S)] @ K[S] - Cc[S]).max() < 1e-8
    # model layout (Ksysid.m:1258-1262)
    m = k.model
    assert m["A"].shape == (28, 28) and m["B"].shape == (28, 84) and m["C"].shape == (6, 28)
    err = k.validate(trial=0)["error"]
    assert err["rmse"].shape == (6,) and np.all(np.isfinite(err["rmse"]))


def test_l1ball_projection():
    rng = np.random.default_rng(1)
    X = rng.standard_normal((7, 5))
    Pj = O.l1ball_project(X, 3.0)
    assert abs(np.abs(Pj).sum() - 3.0) < 1e-12
    assert np.array_equal(O.l1ball_project(X, 1e3), X)


def test_l1ball_qp_kkt():
    rng = np.random.default_rng(2)
    Px = rng.standard_normal((400, 12))
    Py = Px @ rng.standard_normal((12, 12)) * 0.3 + 0.05 * rng.standard_normal((400, 12))
    G, C = O.gram(Px, Py)
    Kls = np.linalg.solve(G, C)
    t = 0.4 * np.abs(Kls).sum()
    K, info = O.solve_l1ball_qp(G, C, t)
    assert info["active"] and abs(np.abs(K).sum() - t) < 1e-9 * t
    grad = G @ K - C
    lam = info["lam"]
    assert np.all(np.abs(grad[K == 0]) <= lam * (1 + 1e-8))
    assert np.abs(grad[K != 0] + lam * np.sign(K[K != 0])).max() < 1e-7 * max(1, lam)
    # objective no worse than a projected-gradient reference run to convergence
    L = np.linalg.eigvalsh(G)[-1]
    Kp = np.zeros_like(K)
    for _ in range(20000):
        Kp = O.l1ball_project(Kp - (G @ Kp - C) / L, t)
    assert O.qp_objective(G, C, K) <= O.qp_objective(G, C, Kp) + 1e-9 * abs(O.qp_objective(G, C, Kp))
    # inactive budget -> least squares
    K2, info2 = O.solve_l1ball_qp(G, C, 10 * np.abs(Kls).sum())
    assert not info2["active"] and np.allclose(K2, Kls)


def test_delay_constraint_pattern():
    # nd = 1: zeta+ = [y+; y; u] so columns n..nzeta-1 of K copy y and u (Ksysid.m:1139-1164)
    n, m, nd, N = 2, 1, 1, 7
    c0, c1, tgt = O.delay_constraint_targets(N, n, m, nd)
    assert (c0, c1) == (2, 5) and tgt.shape == (N + m, 3)
    assert tgt[0, 0] == 1 and tgt[1, 1] == 1 and tgt[N, 2] == 1 and tgt.sum() == 3


def test_config1_oracle_against_extended_precision_truth(arm_data):
    """K is 'parity unpinned' (no reference output exists), so the oracle's dgeqp3 solution is checked against
    the same basic solution computed in x87 extended precision: the float64 oracle is within cond * eps of the truth."""
    k = O.KsysidOracle(arm_data, model_type="bilinear", obs_type=["poly"], obs_degree=[2]).train_models()
    koop = k.koopData[0]
    basic = koop["info"]["perm"][:koop["info"]["rank"]]
    truth = O.basic_solution_extended(koop["Px"], koop["Py"], basic)
    err = float(np.linalg.norm((koop["K"] - truth).astype(np.float64)) / np.linalg.norm(truth.astype(np.float64)))
    assert err < 1e-11, err
    # the truth satisfies the normal equations on the basic set far better than float64 could
    Px, Py = koop["Px"].astype(np.longdouble), koop["Py"].astype(np.longdouble)
    res = Px[:, basic].T @ (Px[:, basic] @ truth[basic] - Py)
    assert float(np.abs(res).max()) < 1e-12


def test_continuous_time_generator_roundtrip():
    """time_type = 'continuous' (Ksysid.m:1186-1190): UT = logm(K' + 1e-12 I) / Ts, so expm(Ts UT) gives K' back; the host
    mirror's _UT and the oracle agree (both are the same few lines over scipy.linalg.logm — the reference calls MATLAB's)."""
    import scipy.linalg as sla
    from koopfit.ksysid import Ksysid
    rng = np.random.default_rng(4)
    N, m, Ts = 6, 2, 0.05
    A0 = sla.expm(Ts * (rng.standard_normal((N, N)) - 2 * np.eye(N)))          # a discrete-time map with a real logarithm
    K = np.zeros((N + m, N + m))
    K[:N, :N] = A0.T
    K[N:, :N] = 0.1 * rng.standard_normal((m, N))
    K[N:, N:] = np.eye(m)
    UT = O.continuous_UT(K, Ts)
    assert np.abs(np.imag(UT)).max() < 1e-12
    assert np.abs(sla.expm(Ts * np.real(UT)) - K.T).max() < 1e-9
    ks = Ksysid.__new__(Ksysid)                      # host logic only: no GPU context
    ks.time_type, ks.params = "continuous", {"Ts": Ts}
    assert np.abs(ks._UT(K) - np.real(UT)).max() < 1e-13
    ks.time_type = "discrete"
    assert np.array_equal(ks._UT(K), K.T)


def test_host_mirror_preprocessing_matches_oracle(arm_data):
    """The host side of the Ksysid mirror (merge / scale / zeta: Ksysid.m:180-229, 380-401, 868-907) against the oracle's
    restatement, without a GPU (module-level helpers only)."""
    from koopfit import ksysid as KS
    merged = KS.merge_trials(arm_data["train"])
    want = O.merge_trials(arm_data["train"])
    assert all(np.array_equal(merged[k], want[k]) for k in ("t", "y", "u"))
    sc = KS._scale_factors(merged)
    _, so = O.get_scale(want)
    assert all(np.array_equal(sc[k], so[k]) for k in ("y_offset", "y_factor", "u_offset", "u_factor"))
    flat = dict(merged)
    flat["y"] = merged["y"].copy()
    flat["y"][:, 2] = 7.0                                   # constant column: factor 1 (Ksysid.m:196-198)
    assert KS._scale_factors(flat)["y_factor"][2] == 1.0 and KS._scale_factors(flat)["y_offset"][2] == 7.0
    for nd in (0, 1, 3):
        z, uz = KS._zeta(merged["y"], merged["u"], nd)
        zo, uo = O.get_zeta(want, nd)
        assert np.array_equal(z, zo) and np.array_equal(uz, uo)
        assert z.shape == (merged["y"].shape[0] - nd, 6 * (nd + 1) + 3 * nd)
