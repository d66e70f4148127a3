"""Multi-GPU inside the library: kf_create_multi / kf_fit_multi (one process, a thread
per device, NCCL owned by the library) and kf_comm_init_rank (one process per GPU).  The single-device cases run on
any box; the 2-device cases are skipped below two GPUs (`gpurun --gpus 2`).

Bars: kf_fit_multi on 2 devices == kf_fit on 1 (G, C to 1e-14, K to 1e-10); the refined Gram route and the column-split
lasso sweep give the single-GPU answers."""
import os
import subprocess
import sys

import numpy as np
import pytest

import koopfit
import oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def relF(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def ngpus():
    import torch
    return torch.cuda.device_count()


def synth(M, n, m, seed=0):
    rng = np.random.default_rng(seed)
    alpha = 2 * rng.random((M, n)) - 1
    u = 2 * rng.random((M, m)) - 1
    A0 = 0.9 * np.linalg.qr(rng.standard_normal((n, n)))[0]
    beta = np.clip(alpha @ A0.T + 0.2 * u @ rng.standard_normal((m, n)) + 0.1 * alpha * u[:, :1] + 0.01 * rng.standard_normal((M, n)), -1, 1)
    return alpha, beta, u


def test_fit_multi_on_one_device_equals_fit(fitter):
    """ndev = 1: no communicator, same code path as kf_fit (host shard = everything)."""
    alpha, beta, u = synth(9000, 4, 2, seed=3)
    basis = koopfit.Basis(["poly", "fourier_sparser"], [2, 2], 4)
    one = fitter.fit(basis, "bilinear", alpha, beta, u, want_gram=True, want_regressors=True, ls_method="gram")
    mf = koopfit.MultiFitter([0])
    try:
        assert mf.ndev == 1
        res = mf.fit(basis, "bilinear", alpha, beta, u, want_gram=True, want_regressors=True, ls_method="gram")
    finally:
        mf.close()
    assert np.array_equal(res["G"], one["G"]) and np.array_equal(res["K"], one["K"])
    assert np.array_equal(res["Px"], one["Px"]) and np.array_equal(res["Py"], one["Py"])


def test_host_shard_copy_blocks_cover_ragged_sizes(fitter):
    """kf_fit copies the host shard in blocks of whole panel chunks and contracts each block as it arrives: sizes that are
    not multiples of the chunk, a tiny chunk (many blocks) and a single chunk must all give the same Gram as the oracle."""
    basis = koopfit.Basis(["poly"], [2], 3)
    prog = O.build_program(["poly"], [2], 3)
    for M, chunk in ((257, 256), (5000, 256), (40000, 256), (12345, 1024), (700, 0)):
        alpha, beta, u = synth(M, 3, 2, seed=M)
        Px, Py = O.build_regressors("linear", prog, alpha, beta, u)
        fitter.set_option("chunk", chunk)
        try:
            res = fitter.fit(basis, "linear", alpha, beta, u, want_gram=True, ls_method="gram")
        finally:
            fitter.set_option("chunk", 0)
        assert relF(res["G"], Px.T @ Px) < 1e-13 and relF(res["C"], Px.T @ Py) < 1e-13, (M, chunk)
        assert relF(res["K"], O.mldivide(Px, Py)) < 1e-9


@pytest.mark.skipif("ngpus() < 2")
def test_fit_multi_two_devices_equals_single(fitter, arm_data):
    """kf_fit_multi on 2 devices == kf_fit on 1: G, C to 1e-14, K to 1e-10; rank-deficient config 1 keeps its basic set;
    every device writes its own rows of Px / Py."""
    mf = koopfit.MultiFitter([0, 1])
    try:
        assert mf.ndev == 2
        # synthetic, all three regressor layouts
        alpha, beta, u = synth(30001, 4, 2, seed=5)
        for model in ("linear", "bilinear", "nonlinear"):
            nv = 4 + (2 if model == "nonlinear" else 0)
            cen = 2 * np.random.default_rng(1).random((nv, 9)) - 1
            basis = koopfit.Basis(["poly", "gaussian"], [2, 9], nv, cen)
            one = fitter.fit(basis, model, alpha, beta, u, want_gram=True, want_regressors=True, ls_method="gram")
            two = mf.fit(basis, model, alpha, beta, u, want_gram=True, want_regressors=True, ls_method="gram")
            assert relF(two["G"], one["G"]) < 1e-14 and relF(two["C"], one["C"]) < 1e-14
            assert two["rank"] == one["rank"] and relF(two["K"], one["K"]) < 1e-10
            assert np.array_equal(two["Px"], one["Px"]) and np.array_equal(two["Py"], one["Py"])
        # config 1 (rank 100 of 112): same basic set, K to 1e-10 of the single-GPU Gram route and 1e-9 of the oracle
        k = O.KsysidOracle(arm_data, model_type="bilinear", obs_type=["poly"], obs_degree=[2]).train_models()
        koop = k.koopData[0]
        basis = koopfit.Basis(["poly"], [2], 6)
        one = fitter.fit(basis, "bilinear", k.pairs["alpha"], k.pairs["beta"], k.pairs["u"], ls_method="gram")
        two = mf.fit(basis, "bilinear", k.pairs["alpha"], k.pairs["beta"], k.pairs["u"])          # AUTO -> Gram on a multi-GPU fit
        assert two["info"]["ls_method_used"] == 1 and two["rank"] == 100
        assert set(two["perm"][:100].tolist()) == set(koop["info"]["perm"][:100].tolist())
        assert relF(two["K"], one["K"]) < 1e-10 and relF(two["K"], koop["K"]) < 1e-9
        with pytest.raises(koopfit.KoopfitError):
            mf.fit(basis, "bilinear", k.pairs["alpha"], k.pairs["beta"], k.pairs["u"], ls_method="qr")
    finally:
        mf.close()


@pytest.mark.skipif("ngpus() < 2")
def test_fit_multi_refined_gram_route_config2a(arm_data):
    """The ill-conditioned config 2a sharded over 2 devices: the refinement passes (one extra all-reduce each) run inside
    kf_fit_multi and reach 1e-9 with mldivide's basic set."""
    k = O.KsysidOracle(arm_data, model_type="linear", obs_type=["poly"], obs_degree=[3], delays=1)
    koop = O.get_koopman("linear", k.prog, k.pairs, lasso=1e6, N=k.N, n=k.n, nd=1)
    basis = koopfit.Basis(["poly"], [3], 15)
    mf = koopfit.MultiFitter([0, 1])
    try:
        res = mf.fit(basis, "linear", k.pairs["alpha"], k.pairs["beta"], k.pairs["u"])
    finally:
        mf.close()
    assert res["info"]["refine_passes"] >= 1 and res["rank"] == 723
    assert set(res["perm"][:723].tolist()) == set(koop["info"]["perm"][:723].tolist())
    assert relF(res["K"], koop["K"]) < 1e-9


@pytest.mark.skipif("ngpus() < 2")
def test_fit_multi_lasso_sweep_column_split(fitter):
    """A lasso vector through the exact active-set solver on 2 devices: the sweep is split by columns of K inside the
    library (device-side reduction of the step scalars, column blocks exchanged over NCCL); objectives, l1 norms and the
    certified gaps equal the single-GPU sweep, K to 1e-9."""
    alpha, beta, u = synth(6000, 3, 1, seed=11)
    basis = koopfit.Basis(["fourier"], [2], 3)            # N = 3 + 125 = 128 -> bilinear P = 256
    P = fitter.dims(basis, "bilinear", 1)[2]
    ls = fitter.fit(basis, "bilinear", alpha, beta, u, ls_method="gram")
    l1 = np.abs(ls["K"]).sum()
    budgets = np.array([0.02, 0.1, 0.3, 0.6]) * l1
    fitter.set_option("qp_method", 2)
    try:
        one = fitter.fit(basis, "bilinear", alpha, beta, u, least_squares=False, t=budgets, psd_shift="never")
    finally:
        fitter.set_option("qp_method", 0)
    mf = koopfit.MultiFitter([0, 1])
    try:
        mf.set_option("qp_method", 2)
        two = mf.fit(basis, "bilinear", alpha, beta, u, least_squares=False, t=budgets, psd_shift="never")
    finally:
        mf.close()
    assert one["info"]["qp_capped"] == 0 and two["info"]["qp_capped"] == 0
    for i in range(budgets.size):
        assert abs(two["objective"][i] - one["objective"][i]) <= 1e-10 * abs(one["objective"][i])
        assert abs(two["l1norm"][i] - budgets[i]) <= 1e-9 * budgets[i]
        assert two["qp_gap"][i] <= 1e-8 * abs(two["objective"][i])
        assert relF(two["K_all"][:, :, i], one["K_all"][:, :, i]) < 1e-9
    assert two["K_all"].shape == (P, P, budgets.size)


@pytest.mark.skipif("ngpus() < 2")
def test_one_process_per_gpu_comm_init_rank():
    """kf_comm_init_rank: two processes (torchrun), each with its own shard through kf_fit; every rank gets the full-data K."""
    script = os.path.join(ROOT, "tests", "mp_comm_check.py")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", script]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert r.stdout.count("RANK_OK") == 2, r.stdout[-2000:] + r.stderr[-2000:]
