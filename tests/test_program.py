"""CPU tests of the host logic: the C++ dictionary compiler (the one libkoopfit.so uses)
against the oracle, and the C-ABI library's exported symbols."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import koopfit
import oracle as O
from koopfit import _abi as A

CASES = [
    (["poly"], [2], 6), (["poly"], [3], 15), (["poly"], [3], 18), (["poly"], [1], 6), (["poly"], [13], 1),
    (["hermite"], [3], 4), (["fourier"], [4], 3), (["fourier_sparser"], [3], 3), (["gaussian"], [8], 3),
    (["poly", "gaussian"], [3, 569], 12),
    (["poly", "fourier", "hermite", "gaussian", "fourier_sparser"], [2, 1, 2, 5, 2], 3),
]


def _centres(types, degs, nv, seed=0):
    ng = sum(d for t, d in zip(types, degs) if t == "gaussian")
    return 2 * np.random.default_rng(seed).random((nv, ng)) - 1 if ng else None


@pytest.mark.parametrize("types,degs,nv", CASES)
def test_compiler_matches_oracle(hostlift, types, degs, nv):
    cen = _centres(types, degs, nv)
    b = A.Basis(types, degs, nv, cen)
    prog = O.build_program(types, degs, nv, cen)
    nf, N = C.c_int(), C.c_int()
    assert hostlift.hostlift_dims(b.ref(), C.byref(nf), C.byref(N)) == 0
    assert nf.value == prog.n_full == N.value
    kinds = np.zeros(nf.value, np.int32); a_ = kinds.copy(); b_ = kinds.copy(); cs = np.zeros(nf.value)
    ip = lambda x: x.ctypes.data_as(A.c_int_p)
    assert hostlift.hostlift_ops(b.ref(), ip(kinds), ip(a_), ip(b_), A.dptr(cs)) == 0
    got = [(int(k), int(x), int(y), float(c)) for k, x, y, c in zip(kinds, a_, b_, cs)]
    assert got == [(k, a, bb, float(c)) for k, a, bb, c in prog.ops]
    # values: the device evaluator's source compiled for the host
    rows = 257
    V = np.asfortranarray(2 * np.random.default_rng(1).random((rows, nv)) - 1)
    out = np.zeros((rows, nf.value), order="F")
    assert hostlift.hostlift_full(b.ref(), C.c_longlong(rows), A.dptr(V), A.dptr(out)) == 0
    F = O.lift_full(prog, V)
    exact = [j for j, op in enumerate(prog.ops) if op[0] in (O.OP_VAR, O.OP_CONST, O.OP_HERM)]
    assert np.array_equal(out[:, exact], F[:, exact])
    if all(t in ("poly", "hermite") for t in types):
        assert np.array_equal(out, F)                      # pure products: bit-identical
    assert np.abs(out - F).max() <= 4 * np.finfo(float).eps


def test_block_tables_match_partitions(hostlift):
    # same rows the reference enumerates (partitions.m:206-219)
    prog = O.build_program(["poly"], [3], 4)
    rows = np.concatenate([O.partitions_ones(k, 4) for k in (1, 2, 3)])
    for j, r in enumerate(rows[4:]):
        kind, a, b, _ = prog.ops[4 + j]
        assert kind == O.OP_MUL


def test_basis_validation():
    with pytest.raises(ValueError):
        A.Basis(["poly", "fourier"], [2], 3)                 # Ksysid.m:465-467
    with pytest.raises(ValueError):
        A.Basis(["gaussian"], [3], 3)                        # centres are an input (Ksysid.m:803)


def test_library_exports_every_declared_symbol():
    """libkoopfit.so loads and exports exactly what include/koopfit.h declares (no compute calls)."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "koopfit.h")).read()
    declared = set(re.findall(r"\b(kf_[A-Za-z_0-9]+)\s*\(", hdr)) - {"kf_ctx"}
    assert declared == set(A.EXPORTS), declared ^ set(A.EXPORTS)
    if not os.path.exists(A.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    lib = A.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.kf_version() == 200
    # dictionary dimensions are host-only logic
    b = A.Basis(["poly"], [2], 6)
    nf, N, P = C.c_int(), C.c_int(), C.c_int()
    assert lib.kf_basis_dims(b.ref(), A.KF_BILINEAR, 3, C.byref(nf), C.byref(N), C.byref(P)) == 0
    assert (nf.value, N.value, P.value) == (28, 28, 112)
    rows, cols = C.c_int(), C.c_int()
    assert lib.kf_block_table(A.KF_POLY, 2, 3, C.byref(rows), C.byref(cols), None) == 0
    tab = np.zeros(rows.value * cols.value, np.int32)
    assert lib.kf_block_table(A.KF_POLY, 2, 3, C.byref(rows), C.byref(cols), tab.ctypes.data_as(A.c_int_p)) == 0
    want = np.concatenate([O.partitions_ones(1, 3), O.partitions_ones(2, 3)])
    assert np.array_equal(tab.reshape(rows.value, cols.value), want)


def test_no_cuda_device_fails_loudly():
    """Without a GPU kf_create must fail (no CPU fallback)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(koopfit.KoopfitError):
        koopfit.Fitter(device=0)


@pytest.mark.parametrize("types,degs,nv,max_slots", [
    (["poly"], [4], 8, 120), (["poly"], [3], 12, 64), (["poly", "gaussian"], [3, 569], 12, 440),
    (["fourier"], [4], 3, 100), (["fourier_sparser"], [3], 4, 40), (["hermite", "poly"], [3, 2], 4, 30),
    (["poly"], [13], 1, 6), (["gaussian"], [50], 5, 16),
])
def test_lift_feature_groups(hostlift, types, degs, nv, max_slots):
    """The dependency-closed feature groups of the materialising lift (program.cpp:kf_build_lift_groups, evaluated on the host
    exactly as kf_lift_tile_kernel does): operands are always written by an earlier level of the same group, every feature
    is stored exactly once, groups respect the slot budget, and the group-wise evaluation reproduces the plain one bit for bit."""
    rng = np.random.default_rng(1)
    ng = sum(d for t, d in zip(types, degs) if t == "gaussian")
    cen = 2 * rng.random((nv, ng)) - 1 if ng else None
    b = A.Basis(types, degs, nv, centres=cen)
    nf, N = C.c_int(), C.c_int()
    assert hostlift.hostlift_dims(b.ref(), C.byref(nf), C.byref(N)) == 0
    rows = 7
    V = np.asfortranarray(2 * rng.random((rows, nv)) - 1)
    plain = np.zeros((rows, nf.value), order="F")
    assert hostlift.hostlift_full(b.ref(), C.c_longlong(rows), A.dptr(V), A.dptr(plain)) == 0
    for single in (0, 1):
        out = np.full((rows, nf.value), np.nan, order="F")
        ngroups, used = C.c_int(), C.c_int()
        rc = hostlift.hostlift_groups(b.ref(), max_slots, single, C.c_longlong(rows), A.dptr(V), A.dptr(out), C.byref(ngroups), C.byref(used))
        assert rc == 0, rc
        assert np.array_equal(out, plain)
        if single:
            assert ngroups.value == 1 and used.value == nf.value
        else:
            assert ngroups.value >= 1 and (used.value <= max_slots or ngroups.value == nf.value - nv or used.value <= nv + 1 + 2 * max(degs))
            if nf.value > 4 * max_slots:
                assert ngroups.value >= 2


@pytest.mark.parametrize("types,degs,nv,max_slots,max_rows", [
    (["poly"], [4], 8, 120, 0), (["poly"], [3], 12, 64, 100), (["poly", "gaussian"], [3, 569], 12, 128, 0),
    (["poly", "gaussian"], [3, 569], 12, 40, 200), (["fourier"], [4], 3, 100, 0), (["fourier_sparser"], [3], 4, 40, 17),
    (["hermite", "poly"], [3, 2], 4, 30, 0), (["poly"], [13], 1, 6, 0), (["gaussian"], [50], 5, 16, 8),
])
def test_lift_row_groups(hostlift, types, degs, nv, max_slots, max_rows):
    """Row groups of the streaming lift kernel (program.cpp:kf_build_lift_rowgroups, evaluated on the host exactly as
    kf_lift_stream_kernel does): only the features a group reads hold a slot, operands are written by an earlier level, the
    groups tile the output rows in order, and the evaluation reproduces the plain one bit for bit."""
    rng = np.random.default_rng(1)
    ng = sum(d for t, d in zip(types, degs) if t == "gaussian")
    cen = 2 * rng.random((nv, ng)) - 1 if ng else None
    b = A.Basis(types, degs, nv, centres=cen)
    nf, N = C.c_int(), C.c_int()
    assert hostlift.hostlift_dims(b.ref(), C.byref(nf), C.byref(N)) == 0
    rows = 7
    V = np.asfortranarray(2 * rng.random((rows, nv)) - 1)
    plain = np.zeros((rows, nf.value), order="F")
    assert hostlift.hostlift_full(b.ref(), C.c_longlong(rows), A.dptr(V), A.dptr(plain)) == 0
    out = np.full((rows, nf.value), np.nan, order="F")
    ngroups, used = C.c_int(), C.c_int()
    rc = hostlift.hostlift_rowgroups(b.ref(), max_slots, max_rows, C.c_longlong(rows), A.dptr(V), A.dptr(out), C.byref(ngroups), C.byref(used))
    assert rc == 0, rc
    assert np.array_equal(out, plain)
    assert used.value <= max(max_slots, nv + 2 * max(degs) + 2)      # a single row's closure may exceed the budget
    if max_rows:
        assert ngroups.value >= -(-nf.value // max_rows)
    if types == ["poly", "gaussian"] and max_rows == 0:
        assert ngroups.value == 1 and used.value <= nv + 78          # only (some of) the degree-2 monomials are read by other features


def test_lift_row_groups_random_dictionaries(hostlift):
    """Randomised sweep over dictionaries, slot budgets and row limits: the row-group evaluation (what kf_lift_stream_kernel runs)
    always reproduces the plain evaluation bit for bit and its invariants hold (hostlift_rowgroups returns 0)."""
    rng = np.random.default_rng(2024)
    kinds = ["poly", "hermite", "fourier_sparser", "gaussian", "fourier"]
    for trial in range(40):
        nv = int(rng.integers(1, 7))
        nb = int(rng.integers(1, 4))
        types = [kinds[int(rng.integers(0, len(kinds)))] for _ in range(nb)]
        degs = []
        for t in types:
            if t == "gaussian":
                degs.append(int(rng.integers(1, 40)))
            elif t == "fourier":
                degs.append(int(rng.integers(1, 3)) if nv <= 3 else 1)
            else:
                degs.append(int(rng.integers(1, 5 if nv <= 4 else 4)))
        ng = sum(d for t, d in zip(types, degs) if t == "gaussian")
        cen = 2 * rng.random((nv, ng)) - 1 if ng else None
        b = A.Basis(types, degs, nv, centres=cen)
        nf, N = C.c_int(), C.c_int()
        assert hostlift.hostlift_dims(b.ref(), C.byref(nf), C.byref(N)) == 0
        if nf.value > 6000:
            continue
        rows = 3
        V = np.asfortranarray(2 * rng.random((rows, nv)) - 1)
        plain = np.zeros((rows, nf.value), order="F")
        assert hostlift.hostlift_full(b.ref(), C.c_longlong(rows), A.dptr(V), A.dptr(plain)) == 0
        max_slots = int(rng.integers(nv + 2, nv + 200))
        max_rows = int(rng.choice([0, 0, 7, 64, 500]))
        out = np.full((rows, nf.value), np.nan, order="F")
        ngroups, used = C.c_int(), C.c_int()
        rc = hostlift.hostlift_rowgroups(b.ref(), max_slots, max_rows, C.c_longlong(rows), A.dptr(V), A.dptr(out), C.byref(ngroups), C.byref(used))
        assert rc == 0, (rc, types, degs, nv, max_slots, max_rows)
        assert np.array_equal(out, plain), (types, degs, nv, max_slots, max_rows)


def test_mex_shim_compiles_against_the_abi():
    """matlab/koopfit_mex.cpp cannot run here (no MATLAB / Octave), but it must at least COMPILE against include/koopfit.h
    and a stub of the MEX API (tests/mex_stub/mex.h): catches ABI drift (renamed fields, changed signatures) in the shim."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run(["g++", "-std=c++11", "-fsyntax-only", "-Wall", "-I", os.path.join(root, "tests", "mex_stub"),
                        "-I", os.path.join(root, "include"), os.path.join(root, "matlab", "koopfit_mex.cpp")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    src = open(os.path.join(root, "matlab", "koopfit_mex.cpp")).read()
    for entry in ("kf_fit_multi", "kf_fit_series", "kf_fit_batch", "kf_rollout", "kf_lift", "kf_pca", "kf_mldivide", "kf_mpc_costB_bilinear"):
        assert entry in src, entry
