"""Round-2 probe (CPU only, NumPy): could the Gram contraction leave the FP64 pipe without losing parity?

Ozaki-style error-free slicing: every row of the lifted panel Z (feature x snapshots of one chunk) is scaled by a power of two
and cut into s signed 7-bit slices (INT8 operands), Z = sum_p 2^(-6-7p) sigma S_p.  The slice products S_p S_q' are exact in
INT32 for chunks of up to 2^19 snapshots (|S| <= 64), so G = Z Z' = sum over p + q < s of 2^(-12-7(p+q)) sigma sigma' (S_p S_q'),
i.e. s (s + 1) / 2 INT8 GEMMs with INT32 accumulation (the tcgen05 kind::i8 path on sm_100a) plus an FP64 recombination.
This script measures, on a config-5-like panel, the error of that reconstruction against an extended-precision Gram and sets it
beside the error of the plain FP64 product — the parity bar of the north_star is 1e-9 on K, i.e. ~1e-13 on G here.
Nothing in the product uses this; it only quantifies the round-2 candidate named in DESIGN.md §7b."""
import json, os, sys, time
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle as O

rng = np.random.default_rng(0)
M, nz = 4096, 12
cen = 2 * rng.random((nz, 160)) - 1
prog = O.build_program(["poly", "gaussian"], [2, 160], nz, cen)          # 12 + 78 + 160 + 1 = 251 features
Z = O.lift(prog, 2 * rng.random((M, nz)) - 1).T.copy()                   # features x snapshots
N = Z.shape[0]
truth = (Z.astype(np.longdouble) @ Z.astype(np.longdouble).T)
nrm = float(np.linalg.norm(truth.astype(np.float64)))
G64 = Z @ Z.T
out = {"N": N, "M": M, "fp64_relF": float(np.linalg.norm((G64 - truth).astype(np.float64)) / nrm), "slices": []}
print("FP64 product   relF error vs extended precision: %.2e" % out["fp64_relF"])

sigma = 2.0 ** np.ceil(np.log2(np.abs(Z).max(axis=1)))                   # power-of-two row scales: exact
X = Z / sigma[:, None]
for s in (5, 6, 7, 8, 9):
    t0 = time.time()
    R = X.copy()
    S = []
    for p in range(s):
        T = np.rint(R * 64.0)                                            # |T| <= 64: fits INT8
        S.append(T.astype(np.int64))
        R = (R - T / 64.0) * 128.0                                       # exact in binary floating point
    acc = np.zeros((N, N), dtype=np.longdouble)
    nprod = 0
    for d in range(s):                                                   # p + q = d, all products exact integers (< 2^31)
        Id = np.zeros((N, N), dtype=np.int64)
        for p in range(d + 1):
            q = d - p
            Id += S[p] @ S[q].T
            nprod += 1
        assert np.abs(Id).max() < 2 ** 62
        acc += Id.astype(np.float64).astype(np.longdouble) * np.longdouble(2.0) ** (-12 - 7 * d)
    Goz = (acc * sigma[:, None] * sigma[None, :]).astype(np.float64)     # final rounding to FP64, as the kernel would store it
    err = float(np.linalg.norm((Goz - truth).astype(np.float64)) / nrm)
    worst = float(np.abs((Goz - truth).astype(np.float64) / np.sqrt(np.outer(np.diag(G64), np.diag(G64)))).max())
    out["slices"].append({"s": s, "int8_gemms": nprod, "relF": err, "max_err_over_norms": worst, "seconds": round(time.time() - t0, 1)})
    print("s = %d slices (%2d INT8 GEMMs): relF %.2e, max |err_ij| / (|z_i| |z_j|) %.2e" % (s, nprod, err, worst))
json.dump(out, open(os.path.join(ROOT, "profiles", "r01_ozaki_probe.json"), "w"), indent=1)
