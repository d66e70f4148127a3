"""CPU oracle for the Ksysid EDMD fit (TEST INFRASTRUCTURE ONLY).

This package is a NumPy/SciPy restatement of the hot path of
roahmlab/koopman-realizations (`Ksysid.m`).  It is the checker for the CUDA
path: only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py` may import it.  Nothing under
`koopman-realizations_b200/` imports it, and the product path raises when the
CUDA library is missing instead of falling back to this code.

Parity status:
  * scaling + snapshot set + polynomial dictionary + PCA + econ layout are
    PINNED against the reference's own shipped lifted states
    (`res_lin.Z`, `res_bilin.Z`; tests/golden/arm_blockM_*.npz).
  * G, C, K, A, B, lasso objective, validation RMSE: PARITY UNPINNED — the
    reference ships no model files (.MISSING_LARGE_BLOBS:14-16) and MATLAB /
    Octave are absent here, so those are defined by this restatement
    (LAPACK dgeqp3 = the routine behind MATLAB's `\\`).
"""
from .ksysid_oracle import *  # noqa: F401,F403
