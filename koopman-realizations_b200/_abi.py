"""ctypes mirror of include/koopfit.h and the loader for libkoopfit.so.

There is NO CPU fallback: `load()` raises if the CUDA library has not been built
(`python -c "import __graft_entry__ as g; g.build()"`), and `kf_create` fails
without a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libkoopfit.so")

KF_LINEAR, KF_BILINEAR, KF_NONLINEAR = 0, 1, 2
MODEL_CODE = {"linear": KF_LINEAR, "bilinear": KF_BILINEAR, "nonlinear": KF_NONLINEAR}
KF_POLY, KF_FOURIER, KF_FOURIER_SPARSER, KF_GAUSSIAN, KF_HERMITE = range(5)
OBS_CODE = {"poly": KF_POLY, "fourier": KF_FOURIER, "fourier_sparser": KF_FOURIER_SPARSER,
            "gaussian": KF_GAUSSIAN, "hermite": KF_HERMITE}
KF_LS_AUTO, KF_LS_GRAM, KF_LS_QR = 0, 1, 2
KF_PSD_AS_REFERENCE, KF_PSD_NEVER, KF_PSD_ALWAYS = 0, 1, 2
KF_EAGAIN, KF_ECOMM = 6, 7
ERRORS = {1: "KF_EINVAL", 2: "KF_ECUDA", 3: "KF_ENOMEM", 4: "KF_ENUMERIC", 5: "KF_EUNSUPPORTED", 6: "KF_EAGAIN", 7: "KF_ECOMM"}

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)


class kf_block(C.Structure):
    _fields_ = [("type", C.c_int), ("degree", C.c_int), ("centres", c_double_p)]


class kf_basis(C.Structure):
    _fields_ = [("nv", C.c_int), ("nblocks", C.c_int), ("blocks", C.POINTER(kf_block)),
                ("pcs", c_double_p), ("n_pcs", C.c_int)]


class kf_problem(C.Structure):
    _fields_ = [("M", C.c_longlong), ("nzeta", C.c_int), ("m", C.c_int), ("model", C.c_int),
                ("alpha", C.c_void_p), ("beta", C.c_void_p), ("u", C.c_void_p),
                ("pc_cols", C.c_int), ("nw", C.c_int), ("w", C.c_void_p)]


def regressor_width(model_type, N, m, nw=0):
    """P of the regressor (Ksysid.m:1019-1028): N (nw+1) + m | N (nw+1) (m+1) | N (nw+1)."""
    NL = N * (nw + 1)
    return {"linear": NL + m, "bilinear": NL * (m + 1), "nonlinear": NL}[model_type]


class kf_solve(C.Structure):
    _fields_ = [("least_squares", C.c_int), ("ls_method", C.c_int), ("pivot_tol", C.c_double),
                ("nt", C.c_int), ("t", c_double_p), ("psd_shift", C.c_int),
                ("delay_constraint", C.c_int), ("n", C.c_int), ("nd", C.c_int),
                ("qp_max_iter", C.c_int), ("qp_tol", C.c_double)]


class kf_info(C.Structure):
    _fields_ = [("rank", C.c_int), ("ls_method_used", C.c_int), ("passes", C.c_int),
                ("psd_shift_applied", C.c_int), ("min_pivot", C.c_double), ("max_pivot", C.c_double),
                ("t_lift_gram_ms", C.c_double), ("t_solve_ms", C.c_double), ("t_total_ms", C.c_double),
                ("qp_capped", C.c_int), ("reserved", C.c_int), ("cond_est", C.c_double),
                ("refine_passes", C.c_int), ("refine_capped", C.c_int)]


class kf_result(C.Structure):
    _fields_ = [("K", c_double_p), ("G", c_double_p), ("C", c_double_p), ("Px", c_double_p),
                ("Py", c_double_p), ("perm", c_int_p), ("objective", c_double_p),
                ("l1norm", c_double_p), ("qp_iters", c_int_p), ("qp_gap", c_double_p), ("info", kf_info)]


class kf_series(C.Structure):
    _fields_ = [("T", C.c_longlong), ("n", C.c_int), ("m", C.c_int), ("nd", C.c_int), ("model", C.c_int),
                ("t", c_double_p), ("y", c_double_p), ("u", c_double_p), ("prescaled", C.c_int), ("pc_cols", C.c_int)]


class kf_scale(C.Structure):
    _fields_ = [("y_offset", c_double_p), ("y_factor", c_double_p), ("u_offset", c_double_p), ("u_factor", c_double_p),
                ("M", C.c_longlong)]


class kf_model(C.Structure):
    _fields_ = [("model", C.c_int), ("n", C.c_int), ("m", C.c_int), ("nzeta", C.c_int), ("N", C.c_int),
                ("A", c_double_p), ("B", c_double_p), ("F", c_double_p)]


ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, c_double_p, C.c_int, C.c_int)


class KoopfitError(RuntimeError):
    pass


def dptr(a):
    """double* of a float64 numpy array (must stay alive while used)."""
    return a.ctypes.data_as(c_double_p)


def fcol(a):
    """float64, column-major (MATLAB layout) copy/view of a 2-D array."""
    return np.asfortranarray(np.asarray(a, dtype=np.float64))


class Basis:
    """Owns the ctypes kf_basis and the numpy buffers it points into.

    obs_type / obs_degree as in the Ksysid constructor (Ksysid.m:20-21, 465);
    centres: (nv, total gaussian degree) — the reference's zeta0 (Ksysid.m:803);
    pcs: basis.pcs after dim_red (Ksysid.m:1507-1510).
    """

    def __init__(self, obs_type, obs_degree, nv, centres=None, pcs=None):
        if isinstance(obs_type, str):
            obs_type = [obs_type]
        obs_degree = [int(d) for d in np.atleast_1d(obs_degree)]
        if len(obs_type) != len(obs_degree):
            raise ValueError("inputs must be of the same size")           # Ksysid.m:465-467
        self.obs_type, self.obs_degree, self.nv = list(obs_type), obs_degree, int(nv)
        known = [(t, d) for t, d in zip(obs_type, obs_degree) if t in OBS_CODE]   # unknown types ignored (486-501)
        self._blocks = (kf_block * max(1, len(known)))()
        self._keep = []
        used = 0
        for i, (t, d) in enumerate(known):
            self._blocks[i].type = OBS_CODE[t]
            self._blocks[i].degree = d
            if t == "gaussian":
                if centres is None:
                    raise ValueError("gaussian block needs centres of shape (nv, degree)")
                c = np.ascontiguousarray(np.asarray(centres, dtype=np.float64)[:, used:used + d].T)  # (d, nv) C == (nv, d) F
                if c.shape != (d, nv):
                    raise ValueError("gaussian centres must have shape (nv, >= total gaussian degree)")
                used += d
                self._keep.append(c)
                self._blocks[i].centres = dptr(c)
        self.centres = None if centres is None else np.asarray(centres, dtype=np.float64)
        self.struct = kf_basis()
        self.struct.nv = nv
        self.struct.nblocks = len(known)
        self.struct.blocks = self._blocks
        self.pcs = None
        if pcs is not None:
            self.set_pcs(pcs)

    def set_pcs(self, pcs):
        self.pcs = fcol(pcs)
        self.struct.pcs = dptr(self.pcs)
        self.struct.n_pcs = self.pcs.shape[1]

    def ref(self):
        return C.byref(self.struct)


_lib = None


def load():
    """Load libkoopfit.so (raises if it is not built — no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise KoopfitError(f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'`. "
                           "There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    vp, i, ll, d = C.c_void_p, C.c_int, C.c_longlong, C.c_double
    P = C.POINTER
    sigs = {
        "kf_create": (i, [P(vp), i]),
        "kf_destroy": (None, [vp]),
        "kf_last_error": (C.c_char_p, [vp]),
        "kf_version": (i, []),
        "kf_basis_dims": (i, [P(kf_basis), i, i, c_int_p, c_int_p, c_int_p]),
        "kf_block_table": (i, [i, i, i, c_int_p, c_int_p, c_int_p]),
        "kf_lift": (i, [vp, P(kf_basis), ll, c_double_p, c_double_p]),
        "kf_pca": (i, [vp, P(kf_basis), ll, c_double_p, c_double_p, c_double_p, c_double_p]),
        "kf_fit": (i, [vp, P(kf_basis), P(kf_problem), P(kf_solve), P(kf_result)]),
        "kf_fit_dev": (i, [vp, P(kf_basis), P(kf_problem), P(kf_solve), P(kf_result)]),
        "kf_fit_series": (i, [vp, P(kf_basis), P(kf_series), P(kf_solve), P(kf_scale), P(kf_result)]),
        "kf_series_pairs": (ll, [ll, i, c_double_p]),
        "kf_fit_batch": (i, [vp, i, P(P(kf_basis)), P(kf_problem), P(kf_solve), P(kf_result)]),
        "kf_rollout": (i, [vp, P(kf_basis), i, P(kf_model), i, c_int_p, P(c_double_p), P(c_double_p), i, P(c_double_p)]),
        "kf_mpc_costB_bilinear": (i, [vp, i, i, i, i, c_double_p, c_double_p, i, c_double_p, c_double_p]),
        "kf_set_qp_partition": (i, [vp, i, i, ALLREDUCE_FN, vp]),
        "kf_mldivide": (i, [vp, ll, i, i, c_double_p, c_double_p, c_double_p, c_int_p, c_int_p]),
        "kf_accumulate_dev": (i, [vp, P(kf_basis), P(kf_problem), i]),
        "kf_regressors_dev": (i, [vp, P(kf_basis), P(kf_problem), vp, ll]),
        "kf_lift_dev": (i, [vp, P(kf_basis), ll, vp, vp]),
        "kf_accum_buffer": (i, [vp, P(vp), P(C.c_size_t)]),
        "kf_solve_dev": (i, [vp, P(kf_solve), P(kf_result)]),
        "kf_sync": (i, [vp]),
        "kf_stream": (vp, [vp]),
        "kf_counters": (i, [vp, c_double_p, P(ll), i]),
        "kf_engine_info": (i, [vp, c_int_p, c_double_p, c_double_p, P(ll), c_int_p, i]),
        "kf_last_times": (i, [vp, c_double_p, c_double_p, c_double_p]),
        "kf_set_option": (i, [vp, C.c_char_p, d]),
        "kf_create_multi": (i, [P(vp), c_int_p, i]),
        "kf_destroy_multi": (None, [vp]),
        "kf_multi_size": (i, [vp]),
        "kf_multi_ctx": (vp, [vp, i]),
        "kf_multi_last_error": (C.c_char_p, [vp]),
        "kf_multi_set_option": (i, [vp, C.c_char_p, d]),
        "kf_fit_multi": (i, [vp, P(kf_basis), P(kf_problem), P(kf_solve), P(kf_result)]),
        "kf_comm_unique_id": (i, [vp, C.c_size_t]),
        "kf_comm_init_rank": (i, [vp, i, i, vp, C.c_size_t]),
        "kf_comm_destroy": (i, [vp]),
        "kf_comm_info": (i, [vp, c_int_p, c_int_p, c_int_p]),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(lib, name)      # AttributeError if the header and the library disagree
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


EXPORTS = ["kf_create", "kf_destroy", "kf_last_error", "kf_version", "kf_basis_dims", "kf_block_table",
           "kf_lift", "kf_pca", "kf_fit", "kf_fit_dev", "kf_fit_series", "kf_series_pairs", "kf_fit_batch", "kf_rollout", "kf_mpc_costB_bilinear", "kf_set_qp_partition", "kf_mldivide", "kf_accumulate_dev", "kf_regressors_dev", "kf_lift_dev", "kf_accum_buffer", "kf_solve_dev", "kf_sync",
           "kf_stream", "kf_counters", "kf_engine_info", "kf_last_times", "kf_set_option",
           "kf_create_multi", "kf_destroy_multi", "kf_multi_size", "kf_multi_ctx", "kf_multi_last_error", "kf_multi_set_option", "kf_fit_multi",
           "kf_comm_unique_id", "kf_comm_init_rank", "kf_comm_destroy", "kf_comm_info"]
