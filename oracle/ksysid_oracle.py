"""NumPy/SciPy restatement of the Ksysid EDMD fit — TEST INFRASTRUCTURE ONLY.

Every function cites the lines of /root/reference/Ksysid.m (or partitions.m)
it follows.  See oracle/__init__.py for who may import this and for the parity
status (what is pinned by the reference's own golden vectors and what is not).

Conventions
-----------
* All matrices are float64, row = snapshot (as in the reference).
* A *feature program* describes the lifted vector psi(v) of length N over the
  variable vector v (v = zeta for linear/bilinear, [zeta; u] for nonlinear):
  psi = [v ; block_1 ; ... ; 1]   (Ksysid.m:484-505).
  Each feature is one op:  VAR i | CONST | MUL a b | COS i c | SIN i c |
  HERM i k | GAUSS k.  MUL multiplies two *earlier* features, which fixes the
  floating-point association:
     feature(row) = feature(row with its LAST non-zero entry zeroed)
                    * primitive(LAST non-zero entry)
  and a pure power v_i^k = v_i^(k-1) * v_i.  This mirrors get_monomial's
  left-to-right product x(1)^e(1) * x(2)^e(2) * ... (Ksysid.m:687-690) with
  integer powers as repeated multiplication.  The CUDA lift kernel follows the
  same association, so polynomial / Hermite features are bit-identical and
  sin/cos/exp features differ only by the libm-vs-CUDA ulp.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np
import scipy.linalg as sla

__all__ = [
    "basic_solution_extended", "continuous_UT",
    "load_data4sysid", "load_rand_systems", "merge_trials", "get_scale", "scale_data",
    "get_zeta", "get_snapshot_pairs", "partitions_ones", "FeatureProgram", "build_program",
    "lift", "lift_full", "build_regressors", "regressor_width", "mldivide", "gram", "qp_objective",
    "solve_l1ball_qp", "l1ball_project", "get_koopman", "get_model", "get_BLmodel",
    "get_NLmodel", "val_model", "val_BLmodel", "val_NLmodel", "get_error", "pca_matlab",
    "econ_reduce", "KsysidOracle", "OP_VAR", "OP_CONST", "OP_MUL", "OP_COS", "OP_SIN",
    "OP_HERM", "OP_GAUSS", "needs_psd_shift", "delay_constraint_targets", "load_lift", "mpc_costB_bilinear",
]

OP_VAR, OP_CONST, OP_MUL, OP_COS, OP_SIN, OP_HERM, OP_GAUSS = range(7)
TWO_PI = 2.0 * math.pi


# --------------------------------------------------------------------------- data
def _as2d(x):
    x = np.asarray(x, dtype=np.float64)
    return x.reshape(-1, 1) if x.ndim == 1 else x


def _trial(s):
    """One trial struct -> dict(t (T,), y (T,n), u (T,m)); Ksysid.m:46-59."""
    return {"t": np.asarray(s.t, dtype=np.float64).reshape(-1), "y": _as2d(s.y), "u": _as2d(s.u)}


def _trial_list(x):
    if isinstance(x, np.ndarray):
        return [_trial(s) for s in x.reshape(-1)]
    return [_trial(x)]


def load_data4sysid(path):
    """Load a `data4sysid` .mat (struct with cell arrays train/val; Ksysid.m:46-53)."""
    import scipy.io as sio

    d = sio.loadmat(path, squeeze_me=True, struct_as_record=False)
    if "train" in d:
        return {"train": _trial_list(d["train"]), "val": _trial_list(d["val"])}
    if "data4sysid" in d:
        s = d["data4sysid"]
        return {"train": _trial_list(s.train), "val": _trial_list(s.val)}
    raise ValueError(f"{path}: no train/val fields")


def load_rand_systems(path):
    """Load an `rsys-all_*.mat`: cell `data4sysid_all{i}` (Rsys.m:195-215)."""
    import scipy.io as sio

    d = sio.loadmat(path, squeeze_me=True, struct_as_record=False)
    allsys = d["data4sysid_all"]
    allsys = allsys.reshape(-1) if isinstance(allsys, np.ndarray) else [allsys]
    return [{"train": _trial_list(s.train), "val": _trial_list(s.val)} for s in allsys]


def merge_trials(trials):
    """Vertical concatenation of every numeric field (Ksysid.m:380-401); the load w rides along when every trial has one."""
    keys = ("t", "y", "u") + (("w",) if all("w" in tr for tr in trials) else ())
    return {k: np.concatenate([np.asarray(tr[k], dtype=np.float64).reshape(len(tr["t"]), -1) if k != "t" else tr[k]
                               for tr in trials], axis=0) for k in keys}


def get_scale(data):
    """Scale merged train data into [-1,1] per column (Ksysid.m:187-210)."""
    sc = {}
    out = {"t": data["t"]}
    for k in ("y", "u"):
        mn, mx = data[k].min(axis=0), data[k].max(axis=0)
        dc = (mx + mn) / 2.0
        s = (mx - mn) / 2.0
        s = np.where(s == 0, 1.0, s)
        sc[k + "_offset"], sc[k + "_factor"] = dc, s
        out[k] = (data[k] - dc) / s
    if "w" in data:      # Ksysid.m:246-265: a constant load is only shifted to zero, a varying one scaled into [-1, 1]
        mn, mx = data["w"].min(axis=0), data["w"].max(axis=0)
        sc["w_offset"] = (mx + mn) / 2.0
        sc["w_factor"] = np.where(mn != mx, (mx - mn) / 2.0, 1.0)
        out["w"] = (data["w"] - sc["w_offset"]) / sc["w_factor"]
    return out, sc


def scale_data(trial, sc, down=True):
    """Scale a trial with the train factors (Ksysid.m:308-343)."""
    if down:
        out = {"t": trial["t"], "y": (trial["y"] - sc["y_offset"]) / sc["y_factor"],
               "u": (trial["u"] - sc["u_offset"]) / sc["u_factor"]}
        if "w" in trial and "w_offset" in sc:            # Ksysid.m:328-330
            out["w"] = (np.asarray(trial["w"], dtype=np.float64).reshape(len(trial["t"]), -1) - sc["w_offset"]) / sc["w_factor"]
        return out
    return {"t": trial["t"], "y": trial["y"] * sc["y_factor"] + sc["y_offset"],
            "u": trial["u"] * sc["u_factor"] + sc["u_offset"]}


def get_zeta(data, nd):
    """zeta_i = [y_i, y_{i-1}..y_{i-nd}, u_{i-1}..u_{i-nd}], uzeta_i = u_i, i>=nd (Ksysid.m:868-907)."""
    y, u = data["y"], data["u"]
    T = y.shape[0]
    if nd == 0:
        return y.copy(), u.copy()
    cols = [y[nd:T]]
    cols += [y[nd - j:T - j] for j in range(1, nd + 1)]
    cols += [u[nd - j:T - j] for j in range(1, nd + 1)]
    return np.concatenate(cols, axis=1), u[nd:T].copy()


def get_snapshot_pairs(data, nd, snapshots=np.inf):
    """Snapshot pairs from (merged, scaled) time series (Ksysid.m:941-982).

    Pairs straddling a trial boundary (before.t >= after.t) are dropped (948);
    num_max = kept-1 (960), so the LAST kept pair is never used; with
    snapshots=Inf the reference takes a random permutation of the first num_max
    kept pairs (974-975) — order does not affect G, C or K, so natural order is
    returned.  A finite `snapshots` would need MATLAB's mlfg6331_64 stream.
    """
    zeta, uzeta = get_zeta(data, nd)
    t = data["t"]
    tb, ta = t[nd:-1], t[nd + 1:]
    good = np.nonzero(tb < ta)[0]
    num_max = good.size - 1
    if np.isfinite(snapshots) and snapshots <= num_max - 1:
        raise NotImplementedError("snapshots < all needs MATLAB RandStream('mlfg6331_64') (Ksysid.m:974)")
    idx = good[:num_max]
    pairs = {"alpha": zeta[:-1][idx], "beta": zeta[1:][idx], "u": uzeta[:-1][idx]}
    if "w" in data:                                      # wzeta(1:end-1) at the kept points (Ksysid.m:953-957, 980-982)
        pairs["w"] = data["w"][nd:][:-1][idx]
    return pairs


# --------------------------------------------------------------------------- dictionaries
def partitions_ones(total, nvars):
    """partitions(total, ones(1,nvars)) row order (partitions.m:206-219): the LAST
    variable's exponent ascends in the outermost loop, recursively."""
    if total == 0:
        return np.zeros((1, nvars), dtype=np.int64)
    if nvars == 1:
        return np.array([[total]], dtype=np.int64)
    rows = []
    for i in range(total + 1):
        sub = partitions_ones(total - i, nvars - 1)
        rows.append(np.concatenate([sub, np.full((sub.shape[0], 1), i, dtype=np.int64)], axis=1))
    return np.concatenate(rows, axis=0)


@dataclass
class FeatureProgram:
    nv: int                      # number of variables lifted (nzeta, or nzeta+m for nonlinear)
    ops: list = field(default_factory=list)     # (kind, a, b, c) per feature
    centres: np.ndarray | None = None           # (nv, ngauss) gaussian centres, all blocks concatenated
    pcs: np.ndarray | None = None               # (N_full, k) PCA coefficients when dim_red
    blocks: list = field(default_factory=list)  # [(type, degree, first, count)]

    @property
    def n_full(self):
        return len(self.ops)

    @property
    def N(self):
        """Lifted dimension seen by the fit (Ksysid.m:534, 1511-1517)."""
        return self.n_full if self.pcs is None else self.nv + self.pcs.shape[1] + 1


def _chain_block(prog, rows, prim, index=None, skip=0):
    """Append one feature per row (rows[skip:]) using the 'last non-zero entry' product rule.

    rows: int array (R, L).  prim(pos, val, index) -> op tuple of the primitive for a row
    whose only non-zero entry is (pos, val).  index: dict row-tuple -> feature index,
    pre-seeded with rows that already exist as features.
    """
    index = {} if index is None else index
    for r in rows[skip:]:
        key = tuple(int(x) for x in r)
        nz = [i for i, x in enumerate(key) if x != 0]
        last = nz[-1]
        if len(nz) == 1:
            op = prim(last, key[last], index)
        else:
            head = list(key)
            head[last] = 0
            tail = [0] * len(key)
            tail[last] = key[last]
            op = (OP_MUL, index[tuple(head)], index[tuple(tail)], 0.0)
        index[key] = len(prog.ops)
        prog.ops.append(op)
    return index


def build_program(obs_type, obs_degree, nv, centres=None):
    """Feature program for psi = [v; blocks...; 1] (Ksysid.m:484-505).

    obs_type: list of 'poly'|'fourier'|'fourier_sparser'|'gaussian'|'hermite';
    obs_degree: same-length ints (465).  centres: (nv, sum of gaussian degrees)
    array of gaussian centres — the reference draws them from MATLAB's global
    rand (803), so they are an INPUT here.
    """
    if isinstance(obs_type, str):
        obs_type = [obs_type]
    obs_degree = [int(d) for d in np.atleast_1d(obs_degree)]
    if len(obs_type) != len(obs_degree):
        raise ValueError("inputs must be of the same size")
    prog = FeatureProgram(nv=nv)
    for i in range(nv):
        prog.ops.append((OP_VAR, i, 0, 0.0))
    gauss_used = 0
    for typ, deg in zip(obs_type, obs_degree):
        first = len(prog.ops)
        if typ == "poly":
            # def_polyLift 645-648 rows; the first nv (degree-1) are v itself and are dropped (488)
            rows = np.concatenate([partitions_ones(k, nv) for k in range(1, deg + 1)], axis=0)
            index = {tuple(1 if k == i else 0 for k in range(nv)): i for i in range(nv)}

            def prim(pos, val, idx):
                lower = [0] * nv
                lower[pos] = val - 1
                return (OP_MUL, idx[tuple(lower)], pos, 0.0)        # v^k = v^(k-1) * v

            _chain_block(prog, rows, prim, index=index, skip=nv if deg >= 1 else 0)
        elif typ == "hermite":
            # def_hermiteLift 848-856: same exponent rows, nothing dropped, product of hermiteH
            rows = np.concatenate([partitions_ones(k, nv) for k in range(1, deg + 1)], axis=0)
            _chain_block(prog, rows, lambda pos, val, idx: (OP_HERM, pos, val, 0.0))
        elif typ == "fourier":
            # def_fourierLift 708-724: kron over variables of [1, cos(2 pi j v), sin(2 pi j v)...],
            # last variable fastest, constant (index 0) removed.
            base = 1 + 2 * deg
            total = base ** nv
            rows = np.zeros((total - 1, nv), dtype=np.int64)
            idx = np.arange(1, total)
            for i in range(nv - 1, -1, -1):
                rows[:, i] = idx % base
                idx //= base

            def prim(pos, val, idx_):
                j = (val + 1) // 2
                return ((OP_COS if val % 2 == 1 else OP_SIN), pos, 0, TWO_PI * j)

            _chain_block(prog, rows, prim)
        elif typ == "fourier_sparser":
            # def_fourierLift_sparser 747-787: multiplier rows over 2*nv slots (sin slots, then cos slots)
            rows = np.concatenate([partitions_ones(k, 2 * nv) for k in range(1, deg + 1)], axis=0)

            def prim(pos, val, idx_):
                if pos < nv:
                    return (OP_SIN, pos, 0, TWO_PI * val)
                return (OP_COS, pos - nv, 0, TWO_PI * val)

            _chain_block(prog, rows, prim)
        elif typ == "gaussian":
            # def_gaussianLift 803-806: exp(-||v - c_k||^2)
            if centres is None or centres.shape[0] != nv or centres.shape[1] < gauss_used + deg:
                raise ValueError("gaussian block needs centres of shape (nv, >=degree)")
            for k in range(deg):
                prog.ops.append((OP_GAUSS, gauss_used + k, 0, 0.0))
            gauss_used += deg
        else:
            # the reference silently ignores unknown types (486-501)
            pass
        prog.blocks.append((typ, deg, first, len(prog.ops) - first))
    prog.ops.append((OP_CONST, 0, 0, 1.0))
    if gauss_used:
        prog.centres = np.ascontiguousarray(centres[:, :gauss_used], dtype=np.float64)
    return prog


def _hermite(k, x):
    """Physicists' H_k by the recurrence H_{j+1} = (2x) H_j - (2j) H_{j-1}, every
    product and difference rounded separately (no FMA)."""
    h0 = np.ones_like(x)
    if k == 0:
        return h0
    tx = 2.0 * x
    h1 = tx
    for j in range(1, k):
        h0, h1 = h1, tx * h1 - (2.0 * j) * h0
    return h1


def lift_full(prog, V):
    """Evaluate the full dictionary on rows of V (M, nv) -> (M, N_full)."""
    V = np.asarray(V, dtype=np.float64)
    if V.ndim == 1:
        V = V[None, :]
    M = V.shape[0]
    F = np.empty((M, prog.n_full), dtype=np.float64)
    for j, (kind, a, b, c) in enumerate(prog.ops):
        if kind == OP_VAR:
            F[:, j] = V[:, a]
        elif kind == OP_CONST:
            F[:, j] = c
        elif kind == OP_MUL:
            F[:, j] = F[:, a] * F[:, b]
        elif kind == OP_COS:
            F[:, j] = np.cos(c * V[:, a])
        elif kind == OP_SIN:
            F[:, j] = np.sin(c * V[:, a])
        elif kind == OP_HERM:
            F[:, j] = _hermite(b, V[:, a])
        elif kind == OP_GAUSS:
            acc = np.zeros(M)
            for i in range(prog.nv):
                d = V[:, i] - prog.centres[i, a]
                acc = acc + d * d
            F[:, j] = np.exp(-acc)
        else:
            raise ValueError(kind)
    return F


def lift(prog, V):
    """lift.econ_full: the full dictionary, or [v; pcs' psi_full(v); 1] after
    dimension reduction (Ksysid.m:1443-1491 vs 1614-1618)."""
    F = lift_full(prog, V)
    if prog.pcs is None:
        return F
    V = np.asarray(V, dtype=np.float64)
    if V.ndim == 1:
        V = V[None, :]
    return np.concatenate([V, F @ prog.pcs, np.ones((F.shape[0], 1))], axis=1)


def regressor_width(model_type, N, m, nw=0):
    """P: N+m linear, N(m+1) bilinear, N nonlinear, with N -> N (nw+1) for a loaded model (Ksysid.m:1019-1028)."""
    NL = N * (nw + 1)
    return {"linear": NL + m, "bilinear": NL * (m + 1), "nonlinear": NL}[model_type]


def load_lift(psi, w):
    """lift.full_loaded: [psi; w_1 psi; ...; w_nw psi] = [1; w] (x) psi per row (Ksysid.m:594-599, 1607-1611)."""
    if w is None or np.size(w) == 0:
        return psi
    w = np.asarray(w, dtype=np.float64).reshape(psi.shape[0], -1)
    return np.concatenate([psi] + [w[:, c:c + 1] * psi for c in range(w.shape[1])], axis=1)


def build_regressors(model_type, prog, alpha, beta, u, w=None):
    """Px, Py of get_Koopman's lift loop (Ksysid.m:1030-1065); same u (and, for a loaded model, same w) on both sides."""
    if model_type == "nonlinear":
        return (load_lift(lift(prog, np.concatenate([alpha, u], axis=1)), w),
                load_lift(lift(prog, np.concatenate([beta, u], axis=1)), w))
    psx, psy = load_lift(lift(prog, alpha), w), load_lift(lift(prog, beta), w)
    if model_type == "linear":
        return np.concatenate([psx, u], axis=1), np.concatenate([psy, u], axis=1)
    if model_type == "bilinear":   # [psi; kron(I_m, psi) u] (510-511)
        m = u.shape[1]
        return (np.concatenate([psx] + [u[:, k:k + 1] * psx for k in range(m)], axis=1),
                np.concatenate([psy] + [u[:, k:k + 1] * psy for k in range(m)], axis=1))
    raise ValueError("Invalid model_type chosen. Must be linear, bilinear, or nonlinear.")


# --------------------------------------------------------------------------- solve
def mldivide(A, B, return_info=False):
    """MATLAB `A \\ B` for a tall/rectangular A (Ksysid.m:1069, 1216): Householder QR
    with column pivoting (LAPACK dgeqp3), numerical rank r from |diag R| against
    tol = max(size(A)) * eps(|R11|), BASIC solution (zeros on the non-pivot columns)."""
    A = np.asarray(A, dtype=np.float64)
    B = np.asarray(B, dtype=np.float64)
    Q, R, perm = sla.qr(A, mode="economic", pivoting=True)
    d = np.abs(np.diag(R))
    tol = max(A.shape) * np.spacing(d[0]) if d.size else 0.0
    r = int(np.sum(d > tol))
    X = np.zeros((A.shape[1], B.shape[1]))
    if r:
        X[perm[:r]] = sla.solve_triangular(R[:r, :r], Q[:, :r].T @ B)
    if return_info:
        return X, {"rank": r, "perm": perm, "diagR": d, "tol": tol}
    return X


def basic_solution_extended(A, B, basic):
    """Ground truth for the rank-deficient `A \\ B`: the least-squares solution on a GIVEN basic column set,
    by Householder QR in x87 extended precision (np.longdouble, 64-bit mantissa), zeros elsewhere.  With cond(A_basic)
    up to ~1e7 the float64 solvers carry errors of cond * 1e-16; this one is ~2000x closer to the exact answer, so it
    decides which of two float64 answers (oracle dgeqp3, GPU) is nearer the truth.  Pure NumPy, small cases only."""
    basic = np.asarray(basic, dtype=int)
    R = np.asarray(A, dtype=np.float64)[:, basic].astype(np.longdouble)
    Y = np.asarray(B, dtype=np.float64).astype(np.longdouble)
    M, r = R.shape
    for j in range(r):
        x = R[j:, j].copy()
        nrm = np.sqrt(np.sum(x * x))
        if nrm == 0:
            continue
        alpha = -nrm if x[0] >= 0 else nrm
        x[0] -= alpha
        v = x / np.sqrt(np.sum(x * x))
        R[j:, j:] -= 2 * np.outer(v, v @ R[j:, j:])
        Y[j:] -= 2 * np.outer(v, v @ Y[j:])
    X = np.zeros((r, Y.shape[1]), dtype=np.longdouble)
    for i in range(r - 1, -1, -1):
        X[i] = (Y[i] - R[i, i + 1:r] @ X[i + 1:]) / R[i, i]
    out = np.zeros((A.shape[1], Y.shape[1]), dtype=np.longdouble)
    out[basic] = X
    return out


def gram(Px, Py):
    """G = Px'Px (Ksysid.m:1114), C = Px'Py (1125)."""
    return Px.T @ Px, Px.T @ Py


def qp_objective(G, C, K):
    """0.5 tr(K'GK) - tr(C'K): the QP cost 0.5 x'Hx + f'x of Ksysid.m:1126-1132 at x=[K+;K-]."""
    return 0.5 * float(np.sum(K * (G @ K))) - float(np.sum(C * K))


def needs_psd_shift(G):
    """`if any(eig(G) < 0)` branch (Ksysid.m:1117-1120): +1e-6 I."""
    return bool(np.any(np.linalg.eigvalsh(G) < 0))


def l1ball_project(X, t):
    """Euclidean projection of X onto {||vec X||_1 <= t} (exact, sort-based)."""
    a = np.abs(X).ravel()
    if a.sum() <= t:
        return X.copy()
    s = np.sort(a)[::-1]
    cs = np.cumsum(s)
    k = np.nonzero(s * np.arange(1, s.size + 1) > (cs - t))[0][-1]
    theta = (cs[k] - t) / (k + 1.0)
    return np.sign(X) * np.maximum(np.abs(X) - theta, 0.0)


def delay_constraint_targets(N, n, m, nd):
    """Equality constraints Ksysid.m:1139-1164 (linear model, nd>=1), replicated AS WRITTEN.

    Constrains vec(K) entries n*Nm .. Nm*(n*(nd+1)+m*nd)-1 (0-based), i.e. columns
    n .. nzeta-1 of K, to a 0/1 pattern.  Returns (col0, col1, target) with target of
    shape (Nm, col1-col0): K[:, col0:col1] must equal target.
    """
    Nm = N + m
    nnd, mnd = n * nd, m * nd
    bd = np.zeros(Nm * (nnd + mnd))
    for i in range(1, nnd + 1):
        bd[(Nm + 1) * (i - 1)] = 1.0
    for i in range(1, m + 1):
        bd[Nm * nnd + N + (Nm + 1) * (i - 1)] = 1.0
    for i in range(1, m * (nd - 1) + 1):
        bd[Nm * (nnd + m) + nnd + (Nm + 1) * (i - 1)] = 1.0
    return n, n * (nd + 1) + mnd, bd.reshape(nnd + mnd, Nm).T.copy()


def _l1ball_qp_homotopy(G, Cf, t_free):
    """Exact solution of  min 0.5 tr(K'GK) - tr(Cf'K)  s.t. ||K||_1 = t_free  by per-column lasso paths."""
    from sklearn.linear_model import lars_path_gram

    P, Pc = Cf.shape
    paths = []
    knots = [0.0]
    for j in range(Pc):
        if not np.any(Cf[:, j]):
            paths.append(None)
            continue
        alphas, _, coefs = lars_path_gram(Xy=Cf[:, j].copy(), Gram=G.copy(), n_samples=1, method="lasso",
                                          alpha_min=0.0, max_iter=20 * P, return_path=True)
        paths.append((alphas[::-1].copy(), coefs[:, ::-1].copy()))       # increasing lam for np.interp
        knots.extend(alphas.tolist())
    knots = np.unique(np.asarray(knots))                                  # ascending

    def K_at(lam):
        K = np.zeros((P, Pc))
        for j, pth in enumerate(paths):
            if pth is None:
                continue
            al, co = pth
            if lam >= al[-1]:
                continue
            i = np.searchsorted(al, lam, side="right")                   # al[i-1] <= lam < al[i]
            if i == 0:
                K[:, j] = co[:, 0]
            else:
                w = (lam - al[i - 1]) / (al[i] - al[i - 1])
                K[:, j] = (1 - w) * co[:, i - 1] + w * co[:, i]
        return K

    l1 = np.array([np.abs(K_at(x)).sum() for x in knots])                 # non-increasing in lam
    if l1[0] < t_free:
        raise RuntimeError("budget inactive")
    i = np.nonzero(l1 >= t_free)[0][-1]                                   # l1[i] >= t > l1[i+1]
    if i + 1 >= len(knots):
        return K_at(knots[i]), float(knots[i])
    lo, hi = knots[i], knots[i + 1]
    lam = lo + (l1[i] - t_free) / (l1[i] - l1[i + 1]) * (hi - lo)        # linear on the segment
    return K_at(lam), float(lam)


def solve_l1ball_qp(G, C, t, fixed=None, tol=1e-13, max_outer=200, verbose=False, use_lars=None):
    """Exact solver for   min 0.5 tr(K'GK) - tr(C'K)  s.t. ||vec K||_1 <= t
    (the QP of Ksysid.m:1095-1176 in un-split variables; `quadprog` stand-in).

    KKT: there is lam >= 0 with every column k_j solving the penalised problem
    min 0.5 k'Gk - c_j'k + lam ||k||_1 and (lam = 0, ||K||_1 <= t) or ||K||_1 = t.
    For fixed lam all columns are solved together by cyclic coordinate descent
    (vectorised across columns) followed by an exact active-set polish
    (k_S = G_SS^{-1}(c_S - lam sign_S)); lam is found by bracketing + Illinois
    secant on phi(lam) = ||K(lam)||_1 - t, then a final exact step that solves for lam
    on the fixed sign pattern.

    fixed = (col0, col1, target): columns col0:col1 of K are pinned to `target`
    (delay_constraint_targets) — they still count towards the L1 budget as in the
    reference, where the budget row covers all 2 Nm^2 split variables (1136).
    """
    G = np.asarray(G, dtype=np.float64)
    C = np.asarray(C, dtype=np.float64)
    P, Pc = C.shape
    free = np.ones(Pc, dtype=bool)
    Kfix = None
    if fixed is not None:
        c0, c1, target = fixed
        free[c0:c1] = False
        Kfix = target
        t_free = t - np.abs(target).sum()
        if t_free < 0:
            raise ValueError("L1 budget smaller than the pinned delay entries: QP infeasible")
    else:
        t_free = t
    Cf = C[:, free]
    dG = np.diag(G).copy()

    def assemble(Kf):
        K = np.zeros((P, Pc))
        K[:, free] = Kf
        if Kfix is not None:
            K[:, ~free] = Kfix
        return K

    # unconstrained minimiser (min-norm if singular) — feasible => done (lam = 0)
    try:
        cf = sla.cho_factor(G)
        K0 = sla.cho_solve(cf, Cf)
    except sla.LinAlgError:
        K0 = np.linalg.lstsq(G, Cf, rcond=None)[0]
    if np.abs(K0).sum() <= t_free:
        return assemble(K0), {"lam": 0.0, "active": False}

    # Exact homotopy for small problems: each column's lasso path k_j(lam) is piecewise linear in lam
    # (LARS with the lasso modification on the Gram), so ||K(lam)||_1 is piecewise linear too and the
    # multiplier with ||K||_1 = t is found exactly on the union of the knots.
    if use_lars is None:
        use_lars = P <= 160
    if use_lars:
        try:
            Kl, lam_l = _l1ball_qp_homotopy(G, Cf, t_free)
            grad = G @ Kl - Cf
            viol = np.max(np.where(Kl == 0, np.abs(grad) - lam_l, 0.0))
            act = np.abs(grad[Kl != 0] + lam_l * np.sign(Kl[Kl != 0])).max() if np.any(Kl != 0) else 0.0
            if viol <= 1e-8 * max(lam_l, 1.0) and act <= 1e-8 * max(lam_l, 1.0) and abs(np.abs(Kl).sum() - t_free) <= 1e-10 * t_free:
                return assemble(Kl), {"lam": lam_l, "active": True, "method": "homotopy"}
        except Exception:
            pass

    def polish(K, lam):
        """Exact solve on the current support/sign pattern, column by column."""
        out = np.zeros_like(K)
        ok = True
        for j in range(K.shape[1]):
            S = np.nonzero(K[:, j])[0]
            if S.size == 0:
                continue
            sg = np.sign(K[S, j])
            try:
                kS = np.linalg.solve(G[np.ix_(S, S)], Cf[S, j] - lam * sg)
            except np.linalg.LinAlgError:
                kS = K[S, j]; ok = False
            if np.any(np.sign(kS) != sg):
                ok = False
                kS = K[S, j]
            out[S, j] = kS
        return out, ok

    def cd(lam, K, sweeps=3000):
        """Cyclic coordinate descent on all columns at once, covariance form."""
        R = Cf - G @ K                       # residual correlation  c - G k
        for it in range(sweeps):
            dmax = 0.0
            for i in range(P):
                if dG[i] <= 0:
                    continue
                old = K[i]
                rho = R[i] + dG[i] * old
                new = np.sign(rho) * np.maximum(np.abs(rho) - lam, 0.0) / dG[i]
                delta = new - old
                if np.any(delta != 0):
                    R -= np.outer(G[:, i], delta)
                    K[i] = new
                    dmax = max(dmax, float(np.max(np.abs(delta))))
            if dmax < tol * max(1.0, float(np.max(np.abs(K)))):
                break
        return K

    def solve_lam(lam, Kstart):
        K = cd(lam, Kstart.copy())
        Kp, ok = polish(K, lam)
        # accept the polish only if KKT holds on the inactive set
        if ok:
            grad = G @ Kp - Cf
            viol = np.max(np.where(Kp == 0, np.abs(grad) - lam, 0.0))
            if viol <= 1e-9 * max(lam, 1.0):
                K = Kp
        return K

    lam_hi = float(np.max(np.abs(Cf)))            # K(lam_hi) = 0
    lam_lo, K_lo = 0.0, None
    phi_hi = -t_free
    Kcur = np.zeros_like(Cf)
    lam = lam_hi
    # bracket: shrink lam geometrically until ||K||_1 > t
    while True:
        lam *= 0.5
        Kcur = solve_lam(lam, Kcur)
        phi = np.abs(Kcur).sum() - t_free
        if phi > 0:
            lam_lo, phi_lo, K_lo = lam, phi, Kcur
            break
        lam_hi, phi_hi = lam, phi
        if lam < 1e-300:
            return assemble(Kcur), {"lam": lam, "active": True}
    # Illinois / regula falsi on lam in [lam_lo (phi>0), lam_hi (phi<0)]
    side = 0
    for it in range(max_outer):
        lam = (lam_lo * phi_hi - lam_hi * phi_lo) / (phi_hi - phi_lo)
        if not (lam_lo < lam < lam_hi):
            lam = 0.5 * (lam_lo + lam_hi)
        Kcur = solve_lam(lam, K_lo)
        phi = np.abs(Kcur).sum() - t_free
        if verbose:
            print(f"  outer {it}: lam={lam:.6e} phi={phi:.3e}")
        if abs(phi) <= 1e-12 * max(t_free, 1.0):
            break
        if phi > 0:
            lam_lo, phi_lo, K_lo = lam, phi, Kcur
            if side == 1:
                phi_hi *= 0.5
            side = 1
        else:
            lam_hi, phi_hi = lam, phi
            if side == -1:
                phi_lo *= 0.5
            side = -1
        if lam_hi - lam_lo <= 1e-15 * lam_hi:
            break
    # final exact step: on the fixed sign pattern k_S(lam) is affine in lam, so solve
    # sum_j sign_S'(a_j - lam b_j) = t exactly.
    num, den = 0.0, 0.0
    sols = []
    for j in range(Kcur.shape[1]):
        S = np.nonzero(Kcur[:, j])[0]
        if S.size == 0:
            sols.append(None)
            continue
        sg = np.sign(Kcur[S, j])
        GS = G[np.ix_(S, S)]
        a = np.linalg.solve(GS, Cf[S, j])
        b = np.linalg.solve(GS, sg)
        num += sg @ a
        den += sg @ b
        sols.append((S, sg, a, b))
    if den > 0:
        lam_x = (num - t_free) / den
        Kx = np.zeros_like(Kcur)
        good = lam_x > 0
        for j, s in enumerate(sols):
            if s is None:
                continue
            S, sg, a, b = s
            kS = a - lam_x * b
            if np.any(np.sign(kS) != sg):
                good = False
                break
            Kx[S, j] = kS
        if good:
            grad = G @ Kx - Cf
            viol = np.max(np.where(Kx == 0, np.abs(grad) - lam_x, 0.0))
            if viol <= 1e-9 * max(lam_x, 1.0):
                Kcur, lam = Kx, lam_x
    return assemble(Kcur), {"lam": lam, "active": True}


def get_koopman(model_type, prog, pairs, lasso=1e6, N=None, n=None, nd=0, psd_shift="as_reference", loaded=False, ls_branch=None):
    """get_Koopman (Ksysid.m:987-1092) for one lasso value.

    lasso >= 1e6 -> K = Px \\ Py (1068-1069); otherwise the L1-ball QP with
    t = lasso * N (996) on G (+1e-6 I if any eigenvalue is negative, 1117-1120)
    and, for linear models with delays, the pinned delay columns (1139-1164).
    """
    w = pairs.get("w") if loaded else None
    Px, Py = build_regressors(model_type, prog, pairs["alpha"], pairs["beta"], pairs["u"], w)
    N = prog.N if N is None else N
    koop = {"Px": Px, "Py": Py, "u": pairs["u"], "alpha": pairs["alpha"], "N": N, "nw": 0 if w is None else w.shape[1]}
    if w is not None:
        koop["w"] = w
    if ls_branch is None:                                # the caller decides from the whole lasso property (Ksysid.m:1068);
        ls_branch = bool(np.all(np.atleast_1d(lasso) >= 1e6))   # a bare call decides from the value it was given
    if ls_branch:
        K, info = mldivide(Px, Py, return_info=True)
        koop.update(K=K, info=info)
        return koop
    G, C = gram(Px, Py)
    shifted = needs_psd_shift(G) if psd_shift == "as_reference" else (psd_shift == "always")
    if shifted:
        G = G + 1e-6 * np.eye(G.shape[0])
    fixed = None
    if model_type == "linear" and nd >= 1:
        m = pairs["u"].shape[1]
        fixed = delay_constraint_targets(N, n, m, nd)
    K, info = solve_l1ball_qp(G, C, float(lasso) * N, fixed=fixed)
    info.update(shifted=shifted, objective=qp_objective(G, C, K), l1=float(np.abs(K).sum()))
    koop.update(K=K, info=info, G=G, C=C)
    return koop


# --------------------------------------------------------------------------- models
def continuous_UT(K, Ts):
    """(1/Ts) logm(K' + 1e-12 I): the continuous-time generator (Ksysid.m:1186-1187, 1245-1246; 1310 for K itself)."""
    return sla.logm(K.T + 1e-12 * np.eye(K.shape[0])) / Ts


def get_model(koop, n, Ts=None):
    """Linear model A,B,C with the projection M = (L \\ R)' (Ksysid.m:1179-1235); Ts given = continuous time:
    A, B from the matrix logarithm and NOT projected (1220-1222)."""
    K, N = koop["K"], koop["N"] * (koop.get("nw", 0) + 1)       # N (nw+1) for a loaded model (Ksysid.m:1199-1203)
    UT = K.T if Ts is None else continuous_UT(K, Ts)
    A, B = UT[:N, :N], UT[:N, N:]
    Cy = np.concatenate([np.eye(n), np.zeros((n, N - n))], axis=1)
    Pxs, Pys, U = koop["Px"][:, :N], koop["Py"][:, :N], koop["u"]
    L = Pxs @ A.T + U @ B.T
    Mt = mldivide(L, Pys)
    Mp = Mt.T
    if Ts is not None:
        return {"A": A, "B": B, "C": Cy, "M": Mp, "K": K}
    return {"A": Mp @ A, "B": Mp @ B, "C": Cy, "M": Mp, "K": K}


def get_BLmodel(koop, n, Ts=None):
    """Bilinear model: A, B=[B_1..B_m], Beta(z)=B kron(I_m,z) (Ksysid.m:1238-1295)."""
    K, N = koop["K"], koop["N"] * (koop.get("nw", 0) + 1)
    UT = K.T if Ts is None else continuous_UT(K, Ts)
    A, B = UT[:N, :N], UT[:N, N:]
    Cy = np.concatenate([np.eye(n), np.zeros((n, N - n))], axis=1)
    return {"A": A, "B": B, "C": Cy, "K": K}


def get_NLmodel(koop, nzeta, n):
    """Nonlinear model F(zeta,u) = K(:,1:nzeta)' psi([zeta;u]), C = I_n (Ksysid.m:1298-1341)."""
    return {"F": koop["K"][:, :nzeta].T.copy(), "C": np.eye(n), "K": koop["K"]}


# --------------------------------------------------------------------------- validation
def get_error(ysim, yreal, treal):
    """Ksysid.m:1886-1891 (scaled units)."""
    d = ysim - yreal
    T = len(treal)
    rmse = np.sqrt(np.sum(d ** 2, axis=0) / T)
    return {"mean": np.mean(np.abs(d), axis=0), "rmse": rmse,
            "nrmse": rmse / np.abs(yreal.max(axis=0) - yreal.min(axis=0)),
            "euclid_mean": float(np.sum(np.sqrt(np.sum(d ** 2, axis=1))) / T)}


def _val_setup(val, nd):
    zetareal, _ = get_zeta(val, nd)
    return val["t"][nd:], val["y"][nd:], val["u"][nd:], zetareal


def val_model(model, prog, val, nd, nzeta, loaded=False):
    """Open-loop rollout z+ = A z + B u (Ksysid.m:1623-1714, discrete).  Loaded: the lifted state is re-expanded with the
    ACTUAL load every step, znow = kron(I, z(1:N)) [1; w_j] (1666)."""
    treal, yreal, ureal, zetareal = _val_setup(val, nd)
    A, B, C = model["A"], model["B"], model["C"]
    T = len(treal)
    ysim = np.zeros_like(yreal)
    ysim[0] = yreal[0]
    z = lift(prog, zetareal[0])[0]
    if loaded:
        wreal = np.asarray(val["w"], dtype=np.float64).reshape(len(val["t"]), -1)[nd:]
        N = z.size
        z = np.kron(np.concatenate([[1.0], wreal[0]]), z)
        for j in range(T - 1):
            znow = np.kron(np.concatenate([[1.0], wreal[j]]), z[:N])
            z = A @ znow + B @ ureal[j]
            ysim[j + 1] = C @ z
        return {"y": ysim, "yreal": yreal, "error": get_error(ysim, yreal, treal)}
    for j in range(T - 1):
        z = A @ z + B @ ureal[j]
        ysim[j + 1] = C @ z
    return {"y": ysim, "yreal": yreal, "error": get_error(ysim, yreal, treal)}


def val_BLmodel(model, prog, val, nd, nzeta):
    """z+ = A z + Beta(z) u (Ksysid.m:1717-1812, discrete, unloaded)."""
    treal, yreal, ureal, zetareal = _val_setup(val, nd)
    A, B, C = model["A"], model["B"], model["C"]
    T = len(treal)
    m = ureal.shape[1]
    ysim = np.zeros_like(yreal)
    ysim[0] = yreal[0]
    z = lift(prog, zetareal[0])[0]
    for j in range(T - 1):
        Beta = B @ np.kron(np.eye(m), z[:, None])
        z = A @ z + Beta @ ureal[j]
        ysim[j + 1] = C @ z
    return {"y": ysim, "yreal": yreal, "error": get_error(ysim, yreal, treal)}


def val_NLmodel(model, prog, val, nd, nzeta, n):
    """zeta+ = F(zeta, u) (Ksysid.m:1815-1879, discrete, unloaded)."""
    treal, yreal, ureal, zetareal = _val_setup(val, nd)
    F = model["F"]
    T = len(treal)
    zeta = zetareal[0].copy()
    ysim = np.zeros_like(yreal)
    ysim[0] = zeta[:n]
    for j in range(T - 1):
        zeta = F @ lift(prog, np.concatenate([zeta, ureal[j]]))[0]
        ysim[j + 1] = zeta[:n]
    return {"y": ysim, "yreal": yreal, "error": get_error(ysim, yreal, treal)}


# --------------------------------------------------------------------------- consumer side (Kmpc)
def mpc_costB_bilinear(A, B, z, horizon):
    """Kmpc.get_costB_bilinear (Kmpc.m:569-596): Bcol block i = A^(i-1) Beta(z_i), Beta(z) = B kron(I_m, z)
    (Ksysid.m:1288-1289; z_i = z(i,:) if z has several rows, else z(1,:)); every further block column is the previous one
    shifted down by N rows (Lshift, 586-594).  Returns the N (h+1) x m h matrix."""
    A, B = np.asarray(A, dtype=np.float64), np.asarray(B, dtype=np.float64)
    z = np.atleast_2d(np.asarray(z, dtype=np.float64))
    N = A.shape[0]
    m = B.shape[1] // N
    Bcol = np.zeros((N * (horizon + 1), m))
    Ap = np.eye(N)
    for i in range(1, horizon + 1):
        zi = z[i - 1] if z.shape[0] > 1 else z[0]
        Bmodel = B @ np.kron(np.eye(m), zi.reshape(-1, 1))
        Bcol[N * i:N * (i + 1)] = Ap @ Bmodel
        Ap = A @ Ap
    out = np.zeros((N * (horizon + 1), m * horizon))
    out[:, :m] = Bcol
    for c in range(1, horizon):
        out[N:, c * m:(c + 1) * m] = out[:-N, (c - 1) * m:c * m]
    return out


# --------------------------------------------------------------------------- dimension reduction
def pca_matlab(X):
    """MATLAB pca(X) coefficients + explained (Ksysid.m:1498): centred SVD, each
    coefficient column's largest-|.| entry made positive, explained = 100 var/sum."""
    Xc = X - X.mean(axis=0)
    _, s, Vt = np.linalg.svd(Xc, full_matrices=False)
    coeff = Vt.T.copy()
    for j in range(coeff.shape[1]):
        i = np.argmax(np.abs(coeff[:, j]))
        if coeff[i, j] < 0:
            coeff[:, j] = -coeff[:, j]
    lat = s ** 2 / (X.shape[0] - 1)
    return coeff, 100.0 * lat / lat.sum()


def econ_reduce(prog, model_type, pairs):
    """lift_snapshots + get_econ_observables with dim_red (Ksysid.m:1394-1432, 1495-1517)."""
    V = pairs["alpha"] if model_type != "nonlinear" else np.concatenate([pairs["alpha"], pairs["u"]], axis=1)
    coeff, explained = pca_matlab(lift_full(prog, V))
    num = 1
    while explained[:num].sum() < 99:
        num += 1
    prog.pcs = coeff[:, :num].copy()
    return prog


# --------------------------------------------------------------------------- class-level restatement
class KsysidOracle:
    """Constructor + train_models + validation of the reference class, CPU only.

    Mirrors Ksysid(data4sysid, 'model_type',..,'obs_type',..,'obs_degree',..,'delays',..,
    'lasso',..,'snapshots',..,'dim_red',..) (Ksysid.m:37-144) for time_type='discrete',
    loaded=false.
    """

    def __init__(self, data4sysid, model_type="linear", obs_type=("poly",), obs_degree=(1,),
                 delays=0, lasso=1e6, snapshots=np.inf, dim_red=False, centres=None):
        if model_type not in ("linear", "bilinear", "nonlinear"):
            raise ValueError("Invalid model_type chosen. Must be linear, bilinear, or nonlinear.")
        tr0 = data4sysid["train"][0]
        self.n, self.m = tr0["y"].shape[1], tr0["u"].shape[1]
        self.nd = int(delays)
        self.nzeta = self.n * (self.nd + 1) + self.m * self.nd
        self.model_type = model_type
        lasso = np.atleast_1d(np.asarray(lasso, dtype=np.float64)).copy()
        if np.all(np.isinf(lasso)):                       # parse_args 154-156: `if obj.lasso == Inf` is true only when ALL
            lasso = np.array([1e6])                       # entries are Inf, and the property then becomes the scalar 1e6
        self.lasso = lasso
        merged = merge_trials(data4sysid["train"])
        self.traindata, self.scale = get_scale(merged)
        self.valdata = [scale_data(v, self.scale) for v in data4sysid["val"]]
        self.pairs = get_snapshot_pairs(self.traindata, self.nd, snapshots)
        nv = self.nzeta + (self.m if model_type == "nonlinear" else 0)
        self.prog = build_program(list(obs_type), list(np.atleast_1d(obs_degree)), nv, centres)
        if dim_red:
            econ_reduce(self.prog, model_type, self.pairs)
        self.N = self.prog.N

    def train_models(self, lasso=None):
        """train_models (Ksysid.m:1344-1389): one full fit per lasso value; model = candidates[0].
        The LS/QP branch test is on the whole lasso PROPERTY (1068)."""
        lasso = self.lasso if lasso is None else np.atleast_1d(np.asarray(lasso, dtype=np.float64))
        ls_branch = bool(np.all(self.lasso >= 1e6))       # the test is on the PROPERTY (1068): a mixed vector sends EVERY entry,
        self.koopData, self.candidates = [], []           # even one >= 1e6, through the QP with t = lasso(i) * N
        for lam in lasso:
            koop = get_koopman(self.model_type, self.prog, self.pairs, lasso=lam, N=self.N, n=self.n, nd=self.nd,
                               ls_branch=ls_branch)
            if self.model_type == "nonlinear":
                mdl = get_NLmodel(koop, self.nzeta, self.n)
            elif self.model_type == "bilinear":
                mdl = get_BLmodel(koop, self.n)
            else:
                mdl = get_model(koop, self.n)
            mdl["lasso"] = lam
            self.koopData.append(koop)
            self.candidates.append(mdl)
        self.model = self.candidates[0]
        return self

    def validate(self, model=None, trial=0):
        model = self.model if model is None else model
        val = self.valdata[trial]
        if self.model_type == "nonlinear":
            return val_NLmodel(model, self.prog, val, self.nd, self.nzeta, self.n)
        if self.model_type == "bilinear":
            return val_BLmodel(model, self.prog, val, self.nd, self.nzeta)
        return val_model(model, self.prog, val, self.nd, self.nzeta)
