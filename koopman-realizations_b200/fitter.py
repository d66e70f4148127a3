"""Thin ctypes driver of the C ABI (include/koopfit.h): one `Fitter` = one kf_ctx = one GPU.

This is the Python stand-in for the MEX shim described in INTEGRATION.md: it only
marshals column-major double buffers across the ABI; all arithmetic of the fit
(lift, Gram, solve) happens in libkoopfit.so on the GPU.  No CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _abi as A


class Fitter:
    def __init__(self, device=0):
        self.lib = A.load()
        self.ctx = C.c_void_p()
        rc = self.lib.kf_create(C.byref(self.ctx), int(device))
        if rc:
            raise A.KoopfitError(f"kf_create failed ({A.ERRORS.get(rc, rc)}): "
                                 f"{self.lib.kf_last_error(None).decode()}")
        self.device = device

    # ------------------------------------------------------------------ plumbing
    def close(self):
        if getattr(self, "ctx", None) and self.ctx.value:
            self.lib.kf_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc:
            raise A.KoopfitError(f"{what} failed ({A.ERRORS.get(rc, rc)}): {self.lib.kf_last_error(self.ctx).decode()}")

    def set_option(self, name, value):
        self._check(self.lib.kf_set_option(self.ctx, name.encode(), float(value)), "kf_set_option")
        if name == "qp_method":
            self._qp_method = int(value)

    def set_qp_partition(self, col_lo, col_hi, allreduce=None):
        """Column partition of the exact active-set QP solver (kf_set_qp_partition): this rank solves the columns
        [col_lo, col_hi) of K; `allreduce(values: np.ndarray, op)` must reduce `values` in place over the ranks
        (op 0 = sum, 1 = max).  col_hi <= col_lo switches the partition off."""
        if col_hi > col_lo:
            def hook(_user, ptr, n, op):
                try:
                    allreduce(np.ctypeslib.as_array(ptr, shape=(n,)), int(op))
                    return 0
                except Exception:       # an exception must not cross the C frame
                    import traceback
                    traceback.print_exc()
                    return 1
            self._qp_hook = A.ALLREDUCE_FN(hook)        # keep the trampoline alive
        else:
            self._qp_hook = A.ALLREDUCE_FN()
        self._check(self.lib.kf_set_qp_partition(self.ctx, int(col_lo), int(col_hi), self._qp_hook, None), "kf_set_qp_partition")

    def counters(self, reset=False):
        f, n = C.c_double(), C.c_longlong()
        self._check(self.lib.kf_counters(self.ctx, C.byref(f), C.byref(n), int(reset)), "kf_counters")
        return f.value, n.value

    def engine_info(self, reset=False):
        """Gram engine of the last accumulation (1 FP64 DMMA, 2 INT8 tensor cores) and its INT8 operation counters."""
        e, n, s_ = C.c_int(), C.c_longlong(), C.c_int()
        ops, per = C.c_double(), C.c_double()
        self._check(self.lib.kf_engine_info(self.ctx, C.byref(e), C.byref(ops), C.byref(per), C.byref(n), C.byref(s_), int(reset)), "kf_engine_info")
        return {"engine": e.value, "i8_ops": ops.value, "i8_ops_per_launch": per.value, "gram_launches": n.value, "gram_sampled": s_.value}

    def last_times(self):
        a, b, c = C.c_double(), C.c_double(), C.c_double()
        self._check(self.lib.kf_last_times(self.ctx, C.byref(a), C.byref(b), C.byref(c)), "kf_last_times")
        return {"lift_gram_ms": a.value, "gram_kernel_ms": b.value, "solve_ms": c.value}

    def sync(self):
        self._check(self.lib.kf_sync(self.ctx), "kf_sync")

    @property
    def stream(self):
        return self.lib.kf_stream(self.ctx)

    # ------------------------------------------------------------------ dictionary
    def dims(self, basis, model, m):
        nf, N, P = C.c_int(), C.c_int(), C.c_int()
        rc = self.lib.kf_basis_dims(basis.ref(), A.MODEL_CODE[model], int(m), C.byref(nf), C.byref(N), C.byref(P))
        if rc:
            raise A.KoopfitError(f"kf_basis_dims failed: {self.lib.kf_last_error(None).decode()}")
        return nf.value, N.value, P.value

    def lift(self, basis, V):
        """lift.econ_full on the rows of V (rows x nv) -> (rows x N)."""
        V = A.fcol(np.atleast_2d(V))
        rows = V.shape[0]
        _, N, _ = self.dims(basis, "nonlinear", 0)
        out = np.empty((rows, N), order="F")
        self._check(self.lib.kf_lift(self.ctx, basis.ref(), rows, A.dptr(V), A.dptr(out)), "kf_lift")
        return out

    def pca(self, basis, V):
        """MATLAB `pca` of the lifted rows of V (Ksysid.m:1498) on the GPU: returns mu (n_full), latent (decreasing), coeff
        (n_full x n_full, MATLAB sign convention).  `basis` must be the full dictionary (no pcs)."""
        V = A.fcol(np.atleast_2d(V))
        rows = V.shape[0]
        nf, _, _ = self.dims(basis, "nonlinear", 0)
        mu, latent = np.empty(nf), np.empty(nf)
        coeff = np.empty((nf, nf), order="F")
        self._check(self.lib.kf_pca(self.ctx, basis.ref(), rows, A.dptr(V), A.dptr(mu), A.dptr(latent), A.dptr(coeff)), "kf_pca")
        return mu, latent, coeff

    # ------------------------------------------------------------------ the fit
    @staticmethod
    def _solve_struct(least_squares=True, ls_method="auto", pivot_tol=0.0, t=None, psd_shift="as_reference",
                      delay_constraint=False, n=0, nd=0, qp_max_iter=0, qp_tol=0.0):
        sv = A.kf_solve()
        sv.least_squares = int(bool(least_squares))
        sv.ls_method = {"auto": A.KF_LS_AUTO, "gram": A.KF_LS_GRAM, "qr": A.KF_LS_QR}[ls_method]
        sv.pivot_tol = float(pivot_tol)
        keep = None
        if not least_squares:
            keep = np.ascontiguousarray(np.atleast_1d(np.asarray(t, dtype=np.float64)))
            sv.nt = keep.size
            sv.t = A.dptr(keep)
        sv.psd_shift = {"as_reference": 0, "never": 1, "always": 2}[psd_shift]
        sv.delay_constraint = int(bool(delay_constraint))
        sv.n, sv.nd = int(n), int(nd)
        sv.qp_max_iter = int(qp_max_iter)
        sv.qp_tol = float(qp_tol)
        return sv, keep

    def _run_fit(self, call, what, basis, model_type, M, nzeta, m, pa, pb, pu, want_gram, want_regressors, pc_cols, solve_kw,
                 rows_out=None, pw=None, nw=0):
        _, N, _ = self.dims(basis, model_type, m)
        P = A.regressor_width(model_type, N, m, nw)
        Pc = int(pc_cols) if 0 < int(pc_cols) < P else P
        pr = A.kf_problem(M=M, nzeta=nzeta, m=m, model=A.MODEL_CODE[model_type], alpha=pa, beta=pb, u=pu, pc_cols=int(pc_cols),
                          nw=int(nw), w=pw)
        sv, keep_t = self._solve_struct(**solve_kw)
        nt = max(1, sv.nt) if not sv.least_squares else 1
        res = A.kf_result()
        K = np.zeros((P, Pc, nt), order="F")
        perm = np.zeros(P, dtype=np.int32)
        obj, l1, gap = np.zeros(nt), np.zeros(nt), np.zeros(nt)
        iters = np.zeros(nt, dtype=np.int32)
        res.K, res.perm = A.dptr(K), perm.ctypes.data_as(A.c_int_p)
        res.objective, res.l1norm, res.qp_iters = A.dptr(obj), A.dptr(l1), iters.ctypes.data_as(A.c_int_p)
        res.qp_gap = A.dptr(gap)
        out = {}
        if want_gram:
            out["G"], out["C"] = np.zeros((P, P), order="F"), np.zeros((P, P), order="F")
            res.G, res.C = A.dptr(out["G"]), A.dptr(out["C"])
        if want_regressors:
            rows = M if rows_out is None else rows_out
            out["Px"], out["Py"] = np.zeros((rows, P), order="F"), np.zeros((rows, P), order="F")
            res.Px, res.Py = A.dptr(out["Px"]), A.dptr(out["Py"])
        call(basis.ref(), C.byref(pr), C.byref(sv), C.byref(res))
        info = {f: getattr(res.info, f) for f, _ in A.kf_info._fields_}
        out.update(K=K[:, :, 0] if nt == 1 else K, K_all=K, rank=info["rank"], perm=perm, info=info, N=N, P=P,
                   objective=obj, l1norm=l1, qp_iters=iters, qp_gap=gap)
        return out

    def fit(self, basis, model_type, alpha, beta, u, want_gram=False, want_regressors=False, pc_cols=0, w=None, **solve_kw):
        """get_Koopman for host snapshot pairs (Ksysid.m:987-1092): returns dict with K (P x P, or
        P x P x nt for a budget vector), rank, perm, info and optionally G, C, Px, Py.
        pc_cols > 0 (opt-in fast mode, least squares only): only the first pc_cols columns of K are computed.
        w (M x nw): the loads of a `loaded` model (snapshotPairs.w): the lifted state becomes [1; w] (x) psi.
        On a context that joined a communicator (comm_init) alpha / beta / u are THIS RANK's shard."""
        alpha, beta, u = A.fcol(alpha), A.fcol(beta), A.fcol(u)
        M, nzeta = alpha.shape
        m = u.shape[1]
        w = None if w is None or np.size(w) == 0 else A.fcol(np.asarray(w, dtype=np.float64).reshape(M, -1))

        def call(b, pr, sv, res):
            self._check(self.lib.kf_fit(self.ctx, b, pr, sv, res), "kf_fit")
        return self._run_fit(call, "kf_fit", basis, model_type, M, nzeta, m, alpha.ctypes.data, beta.ctypes.data, u.ctypes.data,
                             want_gram, want_regressors, pc_cols, solve_kw, pw=None if w is None else w.ctypes.data,
                             nw=0 if w is None else w.shape[1])

    def fit_dev(self, basis, model_type, M, nzeta, m, alpha_ptr, beta_ptr, u_ptr, want_gram=False, pc_cols=0, **solve_kw):
        """kf_fit_dev: the whole fit from DEVICE-resident snapshot pairs (pointers from torch tensors' data_ptr(); column-major
        M x nzeta, i.e. contiguous torch tensors of shape (nzeta, M)); on a multi-GPU context the all-reduce and any
        refinement passes run inside the library."""
        def call(b, pr, sv, res):
            self._check(self.lib.kf_fit_dev(self.ctx, b, pr, sv, res), "kf_fit_dev")
        return self._run_fit(call, "kf_fit_dev", basis, model_type, int(M), int(nzeta), int(m), int(alpha_ptr), int(beta_ptr), int(u_ptr),
                             want_gram, False, pc_cols, solve_kw)

    # ------------------------------------------------------------------ multi-GPU: one process per GPU
    @staticmethod
    def comm_unique_id():
        """128-byte NCCL id (kf_comm_unique_id); draw it on rank 0 and broadcast it to the other ranks."""
        lib = A.load()
        buf = C.create_string_buffer(128)
        rc = lib.kf_comm_unique_id(buf, 128)
        if rc:
            raise A.KoopfitError("kf_comm_unique_id failed: NCCL unavailable")
        return buf.raw

    def comm_init(self, nranks, rank, unique_id):
        """Join the library-owned NCCL communicator (kf_comm_init_rank); collective over all ranks."""
        buf = C.create_string_buffer(bytes(unique_id), 128)
        self._check(self.lib.kf_comm_init_rank(self.ctx, int(nranks), int(rank), buf, 128), "kf_comm_init_rank")

    def comm_info(self):
        n, r, v = C.c_int(), C.c_int(), C.c_int()
        self._check(self.lib.kf_comm_info(self.ctx, C.byref(n), C.byref(r), C.byref(v)), "kf_comm_info")
        return {"nranks": n.value, "rank": r.value, "nccl_version": v.value}

    def fit_series(self, basis, model_type, t, y, u, nd=0, prescaled=False, want_gram=False, want_regressors=False, pc_cols=0,
                   **solve_kw):
        """The fit from the raw merged training series (kf_fit_series): scaling (Ksysid.m:180-229), delay embedding
        (868-907) and snapshot pairs (910-984, snapshots = Inf) are built on the device.  Returns the dict of `fit`
        plus scale = {y_offset, y_factor, u_offset, u_factor} and M."""
        t = np.ascontiguousarray(np.asarray(t, dtype=np.float64).ravel())
        y = A.fcol(np.asarray(y, dtype=np.float64).reshape(t.size, -1))
        u = np.zeros((t.size, 0), order="F") if u is None or np.size(u) == 0 else A.fcol(np.asarray(u, dtype=np.float64).reshape(t.size, -1))
        T, n = y.shape
        m = u.shape[1]
        _, N, P = self.dims(basis, model_type, m)
        Pc = int(pc_cols) if 0 < int(pc_cols) < P else P
        ser = A.kf_series(T=T, n=n, m=m, nd=int(nd), model=A.MODEL_CODE[model_type], t=A.dptr(t), y=A.dptr(y), u=A.dptr(u),
                          prescaled=int(bool(prescaled)), pc_cols=int(pc_cols))
        sv, keep_t = self._solve_struct(**solve_kw)
        nt = max(1, sv.nt) if not sv.least_squares else 1
        res = A.kf_result()
        K = np.zeros((P, Pc, nt), order="F")
        perm = np.zeros(P, dtype=np.int32)
        obj, l1, gap = np.zeros(nt), np.zeros(nt), np.zeros(nt)
        iters = np.zeros(nt, dtype=np.int32)
        res.K, res.perm = A.dptr(K), perm.ctypes.data_as(A.c_int_p)
        res.objective, res.l1norm, res.qp_iters, res.qp_gap = A.dptr(obj), A.dptr(l1), iters.ctypes.data_as(A.c_int_p), A.dptr(gap)
        out = {}
        if want_gram:
            out["G"], out["C"] = np.zeros((P, P), order="F"), np.zeros((P, P), order="F")
            res.G, res.C = A.dptr(out["G"]), A.dptr(out["C"])
        if want_regressors:
            M = int(self.lib.kf_series_pairs(T, int(nd), A.dptr(t)))
            out["Px"], out["Py"] = np.zeros((M, P), order="F"), np.zeros((M, P), order="F")
            res.Px, res.Py = A.dptr(out["Px"]), A.dptr(out["Py"])
        sc = A.kf_scale()
        scale = {k: np.zeros(n if k[0] == "y" else m) for k in ("y_offset", "y_factor", "u_offset", "u_factor")}
        for k, v in scale.items():
            setattr(sc, k, A.dptr(v))
        self._check(self.lib.kf_fit_series(self.ctx, basis.ref(), C.byref(ser), C.byref(sv), C.byref(sc), C.byref(res)), "kf_fit_series")
        info = {f: getattr(res.info, f) for f, _ in A.kf_info._fields_}
        out.update(K=K[:, :, 0] if nt == 1 else K, K_all=K, rank=info["rank"], perm=perm, info=info, N=N, P=P,
                   objective=obj, l1norm=l1, qp_iters=iters, qp_gap=gap, scale=scale, M=int(sc.M))
        return out

    def fit_batch(self, problems):
        """Many independent fits in one call (kf_fit_batch).  `problems`: list of dicts with keys basis, model_type,
        alpha, beta, u and optional solve keywords (as Fitter.fit).  Small problems (P <= 32) run concurrently on the GPU, one
        CTA each: least-squares fits, and QP fits whose LS solution lies inside every budget (then it is the QP minimiser);
        the rest are solved one after the other.  Returns a list of dicts
        (K, rank, perm, info [, objective, l1norm])."""
        n = len(problems)
        bases = (C.POINTER(A.kf_basis) * n)()
        probs = (A.kf_problem * n)()
        solves = (A.kf_solve * n)()
        outs = (A.kf_result * n)()
        keep, results = [], []
        for i, pb in enumerate(problems):
            alpha, beta, u = pb["alpha"], pb["beta"], pb["u"]
            if not (isinstance(alpha, np.ndarray) and alpha.flags.f_contiguous and alpha.dtype == np.float64):
                alpha = A.fcol(alpha)
            if not (isinstance(beta, np.ndarray) and beta.flags.f_contiguous and beta.dtype == np.float64):
                beta = A.fcol(beta)
            if not (isinstance(u, np.ndarray) and u.flags.f_contiguous and u.dtype == np.float64):
                u = A.fcol(u)
            M, nzeta = alpha.shape
            m = u.shape[1]
            _, N, P = self.dims(pb["basis"], pb["model_type"], m)
            bases[i] = C.pointer(pb["basis"].struct)
            probs[i] = A.kf_problem(M=M, nzeta=nzeta, m=m, model=A.MODEL_CODE[pb["model_type"]],
                                    alpha=alpha.ctypes.data, beta=beta.ctypes.data, u=u.ctypes.data)
            kw = {k: v for k, v in pb.items() if k not in ("basis", "model_type", "alpha", "beta", "u")}
            sv, keep_t = self._solve_struct(**kw)
            solves[i] = sv
            nt = max(1, sv.nt) if not sv.least_squares else 1
            K = np.zeros((P, P, nt), order="F")
            perm = np.zeros(P, dtype=np.int32)
            obj, l1, gap = np.zeros(nt), np.zeros(nt), np.zeros(nt)
            iters = np.zeros(nt, dtype=np.int32)
            outs[i].K, outs[i].perm = A.dptr(K), perm.ctypes.data_as(A.c_int_p)
            outs[i].objective, outs[i].l1norm, outs[i].qp_iters = A.dptr(obj), A.dptr(l1), iters.ctypes.data_as(A.c_int_p)
            outs[i].qp_gap = A.dptr(gap)
            keep.append((alpha, beta, u, keep_t, sv))
            results.append(dict(K_all=K, perm=perm, objective=obj, l1norm=l1, qp_iters=iters, qp_gap=gap, N=N, P=P, nt=nt))
        self._check(self.lib.kf_fit_batch(self.ctx, n, bases, probs, solves, outs), "kf_fit_batch")
        for i, r in enumerate(results):
            info = {f: getattr(outs[i].info, f) for f, _ in A.kf_info._fields_}
            r.update(K=r["K_all"][:, :, 0] if r["nt"] == 1 else r["K_all"], rank=info["rank"], info=info)
        return results

    def rollout(self, basis, model_type, n, m, nzeta, models, trials, nout=0):
        """Open-loop validation rollouts on the GPU (kf_rollout; val_model / val_BLmodel / val_NLmodel,
        Ksysid.m:1623-1879).  models: list of dicts with A, B (linear, bilinear) or F (nonlinear: K(:,1:nzeta)');
        trials: list of (zeta0 (nzeta,), u (T x m)).  Returns out[c][k] = (T_k x nout) simulated state rows
        (nout = n: y; nzeta: zeta; N: z) for candidate c on trial k."""
        nc, nt = len(models), len(trials)
        _, N, _ = self.dims(basis, model_type, m)
        nout = int(nout) if nout else int(n)
        mds = (A.kf_model * nc)()
        keep = []
        for c, md in enumerate(models):
            mds[c].model = A.MODEL_CODE[model_type]
            mds[c].n, mds[c].m, mds[c].nzeta, mds[c].N = int(n), int(m), int(nzeta), int(N)
            want = {"linear": {"A": (N, N), "B": (N, m)}, "bilinear": {"A": (N, N), "B": (N, N * m)},
                    "nonlinear": {"F": (nzeta, N)}}[model_type]
            for name, shape in want.items():
                if md.get(name) is None or np.shape(md[name]) != shape:
                    raise A.KoopfitError(f"rollout: model {c} needs {name} of shape {shape}, got {np.shape(md.get(name))}")
                arr = A.fcol(md[name])
                keep.append(arr)
                setattr(mds[c], name, A.dptr(arr))
        T = (C.c_int * nt)()
        z0 = (A.c_double_p * nt)()
        up = (A.c_double_p * nt)()
        yp = (A.c_double_p * (nt * nc))()
        for k, (zeta0, u) in enumerate(trials):
            zz = np.ascontiguousarray(np.asarray(zeta0, dtype=np.float64).ravel())
            uu = A.fcol(np.asarray(u, dtype=np.float64).reshape(-1, m))
            keep += [zz, uu]
            T[k] = uu.shape[0]
            z0[k], up[k] = A.dptr(zz), A.dptr(uu)
        out = []
        for c in range(nc):
            row = []
            for k in range(nt):
                yy = np.zeros((T[k], nout), order="F")
                row.append(yy)
                yp[c * nt + k] = A.dptr(yy)
            out.append(row)
        self._check(self.lib.kf_rollout(self.ctx, basis.ref(), nc, mds, nt, T, z0, up, nout, yp), "kf_rollout")
        return out

    def mpc_costB_bilinear(self, Amat, Bmat, z, horizon):
        """Kmpc.get_costB_bilinear (Kmpc.m:569-596) on the GPU.  Amat (N x N), Bmat (N x N m) of the bilinear model; z: (N,),
        (1, N) or (horizon, N) lifted state(s) along the horizon, or a batch (nbatch, nz, N).  Returns B of shape
        (N (h+1), m h), or (nbatch, N (h+1), m h) for a batch."""
        Amat, Bmat = A.fcol(Amat), A.fcol(Bmat)
        N = Amat.shape[0]
        m = Bmat.shape[1] // N
        z = np.asarray(z, dtype=np.float64)
        batched = z.ndim == 3
        zb = z if batched else np.atleast_2d(z)[None, :, :]
        nb, nz = zb.shape[0], zb.shape[1]
        zf = np.ascontiguousarray(np.transpose(zb, (0, 2, 1)))          # per problem: (N, nz) C-order == (nz x N) column-major
        out = np.zeros((nb, m * horizon, N * (horizon + 1)))             # per problem: C-order (cols, rows) == column-major
        self._check(self.lib.kf_mpc_costB_bilinear(self.ctx, N, m, int(horizon), nb, A.dptr(Amat), A.dptr(Bmat), nz, A.dptr(zf),
                                                   A.dptr(out)), "kf_mpc_costB_bilinear")
        res = np.transpose(out, (0, 2, 1))
        return res if batched else res[0]

    def mldivide(self, Amat, Bmat):
        """MATLAB `A \\ B` (QRCP basic solution) on the GPU; returns X, rank, perm."""
        Amat, Bmat = A.fcol(Amat), A.fcol(Bmat)
        M, P = Amat.shape
        Pc = Bmat.shape[1]
        X = np.zeros((P, Pc), order="F")
        perm = np.zeros(P, dtype=np.int32)
        rank = C.c_int()
        self._check(self.lib.kf_mldivide(self.ctx, M, P, Pc, A.dptr(Amat), A.dptr(Bmat), A.dptr(X),
                                         perm.ctypes.data_as(A.c_int_p), C.byref(rank)), "kf_mldivide")
        return X, rank.value, perm

    # ------------------------------------------------------------------ staged / device-resident API
    def accumulate_dev(self, basis, model_type, M, nzeta, m, alpha_ptr, beta_ptr, u_ptr, reset=True, pc_cols=0):
        """Partial Gram of a device-resident shard (pointers from torch tensors' data_ptr());
        column-major M x nzeta / M x m, i.e. torch tensors of shape (nzeta, M) contiguous."""
        pr = A.kf_problem(M=int(M), nzeta=int(nzeta), m=int(m), model=A.MODEL_CODE[model_type],
                          alpha=int(alpha_ptr), beta=int(beta_ptr), u=int(u_ptr), pc_cols=int(pc_cols))
        self._check(self.lib.kf_accumulate_dev(self.ctx, basis.ref(), C.byref(pr), int(reset)), "kf_accumulate_dev")
        if reset or getattr(self, "_acc_dims", None) is None:
            # the dimensions kf_solve_dev will write with: the host buffers of solve_dev are sized from THESE, never from the
            # caller's idea of P / pc_cols (a mismatch would overflow them — the ABI carries no buffer sizes)
            _, _, P = self.dims(basis, model_type, int(m))
            self._acc_dims = (P, int(pc_cols) if 0 < int(pc_cols) < P else P)
            self._solve_state = None

    def regressors_dev(self, basis, model_type, M, nzeta, m, alpha_ptr, beta_ptr, u_ptr, out_ptr, ld=None):
        """Lift-only mode on device buffers (kf_regressors_dev): [Px | Py] (M x 2P, column-major, ld >= M) written to
        out_ptr; asynchronous on the context stream."""
        pr = A.kf_problem(M=int(M), nzeta=int(nzeta), m=int(m), model=A.MODEL_CODE[model_type],
                          alpha=int(alpha_ptr), beta=int(beta_ptr), u=int(u_ptr), pc_cols=0)
        self._check(self.lib.kf_regressors_dev(self.ctx, basis.ref(), C.byref(pr), int(out_ptr), int(ld if ld else M)), "kf_regressors_dev")

    def lift_dev(self, basis, rows, v_ptr, psi_ptr):
        """lift.econ_full on `rows` device points (kf_lift_dev): V (rows x nv) -> Psi (rows x N), column-major."""
        self._check(self.lib.kf_lift_dev(self.ctx, basis.ref(), int(rows), int(v_ptr), int(psi_ptr)), "kf_lift_dev")

    def accum_buffer(self):
        """(device pointer, count of doubles) of the packed partial-Gram accumulator."""
        p, n = C.c_void_p(), C.c_size_t()
        self._check(self.lib.kf_accum_buffer(self.ctx, C.byref(p), C.byref(n)), "kf_accum_buffer")
        return p.value, n.value

    def solve_dev(self, P=None, want_gram=False, pc_cols=None, **solve_kw):
        """Solve from the (all-reduced) accumulator.  The output buffers are sized from the layout recorded by accumulate_dev;
        P / pc_cols, if given, are only checked against it.

        Returns the result dict, or None when the library asks for ANOTHER data pass (KF_EAGAIN: ill-conditioned regressor,
        Gram-route refinement).  The caller then repeats accumulate_dev(..., reset=False) on the same shard(s), all-reduces
        accum_buffer() and calls solve_dev again with the same arguments; `fit_sharded` does this loop."""
        dims = getattr(self, "_acc_dims", None)
        if dims is None:
            raise A.KoopfitError("solve_dev: nothing accumulated (call accumulate_dev first)")
        Pa, Pca = dims
        if P is not None and int(P) != Pa:
            raise A.KoopfitError(f"solve_dev: P = {P} does not match the accumulated regressor width {Pa}")
        if pc_cols is not None and (int(pc_cols) if 0 < int(pc_cols) < Pa else Pa) != Pca:
            raise A.KoopfitError(f"solve_dev: pc_cols = {pc_cols} does not match the accumulated layout ({Pca} columns)")
        P, Pc = Pa, Pca
        st = getattr(self, "_solve_state", None)
        if st is None:                      # first call: allocate; a continuation after KF_EAGAIN reuses the same buffers
            sv, keep_t = self._solve_struct(**solve_kw)
            nt = max(1, sv.nt) if not sv.least_squares else 1
            res = A.kf_result()
            K = np.zeros((P, Pc, nt), order="F")
            perm = np.zeros(P, dtype=np.int32)
            obj, l1, gap = np.zeros(nt), np.zeros(nt), np.zeros(nt)
            iters = np.zeros(nt, dtype=np.int32)
            res.K, res.perm = A.dptr(K), perm.ctypes.data_as(A.c_int_p)
            res.objective, res.l1norm, res.qp_iters = A.dptr(obj), A.dptr(l1), iters.ctypes.data_as(A.c_int_p)
            res.qp_gap = A.dptr(gap)
            out = {}
            if want_gram:
                out["G"], out["C"] = np.zeros((P, P), order="F"), np.zeros((P, P), order="F")
                res.G, res.C = A.dptr(out["G"]), A.dptr(out["C"])
            st = dict(sv=sv, keep_t=keep_t, nt=nt, res=res, K=K, perm=perm, obj=obj, l1=l1, gap=gap, iters=iters, out=out)
        rc = self.lib.kf_solve_dev(self.ctx, C.byref(st["sv"]), C.byref(st["res"]))
        if rc == A.KF_EAGAIN:
            self._solve_state = st
            return None
        self._solve_state = None
        self._acc_dims = None
        self._check(rc, "kf_solve_dev")
        res, K, nt, out = st["res"], st["K"], st["nt"], st["out"]
        info = {f: getattr(res.info, f) for f, _ in A.kf_info._fields_}
        out.update(K=K[:, :, 0] if nt == 1 else K, K_all=K, rank=info["rank"], perm=st["perm"], info=info, P=P,
                   objective=st["obj"], l1norm=st["l1"], qp_iters=st["iters"], qp_gap=st["gap"])
        return out


class MultiFitter(Fitter):
    """One process, several GPUs (kf_create_multi / kf_fit_multi): what the MEX shim uses to reach every visible B200 from a
    single MATLAB session.  `fit` has Fitter.fit's contract on the FULL host snapshot pairs; the split into per-device shards,
    the NCCL all-reduce of the partial Grams and the column split of a lasso sweep happen inside the library."""

    def __init__(self, devices=None):
        self.lib = A.load()
        self.mc = C.c_void_p()
        if devices is None:
            rc = self.lib.kf_create_multi(C.byref(self.mc), None, 0)
        else:
            arr = (C.c_int * len(devices))(*[int(d) for d in devices])
            rc = self.lib.kf_create_multi(C.byref(self.mc), arr, len(devices))
        if rc:
            raise A.KoopfitError(f"kf_create_multi failed ({A.ERRORS.get(rc, rc)}): {self.lib.kf_last_error(None).decode()}")
        self.ndev = self.lib.kf_multi_size(self.mc)
        self.ctx = C.c_void_p(self.lib.kf_multi_ctx(self.mc, 0))       # device 0's context: dims, options, counters

    def close(self):
        if getattr(self, "mc", None) and self.mc.value:
            self.lib.kf_destroy_multi(self.mc)
            self.mc = C.c_void_p()
            self.ctx = C.c_void_p()

    def set_option(self, name, value):
        rc = self.lib.kf_multi_set_option(self.mc, name.encode(), float(value))
        if rc:
            raise A.KoopfitError(f"kf_multi_set_option({name}) failed")

    def fit(self, basis, model_type, alpha, beta, u, want_gram=False, want_regressors=False, pc_cols=0, w=None, **solve_kw):
        alpha, beta, u = A.fcol(alpha), A.fcol(beta), A.fcol(u)
        M, nzeta = alpha.shape
        m = u.shape[1]
        w = None if w is None or np.size(w) == 0 else A.fcol(np.asarray(w, dtype=np.float64).reshape(M, -1))

        def call(b, pr, sv, res):
            rc = self.lib.kf_fit_multi(self.mc, b, pr, sv, res)
            if rc:
                raise A.KoopfitError(f"kf_fit_multi failed ({A.ERRORS.get(rc, rc)}): {self.lib.kf_multi_last_error(self.mc).decode()}")
        return self._run_fit(call, "kf_fit_multi", basis, model_type, M, nzeta, m, alpha.ctypes.data, beta.ctypes.data, u.ctypes.data,
                             want_gram, want_regressors, pc_cols, solve_kw, pw=None if w is None else w.ctypes.data,
                             nw=0 if w is None else w.shape[1])
