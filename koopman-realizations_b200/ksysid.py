"""Host-side mirror of the reference's `Ksysid` class (Ksysid.m) over the koopfit C ABI.

Same constructor name-value API ('model_type', 'obs_type', 'obs_degree', 'delays', 'lasso',
'snapshots', 'dim_red', ...), same property names (params, lift, basis, model, candidates,
koopData, scaledown, scaleup, traindata, valdata, snapshotPairs) and the same returned
model layout (A, B, C, M / Beta / F_func, K, params, lasso), so a Kmpc/Ksim-style consumer
reads the result unchanged.  What moves to the GPU is the hot path: get_Koopman's lift
loop, Px'Px / Px'Py, `\\` and the L1-ball QP (Ksysid.m:987-1176) and the lasso loop of
train_models (1370-1387) — one kf_fit call.  The MATLAB original keeps this host logic in
MATLAB; here it is Python because MATLAB/Octave are absent.

Pre-processing (merge / scale / zeta / snapshot pairs, Ksysid.m:180-229, 380-401, 868-984) is
available both ways: the host mirror below (the `traindata` / `snapshotPairs` properties of the
reference) and, with `device_preprocess=True`, kf_fit_series, which takes the raw merged series and
builds scale factors, delay embedding and pairs on the GPU (SURVEY §8f "next #3").
"""
from __future__ import annotations

import numpy as np

from . import _abi as A
from .fitter import Fitter


# ------------------------------------------------------------------ host pre-processing
def merge_trials(trials):
    """Ksysid.m:380-401 (every numeric field: the load w rides along when every trial has one)."""
    if isinstance(trials, dict):
        return trials
    out = {k: np.concatenate([np.asarray(t[k], dtype=np.float64) for t in trials], axis=0) for k in ("t", "y", "u")}
    if all("w" in t for t in trials):
        out["w"] = np.concatenate([np.asarray(t["w"], dtype=np.float64).reshape(len(t["t"]), -1) for t in trials], axis=0)
    return out


def _scale_factors(data):
    """Ksysid.m:187-204: centre and half-range per column, unit factor for constant columns; the load w (246-265) is only
    shifted to zero where it is constant and scaled into [-1, 1] where it varies."""
    out = {}
    for k in ("y", "u"):
        mn, mx = data[k].min(axis=0), data[k].max(axis=0)
        half = (mx - mn) / 2.0
        out[k + "_offset"] = (mx + mn) / 2.0
        out[k + "_factor"] = np.where(half == 0, 1.0, half)
    if "w" in data:
        mn, mx = data["w"].min(axis=0), data["w"].max(axis=0)
        out["w_offset"] = (mx + mn) / 2.0
        out["w_factor"] = np.where(mn != mx, (mx - mn) / 2.0, 1.0)
    return out


def _zeta(y, u, nd):
    """Ksysid.m:868-907: zeta_i = [y_i, y_{i-1..i-nd}, u_{i-1..i-nd}], uzeta_i = u_i."""
    T = y.shape[0]
    if nd == 0:
        return y.copy(), u.copy()
    cols = [y[nd:T]] + [y[nd - j:T - j] for j in range(1, nd + 1)] + [u[nd - j:T - j] for j in range(1, nd + 1)]
    return np.concatenate(cols, axis=1), u[nd:T].copy()


class Ksysid:
    """Koopman-based system identification; GPU fit through libkoopfit.so.

    Ksysid(data4sysid, model_type='linear'|'bilinear'|'nonlinear', obs_type=['poly',...],
           obs_degree=[...], delays=0, lasso=Inf|vector, snapshots=Inf, dim_red=None, ...)

    data4sysid: dict with 'train' and 'val' lists of trials, each a dict with t (T,), y (T,n), u (T,m).
    Extra (not in the reference): `centres` — gaussian centres (the reference draws them from MATLAB's
    global rand, Ksysid.m:803; they must be an input here), `device`, `fitter`, `ls_method`.
    """

    def __init__(self, data4sysid, **kw):
        if "train" not in data4sysid or "val" not in data4sysid:
            raise ValueError("Input must have *train* and *val* fields of type cell array")   # Ksysid.m:46-48
        # defaults, Ksysid.m:73-81 (dim_red has NO default there: [] behaves as "reduce", see 137-141)
        self.isupdate = False
        self.obs_type = ["poly"]
        self.obs_degree = [1]
        self.snapshots = np.inf
        self.lasso = [1e6]
        self.delays = 0
        self.model_type = "linear"
        self.loaded = False
        self.time_type = "discrete"
        self.dim_red = None
        extras = {"centres": None, "device": 0, "fitter": None, "ls_method": "auto", "rng_seed": 0, "fast_cols": False,
                  "device_preprocess": False}
        for k, v in kw.items():          # parse_args, Ksysid.m:147-158
            if k in extras:
                extras[k] = v
            else:
                setattr(self, k, v)
        self.lasso = np.atleast_1d(np.asarray(self.lasso, dtype=np.float64)).copy()
        if np.all(np.isinf(self.lasso)):
            self.lasso = np.array([1e6])                     # Ksysid.m:155-157
        if isinstance(self.obs_type, str):
            self.obs_type = [self.obs_type]
        self.obs_degree = [int(d) for d in np.atleast_1d(self.obs_degree)]
        if self.model_type not in ("linear", "bilinear", "nonlinear"):
            raise ValueError("Invalid model_type chosen. Must be linear, bilinear, or nonlinear.")   # 103
        if self.loaded and not all("w" in t for t in data4sysid["train"]):
            raise ValueError("You have specified a loaded system, but your training data does not have the required load field (w)")  # 107-109
        if self.time_type not in ("discrete", "continuous"):
            raise ValueError("time_type must be discrete or continuous")
        self.liftinput = {"linear": 0, "nonlinear": 1, "bilinear": 2}[self.model_type]
        self._ls_method = extras["ls_method"]
        self._rng_seed = int(extras["rng_seed"])            # seed of the `snapshots < Inf` draw (MATLAB's stream is not reproducible here)
        # opt-in: compute only the K columns the model consumes (K(:,1:N) or K(:,1:nzeta)); model['K'] is then P x Pc
        self._fast_cols = bool(extras["fast_cols"])
        # opt-in: train_models hands the raw merged series to the GPU (bilinear / nonlinear models, snapshots = Inf)
        self._device_preprocess = bool(extras["device_preprocess"])

        tr0 = data4sysid["train"][0]
        y0, u0 = np.atleast_2d(np.asarray(tr0["y"], float)), np.atleast_2d(np.asarray(tr0["u"], float))
        self.params = {"n": y0.shape[1], "m": u0.shape[1],
                       "Ts": float(np.mean(np.diff(np.asarray(tr0["t"], float)))), "isfake": False}
        self.params["nd"] = int(self.delays)
        n, m, nd = self.params["n"], self.params["m"], self.params["nd"]
        self.params["nzeta"] = n * (nd + 1) + m * nd
        # merge + scale (Ksysid.m:119-128)
        merged = merge_trials(data4sysid["train"])
        self.params["nw"] = merged["w"].shape[1] if (self.loaded and "w" in merged) else 0      # Ksysid.m:88-93
        self._merged_raw = merged
        sc = _scale_factors(merged)
        self.params["scale"] = sc
        self.scaledown = {"y": lambda y: (y - sc["y_offset"]) / sc["y_factor"], "u": lambda u: (u - sc["u_offset"]) / sc["u_factor"]}
        self.scaleup = {"y": lambda y: y * sc["y_factor"] + sc["y_offset"], "u": lambda u: u * sc["u_factor"] + sc["u_offset"]}
        self.traindata = {"t": merged["t"], "y": self.scaledown["y"](merged["y"]), "u": self.scaledown["u"](merged["u"])}
        if "w" in merged:
            self.scaledown["w"] = lambda w: (w - sc["w_offset"]) / sc["w_factor"]
            self.scaleup["w"] = lambda w: w * sc["w_factor"] + sc["w_offset"]
            self.traindata["w"] = self.scaledown["w"](merged["w"])
        self.valdata = [self.scale_data(v) for v in data4sysid["val"]]
        self.snapshotPairs = self.get_snapshotPairs(self.traindata, self.snapshots)

        # dictionary (def_observables, Ksysid.m:455-536): a numeric descriptor crosses the ABI
        nv = self.params["nzeta"] + (m if self.model_type == "nonlinear" else 0)
        self.basis = A.Basis(self.obs_type, self.obs_degree, nv, centres=extras["centres"])
        self.fitter = extras["fitter"] or Fitter(device=extras["device"])
        self.lift = {"full": lambda v: self._lift(v, econ=False), "econ_full": lambda v: self._lift(v, econ=True),
                     "econ_full_input": self._lift_input}
        if self.loaded:                                       # def_observables_loaded (Ksysid.m:539-626), econ_* (1580-1611)
            self.lift["full_loaded"] = lambda v, w: self._lift_loaded(v, w, econ=False)
            self.lift["econ_full_loaded"] = lambda v, w: self._lift_loaded(v, w, econ=True)
            self.lift["econ_full_loaded_input"] = self._lift_loaded_input
        if self.dim_red is None or self.dim_red:            # `if ~obj.dim_red ... else` with [] -> reduce (137-141)
            self._reduce_dimension()
        self.params["N"] = self.fitter.dims(self.basis, self.model_type, m)[1]
        print(f"Number of basis functions: {self.params['N']}")     # Ksysid.m:143
        self.model = None
        self.candidates = None
        self.koopData = None

    # ------------------------------------------------------------------ data helpers
    def scale_data(self, trial, down=True):
        """Ksysid.m:308-343."""
        f = self.scaledown if down else self.scaleup
        out = {"t": np.asarray(trial["t"], float), "y": f["y"](np.asarray(trial["y"], float)),
               "u": f["u"](np.asarray(trial["u"], float))}
        if "w" in trial and "w" in f:                         # Ksysid.m:328-330
            out["w"] = f["w"](np.asarray(trial["w"], float).reshape(len(out["t"]), -1))
        return out

    def get_zeta(self, data):
        """Ksysid.m:868-907."""
        zeta, uzeta = _zeta(data["y"], data["u"], self.params["nd"])
        out = dict(data)
        out["zeta"], out["uzeta"] = zeta, uzeta
        return out, zeta

    def get_snapshotPairs(self, data, snapshots=np.inf):
        """Ksysid.m:910-984.  Trial-boundary pairs dropped (948); num_max = kept-1 (960).  With
        snapshots=Inf the reference permutes the first num_max kept pairs (974-975); order does not
        change G, C or K, so natural order is kept.  A smaller `snapshots` draws without replacement
        from numpy's generator (MATLAB's mlfg6331_64 stream is not reproducible here)."""
        print("Constructing snapshots...")
        data = merge_trials(data)
        nd = self.params["nd"]
        zeta, uzeta = _zeta(data["y"], data["u"], nd)
        t = data["t"]
        good = np.nonzero(t[nd:-1] < t[nd + 1:])[0]
        num_max = good.size - 1
        idx = good[:num_max]
        if np.isfinite(snapshots) and snapshots <= num_max - 1:
            idx = np.sort(np.random.default_rng(getattr(self, '_rng_seed', 0)).choice(idx, size=int(snapshots), replace=False))
        elif np.isfinite(snapshots):
            print(f"Number of snapshot pairs cannot exceed {num_max}. Taking {num_max} pairs instead.")
        pairs = {"alpha": zeta[:-1][idx], "beta": zeta[1:][idx], "u": uzeta[:-1][idx]}
        if "w" in data:                                       # wzeta(1:end-1) at the kept points (Ksysid.m:953-957, 980-982)
            pairs["w"] = np.asarray(data["w"], float).reshape(len(t), -1)[nd:][:-1][idx]
        return pairs

    # ------------------------------------------------------------------ lifting
    def _lift(self, v, econ=True):
        v = np.asarray(v, dtype=np.float64)
        single = v.ndim == 1
        V = v[None, :] if single else v
        if econ or self.basis.pcs is None:
            out = self.fitter.lift(self.basis, V)
        else:
            full = A.Basis(self.obs_type, self.obs_degree, self.basis.nv, centres=self.basis.centres)
            out = self.fitter.lift(full, V)
        return out[0] if single else out

    def _lift_input(self, zeta, u):
        """lift.econ_full_input (Ksysid.m:1593-1604): [psi; u_1 psi; ...; u_m psi]."""
        z = self._lift(zeta)
        u = np.asarray(u, float)
        return np.concatenate([z] + [u[..., k:k + 1] * z for k in range(u.shape[-1])], axis=-1)

    def _lift_loaded(self, v, w, econ=True):
        """lift.econ_full_loaded (Ksysid.m:1607-1611): [psi; w_1 psi; ...; w_nw psi] = [1; w] (x) psi."""
        z = self._lift(v, econ=econ)
        w = np.asarray(w, float)
        return np.concatenate([z] + [w[..., c:c + 1] * z for c in range(w.shape[-1])], axis=-1)

    def _lift_loaded_input(self, zeta, w, u):
        """lift.econ_full_loaded_input (Ksysid.m:1580-1590): [psi_L; u_1 psi_L; ...; u_m psi_L]."""
        z = self._lift_loaded(zeta, w)
        u = np.asarray(u, float)
        return np.concatenate([z] + [u[..., k:k + 1] * z for k in range(u.shape[-1])], axis=-1)

    def _reduce_dimension(self):
        """lift_snapshots + get_econ_observables with dim_red (Ksysid.m:1394-1432, 1495-1517): `pca` runs on the GPU."""
        print("Performing dimensional reduction...")
        sp = self.snapshotPairs
        # pca of the lifted [alpha, u] (nonlinear) or alpha points, on the GPU: means, centred Gram, Jacobi eigensolver (kf_pca)
        V = np.concatenate([sp["alpha"], sp["u"]], axis=1) if self.model_type == "nonlinear" else sp["alpha"]
        _, lam, vec = self.fitter.pca(self.basis, V)
        explained = 100.0 * lam / lam.sum()
        num_pcs = 1
        while explained[:num_pcs].sum() < 99:               # Ksysid.m:1501-1504
            num_pcs += 1
        self.basis.set_pcs(vec[:, :num_pcs])

    # ------------------------------------------------------------------ the fit
    def get_Koopman(self, snapshotPairs, lasso=None):
        """Ksysid.m:987-1092 — one GPU call; `lasso` may be a vector (all budgets share G, C)."""
        print("Finding Koopman operator approximation...")
        N = self.params["N"]
        lasso = self.lasso if lasso is None else np.atleast_1d(np.asarray(lasso, dtype=np.float64))
        want_reg = self.model_type == "linear"              # get_model needs koopData.Px / Py (1206-1216)
        w = snapshotPairs.get("w") if self.loaded else None  # nw = 0 without the field (Ksysid.m:1006-1011)
        NL = N * ((w.shape[1] if w is not None else 0) + 1)
        if np.all(self.lasso >= 1e6):                       # branch test on the PROPERTY (Ksysid.m:1068)
            pc = (self.params["nzeta"] if self.model_type == "nonlinear" else N) if (self._fast_cols and w is None) else 0
            res = self.fitter.fit(self.basis, self.model_type, snapshotPairs["alpha"], snapshotPairs["beta"],
                                  snapshotPairs["u"], want_regressors=want_reg, least_squares=True, ls_method=self._ls_method,
                                  pc_cols=pc, w=w)
        else:
            res = self.fitter.fit(self.basis, self.model_type, snapshotPairs["alpha"], snapshotPairs["beta"],
                                  snapshotPairs["u"], want_regressors=want_reg, least_squares=False, t=lasso * N,
                                  delay_constraint=(self.model_type == "linear" and self.params["nd"] >= 1),
                                  n=self.params["n"], nd=self.params["nd"], w=w)
        out = []
        # the reference re-runs get_Koopman once per lasso entry (train_models 1370-1387): a vector whose entries are all >= 1e6
        # gives length(lasso) identical least-squares candidates
        ncand = max(res["K_all"].shape[2], len(lasso) if np.all(self.lasso >= 1e6) else 0)
        for i in range(ncand):
            kd = {"K": np.array(res["K_all"][:, :, min(i, res["K_all"].shape[2] - 1)]), "u": snapshotPairs["u"], "alpha": snapshotPairs["alpha"],
                  "info": res["info"], "rank": res["rank"]}
            if w is not None:
                kd["w"] = w                                  # Ksysid.m:1088-1090
            if want_reg:
                kd["Px"], kd["Py"] = res["Px"][:, :NL], res["Py"][:, :NL]     # Ksysid.m:1085-1086: N (nw+1) state columns
            out.append(kd)
        return out

    def _UT(self, K):
        """K' for a discrete model; (1/Ts) logm(K' + 1e-12 I) for a continuous one (Ksysid.m:1186-1190, 1245-1249).
        Needs the whole K (not the fast-columns variant).  A complex logarithm (negative real eigenvalues) is returned as
        MATLAB would; a negligible imaginary part is dropped."""
        if self.time_type != "continuous":
            return K.T
        if K.shape[0] != K.shape[1]:
            raise ValueError("continuous-time models need the full K (fast_cols must be off)")
        from scipy.linalg import logm
        UT = logm(K.T + 1e-12 * np.eye(K.shape[0])) / self.params["Ts"]
        if np.iscomplexobj(UT) and np.abs(UT.imag).max() <= 1e-9 * max(np.abs(UT.real).max(), 1e-300):
            UT = UT.real
        return UT

    def get_model(self, koopData):
        """Ksysid.m:1179-1235: A, B, C and the projection M = (L \\ R)' (applied to A, B only for discrete models)."""
        n = self.params["n"]
        N = self.params["N"] * ((self.params["nw"] if "w" in koopData else 0) + 1)      # N (nw+1), Ksysid.m:1192-1203
        UT = self._UT(koopData["K"])[:N, :]
        Amat, Bmat = UT[:N, :N], UT[:N, N:]
        Cy = np.concatenate([np.eye(n), np.zeros((n, N - n))], axis=1)
        L = koopData["Px"] @ Amat.T + koopData["u"] @ Bmat.T
        Mt, _, _ = self.fitter.mldivide(L, koopData["Py"])          # `L \ R` on the GPU (Ksysid.m:1216)
        Mp = Mt.T
        if self.time_type == "continuous":                  # Ksysid.m:1220-1222: no projection for continuous models
            return {"A": Amat, "B": Bmat, "C": Cy, "M": Mp, "params": self.params, "K": koopData["K"]}
        return {"A": Mp @ Amat, "B": Mp @ Bmat, "C": Cy, "M": Mp, "params": self.params, "K": koopData["K"]}

    def get_BLmodel(self, koopData):
        """Ksysid.m:1238-1282."""
        n, m = self.params["n"], self.params["m"]
        N = self.params["N"] * ((self.params["nw"] if "w" in koopData else 0) + 1)      # Ksysid.m:1251-1259
        UT = self._UT(koopData["K"])[:N, :]
        Amat, Bmat = UT[:N, :N], UT[:N, N:]
        Cy = np.concatenate([np.eye(n), np.zeros((n, N - n))], axis=1)
        return {"A": Amat, "B": Bmat, "Beta": lambda z: Bmat @ np.kron(np.eye(m), np.asarray(z).reshape(-1, 1)),
                "C": Cy, "params": self.params, "K": koopData["K"]}

    def get_NLmodel(self, koopData):
        """Ksysid.m:1298-1341: F(zeta,u) = K(:,1:nzeta)' psi([zeta;u]); C = I_n."""
        nz, n = self.params["nzeta"], self.params["n"]
        Kc = self._UT(koopData["K"]).T if self.time_type == "continuous" else koopData["K"]    # Ksysid.m:1308-1311
        F = np.array(Kc[:, :nz].T)
        if self.loaded and "w" in koopData:                  # F = K(:,1:nzeta)' basis_loaded, F_func(zeta, u, w) (Ksysid.m:1320-1327)
            return {"F_sym": F, "F_func": lambda zeta, u, w: F @ self._lift_loaded(np.concatenate([np.ravel(zeta), np.ravel(u)]), np.ravel(w)),
                    "params": self.params, "C": np.eye(n), "K": koopData["K"]}
        return {"F_sym": F, "F_func": lambda zeta, u: F @ self._lift(np.concatenate([np.ravel(zeta), np.ravel(u)])),
                "params": self.params, "C": np.eye(n), "K": koopData["K"]}

    def get_Koopman_series(self, lasso=None):
        """get_Koopman from the raw merged training series: scaling, zeta and the pairs are built on the GPU
        (kf_fit_series); same outputs as get_Koopman (koopData without Px / Py)."""
        print("Finding Koopman operator approximation...")
        N = self.params["N"]
        lasso = self.lasso if lasso is None else np.atleast_1d(np.asarray(lasso, dtype=np.float64))
        raw = self._merged_raw
        kw = dict(nd=self.params["nd"])
        if np.all(self.lasso >= 1e6):
            pc = (self.params["nzeta"] if self.model_type == "nonlinear" else N) if self._fast_cols else 0
            res = self.fitter.fit_series(self.basis, self.model_type, raw["t"], raw["y"], raw["u"], least_squares=True,
                                         ls_method=self._ls_method, pc_cols=pc, **kw)
        else:
            res = self.fitter.fit_series(self.basis, self.model_type, raw["t"], raw["y"], raw["u"], least_squares=False,
                                         t=lasso * N, **kw)
        self.params["scale_device"] = res["scale"]
        return [{"K": np.array(res["K_all"][:, :, i]), "info": res["info"], "rank": res["rank"], "M": res["M"]}
                for i in range(res["K_all"].shape[2])]

    def train_models(self, lasso=None):
        """Ksysid.m:1344-1389: candidates per lasso value; model = candidates{1}."""
        lasso = self.lasso if lasso is None else np.atleast_1d(np.asarray(lasso, dtype=np.float64))
        if self._device_preprocess and self.model_type != "linear" and not np.isfinite(self.snapshots):
            kds = self.get_Koopman_series(lasso)
        else:
            kds = self.get_Koopman(self.snapshotPairs, lasso)
        extract = {"nonlinear": self.get_NLmodel, "bilinear": self.get_BLmodel, "linear": self.get_model}[self.model_type]
        cands = []
        for i, kd in enumerate(kds):
            mdl = extract(kd)
            mdl["lasso"] = float(lasso[min(i, len(lasso) - 1)])
            cands.append(mdl)
        if len(lasso) < 2:
            self.koopData, self.candidates, self.model = kds[0], cands[0], cands[0]
        else:
            self.koopData, self.candidates, self.model = kds, cands, cands[0]
        return self

    # ------------------------------------------------------------------ validation (Ksysid.m:1623-1898)
    def get_error(self, sim, real):
        d = sim["y"] - real["y"]
        T = len(real["t"])
        rmse = np.sqrt(np.sum(d ** 2, axis=0) / T)
        eu = np.sqrt(np.sum(d ** 2, axis=1))
        return {"abs": np.abs(d), "mean": np.mean(np.abs(d), axis=0), "rmse": rmse,
                "nrmse": rmse / np.abs(real["y"].max(axis=0) - real["y"].min(axis=0)),
                "euclid": eu, "euclid_mean": float(eu.sum() / T)}

    def _val_setup(self, valdata):
        nd = self.params["nd"]
        _, zetareal = self.get_zeta(valdata)
        return valdata["t"][nd:], valdata["y"][nd:], valdata["u"][nd:], zetareal

    def _rollout(self, models, valdatas, nout):
        """All (candidate, trial) open-loop simulations in one kf_rollout call (one CTA each)."""
        setups = [self._val_setup(v) for v in valdatas]
        trials = [(zetareal[0], ureal) for (_, _, ureal, zetareal) in setups]
        p = self.params
        key = {"linear": ("A", "B"), "bilinear": ("A", "B"), "nonlinear": ("F_sym",)}[self.model_type]
        mds = [{("F" if k == "F_sym" else k): mdl[k] for k in key} for mdl in models]
        sims = self.fitter.rollout(self.basis, self.model_type, p["n"], p["m"], p["nzeta"], mds, trials, nout=nout)
        return setups, sims

    def _val_result(self, setup, states):
        treal, yreal, ureal, _ = setup
        n, nz = self.params["n"], self.params["nzeta"]
        sim = {"t": treal, "u": ureal, "y": np.ascontiguousarray(states[:, :n])}
        if states.shape[1] >= nz:
            sim["zeta"] = np.ascontiguousarray(states[:, :nz])
        if self.model_type != "nonlinear" and states.shape[1] in (self.params["N"], self.params["N"] * (self.params["nw"] + 1)):
            sim["z"] = states
        res = {"t": treal, "sim": sim, "real": {"t": treal, "u": ureal, "y": yreal}}
        res["error"] = self.get_error(res["sim"], res["real"])
        return res

    def _val_continuous(self, model, valdata):
        """Continuous-time validation (Ksysid.m:1681-1684, 1777-1780, 1852-1855): every sample interval is integrated
        with RK45 (ode45's pair; MATLAB's default tolerances RelTol 1e-3, AbsTol 1e-6) holding the input.  Host side, as
        in the reference: the GPU rollout kernel covers the discrete models."""
        from scipy.integrate import solve_ivp
        setup = self._val_setup(valdata)
        treal, yreal, ureal, zetareal = setup
        Ts = self.params["Ts"]
        if self.model_type == "nonlinear":
            x = zetareal[0].copy()
            rhs = lambda uj: (lambda t, zeta: model["F_func"](zeta, uj))
        elif self.model_type == "bilinear":
            x = self._lift(zetareal[0])
            rhs = lambda uj: (lambda t, z: model["A"] @ z + model["Beta"](z) @ uj)
        else:
            x = self._lift(zetareal[0])
            rhs = lambda uj: (lambda t, z: model["A"] @ z + model["B"] @ uj)
        states = np.zeros((len(treal), x.size))
        states[0] = x
        for j in range(len(treal) - 1):
            x = solve_ivp(rhs(ureal[j]), (0.0, Ts), x, method="RK45", rtol=1e-3, atol=1e-6).y[:, -1]
            states[j + 1] = x
        return self._val_result(setup, states)

    def _val_loaded(self, model, valdata):
        """Loaded linear / bilinear validation (Ksysid.m:1657-1670, 1751-1764): the lifted state is re-expanded with the ACTUAL
        load every step, znow = kron(I_{nw+1}, z(1:N)) [1; w_j]; z+ = A znow + B u (linear) or A znow + Beta(znow) u (bilinear).
        Host side like the reference (a T-step sequential recursion of one trial)."""
        setup = self._val_setup(valdata)
        treal, yreal, ureal, zetareal = setup
        nd, N, m = self.params["nd"], self.params["N"], self.params["m"]
        wreal = np.asarray(valdata["w"], float).reshape(len(valdata["t"]), -1)[nd:]
        z = self._lift_loaded(zetareal[0], wreal[0])
        states = np.zeros((len(treal), z.size))
        states[0] = z
        for j in range(len(treal) - 1):
            znow = np.kron(np.concatenate([[1.0], wreal[j]]), z[:N])
            if self.model_type == "bilinear":
                z = model["A"] @ znow + (model["B"] @ np.kron(np.eye(m), znow.reshape(-1, 1))) @ ureal[j]
            else:
                z = model["A"] @ znow + model["B"] @ ureal[j]
            states[j + 1] = z
        res = self._val_result(setup, states)
        res["sim"]["w"], res["real"]["w"] = wreal, wreal
        return res

    def val_model(self, model, valdata):
        """Ksysid.m:1623-1722 (discrete): z+ = A z + B u (1685), y = C z; on the GPU (kf_rollout) for unloaded models."""
        if self.time_type == "continuous":
            return self._val_continuous(model, valdata)
        if self.loaded:
            return self._val_loaded(model, valdata)
        setups, sims = self._rollout([model], [valdata], self.params["N"])
        return self._val_result(setups[0], sims[0][0])

    def val_BLmodel(self, model, valdata):
        """Ksysid.m:1725-1820: z+ = A z + Beta(z) u (1783)."""
        if self.time_type == "continuous":
            return self._val_continuous(model, valdata)
        if self.loaded:
            return self._val_loaded(model, valdata)
        setups, sims = self._rollout([model], [valdata], self.params["N"])
        return self._val_result(setups[0], sims[0][0])

    def val_NLmodel(self, model, valdata):
        """Ksysid.m:1823-1879: zeta+ = F_func(zeta, u) (1860), y = zeta(1:n) (1864)."""
        if self.time_type == "continuous":
            return self._val_continuous(model, valdata)
        setups, sims = self._rollout([model], [valdata], self.params["nzeta"])
        return self._val_result(setups[0], sims[0][0])

    def validate_candidates(self, candidates=None, trials=None):
        """Every candidate of a lasso vector (train_models 1370-1387) on every validation trial in ONE GPU call
        (grid = trials x candidates); returns results[c][k] in the val_* layout.  The reference would loop
        valNplot_model over candidates and trials (Ksysid.m:1928-1972)."""
        cands = candidates if candidates is not None else (self.candidates if isinstance(self.candidates, list) else [self.model])
        vals = self.valdata if trials is None else [self.valdata[i] for i in trials]
        if self.time_type == "continuous":
            return [[self._val_continuous(c, v) for v in vals] for c in cands]
        setups, sims = self._rollout(cands, vals, self.params["nzeta"])
        return [[self._val_result(setups[k], sims[c][k]) for k in range(len(vals))] for c in range(len(cands))]

    def valNplot_model(self, trial=None, **_):
        """valNplot_model without the plots (Ksysid.m:1928-1972): validate the chosen model on val trials."""
        trials = range(len(self.valdata)) if trial is None else [trial]
        fn = {"nonlinear": self.val_NLmodel, "bilinear": self.val_BLmodel, "linear": self.val_model}[self.model_type]
        return [fn(self.model, self.valdata[i]) for i in trials]
