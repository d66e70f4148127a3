"""kf_fit_series: get_scale / get_zeta / get_snapshotPairs on the device (Ksysid.m:180-229, 868-984) against the oracle's
host pre-processing followed by the ordinary fit."""
import numpy as np
import pytest

import koopfit
from koopfit.ksysid import Ksysid
import oracle as O

pytestmark = pytest.mark.gpu


def relF(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


@pytest.mark.parametrize("model,nd", [("bilinear", 0), ("bilinear", 2), ("nonlinear", 1), ("linear", 1)])
def test_series_fit_equals_fit_on_host_pairs(fitter, arm_data, model, nd):
    """Ten merged training trials (trial boundaries where t jumps back): scale factors, number of pairs, G, C and K
    from the device pre-processing equal those from the oracle's get_scale -> get_zeta -> get_snapshot_pairs."""
    merged = O.merge_trials(arm_data["train"])
    scaled, sc = O.get_scale(merged)
    pairs = O.get_snapshot_pairs(scaled, nd)
    n, m = merged["y"].shape[1], merged["u"].shape[1]
    nz = n * (nd + 1) + m * nd
    nv = nz + (m if model == "nonlinear" else 0)
    basis = koopfit.Basis(["poly"], [2], nv)
    ref = fitter.fit(basis, model, pairs["alpha"], pairs["beta"], pairs["u"], want_gram=True, ls_method="gram")
    got = fitter.fit_series(basis, model, merged["t"], merged["y"], merged["u"], nd=nd, want_gram=True, ls_method="gram")
    assert got["M"] == pairs["alpha"].shape[0]
    for k in ("y_offset", "y_factor", "u_offset", "u_factor"):
        assert np.array_equal(got["scale"][k], sc[k])
    assert relF(got["G"], ref["G"]) < 1e-15 and relF(got["C"], ref["C"]) < 1e-15
    assert got["rank"] == ref["rank"] and relF(got["K"], ref["K"]) < 1e-12
    # and against the oracle's own numbers
    prog = O.build_program(["poly"], [2], nv)
    Px, Py = O.build_regressors(model, prog, pairs["alpha"], pairs["beta"], pairs["u"])
    G, C = O.gram(Px, Py)
    assert relF(got["G"], G) < 1e-13 and relF(got["C"], C) < 1e-13


def test_series_regressors_and_prescaled(fitter, arm_data):
    """koopData.Px / Py from the series path (M rows known from kf_series_pairs), and the prescaled switch."""
    merged = O.merge_trials(arm_data["train"])
    scaled, sc = O.get_scale(merged)
    pairs = O.get_snapshot_pairs(scaled, 1)
    basis = koopfit.Basis(["poly"], [1], 15)
    got = fitter.fit_series(basis, "linear", merged["t"], merged["y"], merged["u"], nd=1, want_regressors=True)
    prog = O.build_program(["poly"], [1], 15)
    Px, Py = O.build_regressors("linear", prog, pairs["alpha"], pairs["beta"], pairs["u"])
    assert got["Px"].shape == Px.shape and np.abs(got["Px"] - Px).max() < 1e-15 and np.abs(got["Py"] - Py).max() < 1e-15
    Ko = O.mldivide(Px, Py)
    assert relF(got["K"], Ko) < 1e-9
    pre = fitter.fit_series(basis, "linear", scaled["t"], scaled["y"], scaled["u"], nd=1, prescaled=True)
    assert np.all(pre["scale"]["y_factor"] == 1.0) and np.all(pre["scale"]["u_offset"] == 0.0)
    assert relF(pre["K"], got["K"]) < 1e-12


def test_series_edge_cases(fitter):
    """A constant column gets factor 1 (Ksysid.m:196-203); every second sample a trial boundary; too-short series."""
    rng = np.random.default_rng(2)
    T = 5000
    t = np.tile(np.arange(50) * 0.1, T // 50)          # 100 trials of 50 samples
    y = np.cumsum(rng.standard_normal((T, 2)), axis=0) * 0.01
    y[:, 1] = 3.0                                       # constant state
    u = rng.standard_normal((T, 1))
    data = {"t": t, "y": y, "u": u}
    scaled, sc = O.get_scale(data)
    assert sc["y_factor"][1] == 1.0
    pairs = O.get_snapshot_pairs(scaled, 1)
    basis = koopfit.Basis(["poly"], [2], 5)
    got = fitter.fit_series(basis, "bilinear", t, y, u, nd=1, want_gram=True, ls_method="gram")
    assert got["M"] == pairs["alpha"].shape[0] == (T - 2) - 99 - 1
    assert got["scale"]["y_factor"][1] == 1.0 and got["scale"]["y_offset"][1] == 3.0
    prog = O.build_program(["poly"], [2], 5)
    Px, Py = O.build_regressors("bilinear", prog, pairs["alpha"], pairs["beta"], pairs["u"])
    G, C = O.gram(Px, Py)
    assert relF(got["G"], G) < 1e-13 and relF(got["C"], C) < 1e-13
    with pytest.raises(koopfit.KoopfitError):
        fitter.fit_series(basis, "bilinear", t[:2], y[:2], u[:2], nd=1)
    with pytest.raises(koopfit.KoopfitError):
        fitter.fit_series(basis, "bilinear", np.zeros(10), y[:10], u[:10], nd=1)     # t never increases: no pairs


def test_ksysid_device_preprocess(fitter, arm_data):
    """Ksysid(..., device_preprocess=True).train_models gives the same model as the host pre-processing path."""
    kw = dict(model_type="bilinear", obs_type=["poly"], obs_degree=[2], delays=1, dim_red=False, fitter=fitter, ls_method="gram")
    a = Ksysid(arm_data, **kw).train_models()
    b = Ksysid(arm_data, device_preprocess=True, **kw).train_models()
    assert relF(b.model["A"], a.model["A"]) < 1e-12 and relF(b.model["B"], a.model["B"]) < 1e-12
    for k in ("y_offset", "y_factor", "u_offset", "u_factor"):
        assert np.array_equal(b.params["scale_device"][k], b.params["scale"][k])
