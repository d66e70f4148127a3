"""K1 alone: the materialising lift is HBM-bound — 8 (2 nzeta + m) B read + 16 P B written per pair.
Times kf_regressors_dev on device-resident pairs with CUDA events on the library's stream and reports the achieved
ALGORITHMIC GB/s against the measured HBM peak (MEASURED_PEAKS.json).  Cases: a polynomial dictionary (pure data
movement), the config-5 dictionary (569 gaussians: exp() on the FP64 pipe), both kernels (lift_tile = 1 / 0)."""
import json, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import koopfit

peak = 6554.6
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
dev = torch.device("cuda", 0)
fit = koopfit.Fitter(0)
st = torch.cuda.ExternalStream(fit.stream, device=dev)
rng = np.random.default_rng(2)
out = {"hbm_peak_gbs": peak, "cases": []}
# calibration: a pure-write stream (fill) and a copy of the same 8 GiB on this GPU, torch kernels, CUDA events
buf = torch.empty(1 << 30, dtype=torch.float64, device=dev)
buf2 = torch.empty(1 << 30, dtype=torch.float64, device=dev)
cal = {}
for name, fn, nbytes in (("fill_write_only", lambda: buf.fill_(1.5), 8.0 * (1 << 30)), ("copy_read_write", lambda: buf2.copy_(buf), 16.0 * (1 << 30))):
    for _ in range(2):
        fn()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        fn()
    e1.record()
    torch.cuda.synchronize(dev)
    cal[name] = nbytes * 5 / e0.elapsed_time(e1) / 1e6
out["calibration_GB_s"] = cal
print(cal, flush=True)
del buf, buf2
VARIANTS = [(1, 0, 1, 110, 2), (1, 32, 1, 110, 2), (1, 64, 1, 110, 2), (1, 0, 1, 72, 3), (1, 0, 0, 64, 2)]
if len(sys.argv) > 1 and sys.argv[1] == "short":
    VARIANTS = [(1, 0, 1, 110, 2), (1, 0, 0, 64, 2)]
cases = [("poly4_nv12_bilinear_m3", ["poly"], [4], None, 12, 3, "bilinear", 133000),
         ("poly3_nv12_linear_m3", ["poly"], [3], None, 12, 3, "linear", 1050000),
         ("config5_poly3_gauss569_bilinear", ["poly", "gaussian"], [3, 569], 2 * rng.random((12, 569)) - 1, 12, 3, "bilinear", 133000),
         ("config5_dictionary_65536_pairs_aligned", ["poly", "gaussian"], [3, 569], 2 * rng.random((12, 569)) - 1, 12, 3, "bilinear", 65536)]
for name, types, degs, cen, nz, m, model, M in cases:
    basis = koopfit.Basis(types, degs, nz, centres=cen)
    _, N, P = fit.dims(basis, model, m)
    a = torch.rand((nz, M), dtype=torch.float64, device=dev) * 2 - 1
    b = torch.rand((nz, M), dtype=torch.float64, device=dev) * 2 - 1
    u = torch.rand((m, M), dtype=torch.float64, device=dev) * 2 - 1
    o = torch.empty((2 * P, M), dtype=torch.float64, device=dev)
    alg = M * (8.0 * (2 * nz + m) + 16.0 * P)
    rec = {"case": name, "M": M, "N": N, "P": P, "algorithmic_bytes": alg, "output_GB": 16.0 * P * M / 1e9}
    for tile, ls, wide, kb, minb in VARIANTS:
        fit.set_option("lift_minb", minb)
        fit.set_option("lift_tile", tile)
        fit.set_option("lift_ls", ls)
        fit.set_option("lift_wide", wide)
        fit.set_option("lift_smem_kb", kb)
        for _ in range(3):
            fit.regressors_dev(basis, model, M, nz, m, a.data_ptr(), b.data_ptr(), u.data_ptr(), o.data_ptr())
        fit.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        with torch.cuda.stream(st):
            e0.record(st)
            for _ in range(reps):
                fit.regressors_dev(basis, model, M, nz, m, a.data_ptr(), b.data_ptr(), u.data_ptr(), o.data_ptr())
            e1.record(st)
        fit.sync(); torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1) / reps
        key = ((f"stream_kernel_{kb}kb_minb{minb}" if wide else "tile_kernel") + ("" if ls == 0 else f"_ls{ls}")) if tile else "level_kernel"
        rec[key] = {"ms": round(ms, 4), "GB_s": round(alg / ms / 1e6, 1), "frac_of_hbm_peak": round(alg / ms / 1e6 / peak, 4)}
    fit.set_option("lift_ls", 0)
    fit.set_option("lift_tile", 1)
    fit.set_option("lift_wide", 1)
    fit.set_option("lift_smem_kb", 110)
    fit.set_option("lift_minb", 2)
    print(rec, flush=True)
    out["cases"].append(rec)
    del o
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "lift_bw.json"), "w"), indent=1)
