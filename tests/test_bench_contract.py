"""CPU test of bench.py's reference arm: one JSON line on stdout with the contract's keys."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--cpu-sample", "1024"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "edmd_fit_snapshots_per_sec" and d["unit"] == "snapshots/s"
    assert d["higher_is_better"] is True and d["dtype"] == "f64" and d["data"] == "synthetic" and d["vs_baseline"] is None
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"] and d["config"]["P"] == 4096


def test_other_ranks_of_reference_arm_exit_quietly():
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_recorded_gpu_bench_line_has_the_contract_keys():
    """The committed round-end record of `python bench.py` on a B200 (profiles/r01_bench_n1_full_final.json) carries every key
    of the bench contract, with consistent arithmetic (value = snapshots / time, frac = achieved / peak)."""
    d = json.load(open(os.path.join(ROOT, "profiles", "r01_bench_n1_full_final.json")))
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
              "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["n_gpus"] == 1 and d["warmup"] >= 3 and d["scaling"] == "weak" and d["dtype"] == "f64" and d["vs_baseline"] is None
    assert abs(d["value"] - d["config"]["total_snapshots"] / (d["ms_per_step"] * 1e-3)) <= 1e-6 * d["value"]
    r = d["roofline"]
    assert r["bound"] == "tensor" and r["unit"] == "TFLOP/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12 and r["traffic"] > 0
    assert 0.5 < r["frac"] < 1.05
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] == 216 * d["config"]["snapshots_per_gpu"] and e["d2h_bytes_per_step"] == 4096 * 4096 * 8
    assert e["value"] <= d["value"] * 1.02                       # end to end cannot beat the device-resident rate
    c = d["cpu_baseline"]
    assert c["kind"] == "port" and c["cores"] >= 1 and c["value"] > 0 and "sample" in c
    assert d["gpu_launches"] > 1000 and d["clocks"]["sm_mhz"] > 0 and not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    lo = d["lift_only"]
    assert lo["bound"] == "hbm" and lo["unit"] == "GB/s" and abs(lo["frac"] - lo["achieved"] / lo["peak"]) < 1e-12


def test_recorded_round2_bench_line_has_the_contract_keys():
    """The round-2 record (profiles/r02_bench_n1_full.json: INT8 tensor-core Gram engine) against the same contract; the lift-only
    extra must be at or above the 0.75 of the HBM copy peak that VERDICT r1 asked for, and the lasso3a extra fully converged."""
    d = json.loads(open(os.path.join(ROOT, "profiles", "r02_bench_n1_full.json")).read().strip().splitlines()[-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
              "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["n_gpus"] == 1 and d["warmup"] >= 3 and d["scaling"] == "weak" and d["dtype"] == "f64" and d["vs_baseline"] is None
    assert abs(d["value"] - d["config"]["total_snapshots"] / (d["ms_per_step"] * 1e-3)) <= 1e-6 * d["value"]
    r = d["roofline"]
    assert r["bound"] == "tensor" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12 and r["traffic"] > 0 and 0.5 < r["frac"] < 1.05
    assert "oz_gemm_kernel" in r["kernel"] and r["fp64_dgemm_peak_this_run"] > 30
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] == 216 * d["config"]["snapshots_per_gpu"] and e["d2h_bytes_per_step"] == 4096 * 4096 * 8
    assert 0 < e["value"] <= d["value"] * 1.02
    assert d["value"] > 3.0e6                                   # round 1: 1.03e6
    assert d["gpu_launches"] > 1000 and not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    lo = d["lift_only"]
    assert lo["bound"] == "hbm" and abs(lo["frac"] - lo["achieved"] / lo["peak"]) < 1e-12 and lo["frac"] >= 0.75
    la = d["lasso3a"]
    assert la["budgets"] == 64 and la["unconverged"] == 0 and la["worst_rel_gap"] < 1e-8 and la["seconds"] < 5.0
