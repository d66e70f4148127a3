"""Measure the FP64 roofline denominators on this GPU: cuBLAS DGEMM (torch.matmul float64)
burst / sustained, and write gpurun_out/fp64_peak.json.  (first builder task.)"""
import json, os, sys, time
import torch

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
dev = torch.device("cuda:0")
a = torch.randn(n, n, dtype=torch.float64, device=dev)
b = torch.randn(n, n, dtype=torch.float64, device=dev)
for _ in range(3):
    c = a @ b
torch.cuda.synchronize()
best = 1e9
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); c = a @ b; e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
burst = 2 * n ** 3 / best / 1e9
t0 = time.time(); iters = 0
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
while time.time() - t0 < 4.0:
    for _ in range(5):
        c = a @ b
    iters += 5
    torch.cuda.synchronize()
e1.record(); torch.cuda.synchronize()
sust = 2 * n ** 3 * iters / e0.elapsed_time(e1) / 1e9
# syrk-like A^T A in the TN layout the fit uses
at = a.t().contiguous()
for _ in range(3):
    c = at.t() @ a
torch.cuda.synchronize()
out = {"n": n, "dgemm_tflops_burst": burst, "dgemm_tflops_sustained": sust, "gpu": torch.cuda.get_device_name(0)}
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/fp64_peak.json", "w"), indent=1)
print(json.dumps(out))
