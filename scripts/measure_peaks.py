"""Dense GEMM peaks for the MMA kinds the kernels actually issue, measured the way MEASURED_PEAKS.json measures bf16:
torch.matmul 8192^3 (2*N^3 flop), best of 10 (burst) and back to back for 4 s (sustained), CUDA events.
fp16 = the kind::f16 pipe of the assign / kNN kernels; tf32 = the kind::tf32 pipe of the probe / head kernels.
Writes gpurun_out/measured_peaks_extra.json (copy it to profiles/r2_measured_peaks_extra.json)."""
import json
import os
import time

import torch

N = 8192
dev = torch.device("cuda:0")
out = {"gpu_name": torch.cuda.get_device_name(0), "torch": torch.__version__,
       "how": "torch.matmul N=8192 square, 2*N^3 flop; burst = best of 10, sustained = back to back for 4 s"}


def run(a, b):
    for _ in range(3):
        a @ b
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); a @ b; e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 0
    t_end = time.time() + 4.0
    e0.record()
    while time.time() < t_end:
        for _ in range(20):
            a @ b
        n += 20
        torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    sus = e0.elapsed_time(e1) / n
    f = 2.0 * N ** 3
    return f / (best * 1e-3) / 1e12, f / (sus * 1e-3) / 1e12


for name, dt, tf32 in (("fp16", torch.float16, False), ("bf16", torch.bfloat16, False), ("tf32", torch.float32, True)):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    a = torch.randn(N, N, device=dev, dtype=dt)
    b = torch.randn(N, N, device=dev, dtype=dt)
    burst, sus = run(a, b)
    out[f"{name}_tflops"] = round(burst, 1)
    out[f"{name}_tflops_sustained"] = round(sus, 1)
    del a, b
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/measured_peaks_extra.json", "w"), indent=1)
print(json.dumps(out))
