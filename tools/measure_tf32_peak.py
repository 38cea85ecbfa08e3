"""Measure the dense TF32 tensor throughput of this GPU with cuBLAS (torch.matmul, allow_tf32) at 8192^3:
burst = best of 10, sustained = back to back for 4 s -- the same recipe the driver uses for the BF16 figure in
MEASURED_PEAKS.json.  Writes profiles/r02_tf32_peak.json (tracked), which bench.py quotes tensor fractions against.
    gpurun -- python tools/measure_tf32_peak.py && cp gpurun_out/r02_tf32_peak.json profiles/
"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def measure(dtype, n=8192, allow_tf32=True):
    torch.backends.cuda.matmul.allow_tf32 = allow_tf32
    a = torch.randn(n, n, device="cuda", dtype=dtype)
    b = torch.randn(n, n, device="cuda", dtype=dtype)
    c = torch.empty(n, n, device="cuda", dtype=dtype)
    flop = 2.0 * n ** 3
    for _ in range(3):
        torch.matmul(a, b, out=c)
    torch.cuda.synchronize()
    best = 0.0
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b, out=c)
        e1.record()
        torch.cuda.synchronize()
        best = max(best, flop / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    reps = 0
    e0.record()
    while time.perf_counter() - t0 < 4.0:
        for _ in range(20):
            torch.matmul(a, b, out=c)
        reps += 20
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    return best, reps * flop / (e0.elapsed_time(e1) * 1e-3) / 1e12


def main():
    out = {"gpu_name": torch.cuda.get_device_name(0), "torch": torch.__version__,
           "how": "torch.matmul fp32 inputs with allow_tf32=True (cuBLAS TF32 tensor-core GEMM), 8192^3, 2*N^3 FLOP: best of 10 "
                  "(burst) and back to back for 4 s (sustained); bf16 by the same loop for comparison"}
    out["tf32_tflops"], out["tf32_tflops_sustained"] = measure(torch.float32, allow_tf32=True)
    out["bf16_tflops"], out["bf16_tflops_sustained"] = measure(torch.bfloat16)
    out["fp32_simt_tflops"], out["fp32_simt_tflops_sustained"] = measure(torch.float32, n=4096, allow_tf32=False)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "r02_tf32_peak.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    sys.exit(main())
