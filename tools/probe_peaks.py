"""Measures the FP64 roofline denominators that MEASURED_PEAKS.json lacks (cuBLAS DGEMM peak =
the DMMA pipe ceiling; a plain FP64 copy for HBM) plus host facts for the CPU baseline.
Run on the GPU box:  python tools/probe_peaks.py > gpurun_out/fp64_peaks.json"""
import json
import os
import time

import torch


def timed(fn, reps):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e30
    for _ in range(reps):
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def main():
    out = {"gpu": torch.cuda.get_device_name(0), "host_cores": os.cpu_count()}
    try:
        out["host_mem_gb"] = round(os.sysconf("SC_PAGE_SIZE") * os.sysconf("SC_PHYS_PAGES") / 2 ** 30, 1)
    except Exception:
        pass
    n = 8192
    a = torch.randn(n, n, dtype=torch.float64, device="cuda")
    b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    c = torch.empty_like(a)
    for _ in range(3):
        torch.matmul(a, b, out=c)
    ms = timed(lambda: torch.matmul(a, b, out=c), 8)
    out["dgemm_8192_tflops_burst"] = round(2 * n ** 3 / ms / 1e9, 2)
    t0 = time.time()
    k = 0
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    while time.time() - t0 < 3.0:
        torch.matmul(a, b, out=c)
        k += 1
    e1.record()
    torch.cuda.synchronize()
    out["dgemm_8192_tflops_sustained"] = round(2 * n ** 3 * k / e0.elapsed_time(e1) / 1e9, 2)
    # tall-skinny shapes of the hot path through cuBLAS, as a library yardstick
    nn = 1 << 22
    v = torch.randn(111, nn, dtype=torch.float64, device="cuda").t()   # column-major n x 111
    w = torch.randn(111, nn, dtype=torch.float64, device="cuda").t()
    ms = timed(lambda: torch.matmul(v.t(), w), 5)
    out["cublas_gram_111_n4M_ms"] = round(ms, 3)
    out["cublas_gram_111_tflops"] = round(2 * nn * 111 * 111 / ms / 1e9, 2)
    out["cublas_gram_111_gbs"] = round(8 * nn * 222 / ms / 1e6, 1)
    cm = torch.randn(111, 37, dtype=torch.float64, device="cuda")
    ms = timed(lambda: torch.matmul(v, cm), 5)
    out["cublas_blockmul_111x37_n4M_ms"] = round(ms, 3)
    out["cublas_blockmul_gbs"] = round(8 * nn * (111 + 37) / ms / 1e6, 1)
    x = torch.empty(1 << 28, dtype=torch.float64, device="cuda")
    y = torch.empty_like(x)
    ms = timed(lambda: y.copy_(x), 8)
    out["copy_f64_gbs"] = round(2 * x.numel() * 8 / ms / 1e6, 1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
