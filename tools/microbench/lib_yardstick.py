"""Library yardsticks on the GPU box (cuBLAS / cuSOLVER through torch, fp64).

NOT part of the product path: these numbers only size the FP64 roofline denominator
(MEASURED_PEAKS.json has no FP64 entry; SURVEY.md section 8d asks for "best of cuBLAS DGEMM,
a DMMA micro-benchmark and a DFMA micro-benchmark") and give an on-box library comparison
for potrf / trsm.  One JSON line per measurement.
"""
import json
import sys
import time

import torch


def timeit(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def main():
    dev = torch.device("cuda:0")
    out = []
    for n in (2048, 4096, 8192):
        a = torch.randn(n, n, dtype=torch.float64, device=dev)
        b = torch.randn(n, n, dtype=torch.float64, device=dev)
        ms = timeit(lambda: torch.matmul(a, b))
        out.append({"bench": "cublas_dgemm", "n": n, "ms": ms, "tflops": 2.0 * n ** 3 / ms * 1e-9})
        print(json.dumps(out[-1]), flush=True)
    # sustained DGEMM: back-to-back for ~3 s
    n = 8192
    a = torch.randn(n, n, dtype=torch.float64, device=dev)
    b = torch.randn(n, n, dtype=torch.float64, device=dev)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    reps = 80
    e0.record()
    for _ in range(reps):
        torch.matmul(a, b)
    e1.record()
    e1.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(json.dumps({"bench": "cublas_dgemm_sustained", "n": n, "ms": ms,
                      "tflops": 2.0 * n ** 3 / ms * 1e-9}), flush=True)
    del a, b
    for n, batch in ((5000, 1), (2000, 1), (2000, 32), (506, 256), (100, 2048)):
        x = torch.randn(batch, n, n, dtype=torch.float64, device=dev)
        spd = x @ x.transpose(1, 2) + n * torch.eye(n, dtype=torch.float64, device=dev)
        del x
        ms = timeit(lambda: torch.linalg.cholesky(spd), reps=3, warm=1)
        fl = batch * n ** 3 / 3.0
        print(json.dumps({"bench": "cusolver_potrf", "n": n, "batch": batch, "ms": ms,
                          "tflops": fl / ms * 1e-9}), flush=True)
        L = torch.linalg.cholesky(spd)
        rhs = torch.randn(batch, n, 512, dtype=torch.float64, device=dev)
        ms = timeit(lambda: torch.linalg.solve_triangular(L, rhs, upper=False), reps=3, warm=1)
        fl = batch * n * n * 512.0
        print(json.dumps({"bench": "cublas_trsm_512rhs", "n": n, "batch": batch, "ms": ms,
                          "tflops": fl / ms * 1e-9}), flush=True)
        del spd, L, rhs
    # copy bandwidth sanity
    nbytes = 1 << 31
    src = torch.empty(nbytes // 8, dtype=torch.float64, device=dev)
    dst = torch.empty_like(src)
    ms = timeit(lambda: dst.copy_(src))
    print(json.dumps({"bench": "copy", "bytes": nbytes, "ms": ms, "gbs": 2.0 * nbytes / ms * 1e-6}), flush=True)


if __name__ == "__main__":
    main()
