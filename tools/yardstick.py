"""cuBLAS yardsticks on the box (never linked into the product): ZGEMM / CGEMM / DGEMM / SGEMM(+TF32) at 8192^3 and
4096^3 through torch.matmul, CUDA-event timed.  Used only to place the kernels' numbers."""
import json
import sys

import torch


def bench(dtype, n, tf32=False, iters=5):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    a = torch.randn(n, n, device="cuda", dtype=dtype)
    b = torch.randn(n, n, device="cuda", dtype=dtype)
    for _ in range(2):
        (a @ b)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        (a @ b)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    flops = (8.0 if dtype.is_complex else 2.0) * n ** 3
    return flops / ms / 1e9


if __name__ == "__main__":
    out = {}
    for name, dt, tf in [("zgemm", torch.complex128, False), ("dgemm", torch.float64, False),
                         ("cgemm_fp32", torch.complex64, False), ("cgemm_tf32", torch.complex64, True),
                         ("sgemm_fp32", torch.float32, False), ("sgemm_tf32", torch.float32, True)]:
        for n in (4096, 8192):
            try:
                out[f"{name}_{n}"] = round(bench(dt, n, tf), 2)
            except Exception as e:  # noqa
                out[f"{name}_{n}"] = str(e)[:80]
    print(json.dumps(out))
