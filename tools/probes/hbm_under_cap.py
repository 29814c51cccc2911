"""Probe: what HBM bandwidth does a plain copy reach when it runs right behind tensor-heavy work, i.e. while the chip is
power-capped to ~1.3-1.4 GHz?  (The in-bench HBM-bound kernels run at 4.4-5.0 TB/s, the same launches alone under ncu at
5.7-6.0 TB/s with SM clocks of 1.58-1.77 GHz: is that a property of the kernels or of the capped chip?)

Prints one JSON line: copy bandwidth alone, and behind 5 / 20 / 50 ms of bf16 matmul, 20 repetitions each (CUDA events)."""
import json
import subprocess
import threading
import time

import torch


def main():
    dev = torch.device("cuda:0")
    n = 1 << 29                                   # 2 x 1 GiB of bf16: far beyond L2
    a = torch.empty(n, dtype=torch.bfloat16, device=dev).normal_()
    b = torch.empty_like(a)
    x = torch.randn(8192, 8192, dtype=torch.bfloat16, device=dev)
    y = torch.randn(8192, 8192, dtype=torch.bfloat16, device=dev)
    z = torch.empty_like(x)
    clocks = []
    stop = threading.Event()

    def sampler():
        while not stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-i", "0"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                clocks.append((time.time(), float(out[0]), float(out[1])))
            except Exception:
                pass
            time.sleep(0.05)

    th = threading.Thread(target=sampler, daemon=True)
    th.start()

    def run(nmm, reps=20, ncopy=1):
        ts = []
        for _ in range(3):
            b.copy_(a)
        torch.cuda.synchronize()
        t_begin = time.time()
        for _ in range(reps):
            for _ in range(nmm):
                torch.matmul(x, y, out=z)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(ncopy):
                b.copy_(a)
            e1.record()
            ts.append((e0, e1))
        torch.cuda.synchronize()
        t_end = time.time()
        ms = sorted(e0.elapsed_time(e1) / ncopy for e0, e1 in ts)
        mhz = [c[1] for c in clocks if t_begin <= c[0] <= t_end]
        return {"matmuls_before": nmm, "copies": ncopy, "gbs_median": 2 * n * 2 / ms[len(ms) // 2] / 1e6, "gbs_best": 2 * n * 2 / ms[0] / 1e6,
                "sm_mhz_median": sorted(mhz)[len(mhz) // 2] if mhz else None}

    res = [run(0), run(0, ncopy=4), run(8), run(30), run(60, reps=12), run(30, ncopy=4), run(0)]
    stop.set()
    print(json.dumps({"probe": "hbm_under_cap", "bytes_per_copy": 2 * n * 2, "results": res}))


if __name__ == "__main__":
    main()
