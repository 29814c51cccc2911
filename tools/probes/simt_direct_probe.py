"""Probe: SIMT stem kernel, direct form vs staged form (TNB_STEM_SIMT_DIRECT=1/0, one process each), rows-fastest output.
Prints one JSON line per run: GB/s over algorithmic bytes (CUDA events on the context stream, 30 launches)."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import tenet_jl_b200 as tb  # noqa: E402


def main():
    ctx = tb.default_context()
    stream = torch.cuda.ExternalStream(ctx.stream) if ctx.stream else torch.cuda.current_stream()
    out = []
    for dt, nbig, N, K in [(np.complex128, 20, 6, 6), (np.complex64, 22, 8, 4), (np.complex64, 22, 4, 2), (np.complex128, 20, 16, 4)]:
        rng = np.random.default_rng(1)
        a = (rng.standard_normal((2,) * nbig + (K,)) + 0j).astype(dt)
        b = (rng.standard_normal((N, K)) + 0j).astype(dt)
        big = [f"m{i}" for i in range(nbig)]
        ta, tb_ = tb.Tensor(a, big + ["k"]), tb.Tensor(b, ["n", "k"])
        tn = tb.TensorNetwork([ta, tb_])
        o = big + ["n"]
        plan = tb.ContractionPlan(tn, tb.einexpr(tn, output=o), output=o, ctx=ctx)
        kern = plan.step_info(0)["kernel_name"]
        for _ in range(5):
            plan.execute()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 30
        with torch.cuda.stream(stream):
            e0.record()
        for _ in range(reps):
            plan.execute()
        with torch.cuda.stream(stream):
            e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / reps
        byts = np.dtype(dt).itemsize * (1 << nbig) * (K + N)
        out.append({"dtype": np.dtype(dt).name, "M": 1 << nbig, "N": N, "K": K, "kernel": kern, "us": round(us, 2), "gbs": round(byts / us / 1e3, 1)})
        plan.close()
    print(json.dumps({"simt_direct": os.environ.get("TNB_STEM_SIMT_DIRECT", "1"), "results": out}))


if __name__ == "__main__":
    sys.exit(main())
