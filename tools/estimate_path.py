"""Estimate the per-slice and whole-job device time of a committed/candidate path from a DRY plan (no GPU):
per step max(flops / rate, bytes / bandwidth) with the rates bench.py measured for each kernel class on a B200.
Usage: python tools/estimate_path.py sycamore53_m14 path.json [path2.json ...]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import tenet_jl_b200 as tb
from tools.make_paths import network

# measured (profiles/r1_summary.md): flop/s, byte/s, fixed launch cost
RATES = {"c64_tf32x3": (187e12, 3.0e12), "stem_tc": (190e12, 4.2e12), "stem": (40e12, 4.2e12), "stream": (50e12, 4.8e12),
         "generic": (25e12, 1.5e12), "splitk": (25e12, 1.5e12), "c128_dmma": (19e12, 3e12)}
LAUNCH = 4e-6


def estimate(name, fn, verbose=False):
    tn = network(name)
    path = tb.pathfinder.load_path(tn.inds("all"), fn)
    plan = tb.ContractionPlan(tn, path, dry=True)
    t_slice = t_hoist = 0.0
    by = {}
    fl = byts = 0.0
    for s in range(len(path.steps)):
        si = plan.step_info(s)
        r = RATES.get(si["kernel_name"], RATES["generic"])
        t = max(si["flops"] / r[0], si["bytes"] / r[1]) + LAUNCH
        if si["hoisted"]:
            t_hoist += t
        else:
            t_slice += t
            fl += si["flops"]; byts += si["bytes"]
            by[si["kernel_name"]] = by.get(si["kernel_name"], 0.0) + t
    n = plan.nslices
    info = plan.info if hasattr(plan, "info") else {}
    print(f"{os.path.basename(fn)}: slices {n} per-slice {t_slice*1e3:.2f} ms ({fl/1e12:.2f} TFLOP, {byts/1e9:.1f} GB, AI {fl/byts:.1f}) "
          f"total {t_slice*n + t_hoist:.1f} s  -> {fl*n/(t_slice*n+t_hoist)/1e12:.1f} TFLOP/s   " +
          " ".join(f"{k}={v*1e3:.2f}" for k, v in sorted(by.items())) + f"  {info}")
    return t_slice * n + t_hoist


if __name__ == "__main__":
    for fn in sys.argv[2:]:
        estimate(sys.argv[1], fn)
