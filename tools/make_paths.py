"""Generate the committed contraction paths bench.py uses (bench_paths/*.json), so that the benchmark workload is
fixed and path finding (an *input* of the hot path) stays outside the timed region.

    python tools/make_paths.py sycamore53_m14 [--trials N] [--target LOG2_ELEMS]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tenet_jl_b200 as tb  # noqa: E402


def network(name):
    if name in ("sycamore53_m14", "sycamore53_m14_v1", "sycamore53_m14_greedy"):
        tn, _ = tb.workloads.sycamore_amplitude_network(rows=9, cols=6, cycles=14, seed=53, dtype=np.complex64)
        return tn
    if name == "sycamore53_m10":
        tn, _ = tb.workloads.sycamore_amplitude_network(rows=9, cols=6, cycles=10, seed=53, dtype=np.complex64)
        return tn
    if name.startswith("regular3_n") and name.endswith("_d4"):
        n = int(name[len("regular3_n"):-len("_d4")])
        return tb.workloads.random_regular_network(n=n, bond=4, dtype=np.complex64, seed=0)
    if name == "peps6x6_d4":
        return tb.workloads.peps_norm_network(6, 6, D=4, p=2, dtype=np.complex128, seed=4)[0]
    raise SystemExit(f"unknown workload {name}")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("name")
    ap.add_argument("--trials", type=int, default=128)
    ap.add_argument("--target", type=float, default=27.0, help="log2 of the largest intermediate (elements)")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--minimize", default="flops")
    ap.add_argument("--method", default="hyper", choices=["hyper", "greedy"])
    ap.add_argument("--keep", type=int, default=8)
    ap.add_argument("--reconf-size", type=int, default=9)
    ap.add_argument("--slicing", default="greedy", choices=["greedy", "interleaved"])
    ap.add_argument("--prescreen", type=int, default=0, help="give this many greedy trees one reconfiguration round before keeping --keep")
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    tn = network(a.name)
    inputs = [t.inds for t in tn.tensors]
    sizes = tn.sizes()
    t0 = time.time()
    if a.method == "hyper":
        from tenet_jl_b200 import treeopt
        p = treeopt.hyper_search(inputs, sizes, (), ntrials=a.trials, seed=a.seed, target_log2_size=a.target,
                                 reconf_size=a.reconf_size, reconf_rounds=3, keep=a.keep, verbose=True, minimize=a.minimize,
                                 slicing=a.slicing, prescreen=a.prescreen)
    else:
        p = tb.pathfinder.search(inputs, sizes, (), ntrials=a.trials, seed=a.seed, target_log2_size=a.target,
                                 minimize=a.minimize)
    dt = time.time() - t0
    lab = {i: k for k, i in enumerate(tn.inds("all"))}
    out = {"workload": a.name, "ntensors": len(inputs), "steps": [list(s) for s in p.steps],
           "sliced": [lab[i] for i in p.sliced], "log2_macs_per_slice": p.log2_macs,
           "log2_max_size": p.log2_max_size, "nslices_log2": float(np.log2(p.nslices)),
           "search": {"trials": a.trials, "seed": a.seed, "seconds": round(dt, 1), **p.info}}
    os.makedirs(os.path.join(ROOT, "bench_paths"), exist_ok=True)
    fn = a.out or os.path.join(ROOT, "bench_paths", a.name + ".json")
    with open(fn, "w") as f:
        json.dump(out, f)
    print(f"{fn}: per-slice 2^{p.log2_macs:.2f} MACs, peak 2^{p.log2_max_size:.1f} elems, 2^{np.log2(p.nslices):.0f} slices, "
          f"total 2^{p.log2_macs + np.log2(p.nslices):.2f} MACs ({dt:.0f}s)")


if __name__ == "__main__":
    main()
