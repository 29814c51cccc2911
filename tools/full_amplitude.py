"""Contract ALL slices of one or more committed paths of the same network on cuda:0 and print the amplitudes: two
different contraction trees / slicings must give the same number (a full-size parity property: no oracle can run
2^50 MACs).  Usage: python tools/full_amplitude.py sycamore53_m14 sycamore53_m14_v1 [--dtype c128] [--out file.json]
(--dtype c128 runs the complex64 network on the FP64 kernels: the complex128 truth of the amplitude, ~100 s per path.)"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
import tenet_jl_b200 as tb

def main():
    args = sys.argv[1:]
    out = ""
    if "--out" in args:
        i = args.index("--out")
        out = args[i + 1]
        del args[i:i + 2]
    dtype = None
    if "--dtype" in args:
        i = args.index("--dtype")
        dtype = {"c64": np.complex64, "c128": np.complex128}[args[i + 1]]
        del args[i:i + 2]
    names = args
    ctx = tb.default_context(0)
    res = {}
    for name in names:
        tn, path = bench.build_workload(tb, name)
        plan = tb.ContractionPlan(tn, path, ctx=ctx, dtype=dtype)
        plan.zero_output()
        ctx.sync()
        t0 = time.perf_counter()
        plan.execute(0, 1, plan.nslices, accumulate=True)
        ctx.sync()
        dt = time.perf_counter() - t0
        amp = complex(np.asarray(plan.result().parent).reshape(-1)[0])
        info = plan.info
        res[name] = {"amplitude": [amp.real, amp.imag], "seconds": dt, "nslices": plan.nslices, "dtype": str(plan.dtype),
                     "tflops": info["flops_per_slice"] * plan.nslices / dt / 1e12}
        print(f"{name}: amplitude {amp.real:+.9e} {amp.imag:+.9e}j  |a|^2 * 2^53 = {abs(amp) ** 2 * 2.0 ** 53:.4f}  "
              f"{plan.nslices} slices in {dt:.2f} s  ({res[name]['tflops']:.1f} TFLOP/s)", flush=True)
        plan.close()
    if len(names) > 1:
        a0 = complex(*res[names[0]]["amplitude"])
        for n in names[1:]:
            a1 = complex(*res[n]["amplitude"])
            rel = abs(a1 - a0) / max(abs(a0), 1e-300)
            res[n]["rel_diff_vs_" + names[0]] = rel
            print(f"relative difference {n} vs {names[0]}: {rel:.3e}")
    if out:
        json.dump(res, open(out, "w"), indent=1)

if __name__ == "__main__":
    main()
