"""Print a one-screen summary of a bench.py JSON line (+ optional per-step dump): python tools/summarize_bench.py bench.json [steps.json]"""
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
r = d.get("roofline") or {}
print(f"{d['config']['workload']} value {d['value']:.2f} {d['unit']} ms/step {d['ms_per_step']:.1f} slices/step "
      f"{d['config'].get('slices_per_step_per_gpu')} e2e {d['e2e']['value']:.2f} roofline {r.get('kernel')} {r.get('achieved', 0):.0f} "
      f"{r.get('unit')} frac {r.get('frac', 0):.3f} clocks {d.get('clocks', {}).get('sm_mhz')} {d.get('clocks', {}).get('reasons')}")
for k, v in (r.get("kernels") or {}).items():
    print(f"  {k:12s} {v['ms']:9.1f} ms  {v['tflops']:7.1f} TF  {v['gbs']:7.0f} GB/s  {v['launches']} launches")
if len(sys.argv) > 2:
    st = json.load(open(sys.argv[2]))
    tot = sum(s["ms_avg"] for s in st if s["runs"] > 3)
    print(f"per-slice sum {tot:.2f} ms")
    for s in st:
        if s["ms_avg"] > 0.08 and s["runs"] > 3:
            print(f"  {s['step']:3d} {s['M']:9d} {s['N']:9d} {s['K']:5d} {s['kernel_name']:10s} {s['ms_avg']:.3f} ms "
                  f"{s['flops']/s['ms_avg']/1e9:6.1f} TF {s['bytes']/s['ms_avg']/1e6:5.0f} GB/s")
