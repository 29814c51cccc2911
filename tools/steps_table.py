"""Markdown per-step table from a `bench.py --dump-steps` file:  python tools/steps_table.py steps.json "title" [direct_steps] > table.md"""
import json
import sys

st = [s for s in json.load(open(sys.argv[1])) if s["runs"] > 3]
title = sys.argv[2] if len(sys.argv) > 2 else "per-step device time"
direct = set(int(x) for x in sys.argv[3].split(",")) if len(sys.argv) > 3 else set()
tot = sum(s["ms_avg"] for s in st)
print(f"# {title}\n")
print(f"total {tot:.2f} ms/slice over {len(st)} slice-dependent steps (CUDA events on the context stream in bench.py's profiled region, "
      f"`--dump-steps`; `sw_power_cap` active).")
print("bytes = 8*(|A|+|B|+|C|), flops = 8*M*N*K (algorithmic, SURVEY §8d).  `stem_tc` with N or M = 128·n and K ≥ 64 runs on the "
      "CTA-pair kernel, `c64_tf32x3` on the CTA-pair GEMM kernel; epilogue `direct` = rows stored straight from registers, "
      "`staged` = sorted-pattern staging tile (DESIGN.md §4.3).\n")
print("| step | kernel | epilogue | M | N | K | ms | share | TFLOP/s | GB/s | AI flop/B |")
print("|---|---|---|---|---|---|---|---|---|---|---|")
for s in sorted(st, key=lambda s: -s["ms_avg"]):
    if s["ms_avg"] < 0.015:
        continue
    ep = ("direct" if s["step"] in direct else "staged") if s["kernel_name"] == "stem_tc" else ""
    print(f"| {s['step']} | {s['kernel_name']} | {ep} | {s['M']} | {s['N']} | {s['K']} | {s['ms_avg']:.3f} | {100 * s['ms_avg'] / tot:.1f}% | "
          f"{s['flops'] / s['ms_avg'] / 1e9:.1f} | {s['bytes'] / s['ms_avg'] / 1e6:.0f} | {s['flops'] / s['bytes']:.1f} |")
