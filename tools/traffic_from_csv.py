"""profiles/r2_traffic.json from `ncu --metrics ... --csv` launch tables (tools/experiments/r2_final_evidence.sh): for each
named kernel the LONGEST launch matching a regex is taken (the dominant step of the workload), its DRAM read+write bytes
are set against the algorithmic bytes of that step, and the file is stamped with the digest of the kernel sources the
library was built from (bench.py reports `roofline.traffic` only while that digest matches the running build).

    python tools/traffic_from_csv.py WORKLOAD:KERNEL:table.csv:REGEX:M:N:K:dtype_bytes [...] [--out profiles/r2_traffic.json]
"""
import collections
import csv
import importlib.util
import io
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    by = collections.OrderedDict()
    for x in csv.DictReader(io.StringIO("".join(lines))):
        d = by.setdefault(x["ID"], {"name": x["Kernel Name"]})
        d[x["Metric Name"]] = float(x["Metric Value"].replace(",", "")) if x["Metric Value"] not in ("", "n/a") else None
    return list(by.values())


def main():
    args = sys.argv[1:]
    out = os.path.join(ROOT, "profiles", "r2_traffic.json")
    if "--out" in args:
        i = args.index("--out"); out = args[i + 1]; del args[i:i + 2]
    spec = importlib.util.spec_from_file_location("_tnb_build", os.path.join(ROOT, "tenet.jl_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
    res = {"source_digest": mod.source_digest(),
           "note": "longest launch of each kernel in an `ncu --metrics` pass of one bench step on these sources",
           "workloads": {}}
    for a in args:
        wl, kern, path, rx, M, N, K, esz = a.split(":")
        M, N, K, esz = int(M), int(N), int(K), int(esz)
        cand = [l for l in launches(path) if re.search(rx, l["name"])]
        if not cand:
            raise SystemExit(f"no launch matches {rx!r} in {path}")
        l = max(cand, key=lambda l: l["gpu__time_duration.sum"])
        rd, wr = l["dram__bytes_read.sum"], l["dram__bytes_write.sum"]
        alg = esz * (M * K + N * K + M * N)
        res["workloads"].setdefault(wl, {})[kern] = {
            "launch": f"{M} x {N} x {K} ({l['name'][:70]})", "dram_bytes": rd + wr, "dram_read": rd, "dram_write": wr,
            "algorithmic_bytes": alg, "ratio": (rd + wr) / alg, "duration_s_under_ncu": l["gpu__time_duration.sum"] * 1e-9,
            "tensor_pipe_pct": l.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
            "dmma_pipe_pct": l.get("sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active"),
            "sm_clock_mhz": (l.get("sm__cycles_elapsed.avg.per_second") or 0) / 1e6,
            "report": os.path.basename(path)}
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
