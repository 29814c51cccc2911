"""profiles/r2_traffic.json from `ncu --set full` reports: DRAM read+write bytes of the named launches next to their
algorithmic bytes, stamped with the digest of the kernel sources the library was built from (bench.py reports the traffic
only while that digest matches the running build).

    python tools/make_traffic.py WORKLOAD:KERNEL:report.ncu-rep:launch_index:M:N:K:dtype_bytes [...] --out profiles/r2_traffic.json
"""
import csv
import importlib.util
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def raw_rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h = rows[0]
    return h, rows[2:]


def main():
    args = sys.argv[1:]
    out = os.path.join(ROOT, "profiles", "r2_traffic.json")
    if "--out" in args:
        i = args.index("--out"); out = args[i + 1]; del args[i:i + 2]
    spec = importlib.util.spec_from_file_location("_tnb_build", os.path.join(ROOT, "tenet.jl_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
    res = {"source_digest": mod.source_digest(), "workloads": {}}
    for a in args:
        wl, kern, rep, idx, M, N, K, esz = a.split(":")
        h, rows = raw_rows(rep)
        r = rows[int(idx)]
        col = {n: i for i, n in enumerate(h)}
        units = None
        rd, wr = float(r[col["dram__bytes_read.sum"]]), float(r[col["dram__bytes_write.sum"]])
        # ncu prints these in Gbyte / Mbyte depending on magnitude: the unit row is rows[1] of the csv
        out_rows = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
        urow = out_rows[1]
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}
        rd *= scale.get(urow[col["dram__bytes_read.sum"]], 1.0)
        wr *= scale.get(urow[col["dram__bytes_write.sum"]], 1.0)
        M, N, K, esz = int(M), int(N), int(K), int(esz)
        alg = esz * (M * K + N * K + M * N)
        dur = float(r[col["gpu__time_duration.sum"]]) * {"msecond": 1e-3, "usecond": 1e-6, "second": 1.0, "nsecond": 1e-9}.get(urow[col["gpu__time_duration.sum"]], 1e-3)
        res["workloads"].setdefault(wl, {})[kern] = {
            "launch": f"{M} x {N} x {K} ({r[col['Kernel Name']][:60]})", "dram_bytes": rd + wr, "dram_read": rd, "dram_write": wr,
            "algorithmic_bytes": alg, "ratio": (rd + wr) / alg, "duration_s_under_ncu": dur,
            "tensor_pipe_pct": float(r[col["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]]) if "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active" in col else None,
            "report": os.path.basename(rep)}
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
