"""Turn an `ncu --metrics ... --csv --log-file X.csv` launch table into a compact markdown table (one row per launch,
or aggregated per kernel with --by-kernel).  Usage: python tools/ncu_table.py X.csv [--by-kernel] [--min-us 50]"""
import csv
import re
import sys


def load(fn):
    rows = []
    with open(fn, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.reader(lines)
    hdr = next(rd)
    col = {n: i for i, n in enumerate(hdr)}
    for r in rd:
        if len(r) < len(hdr):
            continue
        rows.append(r)
    return col, rows


def short(name):
    name = re.sub(r"<unnamed>::", "", name)
    name = re.sub(r"\(.*$", "", name)
    return name.replace("void ", "")[:60]


def main():
    fn = sys.argv[1]
    by_kernel = "--by-kernel" in sys.argv
    min_us = float(sys.argv[sys.argv.index("--min-us") + 1]) if "--min-us" in sys.argv else 0.0
    col, rows = load(fn)
    # long format: one row per (launch id, metric)
    launches = {}
    for r in rows:
        lid = int(r[col["ID"]])
        d = launches.setdefault(lid, {"kernel": short(r[col["Kernel Name"]]), "grid": r[col.get("Grid Size", 0)] if "Grid Size" in col else ""})
        try:
            v = float(r[col["Metric Value"]].replace(",", ""))
        except ValueError:
            continue
        unit = r[col["Metric Unit"]]
        scale = {"msecond": 1e3, "usecond": 1.0, "nsecond": 1e-3, "second": 1e6, "ms": 1e3, "us": 1.0, "ns": 1e-3, "s": 1e6}.get(unit)
        if r[col["Metric Name"]] == "gpu__time_duration.sum" and scale:
            v *= scale                                     # -> microseconds
        bscale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}.get(unit)
        if r[col["Metric Name"]].startswith("dram__bytes") and bscale:
            v *= bscale
        d[r[col["Metric Name"]]] = v
    L = [d for _, d in sorted(launches.items())]
    tot = sum(d.get("gpu__time_duration.sum", 0.0) for d in L)
    if by_kernel:
        agg = {}
        for d in L:
            a = agg.setdefault(d["kernel"], {"n": 0, "us": 0.0})
            a["n"] += 1; a["us"] += d.get("gpu__time_duration.sum", 0.0)
        print(f"| kernel | launches | total us | share |\n|---|---|---|---|")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
            print(f"| `{k}` | {a['n']} | {a['us']:.0f} | {100 * a['us'] / max(tot, 1e-9):.1f} % |")
        print(f"\ntotal {tot / 1e3:.2f} ms over {len(L)} launches")
        return
    print("| # | kernel | grid | us | dram rd+wr GB | GB/s | tensor % | dmma % | dram % |\n|---|---|---|---|---|---|---|---|---|")
    for i, d in enumerate(L):
        us = d.get("gpu__time_duration.sum", 0.0)
        if us < min_us:
            continue
        by = d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
        print(f"| {i} | `{d['kernel']}` | {d.get('launch__grid_size', 0):.0f} | {us:.0f} | {by / 1e9:.3f} | {by / max(us, 1e-9) / 1e3:.0f} | "
              f"{d.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 0):.1f} | "
              f"{d.get('sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active', 0):.1f} | "
              f"{d.get('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 0):.1f} |")


if __name__ == "__main__":
    main()
