"""Read `ncu -i X.ncu-rep --page source --csv` and print, for launch #n, the SASS instructions with the most
warp-stall samples (and landmark instructions), so warp roles can be told apart.  Usage: ncu_stalls.py rep n [min]"""
import csv, subprocess, sys

def load(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    kernels, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": None, "rows": []}
            kernels.append(cur)
        elif cur is not None and cur["hdr"] is None:
            cur["hdr"] = r
        elif cur is not None:
            cur["rows"].append(r)
    return kernels

if __name__ == "__main__":
    ks = load(sys.argv[1])
    n = int(sys.argv[2]); mn = int(sys.argv[3]) if len(sys.argv) > 3 else 100
    k = ks[n]; h = k["hdr"]
    si, src = h.index("# Samples"), h.index("Source")
    stall = [i for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
    data = [r for r in k["rows"] if len(r) > si]
    tot = sum(int(r[si] or 0) for r in data)
    print(k["name"], "launch", n, "of", len(ks), "samples", tot, "instrs", len(data))
    marks = ("UTC", "LDTM", "UBLKCP", "BAR", "SYNCS", "FENCE", "EXIT", "USETMAXREG", "MEMBAR")
    for i, r in enumerate(data):
        s = r[src].strip(); c = int(r[si] or 0)
        if c >= mn or any(m in s for m in marks):
            st = {h[j][6:]: int(r[j]) for j in stall if r[j] and int(r[j]) > 0}
            st = dict(sorted(st.items(), key=lambda kv: -kv[1])[:3])
            print(f"{i:5d} {c:6d} {100*c/tot:5.1f}%  {s[:64]:64s} {st if c >= mn else ''}")
    if len(sys.argv) > 4:
        bounds = [int(x) for x in sys.argv[4].split(",")]
        for a, b in zip(bounds[:-1], bounds[1:]):
            c = sum(int(r[si] or 0) for r in data[a:b])
            agg = {}
            for r in data[a:b]:
                for j in stall:
                    if r[j] and int(r[j]) > 0: agg[h[j][6:]] = agg.get(h[j][6:], 0) + int(r[j])
            print(f"region {a}-{b}: {c} samples ({100*c/tot:.1f}%)", dict(sorted(agg.items(), key=lambda kv: -kv[1])[:5]))
