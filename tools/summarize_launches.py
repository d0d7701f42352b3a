"""ncu `--metrics gpu__time_duration.sum --csv` launch list -> markdown table of kernel families (launches, total us, share).
usage: summarize_launches.py launches.csv > summary.md"""
import csv, re, sys
rows, started = [], False
with open(sys.argv[1], newline="") as f:
    for line in csv.reader(f):
        if line and line[0] == "ID":
            started, hdr = True, line
            continue
        if started and len(line) == len(hdr):
            rows.append(dict(zip(hdr, line)))
fam = {}
for r in rows:
    if r["Metric Name"] != "gpu__time_duration.sum":
        continue
    n = r["Kernel Name"]
    n = re.sub(r"^void ", "", n)
    n = re.sub(r"\(.*$", "", n)
    n = re.sub(r"^vb::\(anonymous namespace\)::", "vb::", n)
    ns = float(r["Metric Value"]) * (1e3 if r["Metric Unit"] == "us" else 1.0)
    a = fam.setdefault(n, [0, 0.0])
    a[0] += 1
    a[1] += ns
tot = sum(v[1] for v in fam.values())
print("| kernel | launches | total us | share |\n|---|---:|---:|---:|")
for n, (c, ns) in sorted(fam.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{n[:110]}` | {c} | {ns / 1e3:.1f} | {100 * ns / tot:.1f}% |")
print(f"\n{sum(v[0] for v in fam.values())} launches, {tot / 1e6:.3f} ms serialised")
