"""Summarise an `ncu --csv` launch list of one eager training step (tools/step_profile.py under
    ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv ...)
into per-kernel launch counts, time shares and DRAM bytes per launch, and write profiles/r02_ncu_traffic.json (what bench.py reports as
`roofline.traffic` / `hbm_kernels.*.traffic`).  usage: ncu_step_summary.py launches.csv workload out.json out.md"""
import csv, json, re, sys

src, workload, out_json, out_md = sys.argv[1:5]
rows = []
with open(src, newline="") as f:
    lines = [l for l in f if l.startswith('"')]
rd = csv.reader(lines)
hdr = next(rd)
ix = {n: i for i, n in enumerate(hdr)}
per = {}
for r in rd:
    if len(r) < len(hdr):
        continue
    key = (r[ix["ID"]], r[ix["Kernel Name"]])
    d = per.setdefault(key, {})
    v = r[ix["Metric Value"]].replace(",", "")
    try:
        val = float(v)
    except ValueError:
        continue
    unit = r[ix["Metric Unit"]]
    name = r[ix["Metric Name"]]
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
    d[name] = val * scale
fam = {}
def short(k):
    k = re.sub(r"\(.*", "", k)
    k = re.sub(r"^void ", "", k)
    k = k.replace("vb::", "").replace("<unnamed>::", "")
    return k
for (_id, kname), d in per.items():
    f = fam.setdefault(short(kname), dict(launches=0, us=0.0, rd=0.0, wr=0.0))
    f["launches"] += 1
    f["us"] += d.get("gpu__time_duration.sum", 0.0)
    f["rd"] += d.get("dram__bytes_read.sum", 0.0)
    f["wr"] += d.get("dram__bytes_write.sum", 0.0)
tot = sum(f["us"] for f in fam.values())
order = sorted(fam.items(), key=lambda kv: -kv[1]["us"])
md = [f"# ncu launch list of one eager training step, workload `{workload}` (cold-cache, serialised: compare SHARES, not absolutes)", "",
      "| kernel | launches | total us | share | DRAM read + write per launch (MB) |", "|---|---:|---:|---:|---:|"]
for k, f in order:
    md.append(f"| `{k}` | {f['launches']} | {f['us']:.0f} | {100 * f['us'] / tot:.1f} % | {(f['rd'] + f['wr']) / f['launches'] / 1e6:.2f} |")
md.append("")
md.append(f"total {tot / 1e3:.2f} ms over {sum(f['launches'] for f in fam.values())} launches")
open(out_md, "w").write("\n".join(md) + "\n")
def pick(pred):
    sel = [f for k, f in fam.items() if pred(k)]
    n = sum(f["launches"] for f in sel)
    return None if not n else dict(dram_bytes_per_launch=sum(f["rd"] + f["wr"] for f in sel) / n, launches=n, share_of_step=sum(f["us"] for f in sel) / tot,
                                   source=f"{out_md} (ncu dram__bytes_read.sum + dram__bytes_write.sum, average over every launch of the kernel in one eager {workload} step)")
res = dict(workload=workload, gemm=pick(lambda k: k.startswith("gemm_bf16_kernel") or k.startswith("gemm_wgrad_grouped_kernel")), layernorm_fwd=pick(lambda k: k.startswith("ln_fwd")),
           layernorm_bwd=pick(lambda k: k.startswith("ln_bwd")), adamw=pick(lambda k: k.startswith("adamw")),
           attention=pick(lambda k: k.startswith("attn_")))
json.dump(res, open(out_json, "w"), indent=1)
print(json.dumps(res, indent=1))
