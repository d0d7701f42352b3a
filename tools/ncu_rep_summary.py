"""Markdown table of the headline metrics of `ncu --set full` captures (read here with `ncu -i x.ncu-rep --page raw --csv`).
usage: ncu_rep_summary.py title=path.ncu-rep [title=path.ncu-rep ...] > profiles/rNN_ncu_full.md"""
import csv
import io
import subprocess
import sys

COLS = [
    ("gpu__time_duration.sum", "duration"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor instr"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("launch__registers_per_thread", "regs/thread"),
    ("sm__inst_issued.avg.pct_of_peak_sustained_active", "issue active %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "MUFU pipe %"),
    ("launch__grid_size", "grid"),
]


def rows_of(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = list(csv.reader(io.StringIO(out)))
    if len(rd) < 3:
        return [], []
    hdr, units = rd[0], rd[1]
    return hdr, [dict(zip(hdr, r)) for r in rd[2:]], dict(zip(hdr, units))


print("# `ncu --set full --clock-control none --import-source on` captures (read back with `ncu -i ... --page raw --csv`; tools/ncu_rep_summary.py)\n")
for arg in sys.argv[1:]:
    title, path = arg.split("=", 1)
    got = rows_of(path)
    if not got[0]:
        print(f"### {title}\n\n`{path}`: no launches captured\n")
        continue
    hdr, rows, units = got
    have = [(m, t) for m, t in COLS if m in hdr]
    print(f"### {title}\n\n`{path.split('/')[-1]}`\n")
    print("| kernel | " + " | ".join(t for _, t in have) + " |")
    print("|---|" + "---|" * len(have))
    for r in rows:
        name = r.get("Kernel Name", "?")[:70]
        cells = []
        for m, _ in have:
            v, u = r.get(m, ""), units.get(m, "")
            try:
                v = f"{float(v.replace(',', '')):.4g}"
            except ValueError:
                pass
            cells.append(f"{v} {u}".strip())
        print(f"| `{name}` | " + " | ".join(cells) + " |")
    print()
