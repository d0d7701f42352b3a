"""bench.py's `hbm_kernels` measurement alone (LayerNorm forward / backward at the workload's own row count, AdamW at the model's parameter
count).  VAULT_B200_LIB selects the build.  usage: hbm_rates.py [rows] [n_params]"""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 11808
n = int(sys.argv[2]) if len(sys.argv) > 2 else 197_017_344
peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
r = bench.hbm_kernel_rates(torch, peaks, rows, n)
print(os.environ.get("VAULT_B200_LIB", "default").split("/")[-1], json.dumps({k: dict(gbs=round(v["achieved_gbs"]), frac=round(v["frac_of_measured_peak"], 3), us=round(v["us_per_launch"], 2)) for k, v in r.items() if isinstance(v, dict)}), flush=True)
