"""LayerNorm fwd/bwd + colsum + AdamW in isolation at the BASELINE shapes (for ncu --set full and for CUDA-event GB/s)."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vault_b200 import ops, _abi
dev = torch.device("cuda:0")
rows, cols = int(sys.argv[1]) if len(sys.argv) > 1 else 5920, 768
x = torch.randn(rows, cols, device=dev); g = torch.ones(cols, device=dev); b = torch.zeros(cols, device=dev)
dy16 = torch.randn(rows, cols, device=dev).to(torch.bfloat16); dres = torch.randn(rows, cols, device=dev)
dg = torch.zeros(cols, device=dev); db = torch.zeros(cols, device=dev); dc = torch.zeros(cols, device=dev)
flush = torch.empty(512 * 1024 * 1024, device=dev, dtype=torch.uint8)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
def timed(fn, nbytes, name, reps=8):
    fn(); torch.cuda.synchronize()
    tot = 0.0
    for _ in range(reps):
        flush.zero_()                      # evict L2 so the operands come from HBM
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    ms = tot / reps
    print(json.dumps(dict(kernel=name, rows=rows, us=ms * 1e3, gbs=nbytes / ms / 1e6)), flush=True)
y16, _, mean, rstd = ops.layernorm_fwd(x, g, b, 1e-12)
timed(lambda: ops.layernorm_fwd(x, g, b, 1e-12), rows * cols * 6, "ln_fwd f32->bf16")
timed(lambda: ops.layernorm_bwd(None, dy16, x, mean, rstd, g, dres, dg, db, dcolsum=dc), rows * cols * 16, "ln_bwd bf16 dy + resid -> f32 + bf16")
timed(lambda: ops.layernorm_bwd(None, dy16, x, mean, rstd, g, dres, None, None), rows * cols * 16, "ln_bwd (no column sums)")
