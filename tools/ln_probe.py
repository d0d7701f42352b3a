"""GPU probe: LayerNorm fwd/bwd vs torch fp32 + achieved GB/s."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vault_b200 import ops

dev = torch.device("cuda:0")
torch.manual_seed(0)
for rows, cols in ((5920, 768), (37, 128), (11808, 768), (1280, 768)):
    x = torch.randn(rows, cols, device=dev) * 2 + 0.5
    g = 1 + 0.1 * torch.randn(cols, device=dev); b = 0.1 * torch.randn(cols, device=dev)
    y16, y32, mean, rstd = ops.layernorm_fwd(x, g, b, 1e-12, want_bf16=True, want_f32=True)
    xr = x.clone().requires_grad_(True); gr = g.clone().requires_grad_(True); br = b.clone().requires_grad_(True)
    ref = torch.nn.functional.layer_norm(xr, (cols,), gr, br, 1e-12)
    e32 = (y32 - ref).abs().max().item(); e16 = (y16.float() - ref).abs().max().item()
    dy = torch.randn(rows, cols, device=dev); dres = torch.randn(rows, cols, device=dev)
    dy16 = dy.to(torch.bfloat16)
    ref.backward(dy16.float())
    dg = torch.zeros(cols, device=dev); db = torch.zeros(cols, device=dev)
    dx32, dx16 = ops.layernorm_bwd(None, dy16, x, mean, rstd, g, dres, dg, db)
    edx = (dx32 - (xr.grad + dres)).abs().max().item()
    edg = ((dg - gr.grad).abs().max() / gr.grad.abs().max()).item(); edb = ((db - br.grad).abs().max() / br.grad.abs().max()).item()
    # two-input dy path
    dg2 = torch.zeros(cols, device=dev); db2 = torch.zeros(cols, device=dev)
    dx32b, _ = ops.layernorm_bwd(dy - dy16.float(), dy16, x, mean, rstd, g, None, dg2, db2, want_bf16=False)
    xr.grad = None
    torch.nn.functional.layer_norm(xr, (cols,), g, b, 1e-12).backward(dy)
    edx2 = (dx32b - xr.grad).abs().max().item()
    print(json.dumps(dict(case=f"ln {rows}x{cols}", fwd_f32=e32, fwd_bf16=e16, dx=edx, dx_two_input=edx2, dgamma_rel=edg, dbeta_rel=edb,
                          ok=bool(e32 < 1e-4 and e16 < 5e-2 and edx < 1e-3 and edx2 < 1e-3 and edg < 1e-3 and edb < 1e-3))), flush=True)

rows, cols = 5920 * 8, 768  # > L2
x = torch.randn(rows, cols, device=dev); g = torch.ones(cols, device=dev); b = torch.zeros(cols, device=dev)
for _ in range(3): ops.layernorm_fwd(x, g, b, 1e-12)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for _ in range(20): y16, _, mean, rstd = ops.layernorm_fwd(x, g, b, 1e-12)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
print(json.dumps(dict(case="ln_fwd time", rows=rows, ms=ms, gbs=rows * cols * 6 / ms / 1e6)), flush=True)
dy16 = torch.randn(rows, cols, device=dev).to(torch.bfloat16); dres = torch.randn(rows, cols, device=dev)
dg = torch.zeros(cols, device=dev); db = torch.zeros(cols, device=dev)
for _ in range(3): ops.layernorm_bwd(None, dy16, x, mean, rstd, g, dres, dg, db)
torch.cuda.synchronize(); e0.record()
for _ in range(20): ops.layernorm_bwd(None, dy16, x, mean, rstd, g, dres, dg, db)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
print(json.dumps(dict(case="ln_bwd time", rows=rows, ms=ms, gbs=rows * cols * (2 + 4 + 4 + 4 + 2) / ms / 1e6)), flush=True)
