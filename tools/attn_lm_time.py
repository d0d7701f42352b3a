"""Timing of the mma.sync attention kernels at the LM shape (B=32, T=40, 12 heads) with and without probability dropout."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vault_b200 import _abi
dev = torch.device("cuda:0"); lib = _abi.lib(); st = lambda: torch.cuda.current_stream().cuda_stream
for B, S, heads in ((32, 40, 12), (32, 128, 12)):
    H = heads * 64
    qkv = torch.randn(B * S, 3 * H, device=dev).to(torch.bfloat16); mask = torch.ones(B, S, dtype=torch.uint8, device=dev)
    ctx = torch.empty(B * S, H, device=dev, dtype=torch.bfloat16); lse = torch.empty(B, heads, S, device=dev)
    dctx = torch.randn(B * S, H, device=dev).to(torch.bfloat16); dqkv = torch.empty_like(qkv); delta = torch.empty(B, heads, S, device=dev)
    seed_dev = torch.zeros(1, dtype=torch.int64, device=dev)
    res = dict(case=f"B{B} S{S} h{heads}")
    for p in (0.0, 0.1):
        f = lambda: lib.vault_attn_fwd(qkv.data_ptr(), mask.data_ptr(), ctx.data_ptr(), lse.data_ptr(), B, S, heads, p, 7, seed_dev.data_ptr(), 3, st())
        g = lambda: lib.vault_attn_bwd(qkv.data_ptr(), mask.data_ptr(), ctx.data_ptr(), dctx.data_ptr(), lse.data_ptr(), delta.data_ptr(), dqkv.data_ptr(), B, S, heads, p, 7, seed_dev.data_ptr(), 3, st())
        for nm, fn in (("fwd", f), ("bwd", g)):
            gr = torch.cuda.CUDAGraph()
            fn(); torch.cuda.synchronize()
            with torch.cuda.graph(gr):
                for _ in range(10): fn()
            gr.replay(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5): gr.replay()
            e1.record(); torch.cuda.synchronize()
            res[f"{nm}_p{p}_us"] = round(e0.elapsed_time(e1) * 1e3 / 50, 2)
    print(json.dumps(res), flush=True)
