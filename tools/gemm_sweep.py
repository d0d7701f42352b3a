"""Sweep tile-N x cluster mode for every GEMM shape of the training step; checks each variant against the plain one."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vault_b200 import ops
dev = torch.device("cuda:0")
torch.manual_seed(0)
SHAPES = [  # (name, M, N, K, a_mn, b_mn, epi)
    ("vilt qkv", 5920, 2304, 768, 0, 0, ops.EPI_PLAIN_BF16), ("vilt out", 5920, 768, 768, 0, 0, ops.EPI_PLAIN_BF16),
    ("vilt mlp1", 5920, 3072, 768, 0, 0, ops.EPI_PLAIN_BF16), ("vilt mlp2", 5920, 768, 3072, 0, 0, ops.EPI_PLAIN_BF16),
    ("vilt dgrad mlp2", 5920, 3072, 768, 0, 1, ops.EPI_PLAIN_BF16), ("vilt dgrad mlp1", 5920, 768, 3072, 0, 1, ops.EPI_PLAIN_BF16),
    ("vilt dgrad qkv", 5920, 768, 2304, 0, 1, ops.EPI_PLAIN_BF16), ("vilt dgrad out", 5920, 768, 768, 0, 1, ops.EPI_PLAIN_BF16),
    ("vilt wgrad qkv", 2304, 768, 5920, 1, 1, ops.EPI_STORE_F32), ("vilt wgrad mlp1", 3072, 768, 5920, 1, 1, ops.EPI_STORE_F32),
    ("vilt wgrad mlp2", 768, 3072, 5920, 1, 1, ops.EPI_STORE_F32), ("vilt wgrad out", 768, 768, 5920, 1, 1, ops.EPI_STORE_F32),
    ("lm qkv", 1280, 2304, 768, 0, 0, ops.EPI_PLAIN_BF16), ("lm out", 1280, 768, 768, 0, 0, ops.EPI_PLAIN_BF16),
    ("lm mlp1", 1280, 3072, 768, 0, 0, ops.EPI_PLAIN_BF16), ("lm mlp2", 1280, 768, 3072, 0, 0, ops.EPI_PLAIN_BF16),
    ("lm dgrad mlp2", 1280, 3072, 768, 0, 1, ops.EPI_PLAIN_BF16), ("lm dgrad mlp1", 1280, 768, 3072, 0, 1, ops.EPI_PLAIN_BF16),
    ("lm dgrad qkv", 1280, 768, 2304, 0, 1, ops.EPI_PLAIN_BF16), ("lm wgrad mlp1", 3072, 768, 1280, 1, 1, ops.EPI_STORE_F32),
    ("lm wgrad qkv", 2304, 768, 1280, 1, 1, ops.EPI_STORE_F32),
]
only = sys.argv[1:] 
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for name, M, N, K, a_mn, b_mn, epi in SHAPES:
    if only and not any(o in name for o in only): continue
    a = (torch.randn((K, M) if a_mn else (M, K), device=dev) * 0.5).to(torch.bfloat16)
    b = (torch.randn((K, N) if b_mn else (N, K), device=dev) * 0.5).to(torch.bfloat16)
    ref = ops.gemm(a, b, epi, a_mn=bool(a_mn), b_mn=bool(b_mn), block_n=128).float()
    rows = []
    for cl in (0, 1, 2):
        for bn in (256, 128, 64):
            if N < bn: continue
            try:
                out = ops.gemm(a, b, epi, a_mn=bool(a_mn), b_mn=bool(b_mn), block_n=bn, cluster=cl)
                torch.cuda.synchronize()
                err = ((out.float() - ref).abs().max() / ref.abs().max()).item()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    for _ in range(10): ops.gemm(a, b, epi, a_mn=bool(a_mn), b_mn=bool(b_mn), block_n=bn, cluster=cl, out=out)
                g.replay(); torch.cuda.synchronize(); e0.record()
                for _ in range(5): g.replay()
                e1.record(); torch.cuda.synchronize()
                us = e0.elapsed_time(e1) * 1e3 / 50
                rows.append(dict(cl=cl, bn=bn, us=round(us, 2), tflops=round(2.0 * M * N * K / us / 1e6), err=float(f"{err:.1e}")))
            except Exception as ex:
                rows.append(dict(cl=cl, bn=bn, error=repr(ex)[:120]))
                if "CUDA" in repr(ex): print(json.dumps(dict(shape=name, rows=rows))); sys.exit(1)
    best = min((r for r in rows if "us" in r and r["err"] < 1e-2), key=lambda r: r["us"])
    base = min((r for r in rows if "us" in r and r["cl"] == 0), key=lambda r: r["us"])
    print(json.dumps(dict(shape=name, mnk=[M, N, K], best=best, best_nocluster=base, all=rows)), flush=True)
