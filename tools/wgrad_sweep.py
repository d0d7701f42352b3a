"""Sweep tile-N x split-K for the weight-gradient GEMM shapes (both operands MN-major, contraction over tokens)."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vault_b200 import ops
dev = torch.device("cuda:0")
torch.manual_seed(0)
SHAPES = [("vilt qkv", 2304, 768, 5920), ("vilt mlp1", 3072, 768, 5920), ("vilt mlp2", 768, 3072, 5920), ("vilt out", 768, 768, 5920),
          ("lm qkv", 2304, 768, 1280), ("lm mlp1", 3072, 768, 1280), ("lm mlp2", 768, 3072, 1280), ("lm out", 768, 768, 1280), ("patch", 768, 3072, 4608)]
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for name, M, N, K in SHAPES:
    a = (torch.randn(K, M, device=dev) * 0.5).to(torch.bfloat16); b = (torch.randn(K, N, device=dev) * 0.5).to(torch.bfloat16)
    ref = a.float().t() @ b.float()
    rows = []
    for bn in (256, 128):
        for split in (1, 2, 3, 4, 8):
            epi = ops.EPI_STORE_F32 if split == 1 else ops.EPI_ATOMIC_F32
            out = torch.zeros(M, N, device=dev)
            ops.gemm(a, b, epi, a_mn=True, b_mn=True, block_n=bn, split_k=split, out=out)
            err = ((out - ref).abs().max() / ref.abs().max()).item()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for _ in range(10): ops.gemm(a, b, epi, a_mn=True, b_mn=True, block_n=bn, split_k=split, out=out)
            g.replay(); torch.cuda.synchronize(); e0.record()
            for _ in range(5): g.replay()
            e1.record(); torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / 50
            rows.append(dict(bn=bn, split=split, us=round(us, 2), tflops=round(2.0 * M * N * K / us / 1e6), ok=err < 1e-3))
    best = min(rows, key=lambda r: r["us"])
    cur = next(r for r in rows if r["bn"] == 128 and r["split"] == 1)
    print(json.dumps(dict(shape=name, mnk=[M, N, K], best=best, bn128_split1=cur, all=[(r["bn"], r["split"], r["us"]) for r in rows])), flush=True)
