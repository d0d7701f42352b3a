"""Fit T = T_launch + tiles_per_cta * (T_tile + kblocks * T_kb) for the tcgen05 GEMM (BN=256, one m-block per SM)."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vault_b200 import ops
dev = torch.device("cuda:0")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
M = 148 * 128
for epi, en in ((ops.EPI_PLAIN_BF16, "plain_bf16"), (ops.EPI_STORE_F32, "store_f32")):
    for tiles in (1, 2, 4):
        for K in (64, 256, 768, 3072):
            N = 256 * tiles
            a = torch.randn(M, K, device=dev).to(torch.bfloat16); b = torch.randn(N, K, device=dev).to(torch.bfloat16)
            out = ops.gemm(a, b, epi, block_n=256)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for _ in range(10): ops.gemm(a, b, epi, block_n=256, out=out)
            g.replay(); torch.cuda.synchronize(); e0.record()
            for _ in range(5): g.replay()
            e1.record(); torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / 50
            print(json.dumps(dict(epi=en, tiles_per_cta=tiles, kblocks=K // 64, us=round(us, 2), tflops=round(2.0 * M * N * K / us / 1e6))), flush=True)
