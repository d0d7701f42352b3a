"""Launch one tcgen05 GEMM shape a few times (for `ncu --set full -k regex:gemm_bf16_kernel`).  Usage: one_gemm.py M N K epi [a_mn b_mn [split_k block_n colsum]]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vault_b200 import ops
M, N, K, epi = (int(x) for x in sys.argv[1:5])
a_mn, b_mn = (bool(int(x)) for x in sys.argv[5:7]) if len(sys.argv) > 6 else (False, False)
split_k, block_n, colsum = (int(x) for x in sys.argv[7:10]) if len(sys.argv) > 9 else (1, 0, 0)
dev = torch.device("cuda:0")
a = torch.randn((K, M) if a_mn else (M, K), device=dev).to(torch.bfloat16)
b = torch.randn((K, N) if b_mn else (N, K), device=dev).to(torch.bfloat16)
kw = {}
if epi in (0, 1, 2, 6): kw["bias"] = torch.randn(N, device=dev)
if epi == 2: kw["resid"] = torch.randn(M, N, device=dev)
if epi == 4: kw["aux"] = torch.randn(M, N, device=dev).to(torch.bfloat16)
if epi == 1: kw["out2"] = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
if split_k > 1 or block_n: kw.update(split_k=split_k, block_n=block_n)
if colsum: kw["a_colsum"] = torch.zeros(M, device=dev)
if epi == 5: kw["out"] = torch.zeros(M, N, device=dev)
flush = torch.empty(256 * 1024 * 1024, device=dev, dtype=torch.uint8)
for _ in range(6):
    flush.zero_()  # evict L2 between launches
    ops.gemm(a, b, epi, a_mn=a_mn, b_mn=b_mn, **kw)
torch.cuda.synchronize()
