"""Launch a layer's grouped weight-gradient GEMM a few times (for `ncu --set full -k regex:gemm_wgrad_grouped_kernel`).  usage: one_wgrad_group.py [tokens]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vault_b200 import ops
tokens = int(sys.argv[1]) if len(sys.argv) > 1 else 11808
dev = torch.device("cuda:0")
H, I = 768, 3072
probs = []
for n_out, k_in, bias in ((H, I, False), (I, H, True), (H, H, False), (3 * H, H, True)):
    dy = torch.randn(tokens, n_out, device=dev).to(torch.bfloat16); x = torch.randn(tokens, k_in, device=dev).to(torch.bfloat16)
    probs.append((dy, x, torch.zeros(n_out, k_in, device=dev), 2, torch.zeros(n_out, device=dev) if bias else None))
flush = torch.empty(256 * 1024 * 1024, device=dev, dtype=torch.uint8)
for _ in range(5):
    flush.zero_()
    ops.gemm_wgrad_grouped(probs)
torch.cuda.synchronize()
