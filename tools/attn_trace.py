"""Debug: phase stamps of the tcgen05 attention kernels (VAULT_B200_ATTN_TRACE=1).  usage: attn_trace.py"""
import os, sys
os.environ["VAULT_B200_ATTN_TRACE"] = "1"
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vault_b200 import _abi
dev = torch.device("cuda:0"); lib = _abi.lib(); st = lambda: torch.cuda.current_stream().cuda_stream
B, S, heads = 32, 185, 12
H = heads * 64
qkv = torch.randn(B * S, 3 * H, device=dev).to(torch.bfloat16); mask = torch.ones(B, S, dtype=torch.uint8, device=dev)
ctx = torch.empty(B * S, H, device=dev, dtype=torch.bfloat16); lse = torch.empty(B, heads, S, device=dev)
dctx = torch.randn(B * S, H, device=dev).to(torch.bfloat16); dqkv = torch.empty_like(qkv); delta = torch.empty(B, heads, S, device=dev)
for it in range(2):
    lib.vault_attn_fwd(qkv.data_ptr(), mask.data_ptr(), ctx.data_ptr(), lse.data_ptr(), B, S, heads, 0.0, 0, None, 0, st())
for it in range(2):
    lib.vault_attn_bwd(qkv.data_ptr(), mask.data_ptr(), ctx.data_ptr(), dctx.data_ptr(), lse.data_ptr(), delta.data_ptr(), dqkv.data_ptr(), B, S, heads, 0.0, 0, None, 0, st())
torch.cuda.synchronize()
