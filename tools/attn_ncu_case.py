"""One forward + backward attention call at a bench shape (for ncu captures).  usage: attn_ncu_case.py B S heads [impl]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vault_b200 import _abi
B, S, heads = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
impl = int(sys.argv[4]) if len(sys.argv) > 4 else 0
dev = torch.device("cuda:0"); lib = _abi.lib(); H = heads * 64
torch.manual_seed(0)
qkv = (torch.randn(B * S, 3 * H, device=dev) * 0.7).to(torch.bfloat16)
mask = torch.ones(B, S, dtype=torch.uint8, device=dev); mask[:, 20:30] = 0
dctx = (torch.randn(B * S, H, device=dev) * 0.5).to(torch.bfloat16)
ctx = torch.empty(B * S, H, device=dev, dtype=torch.bfloat16); lse = torch.empty(B, heads, S, device=dev)
dq = torch.empty_like(qkv); delta = torch.empty(B, heads, S, device=dev)
_abi.set_attn_impl(impl)
st = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    _abi.check(lib.vault_attn_fwd(qkv.data_ptr(), mask.data_ptr(), ctx.data_ptr(), lse.data_ptr(), B, S, heads, 0.0, 0, None, 0, st), "fwd")
    _abi.check(lib.vault_attn_bwd(qkv.data_ptr(), mask.data_ptr(), ctx.data_ptr(), dctx.data_ptr(), lse.data_ptr(), delta.data_ptr(), dq.data_ptr(), B, S, heads, 0.0, 0, None, 0, st), "bwd")
torch.cuda.synchronize()
