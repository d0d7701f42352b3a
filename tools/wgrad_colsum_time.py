"""In-graph time of the weight-gradient GEMMs of the step with and without the fused bias gradient (a_colsum), next to the stand-alone
column-sum launch it replaces.  Shapes: target workload (11,808 ViLT / 4,096 LM tokens)."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vault_b200 import ops, _abi
dev = torch.device("cuda:0")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
cases = [("vilt_qkv", 2304, 768, 11808, 2, 192), ("vilt_w1", 3072, 768, 11808, 2, 256), ("lm_qkv", 2304, 768, 4096, 2, 192), ("lm_w1", 3072, 768, 4096, 2, 256)]
def timed(fn):
    fn(); g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(10): fn()
    g.replay(); torch.cuda.synchronize(); e0.record()
    for _ in range(5): g.replay()
    e1.record(); torch.cuda.synchronize()
    return round(e0.elapsed_time(e1) * 1e3 / 50, 2)
for name, n_out, k_in, tokens, split, bn in cases:
    dy = torch.randn(tokens, n_out, device=dev).to(torch.bfloat16); x = torch.randn(tokens, k_in, device=dev).to(torch.bfloat16)
    gw = torch.zeros(n_out, k_in, device=dev); gb = torch.zeros(n_out, device=dev)
    st = lambda: torch.cuda.current_stream().cuda_stream
    res = dict(shape=name,
               plain=timed(lambda: ops.gemm(dy, x, ops.EPI_ATOMIC_F32, a_mn=True, b_mn=True, split_k=split, out=gw, block_n=bn)),
               fused=timed(lambda: ops.gemm(dy, x, ops.EPI_ATOMIC_F32, a_mn=True, b_mn=True, split_k=split, out=gw, block_n=bn, a_colsum=gb)),
               colsum=timed(lambda: _abi.call("vault_colsum_bf16", dy.data_ptr(), n_out, gb.data_ptr(), tokens, n_out, st())))
    print(json.dumps(res), flush=True)
