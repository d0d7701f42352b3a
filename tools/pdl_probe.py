"""Does programmatic dependent launch survive torch's CUDA-graph capture?  Chain of small kernels, eager vs graph."""
import os, sys, json, ctypes as C
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vault_b200 import ops, _abi
dev = torch.device("cuda:0")
M, N, K = 1280, 768, 768
a = torch.randn(M, K, device=dev).to(torch.bfloat16); b = torch.randn(N, K, device=dev).to(torch.bfloat16)
outs = [torch.empty(M, N, device=dev, dtype=torch.bfloat16) for _ in range(2)]
x = torch.randn(M, N, device=dev); g = torch.ones(N, device=dev); be = torch.zeros(N, device=dev)
def chain(n=100):
    for i in range(n):
        ops.gemm(a, b, ops.EPI_PLAIN_BF16, out=outs[i & 1])
        ops.layernorm_fwd(x, g, be, 1e-5)
def timeit(fn, reps=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
chain(); torch.cuda.synchronize()
eager = timeit(chain)
gr = torch.cuda.CUDAGraph()
with torch.cuda.graph(gr):
    chain()
graph = timeit(gr.replay)
print(json.dumps(dict(pdl=os.environ.get("VAULT_B200_PDL", "1"), eager_ms=eager, graph_ms=graph, per_pair_us_graph=graph * 10)))
