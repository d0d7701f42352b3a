"""Per-shape GEMM timing inside CUDA graphs (10 launches per graph, the config-3 shapes and epilogues).  VAULT_B200_LIB selects the build."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vault_b200 import ops
dev = torch.device("cuda:0")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
E = ops
MV, ML = int(os.environ.get("VB_M_VILT", "5920")), int(os.environ.get("VB_M_LM", "1280"))  # target shape: VB_M_VILT=11808 VB_M_LM=4096
cases = [("qkv_fwd", MV, 2304, 768, E.EPI_BIAS_BF16, 0), ("out_fwd_resid", MV, 768, 768, E.EPI_BIAS_RESID_F32, 0),
         ("mlp1_fwd_gelu", MV, 3072, 768, E.EPI_BIAS_GELU_BF16, 0), ("mlp2_fwd_resid", MV, 768, 3072, E.EPI_BIAS_RESID_F32, 0),
         ("mlp2_dgrad_dgelu", MV, 3072, 768, E.EPI_DGELU_BF16, 1), ("mlp1_fwd_gelu_grad", MV, 3072, 768, E.EPI_BIAS_GELU_GRAD_BF16, 0),
         ("mlp2_dgrad_mul", MV, 3072, 768, E.EPI_MUL_AUX_BF16, 1), ("lm_gelu_grad", ML, 3072, 768, E.EPI_BIAS_GELU_GRAD_BF16, 0), ("lm_dgrad_mul", ML, 3072, 768, E.EPI_MUL_AUX_BF16, 1), ("mlp1_dgrad", MV, 768, 3072, E.EPI_PLAIN_BF16, 1),
         ("qkv_dgrad", MV, 768, 2304, E.EPI_PLAIN_BF16, 1), ("lm_out_resid", ML, 768, 768, E.EPI_BIAS_RESID_F32, 0),
         ("lm_mlp1_gelu", ML, 3072, 768, E.EPI_BIAS_GELU_BF16, 0), ("lm_dgelu", ML, 3072, 768, E.EPI_DGELU_BF16, 1), ("lm_mlp2_resid", ML, 768, 3072, E.EPI_BIAS_RESID_F32, 0)]
res = {}
tot = 0.0
for name, M, N, K, epi, b_mn in cases:
    a = torch.randn(M, K, device=dev).to(torch.bfloat16)
    b = (torch.randn(K, N, device=dev) if b_mn else torch.randn(N, K, device=dev)).to(torch.bfloat16)
    kw = {}
    if epi in (E.EPI_BIAS_BF16, E.EPI_BIAS_GELU_BF16, E.EPI_BIAS_GELU_GRAD_BF16, E.EPI_BIAS_RESID_F32): kw["bias"] = torch.randn(N, device=dev)
    if epi == E.EPI_BIAS_RESID_F32: kw["resid"] = torch.randn(M, N, device=dev)
    if epi in (E.EPI_DGELU_BF16, E.EPI_MUL_AUX_BF16): kw["aux"] = torch.randn(M, N, device=dev).to(torch.bfloat16)
    if epi in (E.EPI_BIAS_GELU_BF16, E.EPI_BIAS_GELU_GRAD_BF16): kw["out2"] = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    out = ops.gemm(a, b, epi, b_mn=bool(b_mn), **kw)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(10): ops.gemm(a, b, epi, b_mn=bool(b_mn), out=out, **kw)
    g.replay(); torch.cuda.synchronize(); e0.record()
    for _ in range(5): g.replay()
    e1.record(); torch.cuda.synchronize()
    res[name] = round(e0.elapsed_time(e1) * 1e3 / 50, 2)
    tot += res[name]
res["sum"] = round(tot, 1)
print(os.environ.get("VAULT_B200_LIB", "default").split("/")[-1], json.dumps(res), flush=True)
