"""In-graph time of each kernel family of ONE training step: record every C-ABI call of a step, then capture one CUDA graph per
family (same arguments, same buffers) and time its replay with CUDA events."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from transformers import BertConfig, ViltConfig
from vault_b200 import VaultForTMSC, VaultTrainStep, _abi
import bench

dev = torch.device("cuda:0")
torch.manual_seed(0)
B = int(os.environ.get("VB_BATCH", "32"))
W = bench.WORKLOADS[os.environ.get("VB_WORKLOAD", "config3")]
m = VaultForTMSC(ViltConfig(), n_classes=3, vilt_dropout_prob=0.1, bert_config=BertConfig()).to(dev).train()
ts = VaultTrainStep(m, lr=2e-5, use_cuda_graph=False)
batch = {k: v.to(dev) for k, v in bench.synth_batch(torch, B, W["text_len"], tuple(W["image"]), 30522, 3, seed=1, pin=False).items()}
for _ in range(2):
    ts.step(batch)
torch.cuda.synchronize()
counter = _abi.install_counter()
slot = ts._states[next(iter(ts._states))][0]
keep = {}
ob = ts.engine.backward_iter
def bk(tape, a, b, segments=True):
    keep.update(tape.t)
    return ob(tape, a, b, segments=segments)
ts.engine.backward_iter = bk
# keep EVERY tensor allocated during the recorded step alive: the replay reuses the recorded pointers
_alive = []
_orig = {n: getattr(torch, n) for n in ("empty", "zeros", "empty_like", "zeros_like")}
def _wrap(fn):
    def w(*a, **k):
        t = fn(*a, **k); _alive.append(t); return t
    return w
for n, fn in _orig.items(): setattr(torch, n, _wrap(fn))
ts._body(slot)
for n, fn in _orig.items(): setattr(torch, n, fn)
ts.engine.backward_iter = ob
torch.cuda.synchronize()
trace = list(counter.trace)
_abi.uninstall_counter()

fams = {
    "gemm_vilt": lambda n, a: n == "vault_gemm_bf16" and max(a[0].M, a[0].K if a[0].a_mn else 0) >= 4500,
    "gemm_lm": lambda n, a: n == "vault_gemm_bf16" and max(a[0].M, a[0].K if a[0].a_mn else 0) < 4500,
    "gemm_wgrad_grouped_vilt": lambda n, a: n == "vault_gemm_wgrad_grouped" and a[0][0].K >= 4500,
    "gemm_wgrad_grouped_lm": lambda n, a: n == "vault_gemm_wgrad_grouped" and a[0][0].K < 4500,
    "attn_fwd_vilt": lambda n, a: n == "vault_attn_fwd" and int(a[5]) != W["text_len"],
    "attn_bwd_vilt": lambda n, a: n == "vault_attn_bwd" and int(a[8]) != W["text_len"],
    "attn_fwd_lm": lambda n, a: n == "vault_attn_fwd" and int(a[5]) == W["text_len"],
    "attn_bwd_lm": lambda n, a: n == "vault_attn_bwd" and int(a[8]) == W["text_len"],
    "ln_fwd": lambda n, a: n == "vault_layernorm_fwd_drop",
    "ln_bwd": lambda n, a: n == "vault_layernorm_bwd_drop",
    "colsum": lambda n, a: n == "vault_colsum_bf16",
    "other": lambda n, a: n not in ("vault_gemm_bf16", "vault_gemm_wgrad_grouped", "vault_attn_fwd", "vault_attn_bwd", "vault_layernorm_fwd_drop", "vault_layernorm_bwd_drop", "vault_colsum_bf16"),
    "all": lambda n, a: True,
}
res = {}
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for fam, pred in fams.items():
    sub = [(n, a) for n, a in trace if pred(n, a)]
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        _abi.replay_trace(sub, side.cuda_stream)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        _abi.replay_trace(sub, torch.cuda.current_stream().cuda_stream)
    for _ in range(3): g.replay()
    torch.cuda.synchronize(); e0.record()
    for _ in range(10): g.replay()
    e1.record(); torch.cuda.synchronize()
    res[fam] = dict(calls=len(sub), ms=e0.elapsed_time(e1) / 10)
# AdamW alone
ts.engine.adamw_step(1e-5)
torch.cuda.synchronize(); e0.record()
for _ in range(5): ts.engine.adamw_step(1e-5)
e1.record(); torch.cuda.synchronize()
res["adamw"] = dict(calls=1, ms=e0.elapsed_time(e1) / 5)
for k, v in res.items():
    print(json.dumps(dict(family=k, **v)))
