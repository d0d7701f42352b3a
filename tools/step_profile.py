"""One eager (no CUDA graph) training step of the BASELINE workload between cudaProfilerStart/Stop, for
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/step_profile.py
and, without ncu, a per-configuration timing table of every tcgen05 GEMM launch of the step (--gemms)."""
import ctypes as C, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from transformers import BertConfig, ViltConfig
from vault_b200 import VaultForTMSC, VaultTrainStep, _abi
import bench

dev = torch.device("cuda:0")
torch.manual_seed(0)
B = int(os.environ.get("VB_BATCH", "32"))
m = VaultForTMSC(ViltConfig(), n_classes=3, vilt_dropout_prob=0.1, bert_config=BertConfig()).to(dev).train()
ts = VaultTrainStep(m, lr=2e-5, use_cuda_graph=False)
W = bench.WORKLOADS[os.environ.get("VB_WORKLOAD", "config3")]
batch = {k: v.to(dev) for k, v in bench.synth_batch(torch, B, W["text_len"], tuple(W["image"]), 30522, 3, seed=1, pin=False).items()}
for _ in range(2):
    ts.step(batch)
torch.cuda.synchronize()
if "--gemms" in sys.argv:
    counter = _abi.install_counter()
    slot = ts._states[next(iter(ts._states))][0]
    keep = {}
    ob = ts.engine.backward
    def bk(tape, a, b):
        keep.update(tape.t); return ob(tape, a, b)
    ts.engine.backward = bk
    ts._body(slot)
    ts.engine.backward = ob
    torch.cuda.synchronize()
    gemms = list(counter.gemms)
    _abi.uninstall_counter()
    lib = _abi.lib(); st = torch.cuda.current_stream().cuda_stream
    groups = {}
    for g in gemms:
        if isinstance(g, _abi.GemmArgs):
            key = (g.M, g.N, g.K, g.a_mn, g.b_mn, g.epilogue, g.split_k, g.block_n)
        else:  # grouped weight-gradient launch: one row per distinct group shape
            key = ("grouped", tuple((x.M, x.N, x.K) for x in g), 0, 1, 1, 5, g[0].split_k, 256)
        groups.setdefault(key, []).append(g)
    rows = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for key, gs in groups.items():
        for _ in range(2):
            for g in gs: _abi.replay_gemm(g, st)
        torch.cuda.synchronize(); e0.record()
        reps = 5
        for _ in range(reps):
            for g in gs: _abi.replay_gemm(g, st)
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / reps / len(gs)
        M, N, K = key[:3]
        fl = _abi.gemm_flops(gs[0])
        rows.append(dict(M=M, N=N, K=K, a_mn=key[3], b_mn=key[4], epi=key[5], split=key[6], bn=key[7], count=len(gs), us=round(us, 2),
                         tflops=round(fl / us / 1e6, 1), total_us=round(us * len(gs), 1)))
    rows.sort(key=lambda r: -r["total_us"])
    for r in rows: print(json.dumps(r))
    print(json.dumps(dict(total_ms=sum(r["total_us"] for r in rows) / 1e3)))
else:
    torch.cuda.profiler.start()
    ts.step(batch)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
