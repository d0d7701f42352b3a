#!/usr/bin/env python
"""Benchmark of the VAuLT hot path on B200: VaultForTMSC fine-tuning samples/s (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W]                 # this repo's CUDA path (one rank per GPU under torchrun)
    python bench.py --impl reference [--gpus N] [--steps K] [--warmup W]  # the reference path on the box's host cores (CPU)

Workload (default `target`, the shape BASELINE.json's north-star "Target:" sentence names): bert-base + vilt-b32 VaultForTMSC, n_classes 3,
batch 32 per GPU, T=128 text tokens with trailing padding, 384x640 images (240 patches, S=369), synthetic inputs, random-init weights,
dropout active in the LM stack and head as in the reference's model.train().  `--workload config3` (T=40, 384x384, S=185: BASELINE config 3)
/ config4 / config5 select the other named configurations; the default line also carries a short config-3 run under "also".
A "step" = host batch -> H2D -> forward -> CE -> backward -> (gradient all-reduce) -> AdamW -> loss read-back.

Prints ONE JSON line (see the task contract): `value` = device-resident-input throughput, `e2e` = through VaultTrainStep.step()
with pinned HOST batches and per-step loss read-back, plus `roofline` (the tcgen05 GEMM, replayed launch by launch with CUDA
events), `cpu_baseline` (the reference's own VaultForTMSC on the host cores, bounded sample) and `clocks`.  `--impl reference` steps the REAL
reference class (oracle/_ref, copied from /root/reference by oracle/build_ref.py) on the host cores at the same batch and shapes.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "VaultModel fine-tune samples/sec"
# Named workloads of BASELINE.json.  The default (the one the driver runs) is `target`, the shape of the north-star's "Target:" sentence.
# train_gflop = dense-shape algorithmic FLOPs per sample of one training step (SURVEY.md section 8d / BASELINE.md section 4).
WORKLOADS = {
    "config3": dict(desc="VaultForTMSC fine-tune step (BASELINE config 3): fwd + CE + bwd + grad all-reduce + HF-AdamW", lm_kind="bert", freeze_lm=False,
                    lm="bert-base-uncased (random init)", text_len=40, image=[384, 384], patches=144, seq_len=185, train_gflop=119.99),
    "target": dict(desc="VaultForTMSC fine-tune step at the north-star target shape (bert-base + vilt-b32, 384x640 images, 128 text tokens)", lm_kind="bert",
                   freeze_lm=False, lm="bert-base-uncased (random init)", text_len=128, image=[384, 640], patches=240, seq_len=369, train_gflop=272.41),
    "config4": dict(desc="VaultForTMSC fine-tune step (BASELINE config 4): BERTweet-shaped RoBERTa + vilt-b32, max shape, variable-length text masks",
                    lm_kind="roberta", freeze_lm=False, lm="BERTweet-base-shaped RoBERTa (vocab 64001, 130 positions; random init)", text_len=128,
                    image=[384, 640], patches=240, seq_len=369, train_gflop=272.41),
    "config5": dict(desc="VaultForTMSC fine-tune step (BASELINE config 5): frozen LM (forward only) + trainable ViLT", lm_kind="bert", freeze_lm=True,
                    lm="bert-base-uncased (random init, frozen)", text_len=40, image=[384, 384], patches=144, seq_len=185, train_gflop=106.28),
}
COMMON = dict(vilt="vilt-b32 (random init)", head="VaultForTMSC n_classes=3", per_gpu_batch=32)


def workload_config(name):
    w = WORKLOADS[name]
    return dict(lm=w["lm"], **COMMON, text_len=w["text_len"], image=w["image"], patches=w["patches"], seq_len=w["seq_len"], freeze_lm=w["freeze_lm"])


def bench_config(name, B, world):
    """`config` of the JSON line: the workload only (identical in this repo's arm and in the reference arm); how an arm runs it goes to `impl_config`."""
    w = WORKLOADS[name]
    return dict(workload=w["desc"], name=name, **{**workload_config(name), "per_gpu_batch": B}, global_batch=B * world, parallelism=f"dp{world}",
                optimizer="HF-AdamW (transformers 4.48.0 rule, correct_bias=False), lr 2e-5, linear warm-up schedule past warm-up",
                l2="per-step working set (0.44 GB bf16 weights + 0.79 GB fp32 grads + >= 2 GB activations) >> 126 MB L2; 4 rotating input batches")


def train_gflop_valid_tokens(text_lens, patches, freeze_lm):
    """Algorithmic FLOPs per sample of one training step counting VALID tokens only (SURVEY.md section 8d asks for both figures): the same
    closed form as the dense-shape number -- 12 layers x (24 n H^2 + 4 n^2 H), patch projection 2 x P x 3072 x 768, train = 3 x forward
    (patch projection 2 x, a frozen LM 1 x) -- evaluated per sample with n = its number of un-padded text tokens (+ 1 + P for ViLT)."""
    enc = lambda n: 12 * (14155776 * n + 3072 * n * n)
    tot = 0.0
    for t in text_lens:
        tot += (1 if freeze_lm else 3) * enc(t) + 3 * enc(t + 1 + patches) + 2 * 4718592 * patches
    return tot / max(len(text_lens), 1) / 1e9


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="vault_b200", choices=["vault_b200", "reference"])
    ap.add_argument("--batch", type=int, default=COMMON["per_gpu_batch"])
    ap.add_argument("--workload", default="target", choices=sorted(WORKLOADS), help="named BASELINE workload (default: the north-star target shape)")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    ap.add_argument("--skip-roofline", action="store_true")
    return ap.parse_args()


def synth_batch(torch, B, T, hw, vocab, n_classes, seed, pin, pad_id=0):
    """TWITTER-15-shaped synthetic batch: ids ~ U, lengths ~ U{8..T} with trailing pad, pixels ~ N(0,1), mask all ones."""
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(1000, vocab, (B, T), generator=g)
    lens = torch.randint(8, T + 1, (B,), generator=g)
    am = (torch.arange(T)[None, :] < lens[:, None]).long()
    ids = ids * am + pad_id * (1 - am)
    batch = dict(input_ids=ids, attention_mask=am, token_type_ids=torch.zeros_like(ids),
                 pixel_values=torch.randn((B, 3, hw[0], hw[1]), generator=g), pixel_mask=torch.ones((B, hw[0], hw[1]), dtype=torch.long),
                 labels=torch.randint(0, n_classes, (B,), generator=g))
    if pin:
        batch = {k: v.pin_memory() for k, v in batch.items()}
    return batch


class ClockSampler:
    """nvidia-smi clock / throttle-reason sampler running during the timed region (B200_PROFILING.md recipe)."""

    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.gpu)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx = float(r[2])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        sm.sort()
        busy = sm[len(sm) // 2:] if sm else []
        return dict(sm_mhz=(busy[len(busy) // 2] if busy else None), sm_max_mhz=mx, samples=len(sm), reasons=sorted(reasons))


def _hf_lm_config(name):
    from transformers import BertConfig, RobertaConfig

    if WORKLOADS[name]["lm_kind"] == "roberta":  # BERTweet-base shape
        return RobertaConfig(vocab_size=64001, max_position_embeddings=130, type_vocab_size=1, pad_token_id=1, layer_norm_eps=1e-5, bos_token_id=0, eos_token_id=2)
    return BertConfig()


class ReferenceStepper:
    """The reference's fine-tuning step on the host CPU: its own ``VaultForTMSC`` (ref:vault/models/vault/model.py:512-570, loaded from
    oracle/_ref or /root/reference by oracle/ref_loader.py over the installed HF ViLT / BERT modules, transformers==4.48.0 position-embedding
    gate restored), driven exactly as ref:vault/tmsc_utils/trainer.py:353-369 drives it: model(**kwargs) -> CrossEntropyLoss -> zero_grad ->
    backward -> HF-AdamW.step -> loss.item().  Falls back to the pinned oracle port (kind "port") only if the reference files are absent."""

    def __init__(self, torch, workload, B):
        from oracle import ref_loader, synth

        self.torch, self.B = torch, B
        w = WORKLOADS[workload]
        self.w = w
        self.cores = os.cpu_count() or 1
        torch.set_num_threads(self.cores)
        bc = _hf_lm_config(workload)
        self.batch = synth_batch(torch, B, w["text_len"], tuple(w["image"]), bc.vocab_size, 3, seed=1, pin=False, pad_id=bc.pad_token_id or 0)
        self.kind = "reference" if ref_loader.available() else "port"
        if self.kind == "reference":
            from transformers import ViltConfig

            mod = ref_loader.load_reference_module()
            torch.manual_seed(0)
            m = mod.VaultForTMSC(ViltConfig(), n_classes=3, vilt_dropout_prob=0.1, bert_config=bc)
            m.embeddings.text_embeddings.position_embedding_type = "NOT_absolute"  # what from_pretrained does (ref:...model.py:112-116)
            if w["freeze_lm"]:
                m.freeze_lm = True
                for p_ in m.bert.parameters():
                    p_.requires_grad_(False)
            with torch.no_grad():
                m.embeddings.cls_token.normal_(0, 0.02)
                m.embeddings.position_embeddings.normal_(0, 0.02)
            self.model = m.train()
            self.opt = ref_loader.HFAdamW([p_ for p_ in m.parameters() if p_.requires_grad], lr=2e-5, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0,
                                          correct_bias=False)
            self.loss_fn = torch.nn.CrossEntropyLoss()
        else:
            self.d = synth.Dims.bertweet() if w["lm_kind"] == "roberta" else synth.Dims.base()
            self.sd = synth.make_state_dict(self.d, seed=0)
            self.state = None

    def step(self):
        torch = self.torch
        b = self.batch
        if self.kind == "reference":
            logits = self.model(input_ids=b["input_ids"], attention_mask=b["attention_mask"], token_type_ids=b["token_type_ids"],
                                pixel_values=b["pixel_values"], pixel_mask=b["pixel_mask"])
            loss = self.loss_fn(logits, b["labels"])
            self.opt.zero_grad()
            loss.backward()
            self.opt.step()
            return loss.item()
        from oracle import vault_oracle as O

        out = O.train_step(self.sd, self.d, b, lr=2e-5, freeze_lm=self.w["freeze_lm"], state=self.state, train_mode=True)
        self.state = out["state"]
        return float(out["loss"])

    def describe(self, what):
        w = self.w
        src = ("the reference's own VaultForTMSC (oracle/_ref: ref:vault/models/vault/model.py over HF ViLT/BERT) + HF-AdamW" if self.kind == "reference"
               else "oracle port (oracle/vault_oracle.py; reference files absent)")
        return f"{what}: {src}, fwd + CE + bwd + optimizer step, fp32, dropout on, B={self.B}, T={w['text_len']}, {w['image'][0]}x{w['image'][1]}, {self.cores} host threads"


def cpu_baseline(torch, workload, steps=2, warmup=1, B=8):
    """Bounded sample (about 10-20 s of CPU work) of the reference step on the host cores, run next to the GPU measurement."""
    st = ReferenceStepper(torch, workload, B)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        st.step()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    times.sort()
    med = times[len(times) // 2]
    return dict(value=B / med, unit="samples/s", cores=st.cores, kind=st.kind,
                sample=st.describe(f"B={B} rows of the 32-row per-GPU batch, median of {steps} steps after {warmup} warm-up"), ms_per_step=med * 1e3)


def run_reference(args):
    """Reference arm: the reference's own CPU implementation of the path on the host cores, same workload, same per-GPU batch (each step is
    ONE rank's 32-row batch: a CPU has no ranks; under torchrun rank 0 alone runs it), all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch

    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    st = ReferenceStepper(torch, args.workload, args.batch)
    t0 = None
    for i in range(args.warmup + args.steps):
        if i == args.warmup:
            t0 = time.perf_counter()
        st.step()
    dt = time.perf_counter() - t0
    v = args.batch * args.steps / dt
    sample = st.describe(f"every step = one {args.batch}-row per-GPU batch of the workload")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "samples/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": bench_config(args.workload, args.batch, world),
        "impl_config": dict(device="cpu", threads=st.cores, kind=st.kind),
        "cpu_baseline": dict(value=v, unit="samples/s", cores=st.cores, kind=st.kind, sample=sample),
        "e2e": dict(value=v, unit="samples/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
    }), flush=True)


def gemm_roofline(torch, ts, batch_dev, peaks):
    """Replays every tcgen05 GEMM launch of one training step back to back (same shapes, layouts, epilogues, real buffers) between
    two CUDA events on the launching stream: achieved = algorithmic 2*M*N*K of all launches / elapsed."""
    import ctypes as C
    from vault_b200 import _abi

    counter = _abi.install_counter()
    try:
        slot = ts._states[next(iter(ts._states))][0]
        for k, dst in slot.buf.items():
            if k in batch_dev:
                dst.copy_(batch_dev[k])
        # keep the step's activations alive so the recorded pointers stay valid while replaying
        keep = {}
        orig_backward = ts.engine.backward

        def backward_keep(tape, dlhs, dpooled):
            keep.update(tape.t)
            return orig_backward(tape, dlhs, dpooled)

        ts.engine.backward = backward_keep
        ts._body(slot)
        ts.engine.backward = orig_backward
        torch.cuda.synchronize()
        launches, calls, gemms = counter.launches, dict(counter.calls), list(counter.gemms)
    finally:
        _abi.uninstall_counter()
    lib = _abi.lib()
    st = torch.cuda.current_stream().cuda_stream
    # forward-type GEMMs only write outputs, so replay is idempotent except split-K atomics (harmless: gradients are recomputed)
    flops = sum(_abi.gemm_flops(g) for g in gemms)  # grouped weight-gradient launches count all their problems
    reps = 5
    for g in gemms:
        _abi.replay_gemm(g, st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        for g in gemms:
            _abi.replay_gemm(g, st)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    achieved = flops / (ms * 1e-3) / 1e12
    peak = peaks.get("bf16_tflops_sustained") or 1400.0
    del keep
    tr = NCU_TRAFFIC.get("gemm")  # one `ncu --set full` capture of this round (profiles/r02_ncu_traffic.json), per launch; None if not captured
    return dict(bound="tensor", kernel="vb::gemm_bf16_kernel + vb::gemm_wgrad_grouped_kernel (tcgen05/TMEM/TMA)", achieved=achieved, peak=peak, unit="TFLOP/s", frac=achieved / peak,
                traffic=(tr or {}).get("dram_bytes_per_launch"), traffic_note=tr,
                launches_per_step=len(gemms), gemm_ms_per_step=ms, flops_per_step=flops,
                peak_source="MEASURED_PEAKS.json bf16_tflops_sustained (kernel replayed inside a multi-ms dense run)" if "bf16_tflops_sustained" in peaks
                else "fallback (B200_PROFILING.md): 1.4 PFLOP/s sustained"), launches, calls


def _load_ncu_traffic():
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of one `ncu --set full` capture per kernel, taken this round at
    the workload's own launch shapes and committed under profiles/ with the command that produced it.  Not a per-run measurement: a run
    under ncu is never a bench value, so the capture is a separate call; absent file / kernel -> traffic is null."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")))
    except Exception:
        return {}


NCU_TRAFFIC = _load_ncu_traffic()


def hbm_kernel_rates(torch, peaks, rows_vilt, n_params):
    """Achieved HBM GB/s of the bandwidth-bound kernels AT THE WORKLOAD'S REAL LAUNCH SHAPES: LayerNorm forward / backward over the
    ViLT stack's `rows_vilt` = B x S token rows (11,808 at the target shape, 5,920 at config 3), AdamW over the model's `n_params` trainable
    parameters.  Every kernel runs round-robin over independent operand sets whose total size is several times the 126 MB L2, so each
    launch finds its operands in HBM without a flush kernel in between (a memset flush leaves the L2 full of dirty lines whose
    write-back the next launch pays for); all launches of `reps` passes sit between one pair of CUDA events; algorithmic bytes per
    DESIGN.md section 3.  (In the step itself consecutive kernels hand activations over through the L2, so the in-step times are shorter
    than these HBM-resident ones; profiles/ holds the in-graph family times.)"""
    from vault_b200 import _abi

    dev = torch.device("cuda", torch.cuda.current_device())
    rows, cols = rows_vilt, 768
    st = torch.cuda.current_stream().cuda_stream
    lib = _abi.lib()
    g, b = torch.ones(cols, device=dev), torch.zeros(cols, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    peak = peaks.get("hbm_gbs") or 6650.0
    out = {}

    def run(name, sets, fn, nbytes, reps=6):
        for s_ in sets:  # warm-up pass (also fills statistics the backward reads)
            fn(s_)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            for s_ in sets:
                fn(s_)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / (reps * len(sets))
        gbs = nbytes / us / 1e3
        out[name] = dict(achieved_gbs=gbs, frac_of_measured_peak=gbs / peak, algorithmic_bytes=nbytes, us_per_launch=us, operand_sets=len(sets),
                         shape=[rows, cols] if name != "adamw" else [n_params], traffic=(NCU_TRAFFIC.get(name) or {}).get("dram_bytes_per_launch"))

    # enough operand sets that one pass touches > 4 x L2
    nset_f = max(4, int(4 * 126e6 / (rows * cols * 6)) + 1)
    nset_b = max(3, int(4 * 126e6 / (rows * cols * 16)) + 1)
    # LayerNorm forward: fp32 in, bf16 out (+ row statistics): 6 B / element
    fsets = [dict(x=torch.randn(rows, cols, device=dev), y=torch.empty(rows, cols, device=dev, dtype=torch.bfloat16),
                  mean=torch.empty(rows, device=dev), rstd=torch.empty(rows, device=dev)) for _ in range(nset_f)]
    run("layernorm_fwd", fsets, lambda s_: lib.vault_layernorm_fwd(s_["x"].data_ptr(), g.data_ptr(), b.data_ptr(), s_["y"].data_ptr(), None,
                                                                    s_["mean"].data_ptr(), s_["rstd"].data_ptr(), rows, cols, 1e-12, st), rows * cols * 6)
    # LayerNorm backward: x fp32 + dy bf16 + residual gradient fp32 in, dx fp32 + dx bf16 out: 16 B / element
    dg, db, dc = (torch.zeros(cols, device=dev) for _ in range(3))
    bsets = [dict(fsets[i], dy=torch.randn(rows, cols, device=dev).to(torch.bfloat16), dres=torch.randn(rows, cols, device=dev),
                  dx32=torch.empty(rows, cols, device=dev), dx16=torch.empty(rows, cols, device=dev, dtype=torch.bfloat16)) for i in range(min(nset_b, nset_f))]
    run("layernorm_bwd", bsets, lambda s_: lib.vault_layernorm_bwd_drop(None, s_["dy"].data_ptr(), s_["x"].data_ptr(), s_["mean"].data_ptr(), s_["rstd"].data_ptr(),
                                                                         g.data_ptr(), s_["dres"].data_ptr(), s_["dx32"].data_ptr(), s_["dx16"].data_ptr(), dg.data_ptr(),
                                                                         db.data_ptr(), dc.data_ptr(), rows, cols, 0.0, 0, 0.0, 0, 0, None, st), rows * cols * 16)
    del fsets, bsets
    # fused AdamW at the model's own size (197 M parameters = 5.9 GB of state per launch >> L2): 30 B / parameter
    n = int(n_params)
    p_, g_, m_, v_ = (torch.zeros(n, device=dev) for _ in range(4))
    sh = torch.zeros(n, device=dev, dtype=torch.bfloat16)
    run("adamw", [0], lambda _s: lib.vault_adamw_step(p_.data_ptr(), g_.data_ptr(), 0, m_.data_ptr(), v_.data_ptr(), sh.data_ptr(), n, 1e-5, 0.9, 0.999, 1e-8,
                                                      0.0, 0, 1, 1.0, None, st), n * 30, reps=5)
    out["peak_gbs"] = peak
    out["method"] = "the workload's own launch shapes, round-robin over operand sets >> L2, all launches between one CUDA-event pair (no flush kernel)"
    return out


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: vault_b200 has no CPU path (use --impl reference for the CPU reference arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    from transformers import ViltConfig

    from vault_b200 import VaultForTMSC, VaultTrainStep
    from vault_b200.model import set_parameter_requires_grad

    W = WORKLOADS[args.workload]
    torch.manual_seed(0)
    vc = ViltConfig()
    bc = _hf_lm_config(args.workload)
    model = VaultForTMSC(vc, n_classes=3, vilt_dropout_prob=0.1, bert_config=bc)
    if W["freeze_lm"]:  # what VaultMixin.__init__(freeze_lm=True) does (ref:vault/models/vault/model.py:84-87); VaultForTMSC does not forward the flag
        model.freeze_lm = True
        set_parameter_requires_grad(model.bert, False)
    with torch.no_grad():  # HF leaves these at zero under random init
        model.embeddings.cls_token.normal_(0, 0.02)
        model.embeddings.position_embeddings.normal_(0, 0.02)
    model = model.to(dev).train()
    B, T, hw = args.batch, W["text_len"], tuple(W["image"])
    total_sched = 100000
    ts = VaultTrainStep(model, lr=2e-5, total_steps=total_sched, use_cuda_graph=not args.no_graph,
                        overlap_comm=os.environ.get("VB_OVERLAP", "1") == "1", comm_reserve_sms=int(os.environ.get("VB_COMM_RESERVE", "0")), grad_comm_dtype=os.environ.get("VB_COMM_DTYPE", "bf16"))
    ts.step_idx = total_sched // 5  # past warm-up: a non-zero learning rate so AdamW really moves the weights
    NB = 4
    host = [synth_batch(torch, B, T, hw, bc.vocab_size, 3, seed=1000 * rank + i, pin=True, pad_id=bc.pad_token_id or 0) for i in range(NB)]
    devb = [{k: v.to(dev) for k, v in b.items()} for b in host]
    h2d = sum(v.numel() * v.element_size() for v in host[0].values())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(batches, read_loss, steps=None, warmup=None):
        steps_, warm_ = steps or args.steps, warmup or args.warmup
        return _timed(batches, read_loss, steps_, warm_)

    def _timed(batches, read_loss, n_steps, n_warm):
        for i in range(n_warm):
            r = ts.step(batches[i % NB])
        r.loss()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        prev, losses = None, []
        # `ncu --profile-from-start off ... python bench.py` then lists exactly the launches of the device-resident timed region
        prof = os.environ.get("VAULT_B200_PROFILE_RANGE") == "1" and not read_loss
        if prof:
            torch.cuda.profiler.start()
        e0.record()
        for i in range(n_steps):
            r = ts.step(batches[i % NB])
            if read_loss and prev is not None:
                losses.append(prev.loss())
            prev = r
        if read_loss:
            losses.append(prev.loss())
        e1.record()
        barrier()
        if prof:
            torch.cuda.profiler.stop()
        ms = e0.elapsed_time(e1)
        t = torch.tensor([ms], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), losses

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    ms_dev, _ = timed(devb, read_loss=False)
    ms_e2e, losses = timed(host, read_loss=True)
    h2d_copied = int(getattr(ts, "last_h2d_bytes", 0)) or h2d  # what step() really copied (a host pixel mask is reduced to patch-grid sizes first)
    clocks = sampler.stop() if sampler else None
    # BASELINE config 3 (T=40, 384x384, S=185) on the same model and step object, a short run next to the headline shape
    also = None
    if args.workload == "target" and os.environ.get("VB_BENCH_ALSO", "1") == "1":
        W3 = WORKLOADS["config3"]
        host3 = [synth_batch(torch, B, W3["text_len"], tuple(W3["image"]), bc.vocab_size, 3, seed=7000 + 1000 * rank + i, pin=True, pad_id=bc.pad_token_id or 0)
                 for i in range(NB)]
        dev3 = [{k: v.to(dev) for k, v in b.items()} for b in host3]
        ms3, _ = timed(dev3, read_loss=False, steps=10, warmup=4)
        ms3e, _ = timed(host3, read_loss=True, steps=10, warmup=4)
        also = dict(config3=dict(config=bench_config("config3", B, world), steps=10, warmup=4, ms_per_step=ms3 / 10, value=B * world * 10 / (ms3 * 1e-3),
                                 e2e=B * world * 10 / (ms3e * 1e-3), unit="samples/s",
                                 model_tflops_per_gpu=B * 10 / (ms3 * 1e-3) * W3["train_gflop"] / 1e3))
        del host3, dev3

    replicas_identical = None
    if world > 1:
        # after the run every replica must hold the same fp32 weights (synchronize() also consolidates sharded masters): compare a checksum
        ts.synchronize()
        mw = ts.engine.master[:ts.engine.n_train]
        chk = torch.stack([mw.double().sum(), mw[::997].double().abs().sum()])
        allc = [torch.empty_like(chk) for _ in range(world)]
        dist.all_gather(allc, chk)
        replicas_identical = bool(all(torch.equal(allc[0], c) for c in allc[1:]))
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        roof, launches, calls = (None, None, None)
        if not args.skip_roofline:
            roof, launches, calls = gemm_roofline(torch, ts, devb[0], peaks)
        else:
            launches = 0
        hbm = hbm_kernel_rates(torch, peaks, B * W["seq_len"], ts.engine.n_train) if (world == 1 and not args.skip_roofline) else None
        cpu = None
        if world == 1 and not args.skip_cpu_baseline:
            cpu = cpu_baseline(torch, args.workload)
        gb = B * world
        value = gb * args.steps / (ms_dev * 1e-3)
        e2e = gb * args.steps / (ms_e2e * 1e-3)
        train_gflop_per_sample = W["train_gflop"]  # BASELINE.md section 4 (dense-shape FLOPs of the named shape)
        lens = [int(n) for b_ in host for n in b_["attention_mask"].sum(dim=1).tolist()]
        valid_gflop = train_gflop_valid_tokens(lens, W["patches"], W["freeze_lm"])  # padded text positions not counted
        out = {
            "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": bench_config(args.workload, B, world),
            "impl_config": dict(device="cuda", cuda_graph=not args.no_graph, comm_overlap=ts.overlap, grad_comm_dtype=ts.grad_comm, comm=("multimem" if ts.mc is not None else ("nccl" if world > 1 else "none")),
                                gemm_ctas=ts.engine.gemm_max_ctas or ts.engine.sms),
            "e2e": dict(value=e2e, unit="samples/s", h2d_bytes_per_step=h2d_copied, host_batch_bytes=h2d, d2h_bytes_per_step=4, ms_per_step=ms_e2e / args.steps,
                        api="vault_b200.VaultTrainStep.step(pinned host batch) -> StepResult.loss()"),
            "gpu_launches": (launches + 1) * args.steps if launches else None,
            "launches_per_step": dict(total=(launches + 1) if launches else None, by_call=calls),
            "model_tflops_per_gpu": value / world * train_gflop_per_sample / 1e3,
            "model_tflops_per_gpu_valid_tokens": value / world * valid_gflop / 1e3,
            "clocks": clocks,
            "roofline": roof,
            "hbm_kernels": hbm,
            "cpu_baseline": cpu,
            "loss_first_last": [losses[0], losses[-1]] if losses else None,
            "replicas_identical": replicas_identical,
            "also": also,
        }
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
