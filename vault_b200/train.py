"""Data-parallel fine-tuning step for VaultForTMSC on B200s: the fast equivalent of the reference's hot loop body

    batch_to_device -> model(**kwargs) -> CrossEntropyLoss -> zero_grad -> backward -> HF-AdamW.step -> scheduler.step -> loss.item()
    (ref:vault/tmsc_utils/trainer.py:353-369)

as ONE replayed CUDA graph per step (forward, fused CE head, backward) + bucketless flat gradient all-reduce over NCCL (one
process per GPU, batch sharding, no other collective) + ONE fused AdamW launch over the flat parameter range.  Host->device
input copies run on a side stream into double-buffered static inputs so step i+1's copy overlaps step i's compute; the loss is
read back asynchronously (pinned memory) so there is no per-step device synchronisation.

Semantics kept from the reference: CE mean over the (global) batch, transformers==4.48.0 AdamW rule with correct_bias=False,
linear warm-up/decay schedule with lr(0)=0, no gradient clipping, dropout active in the LM stack and the head (p=0.1), ViLT
dropout 0 (SURVEY.md sections 5, 8a).  Image tokens always span the full padded patch grid here (static shapes): padded patches
are masked out of attention, so loss and gradients equal the reference's dynamic-length computation.
"""
from __future__ import annotations

import os
from typing import Dict, Optional

import torch

from . import _abi
from .engine import VaultEngine

_INPUT_KEYS = ("input_ids", "attention_mask", "token_type_ids", "pixel_values", "pixel_mask", "labels")


def allreduce_flat_(flat: torch.Tensor, group=None, bucket_elems: int = 16 * 1024 * 1024, async_op: bool = False):
    """Sum-all-reduce a flat gradient range in fixed-size buckets (the only collective of the path: batch data parallelism).
    The 1/world averaging is folded into the AdamW kernel's grad_scale."""
    import torch.distributed as dist

    works = []
    n = flat.numel()
    for a in range(0, n, bucket_elems):
        w = dist.all_reduce(flat[a:min(n, a + bucket_elems)], group=group, async_op=async_op)
        if async_op:
            works.append(w)
    return works


class _McBuffers:
    """Symmetric-memory home of the engine's flat buffers for the multicast data-parallel path: fp32 masters, bf16 shadow and fp32
    gradients are allocated with torch.distributed._symmetric_memory.empty, exchanged with rendezvous() (CUDA VMM handles + one NVSwitch
    multicast object per buffer) and addressed through `multicast_ptr` by vault_mc_adamw_step."""

    # a rank that waits longer than this at a gradient-range barrier traps (CUDA error) instead of hanging: same order as NCCL's watchdog
    BARRIER_TIMEOUT_MS = int(os.environ.get("VAULT_B200_MC_BARRIER_TIMEOUT_MS", "600000"))

    def __init__(self):
        self.master_mc = self.shadow_mc = self.grad_mc = self.grad16_mc = 0
        self.grad16 = None
        self.handles = []
        self._ptrs = ()

    @classmethod
    def create(cls, engine: VaultEngine, dev, group, required: bool, payload: str = "bf16"):
        import torch.distributed as dist
        try:
            import torch.distributed._symmetric_memory as symm_mem
        except Exception as e:  # pragma: no cover
            if required:
                raise RuntimeError(f"comm='multimem' needs torch.distributed._symmetric_memory: {e}")
            return None
        pg = group if group is not None else dist.group.WORLD
        if dist.get_backend(pg) != "nccl":
            if required:
                raise RuntimeError("comm='multimem' needs an NCCL process group on CUDA devices")
            return None
        self = cls()
        made = []

        def alloc(numel, dt):
            t = symm_mem.empty(numel, dtype=dt, device=dev)
            t.zero_()
            made.append(t)
            return t

        try:
            engine.repack(dev, flat_alloc=alloc)
            bufs = [engine.master, engine.shadow, engine.grad]
            if payload == "bf16":  # gradients travel as a bf16 copy of each finished range (half the link bytes; the switch accumulates in fp32)
                self.grad16 = alloc(engine.grad.numel(), torch.bfloat16)
                bufs.append(self.grad16)
            torch.cuda.synchronize(dev)
            for t in bufs:
                self.handles.append(symm_mem.rendezvous(t, pg))
            ok = all(int(h.multicast_ptr) != 0 for h in self.handles)
        except Exception as e:
            if required:
                raise
            ok = False
            self._why = repr(e)
        # every rank must take the same path
        flag = torch.tensor([1 if ok else 0], device=dev, dtype=torch.int32)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        if int(flag.item()) == 0:
            if required:
                raise RuntimeError("comm='multimem': this system has no NVSwitch multicast support (multicast_ptr == 0 on some rank)")
            engine.repack(dev, flat_alloc=None)
            return None
        self.master_mc, self.shadow_mc, self.grad_mc = (int(h.multicast_ptr) for h in self.handles[:3])
        self.grad16_mc = int(self.handles[3].multicast_ptr) if self.grad16 is not None else 0
        self._ptrs = (engine.master.data_ptr(), engine.shadow.data_ptr(), engine.grad.data_ptr())
        return self

    def check(self, engine: VaultEngine):
        if (engine.master.data_ptr(), engine.shadow.data_ptr(), engine.grad.data_ptr()) != self._ptrs:
            raise RuntimeError("the engine re-packed its parameters (requires_grad changed?) after VaultTrainStep mapped them into multicast "
                               "memory: create a new VaultTrainStep")

    def barrier(self):
        self.handles[2].barrier(channel=0, timeout_ms=self.BARRIER_TIMEOUT_MS)


LOSS_RING = 4096  # steps whose loss may be outstanding (unread) at once


class StepResult:
    """Handle on a step's loss: ``loss()`` waits for that step's device->host copy only.  Every step owns one entry (and one event)
    of a pinned ring of LOSS_RING floats, so handles may be resolved late and in any order -- the reference accumulates
    ``loss.item() * batch_len`` per step (ref:vault/tmsc_utils/trainer.py:369); a caller that defers the read gets the same numbers."""

    def __init__(self, owner, step_no: int, lr: float):
        self._owner, self._step, self.lr = owner, step_no, lr
        self._value: Optional[float] = None

    def loss(self) -> float:
        if self._value is None:
            o = self._owner
            if o.step_idx - self._step > LOSS_RING:
                raise RuntimeError(f"StepResult.loss(): step {self._step} is more than {LOSS_RING} steps old, its read-back entry was reused")
            i = self._step % LOSS_RING
            o._loss_events[i].synchronize()
            self._value = float(o._loss_ring[i])
        return self._value


class _Slot:
    def __init__(self):
        self.buf: Dict[str, torch.Tensor] = {}
        self.ready = torch.cuda.Event()
        self.free = torch.cuda.Event()
        self.graphs = None   # list of (CUDAGraph, grad offset reached when it finishes)
        self.host_hw = False  # the per-sample patch-grid sizes arrive from the host (the pixel mask itself is never copied)
        self.hw_pinned = None


class VaultTrainStep:
    def __init__(self, model, lr: float = 2e-5, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0, correct_bias: bool = False,
                 total_steps: Optional[int] = None, warmup_ratio: float = 0.1, process_group=None, use_cuda_graph: bool = True,
                 dropout: bool = True, overlap_comm: bool = True, comm_reserve_sms: int = 0, grad_comm_dtype: str = "bf16", loss: str = "auto",
                 comm: str = "auto", mc_ctas: int = 0):
        self.model = model
        self.engine: VaultEngine = model.engine
        self.lr, self.betas, self.eps, self.wd, self.correct_bias = lr, betas, eps, weight_decay, correct_bias
        self.total_steps, self.warmup_ratio = total_steps, warmup_ratio
        self.use_graph = use_cuda_graph
        self.dropout = dropout  # False: dropout off everywhere (deterministic parity runs)
        self.n_classes = model.classifier[1].out_features
        # loss of the reference trainer in use: "ce" (Twitter-201x / pre-processed MVSA, int64 labels [B]), "bce" (Bloomberg: one
        # logit, float labels [B], ref:vault/models/vault/trainer.py:42-56), "ce2" (raw MVSA: two label groups, int64 labels [B,2], ref :114-137)
        if loss == "auto":
            loss = "bce" if self.n_classes == 1 else "ce"
        if loss not in ("ce", "bce", "ce2") or (loss == "bce") != (self.n_classes == 1) or (loss == "ce2" and self.n_classes % 2):
            raise ValueError(f"VaultTrainStep: loss={loss!r} does not fit a head with {self.n_classes} outputs")
        self.loss_kind = {"ce": 0, "bce": 1, "ce2": 2}[loss]
        self.head_p = float(model.classifier[0].p)
        self.dev = next(model.parameters()).device
        if self.dev.type != "cuda":
            raise RuntimeError("VaultTrainStep needs the model on a CUDA device (sm_100a); there is no CPU path")
        self.world, self.rank, self.pg = 1, 0, process_group
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized():
            self.world = dist.get_world_size(process_group)
            self.rank = dist.get_rank(process_group)
        # Gradient exchange of the data-parallel step:
        #   "multimem": the flat master / shadow / gradient buffers live in symmetric memory mapped into one NVSwitch multicast object and
        #               each finished gradient range is handled by ONE kernel per rank (vault_mc_adamw_step: in-switch reduce of the rank's
        #               1/world slice, AdamW with the slice's own moments, multicast store of the new weights to every replica);
        #   "nccl":     cast -> NCCL all-reduce -> full-range AdamW on every rank.
        #   "auto":     multimem when the system has multicast support, else nccl.
        if grad_comm_dtype not in ("bf16", "fp32"):
            raise ValueError("grad_comm_dtype must be 'bf16' or 'fp32'")
        comm = os.environ.get("VAULT_B200_COMM", comm)
        if comm not in ("auto", "multimem", "nccl"):
            raise ValueError("comm must be 'auto', 'multimem' or 'nccl'")
        self.mc = None
        if self.world > 1 and comm != "nccl":
            self.mc = _McBuffers.create(self.engine, self.dev, process_group, required=(comm == "multimem"), payload=grad_comm_dtype)
        if self.mc is None:
            self.engine.ensure_packed(self.dev)
        self.mc_ctas = int(os.environ.get("VAULT_B200_MC_CTAS", mc_ctas or 128))
        # multimem: the fp32 masters of the dense projection matrices (read through their bf16 shadow only) are SHARDED by default -- only
        # the shadow of an updated slice is multicast (2 B/parameter instead of 6); everything kernels read in fp32 (biases, LayerNorm,
        # embedding tables, patch projection, pooler, head) is replicated at once.  synchronize() brings every replica's masters up to date
        # before anybody reads the Parameters (evaluation, state_dict).
        self.mc_shard_master = os.environ.get("VAULT_B200_MC_SHARD_MASTER", "1") != "0"
        self._mc_local_bits = self.engine.shadow_only_bitmap() if (self.mc is not None and self.mc_shard_master) else None
        self._mc_segments = set()
        self._mc_stale = False
        # multimem + CUDA graphs: the WHOLE step -- forward, backward, per-range barrier + fused optimizer kernel on a third (comm) stream,
        # end-of-step barrier -- is ONE captured graph: no host launch between gradient ranges, and the dependency chain does not wait for
        # the weight-gradient stream at the range boundaries (only the comm stream does, VaultEngine._cut).
        self.mc_in_graph = self.mc is not None and use_cuda_graph and os.environ.get("VAULT_B200_MC_IN_GRAPH", "1") != "0"
        self._comm_stream = torch.cuda.Stream(device=self.dev) if self.mc is not None else None
        self.engine.refresh_shadow(force=True)
        self.engine.init_opt_state()
        if self.world > 1:
            # identical replicas: rank 0's weights everywhere; decorrelated dropout streams per rank
            dist.broadcast(self.engine.master, src=dist.get_global_rank(process_group, 0) if process_group is not None else 0, group=process_group)
            self.engine.refresh_shadow(force=True)
            self.engine.seed = (self.engine.seed + 0x9E3779B97F4A7C15 * self.rank) & 0xFFFFFFFFFFFFFFFF
        # data-parallel overlap: backward is cut into segments of the reverse-topological gradient layout; each finished
        # range is all-reduced (async, NCCL's stream) while the next segment computes.  The persistent GEMMs then leave a few
        # SMs to the collective instead of queueing behind it.
        # Each reduced range's AdamW update runs right behind its all-reduce on the side stream, under the remaining backward:
        # a finished segment's weights are not read again in this step.
        # (On one GPU the segmentation buys nothing measurable -- AdamW and the GEMMs contend for the same HBM/L2 -- so the step
        # stays one graph there.)
        self.overlap = bool(overlap_comm) and self.world > 1
        # Opt-in experiment (VAULT_B200_SPLIT_LM=1), one GPU: run the LM forward of step i+1 under the ViLT part of step i's AdamW
        # (graph cut after the LM forward, AdamW issued LM range first).  Measured neutral on B200 -- AdamW saturates HBM and
        # slows the concurrent kernels by what it hides -- so it is off by default.
        self.split_lm = (self.world == 1 and self.engine.lm is not None and not getattr(model, "freeze_lm", False)
                         and self.engine._first_off("bert.") is not None and os.environ.get("VAULT_B200_SPLIT_LM", "0") == "1")
        self._ev_lm = self._ev_rest = None
        if self.overlap and comm_reserve_sms > 0:
            self.engine.gemm_max_ctas = max(1, self.engine.sms - comm_reserve_sms)
        # gradient all-reduce payload: "bf16" halves the NVLink bytes (each finished range is cast to a bf16 comm buffer, summed
        # by NCCL, and AdamW reads the bf16 sums); "fp32" reduces the fp32 buffer in place (bit-faithful sum of the ranks' gradients)
        if grad_comm_dtype not in ("bf16", "fp32"):
            raise ValueError("grad_comm_dtype must be 'bf16' or 'fp32'")
        self.grad16 = (torch.empty(self.engine.n_train, device=self.dev, dtype=torch.bfloat16)
                       if (self.world > 1 and grad_comm_dtype == "bf16" and self.mc is None) else None)
        self.grad_comm = grad_comm_dtype if self.world > 1 else "none"
        self.copy_stream = torch.cuda.Stream(device=self.dev)
        self.sched_dev = torch.zeros(2, device=self.dev, dtype=torch.float32)
        self.step_idx = 0
        self._states: Dict[tuple, list] = {}
        self._loss_ring = torch.zeros(LOSS_RING, dtype=torch.float32).pin_memory()
        self._loss_events = [None] * LOSS_RING
        self._pool = None
        self._turn = 0
        self._mc_dirty = False
        self.last_h2d_bytes = 0

    # ------------------------------------------------------------------------------------------------------------
    def lr_at(self, step: int) -> float:
        """get_linear_schedule_with_warmup (HF:optimization.py:101-131) with warm = int(ratio*total) (ref:vault/tmsc_utils/trainer.py:256-280)."""
        if self.total_steps is None:
            return self.lr
        warm = int(self.warmup_ratio * self.total_steps)
        if step < warm:
            return self.lr * float(step) / float(max(1, warm))
        return self.lr * max(0.0, float(self.total_steps - step) / float(max(1, self.total_steps - warm)))

    def _make_slot(self, batch) -> _Slot:
        s = _Slot()
        # A HOST pixel mask is reduced on the host to what the path needs from it -- (valid patch rows, valid patch columns) per sample, the
        # same two counts vault_patch_grid takes from the first patch column / row -- so 8 bytes per sample cross PCIe instead of the int64
        # mask (63 MB of the 157 MB a 32 x 384 x 640 batch would otherwise copy per step).
        s.host_hw = batch.get("pixel_mask") is not None and not batch["pixel_mask"].is_cuda
        for k in _INPUT_KEYS:
            v = batch.get(k)
            if v is None or (k == "pixel_mask" and s.host_hw):
                continue
            dt = torch.float32 if (k == "pixel_values" or (k == "labels" and self.loss_kind == 1)) else torch.int64
            s.buf[k] = torch.empty(v.shape, device=self.dev, dtype=dt)
        B = batch["input_ids"].shape[0]
        s.buf["hw"] = torch.empty((B, 2), device=self.dev, dtype=torch.int32)
        if s.host_hw:
            s.hw_pinned = torch.empty((B, 2), dtype=torch.int32).pin_memory()
        s.buf["loss"] = torch.zeros(1, device=self.dev, dtype=torch.float32)
        s.buf["logits"] = torch.zeros((B, self.n_classes), device=self.dev, dtype=torch.float32)
        s.free.record(torch.cuda.current_stream(self.dev))
        return s

    def _body(self, s: _Slot):
        for _ in self._body_iter(s, segments=False):
            pass

    def _body_iter(self, s: _Slot, segments: bool):
        """forward + CE head + backward of one local batch, all kernels on the current stream (capturable).  Generator: yields
        the gradient offset that is final after each backward segment (see VaultEngine.backward_iter)."""
        eng, lib = self.engine, _abi.lib()
        b = s.buf
        st = eng._stream()
        eng._lib, eng._st = lib, st
        B, T = b["input_ids"].shape
        _, Cc, Hi, Wi = b["pixel_values"].shape
        gh, gw = Hi // eng.patch, Wi // eng.patch
        H = eng.H
        train = self.dropout
        eng.seed_dev.add_(1)
        if "pixel_mask" in b:
            _abi.check(lib.vault_patch_grid(b["pixel_mask"].data_ptr(), 0, b["hw"].data_ptr(), B, Hi, Wi, eng.patch, st), "patch_grid")
        elif s.host_hw:
            pass  # b["hw"] was filled by step()'s host->device copy
        else:
            b["hw"][:, 0] = gh
            b["hw"][:, 1] = gw
        # gradient zero-fill (the accumulated ranges, ~0.8 GB) on the side stream: it runs under the LM forward instead of in front of the backward
        if not (segments and self.split_lm):  # (that opt-in mode ends a graph right after the LM forward: nothing may be in flight on the side stream there)
            ev0 = torch.cuda.Event()
            ev0.record(torch.cuda.current_stream(self.dev))
            eng._side.wait_event(ev0)
            with torch.cuda.stream(eng._side):
                eng.zero_accumulated_grads()
            eng._side_dirty = True
            eng.prezeroed = True
        lhs, pooled, key_mask, tape = yield from eng.forward_iter(b["input_ids"], b.get("attention_mask"), b.get("token_type_ids"), b["pixel_values"],
                                                                  None, training=train, need_grad=True, hw=b["hw"], pmax=gh * gw,
                                                                  split_lm=segments and self.split_lm)
        eng._join_side()  # the zero-fill (and the image branch) have landed before the first gradient is written
        # ---- head: Linear(Dropout(pooled)) -> CE mean; dlogits = (softmax - onehot) / B_local (DP averaging is applied in AdamW's grad_scale)
        p = self.head_p if train else 0.0
        x = pooled
        if p > 0:
            x = torch.empty_like(pooled)
            _abi.check(lib.vault_dropout_f32(pooled.data_ptr(), x.data_ptr(), pooled.numel(), p, eng.seed, eng.seed_dev.data_ptr(), eng.SITE_HEAD, st),
                       "head_dropout")
        n = self.n_classes
        _abi.check(lib.vault_small_linear_fwd(x.data_ptr(), H, eng.w32("classifier.1.weight"), eng.w32("classifier.1.bias"), b["logits"].data_ptr(), B, n,
                                              H, 0, st), "classifier_fwd")
        dlogits = torch.empty((B, n), device=self.dev, dtype=torch.float32)
        _abi.check(lib.vault_head_loss(b["logits"].data_ptr(), b["labels"].data_ptr(), b["loss"].data_ptr(), dlogits.data_ptr(), B, n, self.loss_kind, 1.0,
                                       st), "head_loss")
        dx = torch.empty_like(pooled)
        _abi.check(lib.vault_small_linear_bwd(dlogits.data_ptr(), None, x.data_ptr(), H, eng.w32("classifier.1.weight"), dx.data_ptr(), H, 0,
                                              eng.g32("classifier.1.weight") or None, eng.g32("classifier.1.bias") or None, B, n, H, 0, st),
                   "classifier_bwd")
        if p > 0:
            _abi.check(lib.vault_dropout_f32(dx.data_ptr(), dx.data_ptr(), dx.numel(), p, eng.seed, eng.seed_dev.data_ptr(), eng.SITE_HEAD, st),
                       "head_dropout_bwd")
        yield from eng.backward_iter(tape, None, dx, segments=segments)

    def _after_segment(self, lo: int, reached, hp, cs) -> int:
        """Called after each captured / eager segment of the step; returns the new low end of the not-yet-updated gradient range."""
        if reached == "lm_done":
            if self._ev_rest is not None:
                cs.wait_event(self._ev_rest)  # previous step's AdamW has finished the ViLT range: its weights may be read now
            return lo
        if self.split_lm:
            return reached  # single GPU: AdamW is issued once, after the whole backward (see step())
        self._finish_range(lo, reached, hp)
        return reached

    def synchronize(self):
        """Wait for everything this object has enqueued (incl. the trailing AdamW on the side stream).  Multicast data parallelism with
        sharded masters: also brings every replica's fp32 Parameters up to date (each rank multicasts the slices it owns) -- a COLLECTIVE
        then: every rank must call it (the trainer does, before evaluation and before a checkpoint)."""
        torch.cuda.current_stream(self.dev).wait_stream(self.engine._side)
        if self.mc is not None and self._mc_stale:
            eng, mc = self.engine, self.mc
            mc.check(eng)
            st = torch.cuda.current_stream(self.dev).cuda_stream
            mc.barrier()
            for lo, hi in sorted(self._mc_segments):
                a, b = self._mc_slice(lo, hi)
                if b > a:
                    _abi.call("vault_mc_broadcast_f32", eng.master.data_ptr() + 4 * a, mc.master_mc + 4 * a, b - a, 2 * eng.sms, st)
            mc.barrier()
            self._mc_stale = False
        torch.cuda.synchronize(self.dev)

    def _mc_slice(self, lo: int, hi: int):
        """This rank's share [a, b) of the gradient range [lo, hi): equal slices in units of 8 parameters."""
        n = hi - lo
        per = (-(-n // self.world) + 7) // 8 * 8
        return lo + min(n, per * self.rank), lo + min(n, per * (self.rank + 1))

    def _finish_range(self, lo: int, hi: int, hp):
        """Gradients in [lo, hi) are final on the main stream: all-reduce them (async) and apply AdamW to that range on the side
        stream while the main stream carries on with the rest of backward."""
        if hi <= lo:
            return
        eng = self.engine
        side = eng._side if eng.comm_stream is None else eng.comm_stream
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.dev))
        side.wait_event(ev)
        with torch.cuda.stream(side):
            if self.mc is not None:
                self._mc_range(lo, hi, hp, side)
                return
            works = []
            if self.world > 1:
                if self.grad16 is not None:
                    _abi.call("vault_cast_f32_bf16", eng.grad.data_ptr() + 4 * lo, self.grad16.data_ptr() + 2 * lo, hi - lo, side.cuda_stream)
                    works = allreduce_flat_(self.grad16[lo:hi], self.pg, bucket_elems=1 << 30, async_op=True)
                else:
                    works = allreduce_flat_(eng.grad[lo:hi], self.pg, bucket_elems=1 << 30, async_op=True)
            for w in works:
                w.wait()
            eng.adamw_range(lo, hi, hp["step"], hp["lr"], hp["b1"], hp["b2"], self.eps, self.wd, self.correct_bias, 1.0 / self.world, self.sched_dev,
                            side.cuda_stream, grad16=self.grad16)

    def _mc_range(self, lo: int, hi: int, hp, side):
        """Multicast path, on the side stream: barrier (every rank's gradients of [lo, hi) are final), then the fused kernel on this
        rank's slice of the range.  The barrier is the symmetric-memory handle's (stream-ordered, traps after a timeout instead of
        hanging); the matching end-of-step barrier is issued by step()."""
        eng, mc = self.engine, self.mc
        mc.check(eng)
        n = hi - lo
        a, b = self._mc_slice(lo, hi)
        g16 = mc.grad16 is not None
        if g16:
            _abi.call("vault_cast_f32_bf16", eng.grad.data_ptr() + 4 * lo, mc.grad16.data_ptr() + 2 * lo, n, side.cuda_stream)
        mc.barrier()
        if b > a and os.environ.get("VAULT_B200_MC_SKIP_KERNEL", "0") != "1":  # (diagnostic switch: cast + barriers only)
            st = eng.opt_state
            _abi.call("vault_mc_adamw_step", eng.master.data_ptr() + 4 * a, mc.master_mc + 4 * a,
                      self._mc_local_bits.data_ptr() if self.mc_shard_master else None, a,
                      (mc.grad16_mc + 2 * a) if g16 else (mc.grad_mc + 4 * a), int(g16), st["m"].data_ptr() + 4 * a, st["v"].data_ptr() + 4 * a,
                      mc.shadow_mc + 2 * a, b - a, hp["lr"], hp["b1"], hp["b2"], self.eps, self.wd, int(self.correct_bias), max(1, hp["step"]),
                      1.0 / self.world, self.sched_dev.data_ptr(), self.mc_ctas, side.cuda_stream)
        self._mc_segments.add((lo, hi))
        self._mc_stale = self.mc_shard_master
        self._mc_dirty = True

    def _get_slot(self, batch) -> _Slot:
        key = tuple((k, tuple(batch[k].shape), bool(batch[k].is_cuda)) for k in _INPUT_KEYS if batch.get(k) is not None)
        if key not in self._states:
            self._states[key] = [self._make_slot(batch), self._make_slot(batch)]
        self._turn ^= 1
        return self._states[key][self._turn]

    def step(self, batch: Dict[str, torch.Tensor]) -> StepResult:
        """One optimizer step on this rank's local batch (host tensors -- ideally pinned -- or device tensors)."""
        eng = self.engine
        s = self._get_slot(batch)
        cs = torch.cuda.current_stream(self.dev)
        on_host = not batch["pixel_values"].is_cuda
        if on_host:
            nbytes = 0
            with torch.cuda.stream(self.copy_stream):
                self.copy_stream.wait_event(s.free)
                if s.host_hw:
                    pm, P = batch["pixel_mask"], eng.patch
                    s.free.synchronize()  # the pinned staging pair of this slot is reused: its previous copy must have been consumed
                    s.hw_pinned[:, 0] = (pm[:, ::P, 0] != 0).sum(1)
                    s.hw_pinned[:, 1] = (pm[:, 0, ::P] != 0).sum(1)
                    s.buf["hw"].copy_(s.hw_pinned, non_blocking=True)
                    nbytes += s.hw_pinned.numel() * 4
                for k, dst in s.buf.items():
                    if k in batch and batch[k] is not None:
                        dst.copy_(batch[k], non_blocking=True)
                        nbytes += dst.numel() * dst.element_size()
                s.ready.record(self.copy_stream)
            self.last_h2d_bytes = nbytes
            cs.wait_event(s.ready)
        else:
            for k, dst in s.buf.items():
                if k in batch and batch[k] is not None:
                    dst.copy_(batch[k], non_blocking=True)
        lr = self.lr_at(self.step_idx)
        eng.refresh_shadow()  # host-side check only, unless someone modified the Parameters in place
        n_train = eng.n_train
        eng.opt_state["step"] += 1
        b1, b2 = self.betas
        step_no = eng.opt_state["step"]
        step_size = lr
        if self.correct_bias:
            step_size = lr * (1.0 - b2 ** step_no) ** 0.5 / (1.0 - b1 ** step_no)
        if self._ev_rest is not None:
            cs.wait_event(self._ev_rest)  # split_lm: the previous step's ViLT-range AdamW (side stream) still reads sched_dev
        self.sched_dev.copy_(torch.tensor([step_size, lr * self.wd if self.wd > 0 else 0.0], dtype=torch.float32), non_blocking=True)
        hp = dict(step=step_no, lr=lr, b1=b1, b2=b2)
        if self.use_graph:
            if s.graphs is None:
                self._body(s)  # eager warm-up: sets kernel attributes, sizes the allocator
                torch.cuda.synchronize(self.dev)
                eng.seed_dev.sub_(1)
                s.graphs = []
                it = self._body_iter(s, segments=self.overlap or self.split_lm)
                # the dependency chain (forward; dgrad -> LayerNorm -> attention) is captured on a HIGH-priority stream, the engine's side
                # stream (weight gradients, bias sums) keeps the default priority: pending chain CTAs are placed first, the rest fills in
                cap_stream = torch.cuda.Stream(device=self.dev, priority=-1 if os.environ.get("VAULT_B200_CHAIN_PRIORITY", "0") != "0" else 0)
                done = False
                if self.mc_in_graph:
                    g = torch.cuda.CUDAGraph()
                    eng.comm_stream = self._comm_stream
                    try:
                        with torch.cuda.graph(g, pool=self._pool, stream=cap_stream):
                            lo_c = 0
                            for reached in it:
                                lo_c = self._after_segment(lo_c, reached, hp, cap_stream)
                            self._after_segment(lo_c, n_train, hp, cap_stream)
                            with torch.cuda.stream(self._comm_stream):
                                self.mc.barrier()  # every rank has stored every slice to every replica; gradients may be overwritten
                            evc = torch.cuda.Event()
                            evc.record(self._comm_stream)
                            torch.cuda.current_stream(self.dev).wait_event(evc)
                    finally:
                        eng.comm_stream = None
                    self._mc_dirty = False
                    if self._pool is None:
                        self._pool = g.pool()
                    s.graphs.append((g, "whole_step"))
                    done = True
                while not done:
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, pool=self._pool, stream=cap_stream):
                        try:
                            reached = next(it)
                        except StopIteration:
                            reached, done = n_train, True
                    if self._pool is None:
                        self._pool = g.pool()
                    s.graphs.append((g, reached))
            lo = 0
            if self._ev_lm is not None:
                cs.wait_event(self._ev_lm)  # previous step's AdamW has finished the LM range
            for g, reached in s.graphs:
                g.replay()
                if reached == "whole_step":
                    self._mc_stale = self.mc_shard_master
                    continue
                lo = self._after_segment(lo, reached, hp, cs)
        else:
            lo = 0
            if self._ev_lm is not None:
                cs.wait_event(self._ev_lm)
            for reached in self._body_iter(s, segments=self.overlap or self.split_lm):
                lo = self._after_segment(lo, reached, hp, cs)
            lo = self._after_segment(lo, n_train, hp, cs)
        if on_host:
            s.free.record(cs)
        if self.split_lm:
            # AdamW on the side stream, LM range first; the next step's LM forward waits for _ev_lm only, its ViLT part for _ev_rest
            lm_lo = eng._first_off("bert.")
            side = eng._side
            ev = torch.cuda.Event()
            ev.record(cs)
            side.wait_event(ev)
            with torch.cuda.stream(side):
                eng.adamw_range(lm_lo, n_train, hp["step"], hp["lr"], hp["b1"], hp["b2"], self.eps, self.wd, self.correct_bias, 1.0, self.sched_dev,
                                side.cuda_stream)
                self._ev_lm = torch.cuda.Event()
                self._ev_lm.record(side)
                eng.adamw_range(0, lm_lo, hp["step"], hp["lr"], hp["b1"], hp["b2"], self.eps, self.wd, self.correct_bias, 1.0, self.sched_dev,
                                side.cuda_stream)
                self._ev_rest = torch.cuda.Event()
                self._ev_rest.record(side)
        else:
            if self.mc is not None and self._mc_dirty:
                with torch.cuda.stream(eng._side):
                    self.mc.barrier()  # every rank has stored every slice to every replica; gradients may be overwritten from here on
                self._mc_dirty = False
            ev = torch.cuda.Event()
            ev.record(eng._side)
            cs.wait_event(ev)  # every range's AdamW (and all-reduce) is done before the next step touches weights or gradients
        i = self.step_idx % LOSS_RING
        self._loss_ring[i:i + 1].copy_(s.buf["loss"], non_blocking=True)
        if self._loss_events[i] is None:
            self._loss_events[i] = torch.cuda.Event()
        self._loss_events[i].record(cs)
        res = StepResult(self, self.step_idx, lr)
        self.step_idx += 1
        return res
