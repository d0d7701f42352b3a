"""Host-side engine of the VAuLT hot path on one B200: owns the flat parameter / gradient / bf16-shadow buffers and enqueues
the sm_100a kernels (through the C ABI, include/vault_b200.h) for the whole forward and backward of

    LM (BERT / RoBERTa, post-LN)  ->  ViLT text embed + im2col-free patch embed + sequence assembly
    ->  12 x ViLT layer (pre-LN)  ->  final LayerNorm  ->  pooler

mirroring ref:vault/models/vault/model.py:151-218 (VaultMixin.lm_preprocess / forward) and the HF modules it drives
(HF:models/vilt/modeling_vilt.py:67-675, HF:models/bert/modeling_bert.py:53-453).  Only kernel launches, pointer arithmetic
and buffer allocation happen here; every FLOP is in vault_b200/csrc.  There is no torch-op fallback.

Precision plan: bf16 GEMM/attention operands with fp32 accumulation; the residual stream, LayerNorm statistics, softmax,
pooler/head and all parameter gradients are fp32; master weights fp32 with a bf16 shadow for the tensor-core operands.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import Dict, List, Optional, Tuple

import torch

from . import _abi
from ._abi import (EPI_ATOMIC_BIAS_DROP_F32, EPI_ATOMIC_F32, EPI_BIAS_BF16, EPI_BIAS_F32, EPI_BIAS_GELU_BF16, EPI_BIAS_GELU_GRAD_BF16, EPI_BIAS_RESID_F32,
                   EPI_MUL_AUX_BF16, EPI_PLAIN_BF16, EPI_STORE_F32, GemmArgs)

ALIGN = 64  # elements; every slot starts on a 256-byte (fp32) / 128-byte (bf16) boundary -> TMA- and float4-safe


def _round_up(x: int, a: int) -> int:
    return (x + a - 1) // a * a


class Slot:
    __slots__ = ("name", "off", "numel", "shape", "trainable")

    def __init__(self, name, off, numel, shape, trainable):
        self.name, self.off, self.numel, self.shape, self.trainable = name, off, numel, shape, trainable


class Tape:
    """Activations saved by a training-mode forward for its backward."""

    def __init__(self):
        self.t: Dict[str, torch.Tensor] = {}
        self.meta: Dict[str, object] = {}
        self.done = False
        self.gen = 0  # gradient generation this forward belongs to (VaultEngine._gen at forward time)


class VaultEngine:
    # dropout site ids: (stack, layer, kind)
    SITE_LM_EMB = 1
    SITE_HEAD = 2

    @staticmethod
    def _site(layer: int, kind: int) -> int:  # kind 0 attn-probs, 1 attn-out, 2 ffn-out
        return 16 + layer * 4 + kind

    def __init__(self, model):
        self.model = model  # the nn.Module (VaultModel / VaultForTMSC / a VaultFor* head wrapper) whose Parameters this engine serves
        # head wrappers (ViltForMaskedLM & co.) keep the ViLT trunk in `.vilt`: its parameter names carry that prefix
        self.vilt = getattr(model, "vilt", model)
        self.vp = "vilt." if self.vilt is not model else ""
        cfg = model.config
        self.H = cfg.hidden_size
        self.L = cfg.num_hidden_layers
        self.heads = cfg.num_attention_heads
        self.I = cfg.intermediate_size
        self.patch = cfg.patch_size
        self.grid = cfg.image_size // cfg.patch_size
        self.C = cfg.num_channels
        self.vilt_eps = float(cfg.layer_norm_eps)
        if self.H // self.heads != 64:
            raise RuntimeError("vault_b200 attention kernels are built for head_dim 64")
        if self.H % 128 != 0 or self.H > 1024:
            raise RuntimeError("vault_b200 LayerNorm kernels need hidden_size in {128,256,512,768,1024}")
        self.lm = getattr(model, "bert", None)
        if self.lm is not None:
            lc = self.lm.config
            if lc.hidden_size != self.H or lc.num_attention_heads != self.heads or lc.intermediate_size != self.I:
                raise RuntimeError("vault_b200 needs LM and ViLT of the same width (bert-base / vilt-b32 family)")
            self.lm_L = lc.num_hidden_layers
            self.lm_eps = float(lc.layer_norm_eps)
            self.lm_p = float(lc.hidden_dropout_prob)
            self.lm_p_attn = float(lc.attention_probs_dropout_prob)
            self.lm_roberta_pad = int(lc.pad_token_id) if lc.model_type in ("roberta", "xlm-roberta", "camembert") else -1
            self.lm_word_pad = -1 if lc.pad_token_id is None else int(lc.pad_token_id)
            self.lm_type_vocab = int(lc.type_vocab_size)
        self.device = None
        self.slots: Dict[str, Slot] = {}
        self._sig = None
        self._versions = None
        self._g = GemmArgs()
        self.seed = 0x5EED5EED
        self.seed_dev: Optional[torch.Tensor] = None  # device counter added to the seed (advanced once per training step)
        # what the kernels read the counter from: seed_dev itself (fused train step: forward and backward in one graph), or the per-forward
        # snapshot a Tape carries (autograd path: several forwards may precede one backward, each must regenerate ITS OWN masks)
        self._seed_buf: Optional[torch.Tensor] = None
        self.sms = 0
        self.gemm_max_ctas = 0  # >0: leave SMs free for a concurrently running collective (data-parallel overlap)
        self.dynamic_tiles = os.environ.get("VAULT_B200_DYNAMIC_TILES", "0") == "1"  # measured neutral on B200: opt-in
        self.wgrad_side_stream = os.environ.get("VAULT_B200_WGRAD_SIDE", "1") != "0"
        self.small_m_split_k = os.environ.get("VAULT_B200_SMALL_M_SPLITK", "1") != "0"
        self.comm_stream = None  # see _cut()
        self._wgrad_group = None
        self.group_wgrads = os.environ.get("VAULT_B200_GROUP_WGRADS", "1") != "0"  # A/B switch: 0 = one launch per weight gradient
        self.prezeroed = False
        self.flat_alloc = None  # optional allocator of the flat master / shadow / gradient buffers (see ensure_packed)
        self.fuse_bias_grad = os.environ.get("VAULT_B200_FUSE_BIAS_GRAD", "1") != "0"  # A/B switch: 0 = separate vault_colsum_bf16 launches
        self.patch_wgrad_tma = os.environ.get("VAULT_B200_PATCH_WGRAD_TMA", "1") != "0"  # 0: bf16 im2col + GEMM (A/B switch)
        self._side = None
        self._side_keep = []
        self._side_dirty = False
        # Several forwards may share one autograd backward (ViltForImagesAndTextClassification runs the trunk once per image):
        # the first backward of a generation zero-fills the accumulated gradient ranges, later ones of the same generation add.
        self._gen = 0
        self._accumulate = False
        self._grads_live = False  # attach_grads() has handed out p.grad views since the last zeroing generation

    # ------------------------------------------------------------------------------------------------------------
    # parameter packing
    # ------------------------------------------------------------------------------------------------------------
    def _param_order(self) -> Tuple[List[str], List[str]]:
        """(trainable names in reverse-topological order, static names).  QKV weights/biases are adjacent so their
        concatenation is one contiguous [3H,H] / [3H] slice (fused QKV GEMM and its wgrad write straight into it)."""
        m = self.model
        named = dict(m.named_parameters())
        order: List[str] = []

        def add(*names):
            for n in names:
                if n in named and n not in order:
                    order.append(n)

        vp = self.vp
        if not vp:
            add("classifier.1.weight", "classifier.1.bias")  # VaultForTMSC head (kernels of head.cu)
        add(vp + "pooler.dense.weight", vp + "pooler.dense.bias", vp + "layernorm.weight", vp + "layernorm.bias")
        for i in reversed(range(self.L)):
            p = f"{vp}encoder.layer.{i}."
            a = p + "attention.attention."
            # accumulated-with-atomics slots first (one memset per layer), then the dense weights their wgrad GEMMs overwrite
            add(a + "query.bias", a + "key.bias", a + "value.bias", p + "attention.output.dense.bias")
            add(p + "layernorm_before.weight", p + "layernorm_before.bias", p + "layernorm_after.weight", p + "layernorm_after.bias")
            add(p + "intermediate.dense.bias", p + "output.dense.bias", p + "attention.output.dense.weight")
            add(a + "query.weight", a + "key.weight", a + "value.weight", p + "intermediate.dense.weight", p + "output.dense.weight")
        e = vp + "embeddings."
        add(e + "cls_token", e + "position_embeddings", e + "token_type_embeddings.weight", e + "patch_embeddings.projection.weight",
            e + "patch_embeddings.projection.bias", e + "text_embeddings.token_type_embeddings.weight",
            e + "text_embeddings.LayerNorm.weight", e + "text_embeddings.LayerNorm.bias",
            e + "text_embeddings.position_embeddings.weight", e + "text_embeddings.word_embeddings.weight")
        if self.lm is not None:
            for i in reversed(range(self.lm_L)):
                p = f"bert.encoder.layer.{i}."
                a = p + "attention.self."
                add(a + "query.bias", a + "key.bias", a + "value.bias", p + "attention.output.dense.bias")
                add(p + "attention.output.LayerNorm.weight", p + "attention.output.LayerNorm.bias", p + "output.LayerNorm.weight",
                    p + "output.LayerNorm.bias")
                add(p + "intermediate.dense.bias", p + "output.dense.bias", p + "attention.output.dense.weight")
                add(a + "query.weight", a + "key.weight", a + "value.weight", p + "intermediate.dense.weight", p + "output.dense.weight")
            b = "bert.embeddings."
            add(b + "word_embeddings.weight", b + "position_embeddings.weight", b + "token_type_embeddings.weight",
                b + "LayerNorm.weight", b + "LayerNorm.bias")
        owned = set(order)  # parameters whose gradients this engine's kernels write
        for n in named:  # anything else (the heads of the VaultFor* wrappers): packed behind the trunk, trained by torch autograd
            add(n)
        never = self.never_grad_names()
        train = [n for n in order if named[n].requires_grad and n not in never and n in owned]
        static = [n for n in order if n not in train]
        return train, static

    def never_grad_names(self):
        """ViLT parameters that receive grad=None whenever an LM is attached."""
        if self.lm is None:
            return set()
        s = {self.vp + "embeddings.text_embeddings.word_embeddings.weight"}
        if not self.use_text_pos():
            s.add(self.vp + "embeddings.text_embeddings.position_embeddings.weight")
        return s

    def use_text_pos(self) -> bool:
        # transformers==4.48.0 gate (HF:models/vilt/modeling_vilt.py:240-272): add position embeddings iff "absolute"
        te = self.vilt.embeddings.text_embeddings
        return getattr(te, "position_embedding_type", "absolute") == "absolute"

    def _signature(self):
        return tuple((id(p), p.requires_grad) for p in self.model.parameters()) + (self.use_text_pos(),)

    def ensure_packed(self, device: torch.device):
        if device.type != "cuda":
            raise RuntimeError("vault_b200 runs on CUDA (sm_100a) only -- there is no CPU path")
        sig = self._signature()
        if self._sig == sig and self.device == device and self._params_in_place():
            return
        _abi.check(_abi.lib().vault_check_device(device.index if device.index is not None else torch.cuda.current_device()), "check_device")
        bad = [n for n, p in self.model.named_parameters() if p.dtype != torch.float32]
        if bad:  # every parameter, once per (re)pack: a partially cast module (model.bert.half(), a bf16-loaded LM) must not be re-typed silently
            raise RuntimeError("vault_b200 keeps fp32 master weights (bf16 tensor-core operands are derived): do not cast the module "
                               f"({len(bad)} non-fp32 parameters, e.g. {bad[0]})")
        self.device = device
        self.sms = torch.cuda.get_device_properties(device).multi_processor_count
        named = dict(self.model.named_parameters())
        train, static = self._param_order()
        off = 0
        slots: Dict[str, Slot] = {}
        for n in train:
            slots[n] = Slot(n, off, named[n].numel(), tuple(named[n].shape), True)
            off += _round_up(named[n].numel(), ALIGN)
        self.n_train = off
        for n in static:
            slots[n] = Slot(n, off, named[n].numel(), tuple(named[n].shape), False)
            off += _round_up(named[n].numel(), ALIGN)
        self.n_total = off
        # flat_alloc(numel, dtype) -> zero-filled 1-D tensor: the data-parallel multicast path (train.py) supplies symmetric memory here
        alloc = self.flat_alloc or (lambda numel, dt: torch.zeros(numel, device=device, dtype=dt))
        master = alloc(self.n_total, torch.float32)
        with torch.no_grad():
            for n, s in slots.items():
                view = master[s.off:s.off + s.numel].view(s.shape)
                view.copy_(named[n].detach())
                named[n].data = view  # the nn.Parameter now aliases the flat master buffer
        self.master = master
        self.shadow = alloc(self.n_total, torch.bfloat16)
        self.grad = alloc(max(self.n_train, ALIGN), torch.float32)
        self.slots = slots
        self._sig = sig
        self._ptrs = {n: named[n].data_ptr() for n in slots}
        self._params = [named[n] for n in slots]
        self._grad_views = [(n, named[n], self.grad[s.off:s.off + s.numel].view(s.shape)) for n, s in slots.items() if s.trainable]
        self._versions = None
        self.seed_dev = torch.zeros(1, device=device, dtype=torch.int64)
        self._seed_buf = self.seed_dev
        # gradient ranges that are ACCUMULATED into (atomics): zero-filled at the start of every backward
        self._zero_ranges = self._compute_zero_ranges()
        self.opt_state = None
        self._side = torch.cuda.Stream(device=device, priority=-1 if os.environ.get("VAULT_B200_SIDE_PRIORITY", "0") != "0" else 0)
        if "VAULT_B200_ATTN_IMPL" in os.environ:  # A/B switch: 1 = mma.sync attention everywhere (default: tcgen05 where the shape allows)
            _abi.set_attn_impl(int(os.environ["VAULT_B200_ATTN_IMPL"]))
        self._sched_slots = 512
        self._sched = torch.zeros(2 * self._sched_slots, device=device, dtype=torch.int32)  # dynamic tile-scheduler counters
        self._sched_i = 0

    def repack(self, device: torch.device, flat_alloc=None):
        """Move the flat buffers to memory from `flat_alloc` (values kept; optimizer state dropped)."""
        self.flat_alloc = flat_alloc
        self._sig = None
        self.ensure_packed(device)

    def shadow_only_bitmap(self) -> torch.Tensor:
        """int32 bitmap over the 64-parameter blocks of the trainable flat range: bit = 1 where the block belongs to a dense projection matrix
        of an encoder layer (q/k/v, attention output, MLP), i.e. a parameter the kernels read through its bf16 shadow ONLY (linear_fwd /
        linear_dgrad take w16; nothing takes w32 of these).  Used by the multicast optimizer step to leave those fp32 masters sharded."""
        import re
        pat = re.compile(r"encoder\.layer\.\d+\.(attention\.(attention|self)\.(query|key|value)|attention\.output\.dense|intermediate\.dense|output\.dense)\.weight$")
        nblk = (self.n_train + ALIGN - 1) // ALIGN
        bits = torch.zeros((nblk + 31) // 32 * 32, dtype=torch.bool)
        for n, sl in self.slots.items():
            if sl.trainable and pat.search(n):
                bits[sl.off // ALIGN:(sl.off + sl.numel + ALIGN - 1) // ALIGN] = True
        w = (bits.view(-1, 32).to(torch.int64) << torch.arange(32, dtype=torch.int64)).sum(1)
        w = torch.where(w >= 2 ** 31, w - 2 ** 32, w).to(torch.int32)
        return w.to(self.device)

    def _params_in_place(self) -> bool:
        named = dict(self.model.named_parameters())
        return all(named[n].data_ptr() == p for n, p in self._ptrs.items()) if len(named) == len(self._ptrs) else False

    def refresh_shadow(self, force: bool = False):
        """bf16 shadow <- fp32 masters, if any Parameter was modified in place since the last refresh."""
        v = sum(p._version for p in self._params)
        if force or v != self._versions:
            _abi.call("vault_cast_f32_bf16", self.master.data_ptr(), self.shadow.data_ptr(), self.n_total, self._stream())
            self._versions = v

    def mark_shadow_fresh(self):
        self._versions = sum(p._version for p in self._params)

    # pointers ---------------------------------------------------------------------------------------------------
    def _k(self, name: str) -> str:
        """Trunk-relative ViLT name -> state-dict key of the served module ("vilt." in front for the head wrappers)."""
        if not self.vp or name.startswith("bert.") or name.startswith("classifier.") or name.startswith(self.vp):
            return name
        return self.vp + name

    def has(self, name: str) -> bool:
        return self._k(name) in self.slots

    def w16(self, name):  # bf16 shadow of a weight
        return self.shadow.data_ptr() + 2 * self.slots[self._k(name)].off

    def w32(self, name):
        return self.master.data_ptr() + 4 * self.slots[self._k(name)].off

    def g32(self, name):  # gradient slot, or 0 (NULL) if the parameter is not trainable
        s = self.slots.get(self._k(name))
        if s is None or not s.trainable:
            return 0
        return self.grad.data_ptr() + 4 * s.off

    def grad_view(self, name) -> Optional[torch.Tensor]:
        s = self.slots[self._k(name)]
        if not s.trainable:
            return None
        return self.grad[s.off:s.off + s.numel].view(s.shape)

    def _wgrad_cfg(self, n_out: int, k_out: int, tokens: int) -> Tuple[int, int]:
        """(tile N, split-K) of a weight-gradient GEMM dW[n_out, k_out] = dy^T x (contraction over `tokens`).  Measured on B200
        (tools/wgrad_sweep.py): 128x256 tiles with the token contraction split over several CTAs so that the launch fills the SMs
        (72 tiles -> 2, 18 tiles -> 8) beat un-split 128x128 tiles by 25-30 %; partial sums meet in the zero-filled fp32 gradient
        slot through red.global.add.v4.f32.  The 2304 x 768 QKV gradient is 54 tiles of 128x256 (x2 = 108 CTAs, 73 % of the SMs):
        128x192 tiles make it 72 x 2 = 144.  Score = SM fill x relative MMA rate of the tile width (as in the kernel's own choice)."""
        nkb = (tokens + 63) // 64
        best = (-1.0, 128, 1)
        for bn, rate in ((256, 1.0), (192, 0.93), (128, 0.85)):
            if bn > 128 and (k_out < bn or (bn == 192 and k_out % 192 != 0)):
                continue
            tiles = ((n_out + 127) // 128) * ((k_out + bn - 1) // bn)
            split = 1
            while split * 2 * tiles <= self.sms and split * 2 <= 8 and nkb // (split * 2) >= 2:
                split *= 2
            waves = -(-(tiles * split) // self.sms)
            score = rate * tiles * split / (waves * self.sms)
            if score > best[0] + 1e-6:
                best = (score, bn, split)
        return best[1], best[2]

    def _compute_zero_ranges(self):
        """Every trainable gradient slot is ACCUMULATED into (split-K weight gradients, bias / LayerNorm / embedding atomics), except the
        pooler and classifier (overwritten by their kernels, and written before the trunk's backward starts): one memset of the
        rest of the flat buffer at the start of backward."""
        lo = None
        for n, s in sorted(self.slots.items(), key=lambda kv: kv[1].off):
            if s.trainable and not (n.startswith("classifier.") or n.startswith(self.vp + "pooler.")):
                lo = s.off
                break
        return [(lo, self.n_train)] if lo is not None and lo < self.n_train else []

    # ------------------------------------------------------------------------------------------------------------
    # kernel launch helpers
    # ------------------------------------------------------------------------------------------------------------
    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def _new(self, shape, dtype) -> torch.Tensor:
        return torch.empty(shape, device=self.device, dtype=dtype)

    def gemm(self, A, lda, a_mn, B, ldb, b_mn, M, N, K, epi, out, ldo, bias=0, resid=0, ldr=0, aux=0, ldaux=0, out2=0, ldo2=0, p=0.0, site=0,
             split_k=1, block_n=0, stream=None, a_colsum=0):
        g = self._g
        g.M, g.N, g.K = M, N, K
        g.A, g.lda, g.a_mn = A, lda, a_mn
        g.B, g.ldb, g.b_mn = B, ldb, b_mn
        g.epilogue = epi
        g.bias, g.resid, g.ldr = bias or None, resid or None, ldr
        g.aux, g.ldaux = aux or None, ldaux
        g.out, g.ldo, g.out2, g.ldo2 = out, ldo, out2 or None, ldo2
        g.dropout_p, g.seed, g.site = p, self.seed, site
        g.seed_dev = self._seed_buf.data_ptr() if p > 0.0 else None
        g.split_k, g.block_n, g.max_ctas = split_k, block_n, self.gemm_max_ctas
        g.a_colsum = a_colsum or None
        if self.dynamic_tiles:
            self._sched_i = (self._sched_i + 1) % self._sched_slots
            g.sched = self._sched.data_ptr() + 8 * self._sched_i
        else:
            g.sched = None
        rc = self._lib.vault_gemm_bf16(C.byref(g), stream if stream is not None else self._st)
        if rc:
            _abi.check(rc, "vault_gemm_bf16")

    def linear_fwd(self, x16, M, wname, bname, N, K, epi, out, **kw):
        """y = x W^T + b : A = x [M,K], B = W [N,K] (bf16 shadow), fp32 bias."""
        self.gemm(x16.data_ptr(), K, 0, self.w16(wname), K, 0, M, N, K, epi, out.data_ptr(), N, bias=self.w32(bname), **kw)

    def linear_dgrad(self, dy16, M, wname, N_out, K_in, epi, out, **kw):
        """dx[M,K_in] = dy[M,N_out] W[N_out,K_in] : contraction over N_out, W read un-transposed as the MN-major operand."""
        self.gemm(dy16.data_ptr(), N_out, 0, self.w16(wname), K_in, 1, M, K_in, N_out, epi, out.data_ptr(), K_in, **kw)

    def linear_wgrad(self, dy16, x16, M, wname, bname, N_out, K_in, bias=True):
        """dW[N_out,K_in] = dy^T x (contraction over the M tokens, both operands read un-transposed), db = colsum(dy)
        (bias=False when the kernel that produced dy already accumulated its column sums)."""
        gw = self.g32(wname)
        gb = self.g32(bname) if bias else 0
        if not gw and not gb:
            return
        if self._wgrad_group is not None and gw and N_out % 2 == 0 and K_in % 8 == 0 and (not gb or self.fuse_bias_grad):
            # inside a layer's backward: queued, launched with the layer's other weight gradients as one grouped GEMM (_wgrad_flush)
            self._wgrad_group.append((dy16, x16, M, gw, gb, N_out, K_in))
            return
        st = self._st
        if self.wgrad_side_stream:
            # weight gradients are off the critical path (nothing downstream in backward reads them): run them on a side stream
            # so they fill the SMs the dgrad / LayerNorm / attention chain leaves idle in its tails.  Joined in _join_side().
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))
            self._side.wait_event(ev)
            st = self._side.cuda_stream
            self._side_keep.append((dy16, x16))  # keep the operands alive until the join
            self._side_dirty = True
        if gw:
            bn, split = self._wgrad_cfg(N_out, K_in, M)  # the slot was zero-filled by zero_accumulated_grads()
            # db = column sums of dy = row sums of the GEMM's MN-major A operand: taken from the A tiles in shared memory by the
            # epilogue warps of the same launch (they idle during a weight gradient's long contraction) instead of a second pass over dy
            fused = gb if (self.fuse_bias_grad and N_out % 2 == 0) else 0
            self.gemm(dy16.data_ptr(), N_out, 1, x16.data_ptr(), K_in, 1, N_out, K_in, M,
                      EPI_ATOMIC_F32 if (split > 1 or self._accumulate) else EPI_STORE_F32, gw, K_in, split_k=split, block_n=bn, stream=st,
                      a_colsum=fused)
            if fused:
                gb = 0
        if gb:
            rc = self._lib.vault_colsum_bf16(dy16.data_ptr(), N_out, gb, M, N_out, st)
            if rc:
                _abi.check(rc, "vault_colsum_bf16")

    def _wgrad_begin(self):
        """Start collecting the weight gradients of one transformer layer (same token count = same contraction)."""
        self._wgrad_group = [] if self.group_wgrads else None

    def _group_split(self, tiles: int, k_blocks: int) -> int:
        """Split-K of a grouped weight-gradient launch: the pooled 128x256 tile count times the split should come out as whole waves of SMs
        (a layer's 216 tiles: x2 = 432 = 2.92 waves of 148); the fewest splits among the best fills, at least two k-blocks per split."""
        best, split = -1.0, 1
        for sk in (1, 2, 4, 8):
            if sk > 1 and k_blocks // sk < 2:
                break
            fill = tiles * sk / (-(-(tiles * sk) // self.sms) * self.sms)
            if fill > best + 0.02:
                best, split = fill, sk
        return split

    def _wgrad_flush(self):
        """The queued weight gradients of a layer as ONE persistent launch over their pooled 128x256 tiles (vault_gemm_wgrad_grouped): one launch
        head / tail instead of four, every CTA's epilogue hidden under its next tile, the SMs filled in whole waves."""
        grp, self._wgrad_group = self._wgrad_group, None
        if not grp:
            return
        st = self._st
        if self.wgrad_side_stream:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))
            self._side.wait_event(ev)
            st = self._side.cuda_stream
            self._side_keep.append(tuple(t for g in grp for t in g[:2]))  # keep the operands alive until the join
            self._side_dirty = True
        for a in range(0, len(grp), 4):
            chunk = grp[a:a + 4]
            tiles = sum(-(-n_out // 128) * -(-k_in // 256) for (_, _, _, _, _, n_out, k_in) in chunk)
            split = self._group_split(tiles, -(-chunk[0][2] // 64))
            arr = (GemmArgs * len(chunk))()
            for g, (dy16, x16, M, gw, gb, n_out, k_in) in zip(arr, chunk):
                g.M, g.N, g.K = n_out, k_in, M
                g.A, g.lda, g.a_mn = dy16.data_ptr(), n_out, 1
                g.B, g.ldb, g.b_mn = x16.data_ptr(), k_in, 1
                g.epilogue, g.out, g.ldo = EPI_ATOMIC_F32, gw, k_in
                g.split_k, g.max_ctas = split, self.gemm_max_ctas
                g.a_colsum = gb or None
            rc = self._lib.vault_gemm_wgrad_grouped(arr, len(chunk), st)
            if rc:
                _abi.check(rc, "vault_gemm_wgrad_grouped")

    def _cut(self):
        """A gradient range is complete once the chain AND the weight-gradient stream get here.  Default: the chain waits for the side
        stream (the caller reduces the range from the host between two graph replays).  With `comm_stream` set (the whole step is ONE
        captured graph, train.py) only that stream waits for both -- the dependency chain does not stop at the cut."""
        if self.comm_stream is None:
            self._join_side()
            return
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self.comm_stream.wait_event(ev)
        if self._side_dirty:
            ev2 = torch.cuda.Event()
            ev2.record(self._side)
            self.comm_stream.wait_event(ev2)

    def _join_side(self):
        """Main stream waits for every weight-gradient kernel issued on the side stream so far."""
        if self._side_dirty:
            ev = torch.cuda.Event()
            ev.record(self._side)
            torch.cuda.current_stream(self.device).wait_event(ev)
            self._side_dirty = False
        self._side_keep = []

    def ln_fwd(self, x32, rows, gname, bname, eps, want16=True, want32=False, p=0.0, site=0):
        y16 = self._new((rows, self.H), torch.bfloat16) if want16 else None
        y32 = self._new((rows, self.H), torch.float32) if want32 else None
        stats = self._new((2, rows), torch.float32)
        rc = self._lib.vault_layernorm_fwd_drop(x32.data_ptr(), self.w32(gname), self.w32(bname), y16.data_ptr() if want16 else None,
                                                y32.data_ptr() if want32 else None, stats.data_ptr(), stats.data_ptr() + 4 * rows, rows, self.H, eps,
                                                p, self.seed, self._seed_buf.data_ptr() if p > 0 else None, site, self._st)
        if rc:
            _abi.check(rc, "vault_layernorm_fwd")
        return y16, y32, stats

    def ln_bwd(self, dy32, dy16, x32, stats, rows, gname, bname, dres32=None, want16=True, in_p=0.0, in_site=0, out_p=0.0, out_site=0,
               colsum_to: Optional[str] = None):
        """colsum_to: name of the bias whose gradient is the column sum of this call's bf16 output (fused, saves a pass)."""
        dx32 = self._new((rows, self.H), torch.float32)
        dx16 = self._new((rows, self.H), torch.bfloat16) if want16 else None
        use_seed = in_p > 0 or out_p > 0
        rc = self._lib.vault_layernorm_bwd_drop(dy32.data_ptr() if dy32 is not None else None, dy16.data_ptr() if dy16 is not None else None,
                                                x32.data_ptr(), stats.data_ptr(), stats.data_ptr() + 4 * rows, self.w32(gname),
                                                dres32.data_ptr() if dres32 is not None else None, dx32.data_ptr(),
                                                dx16.data_ptr() if want16 else None, self.g32(gname) or None, self.g32(bname) or None,
                                                (self.g32(colsum_to) or None) if (colsum_to and want16) else None, rows, self.H,
                                                in_p, in_site, out_p, out_site, self.seed, self._seed_buf.data_ptr() if use_seed else None, self._st)
        if rc:
            _abi.check(rc, "vault_layernorm_bwd")
        return dx32, dx16

    def attn_fwd(self, qkv, key_mask, B, S, p=0.0, site=0, want_lse=True):
        ctx = self._new((B * S, self.H), torch.bfloat16)
        lse = self._new((B, self.heads, S), torch.float32) if want_lse else None
        rc = self._lib.vault_attn_fwd(qkv.data_ptr(), key_mask.data_ptr(), ctx.data_ptr(), lse.data_ptr() if want_lse else None, B, S, self.heads, p,
                                      self.seed, self._seed_buf.data_ptr() if p > 0 else None, site, self._st)
        if rc:
            _abi.check(rc, "vault_attn_fwd")
        return ctx, lse

    def attn_bwd(self, qkv, key_mask, ctx, dctx, lse, B, S, p=0.0, site=0):
        dqkv = self._new((B * S, 3 * self.H), torch.bfloat16)
        delta = self._new((B, self.heads, S), torch.float32)
        rc = self._lib.vault_attn_bwd(qkv.data_ptr(), key_mask.data_ptr(), ctx.data_ptr(), dctx.data_ptr(), lse.data_ptr(), delta.data_ptr(),
                                      dqkv.data_ptr(), B, S, self.heads, p, self.seed, self._seed_buf.data_ptr() if p > 0 else None, site, self._st)
        if rc:
            _abi.check(rc, "vault_attn_bwd")
        return dqkv

    # ------------------------------------------------------------------------------------------------------------
    # transformer layers (shared by the LM and ViLT stacks)
    # ------------------------------------------------------------------------------------------------------------
    def _names(self, prefix: str, i: int, vilt: bool):
        p = f"{prefix}encoder.layer.{i}."
        a = p + ("attention.attention." if vilt else "attention.self.")
        return dict(
            qkv_w=a + "query.weight", qkv_b=a + "query.bias", o_w=p + "attention.output.dense.weight", o_b=p + "attention.output.dense.bias",
            w1=p + "intermediate.dense.weight", b1=p + "intermediate.dense.bias", w2=p + "output.dense.weight", b2=p + "output.dense.bias",
            ln1=p + ("layernorm_before" if vilt else "attention.output.LayerNorm"), ln2=p + ("layernorm_after" if vilt else "output.LayerNorm"),
        )

    def _attn_block_fwd(self, x16, M, B, S, nm, key_mask, p_attn, site, save: Optional[dict], li):
        H = self.H
        qkv = self._new((M, 3 * H), torch.bfloat16)
        self.linear_fwd(x16, M, nm["qkv_w"], nm["qkv_b"], 3 * H, H, EPI_BIAS_BF16, qkv)
        ctx, lse = self.attn_fwd(qkv, key_mask, B, S, p=p_attn, site=site, want_lse=save is not None)
        if save is not None:
            save[f"{li}.qkv"], save[f"{li}.ctx"], save[f"{li}.lse"] = qkv, ctx, lse
        return ctx

    def _small_m_split(self, M: int, n_out: int, k_in: int) -> int:
        """Split-K factor for a GEMM with few output tiles and a long contraction (the LM's K=3072 products at 1,280 tokens run as
        30 tiles of 128x256 on 148 SMs): largest power of two that keeps >= 8 k-blocks per split and fits one wave; 1 = do not split."""
        if not self.small_m_split_k:
            return 1
        tiles = -(-M // 128) * -(-n_out // 256)
        split = 1
        while split < 8 and tiles * split * 2 <= self.sms and k_in // 64 // (split * 2) >= 8:
            split *= 2
        return split

    def _mlp_fwd(self, n16, M, nm, resid32, p_out, site, save: Optional[dict], li, resid_inplace: bool = False):
        """resid_inplace: the caller does not need resid32 afterwards, so a split-K second GEMM may accumulate into it."""
        H, I = self.H, self.I
        act = self._new((M, I), torch.bfloat16)
        # training: the epilogue also writes gelu'(pre-activation) -- it shares the exponential and the tail polynomial with the value -- so the
        # backward epilogue of the W2 dgrad only multiplies by it (saved in place of the pre-activation, which nothing else reads)
        dact = self._new((M, I), torch.bfloat16) if save is not None else None
        if dact is not None:
            self.linear_fwd(n16, M, nm["w1"], nm["b1"], I, H, EPI_BIAS_GELU_GRAD_BF16, act, out2=dact.data_ptr(), ldo2=I)
        else:
            self.linear_fwd(n16, M, nm["w1"], nm["b1"], I, H, EPI_BIAS_GELU_BF16, act)
        # training only: atomic accumulation order makes the result run-to-run different in the last fp32 bit; inference stays bit-reproducible
        split = self._small_m_split(M, H, I) if (resid_inplace and save is not None) else 1
        if split > 1:
            y32 = resid32  # y = resid + dropout(act W2^T + b2), the K range cut in `split` parts that add their share atomically
            self.linear_fwd(act, M, nm["w2"], nm["b2"], H, I, EPI_ATOMIC_BIAS_DROP_F32, y32, p=p_out, site=site, split_k=split, block_n=256)
        else:
            y32 = self._new((M, H), torch.float32)
            self.linear_fwd(act, M, nm["w2"], nm["b2"], H, I, EPI_BIAS_RESID_F32, y32, resid=resid32.data_ptr(), ldr=H, p=p_out, site=site)
        if save is not None:
            save[f"{li}.dact"], save[f"{li}.act"] = dact, act
        return y32

    # ---- ViLT (pre-LN) -----------------------------------------------------------------------------------------
    def vilt_layer_fwd(self, i, x32, M, B, S, key_mask, save):
        nm = self._names("", i, True)
        li = f"v{i}"
        n1, _, st1 = self.ln_fwd(x32, M, nm["ln1"] + ".weight", nm["ln1"] + ".bias", self.vilt_eps)
        ctx = self._attn_block_fwd(n1, M, B, S, nm, key_mask, 0.0, 0, save, li)
        h32 = self._new((M, self.H), torch.float32)
        self.linear_fwd(ctx, M, nm["o_w"], nm["o_b"], self.H, self.H, EPI_BIAS_RESID_F32, h32, resid=x32.data_ptr(), ldr=self.H)
        n2, _, st2 = self.ln_fwd(h32, M, nm["ln2"] + ".weight", nm["ln2"] + ".bias", self.vilt_eps)
        y32 = self._mlp_fwd(n2, M, nm, h32, 0.0, 0, save, li)
        if save is not None:
            save[f"{li}.x"], save[f"{li}.st1"], save[f"{li}.n1"] = x32, st1, n1
            save[f"{li}.h"], save[f"{li}.st2"], save[f"{li}.n2"] = h32, st2, n2
        return y32

    def vilt_layer_bwd(self, i, g32, g16, M, B, S, key_mask, sv):
        nm = self._names("", i, True)
        li = f"v{i}"
        H, I = self.H, self.I
        self._wgrad_begin()
        dpre = self._new((M, I), torch.bfloat16)
        self.linear_dgrad(g16, M, nm["w2"], H, I, EPI_MUL_AUX_BF16, dpre, aux=sv[f"{li}.dact"].data_ptr(), ldaux=I)
        self.linear_wgrad(g16, sv[f"{li}.act"], M, nm["w2"], nm["b2"], H, I, bias=False)  # db2: summed by the kernel that produced g16
        dn2 = self._new((M, H), torch.bfloat16)
        self.linear_dgrad(dpre, M, nm["w1"], I, H, EPI_PLAIN_BF16, dn2)
        self.linear_wgrad(dpre, sv[f"{li}.n2"], M, nm["w1"], nm["b1"], I, H)
        g2_32, g2_16 = self.ln_bwd(None, dn2, sv[f"{li}.h"], sv[f"{li}.st2"], M, nm["ln2"] + ".weight", nm["ln2"] + ".bias", dres32=g32,
                                   colsum_to=nm["o_b"])
        dctx = self._new((M, H), torch.bfloat16)
        self.linear_dgrad(g2_16, M, nm["o_w"], H, H, EPI_PLAIN_BF16, dctx)
        self.linear_wgrad(g2_16, sv[f"{li}.ctx"], M, nm["o_w"], nm["o_b"], H, H, bias=False)
        dqkv = self.attn_bwd(sv[f"{li}.qkv"], key_mask, sv[f"{li}.ctx"], dctx, sv[f"{li}.lse"], B, S)
        dn1 = self._new((M, H), torch.bfloat16)
        self.linear_dgrad(dqkv, M, nm["qkv_w"], 3 * H, H, EPI_PLAIN_BF16, dn1)
        self.linear_wgrad(dqkv, sv[f"{li}.n1"], M, nm["qkv_w"], nm["qkv_b"], 3 * H, H)
        self._wgrad_flush()
        below_b2 = self._names("", i - 1, True)["b2"] if i > 0 else None  # this output is the dy of the layer below's MLP-2
        return self.ln_bwd(None, dn1, sv[f"{li}.x"], sv[f"{li}.st1"], M, nm["ln1"] + ".weight", nm["ln1"] + ".bias", dres32=g2_32, colsum_to=below_b2)

    # ---- LM (post-LN, dropout when training) ---------------------------------------------------------------------
    def lm_layer_fwd(self, i, r32, x16, M, B, T, key_mask, save, train):
        nm = self._names("bert.", i, False)
        li = f"l{i}"
        p, pa = (self.lm_p, self.lm_p_attn) if train else (0.0, 0.0)
        ctx = self._attn_block_fwd(x16, M, B, T, nm, key_mask, pa, self._site(i, 0), save, li)
        t32 = self._new((M, self.H), torch.float32)
        self.linear_fwd(ctx, M, nm["o_w"], nm["o_b"], self.H, self.H, EPI_BIAS_RESID_F32, t32, resid=r32.data_ptr(), ldr=self.H, p=p,
                        site=self._site(i, 1))
        a16, a32, st1 = self.ln_fwd(t32, M, nm["ln1"] + ".weight", nm["ln1"] + ".bias", self.lm_eps, want32=True)
        s32 = self._mlp_fwd(a16, M, nm, a32, p, self._site(i, 2), save, li, resid_inplace=True)  # a32 is only this residual
        y16, y32, st2 = self.ln_fwd(s32, M, nm["ln2"] + ".weight", nm["ln2"] + ".bias", self.lm_eps, want32=True)
        if save is not None:
            save[f"{li}.x16"], save[f"{li}.t"], save[f"{li}.st1"], save[f"{li}.a16"] = x16, t32, st1, a16
            save[f"{li}.s"], save[f"{li}.st2"] = s32, st2
        return y32, y16

    def lm_layer_bwd(self, i, g32, gx16, M, B, T, key_mask, sv, train):
        """g32 (+ gx16) = gradient w.r.t. the layer's output y = LN_b(s).  Returns (dt32, gx16') for the layer below."""
        nm = self._names("bert.", i, False)
        li = f"l{i}"
        H, I = self.H, self.I
        p, pa = (self.lm_p, self.lm_p_attn) if train else (0.0, 0.0)
        self._wgrad_begin()
        ds32, ds16 = self.ln_bwd(g32, gx16, sv[f"{li}.s"], sv[f"{li}.st2"], M, nm["ln2"] + ".weight", nm["ln2"] + ".bias", out_p=p,
                                 out_site=self._site(i, 2), colsum_to=nm["b2"])
        dpre = self._new((M, I), torch.bfloat16)
        self.linear_dgrad(ds16, M, nm["w2"], H, I, EPI_MUL_AUX_BF16, dpre, aux=sv[f"{li}.dact"].data_ptr(), ldaux=I)
        self.linear_wgrad(ds16, sv[f"{li}.act"], M, nm["w2"], nm["b2"], H, I, bias=False)
        split = self._small_m_split(M, H, I)
        if split > 1:
            # da = dpre W1 (contraction over 3072) added straight into the fp32 residual-branch gradient, K range split across CTAs
            da16 = None
            self.linear_dgrad(dpre, M, nm["w1"], I, H, EPI_ATOMIC_F32, ds32, split_k=split, block_n=256)
        else:
            da16 = self._new((M, H), torch.bfloat16)
            self.linear_dgrad(dpre, M, nm["w1"], I, H, EPI_PLAIN_BF16, da16)
        self.linear_wgrad(dpre, sv[f"{li}.a16"], M, nm["w1"], nm["b1"], I, H)
        dt32, dt16 = self.ln_bwd(ds32, da16, sv[f"{li}.t"], sv[f"{li}.st1"], M, nm["ln1"] + ".weight", nm["ln1"] + ".bias", out_p=p,
                                 out_site=self._site(i, 1), colsum_to=nm["o_b"])
        dctx = self._new((M, H), torch.bfloat16)
        self.linear_dgrad(dt16, M, nm["o_w"], H, H, EPI_PLAIN_BF16, dctx)
        self.linear_wgrad(dt16, sv[f"{li}.ctx"], M, nm["o_w"], nm["o_b"], H, H, bias=False)
        dqkv = self.attn_bwd(sv[f"{li}.qkv"], key_mask, sv[f"{li}.ctx"], dctx, sv[f"{li}.lse"], B, T, p=pa, site=self._site(i, 0))
        gx = self._new((M, H), torch.bfloat16)
        self.linear_dgrad(dqkv, M, nm["qkv_w"], 3 * H, H, EPI_PLAIN_BF16, gx)
        self.linear_wgrad(dqkv, sv[f"{li}.x16"], M, nm["qkv_w"], nm["qkv_b"], 3 * H, H)
        self._wgrad_flush()
        return dt32, gx

    # ------------------------------------------------------------------------------------------------------------
    # whole-model forward / backward
    # ------------------------------------------------------------------------------------------------------------
    def patch_hw(self, pixel_mask: Optional[torch.Tensor], B, Hi, Wi) -> Tuple[torch.Tensor, int]:
        """Per-sample valid patch grid on the device and Pmax = max_b h_b*w_b (HF:models/vilt/modeling_vilt.py:95-98,130-136).
        A data-dependent output length needs ONE small device->host read, as in the reference; pass pixel_mask=None (all
        valid) or use TrainStep (static shapes) to avoid it."""
        gh, gw = Hi // self.patch, Wi // self.patch
        hw = self._new((B, 2), torch.int32)
        if pixel_mask is None:
            hw[:, 0] = gh
            hw[:, 1] = gw
            return hw, gh * gw
        if pixel_mask.dtype not in (torch.int64, torch.float32):
            pixel_mask = pixel_mask.to(torch.int64)
        pixel_mask = pixel_mask.contiguous()
        _abi.call("vault_patch_grid", pixel_mask.data_ptr(), int(pixel_mask.dtype == torch.float32), hw.data_ptr(), B, Hi, Wi, self.patch, self._st)
        pmax = int((hw[:, 0] * hw[:, 1]).max().item())
        return hw, pmax

    def forward(self, *args, **kwargs):
        """Returns (last_hidden_state fp32 [B,S,H], pooler_output fp32 [B,H] or None, key_mask uint8 [B,S], tape or None)."""
        it = self.forward_iter(*args, **kwargs)
        try:
            while True:
                next(it)
        except StopIteration as e:
            return e.value

    def forward_iter(self, input_ids, attention_mask, token_type_ids, pixel_values, pixel_mask, image_token_type_idx=1, training=False,
                     need_grad=False, hw: Optional[torch.Tensor] = None, pmax: Optional[int] = None, split_lm: bool = False,
                     image_embeds: Optional[torch.Tensor] = None, inputs_embeds: Optional[torch.Tensor] = None,
                     seed_snapshot: Optional[torch.Tensor] = None, hidden_out: Optional[list] = None):
        """Generator form of forward.  hidden_out: a list that receives detached copies of the ViLT encoder's hidden states (embedding output,
        then the output of every layer: what ViltEncoder collects with output_hidden_states=True, HF:models/vilt/modeling_vilt.py:505-540).  With split_lm it yields "lm_done" once the language model's forward is enqueued and before
        anything reads a ViLT parameter, so a caller can start the LM while the previous step's AdamW is still updating the
        ViLT range (VaultTrainStep); the return value (StopIteration.value) is forward()'s tuple."""
        embeds_mode = image_embeds is not None  # pre-embedded image tokens [B,P,H] (+ pixel_mask flattened to [B,P]) instead of pixels
        dev = image_embeds.device if embeds_mode else pixel_values.device
        self.ensure_packed(dev)
        self._lib, self._st = _abi.lib(), self._stream()
        # seed_snapshot: device copy of seed_dev taken for THIS forward; its backward (and the head's) regenerate the masks from it even
        # if later forwards have advanced the live counter in between
        self._seed_buf = seed_snapshot if seed_snapshot is not None else self.seed_dev
        if not split_lm and not torch.cuda.is_current_stream_capturing():
            # an optimizer update issued on the side stream (VaultTrainStep) must have landed before weights are read here
            torch.cuda.current_stream(dev).wait_stream(self._side)
        self.refresh_shadow()
        H = self.H
        # text inputs_embeds [B,T,H] with input_ids=None (ref:vault/models/vault/model.py:170-200): they replace the LM's word-embedding lookup
        # (without an LM: ViLT's own), HF:models/bert/modeling_bert.py:72-112 / RobertaEmbeddings with sequential position ids
        text_embeds_mode = inputs_embeds is not None
        if text_embeds_mode:
            if input_ids is not None:
                raise ValueError("You cannot specify both input_ids and inputs_embeds at the same time")
            if inputs_embeds.dim() != 3 or inputs_embeds.shape[2] != H:
                raise ValueError(f"inputs_embeds {tuple(inputs_embeds.shape)}: need [B, T, {H}]")
            B, T = inputs_embeds.shape[:2]
            inputs_embeds = inputs_embeds.contiguous().float()
        else:
            B, T = input_ids.shape
        if embeds_mode:
            if image_embeds.dim() != 3 or image_embeds.shape[0] != B or image_embeds.shape[2] != self.H:
                raise ValueError(f"image_embeds {tuple(image_embeds.shape)}: need [B={B}, P, {self.H}]")
            image_embeds = image_embeds.contiguous().float()
            P_img = image_embeds.shape[1]
            img_mask = None
            if pixel_mask is not None:
                img_mask = (pixel_mask.reshape(B, -1) != 0).to(torch.uint8).contiguous()  # HF: image_masks = pixel_mask.flatten(1)
                if img_mask.shape[1] != P_img:
                    raise ValueError(f"pixel_mask flattens to {img_mask.shape[1]} positions, image_embeds has {P_img}")
            Cc = Hi = Wi = gh = gw = 0
        else:
            Cc, Hi, Wi = pixel_values.shape[1:]
            if Cc != self.C or Hi % self.patch or Wi % self.patch:
                raise ValueError(f"pixel_values {tuple(pixel_values.shape)}: need {self.C} channels and sides divisible by {self.patch}")
            gh, gw = Hi // self.patch, Wi // self.patch
        if input_ids is not None:
            input_ids = input_ids.contiguous()
            if input_ids.dtype != torch.int64:
                input_ids = input_ids.to(torch.int64)
        if attention_mask is not None:
            attention_mask = attention_mask.contiguous() if attention_mask.dtype == torch.int64 else attention_mask.to(torch.int64)
        if token_type_ids is not None:
            token_type_ids = token_type_ids.contiguous() if token_type_ids.dtype == torch.int64 else token_type_ids.to(torch.int64)
        if not embeds_mode:
            pixel_values = pixel_values.contiguous() if pixel_values.dtype == torch.float32 else pixel_values.float().contiguous()
        tape = Tape() if need_grad else None
        if tape is not None:
            tape.gen = self._gen
            tape.meta["seed_buf"] = self._seed_buf
        sv = tape.t if tape is not None else None
        Mt = B * T
        am_ptr = attention_mask.data_ptr() if attention_mask is not None else None
        tt_ptr = token_type_ids.data_ptr() if token_type_ids is not None else None
        lib, st = self._lib, self._st

        G = gh * gw
        Kp = self.C * self.patch * self.patch
        # ---------------- image branch on the side stream: independent of the (latency-bound) LM forward ----------------
        def image_branch():
            nonlocal st
            main_st = st
            ev0 = torch.cuda.Event()
            ev0.record(torch.cuda.current_stream(self.device))
            self._side.wait_event(ev0)
            st = self._st = self._side.cuda_stream
            patch_out = self._new((B * G, H), torch.float32)
            if self.patch == 32:
                # im2col-free: TF32 tcgen05 GEMM fed by 5-D TMA boxes over the raw NCHW pixels, fp32 master weight
                _abi.check(lib.vault_patch_embed_fwd(pixel_values.data_ptr(), self.w32("embeddings.patch_embeddings.projection.weight"),
                                                     self.w32("embeddings.patch_embeddings.projection.bias"), patch_out.data_ptr(), B, self.C, Hi, Wi,
                                                     self.patch, H, st), "patch_embed_fwd")
                patches = None
            else:
                patches = self._new((B * G, Kp), torch.bfloat16)
                _abi.check(lib.vault_patchify_bf16(pixel_values.data_ptr(), patches.data_ptr(), B, self.C, Hi, Wi, self.patch, st), "patchify")
                self.gemm(patches.data_ptr(), Kp, 0, self.w16("embeddings.patch_embeddings.projection.weight"), Kp, 0, B * G, H, Kp, EPI_BIAS_F32,
                          patch_out.data_ptr(), H, bias=self.w32("embeddings.patch_embeddings.projection.bias"))
            if (sv is not None and patches is None and self.g32("embeddings.patch_embeddings.projection.weight")
                    and not (self.patch_wgrad_tma and lib.vault_patch_embed_wgrad_ok(self.C, Hi, Wi, self.patch, H))):
                # patch grids the im2col-free weight-gradient kernel does not take (a patch row of more than 64 patches; other patch sizes): the
                # projection's wgrad (dW = dpatch^T * patches) then reads a bf16 patch matrix as its MN-major B operand (training only)
                patches = self._new((B * G, Kp), torch.bfloat16)
                _abi.check(lib.vault_patchify_bf16(pixel_values.data_ptr(), patches.data_ptr(), B, self.C, Hi, Wi, self.patch, st), "patchify")
            ev1 = torch.cuda.Event()
            ev1.record(self._side)
            st = self._st = main_st
            return patch_out, patches, ev1

        img = None if (split_lm or embeds_mode) else image_branch()
        # ---------------- text: LM or ViLT word embeddings -> inputs_embeds fp32 [Mt,H] ----------------
        lm_trains = self.lm is not None and not getattr(self.model, "freeze_lm", False) and need_grad
        if self.lm is not None:
            # ref:vault/models/vault/model.py:174-180: the LM sees zero type ids when its type vocabulary has < 2 entries
            lm_tt = None if self.lm_type_vocab < 2 else tt_ptr
            lm_train_mode = training  # ref :189 only disables grad for a frozen LM; dropout follows module.training
            lsv = sv if lm_trains else None
            x_sum = self._new((Mt, H), torch.float32)
            lm_mask = self._new((B, T), torch.uint8)
            if text_embeds_mode:
                pos_off = self._lm_embeds_pos_offset(T)
                _abi.check(lib.vault_vilt_text_embed_fwd(inputs_embeds.data_ptr(), lm_tt, self.w32("bert.embeddings.token_type_embeddings.weight"),
                                                         self.w32("bert.embeddings.position_embeddings.weight") + 4 * H * pos_off, x_sum.data_ptr(),
                                                         B, T, H, st), "lm_embeds_fwd")
                if attention_mask is None:
                    lm_mask.fill_(1)
                else:
                    lm_mask.copy_(attention_mask != 0)
            else:
                _abi.check(lib.vault_lm_embed_fwd(input_ids.data_ptr(), lm_tt, self.w32("bert.embeddings.word_embeddings.weight"),
                                                  self.w32("bert.embeddings.token_type_embeddings.weight"),
                                                  self.w32("bert.embeddings.position_embeddings.weight"), x_sum.data_ptr(), am_ptr, lm_mask.data_ptr(), B, T,
                                                  H, self.lm_roberta_pad, st), "lm_embed_fwd")
            p_emb = self.lm_p if lm_train_mode else 0.0
            x16, r32, st0 = self.ln_fwd(x_sum, Mt, "bert.embeddings.LayerNorm.weight", "bert.embeddings.LayerNorm.bias", self.lm_eps, want32=True,
                                        p=p_emb, site=self.SITE_LM_EMB)
            if lsv is not None:
                lsv["lm.x_sum"], lsv["lm.st0"], lsv["lm.mask"] = x_sum, st0, lm_mask
            for i in range(self.lm_L):
                r32, x16 = self.lm_layer_fwd(i, r32, x16, Mt, B, T, lm_mask, lsv, lm_train_mode)
            lm_out = r32
            if split_lm:
                yield "lm_done"
            text_pos = self.w32("embeddings.text_embeddings.position_embeddings.weight") if self.use_text_pos() else None
            v_sum = self._new((Mt, H), torch.float32)
            _abi.check(lib.vault_vilt_text_embed_fwd(lm_out.data_ptr(), tt_ptr, self.w32("embeddings.text_embeddings.token_type_embeddings.weight"),
                                                     text_pos, v_sum.data_ptr(), B, T, H, st), "vilt_text_embed_fwd")
        elif text_embeds_mode:
            text_pos = self.w32("embeddings.text_embeddings.position_embeddings.weight") if self.use_text_pos() else None
            v_sum = self._new((Mt, H), torch.float32)
            _abi.check(lib.vault_vilt_text_embed_fwd(inputs_embeds.data_ptr(), tt_ptr, self.w32("embeddings.text_embeddings.token_type_embeddings.weight"),
                                                     text_pos, v_sum.data_ptr(), B, T, H, st), "vilt_text_embed_fwd")
        else:
            v_sum = self._new((Mt, H), torch.float32)
            _abi.check(lib.vault_lm_embed_fwd(input_ids.data_ptr(), tt_ptr, self.w32("embeddings.text_embeddings.word_embeddings.weight"),
                                              self.w32("embeddings.text_embeddings.token_type_embeddings.weight"),
                                              self.w32("embeddings.text_embeddings.position_embeddings.weight"), v_sum.data_ptr(), None, None, B, T, H,
                                              -1, st), "vilt_word_embed_fwd")
        _, text_ln, st_t = self.ln_fwd(v_sum, Mt, "embeddings.text_embeddings.LayerNorm.weight", "embeddings.text_embeddings.LayerNorm.bias",
                                       self.vilt_eps, want16=False, want32=True)

        # ---------------- assembly (joins the image branch) ----------------
        if embeds_mode:
            S = T + P_img
            M = B * S
            X = self._new((B, S, H), torch.float32)
            key_mask = self._new((B, S), torch.uint8)
            _abi.check(lib.vault_vilt_assemble_embeds_fwd(text_ln.data_ptr(), image_embeds.data_ptr(), self.w32("embeddings.token_type_embeddings.weight"),
                                                          am_ptr, img_mask.data_ptr() if img_mask is not None else None, X.data_ptr(), key_mask.data_ptr(),
                                                          B, T, P_img, H, int(image_token_type_idx), st), "vilt_assemble_embeds_fwd")
            pmax, patches = P_img, None
        else:
            if img is None:
                img = image_branch()
            patch_out, patches, ev1 = img
            torch.cuda.current_stream(self.device).wait_event(ev1)
            if hw is None:
                hw, pmax = self.patch_hw(pixel_mask, B, Hi, Wi)
            S = T + 1 + pmax
            M = B * S
            X = self._new((B, S, H), torch.float32)
            key_mask = self._new((B, S), torch.uint8)
            _abi.check(lib.vault_vilt_assemble_fwd(text_ln.data_ptr(), patch_out.data_ptr(), self.w32("embeddings.cls_token"),
                                                   self.w32("embeddings.position_embeddings"), self.w32("embeddings.token_type_embeddings.weight"), am_ptr,
                                                   hw.data_ptr(), X.data_ptr(), key_mask.data_ptr(), B, T, pmax, gh, gw, self.grid, H,
                                                   int(image_token_type_idx), st), "vilt_assemble_fwd")
        if sv is not None:
            sv["v_sum"], sv["st_t"], sv["patches"], sv["hw"], sv["key_mask"] = v_sum, st_t, patches, hw, key_mask
            sv["ids"], sv["tt"], sv["am"] = input_ids, token_type_ids, attention_mask
            sv["pixels"] = None if embeds_mode else pixel_values  # the patch projection's weight gradient reads them through TMA (no im2col)
            tape.meta.update(B=B, T=T, S=S, pmax=pmax, gh=gh, gw=gw, Hi=Hi, Wi=Wi, img_type=int(image_token_type_idx), training=training,
                             lm_trains=lm_trains, embeds_mode=embeds_mode, text_embeds_mode=text_embeds_mode)

        # ---------------- ViLT encoder, final LN, pooler ----------------
        x32 = X.view(M, H)
        for i in range(self.L):
            if hidden_out is not None:
                hidden_out.append(x32.view(B, S, H).clone())
            x32 = self.vilt_layer_fwd(i, x32, M, B, S, key_mask, sv)
        if hidden_out is not None:
            hidden_out.append(x32.view(B, S, H).clone())
        _, lhs, st_f = self.ln_fwd(x32, M, "layernorm.weight", "layernorm.bias", self.vilt_eps, want16=False, want32=True)
        pooled = None
        if self.has("pooler.dense.weight"):
            pooled = self._new((B, H), torch.float32)
            _abi.check(lib.vault_small_linear_fwd(lhs.data_ptr(), S * H, self.w32("pooler.dense.weight"), self.w32("pooler.dense.bias"),
                                                  pooled.data_ptr(), B, H, H, 1, st), "pooler_fwd")
        if sv is not None:
            sv["x_final"], sv["st_f"], sv["lhs"], sv["pooled"] = x32, st_f, lhs, pooled
        return lhs.view(B, S, H), pooled, key_mask, tape

    def attach_grads(self, exclude_prefix: Optional[str] = None, only_prefix: Optional[str] = None):
        """p.grad <- view of the flat gradient buffer (no copy).  A foreign p.grad tensor is accumulated into instead."""
        self._grads_live = True
        for n, p, view in self._grad_views:
            if only_prefix is not None and not n.startswith(only_prefix):
                continue
            if exclude_prefix is not None and n.startswith(exclude_prefix):
                continue
            if p.grad is None:
                p.grad = view
            elif p.grad.data_ptr() != view.data_ptr():
                p.grad.add_(view)

    def zero_accumulated_grads(self):
        for a, b in self._zero_ranges:
            self.grad[a:b].zero_()

    def grads_pending(self, prefix: Optional[str] = None, exclude_prefix: Optional[str] = None) -> bool:
        """True if some served Parameter's .grad still IS its view of the flat buffer, i.e. the caller has not dropped the previous
        backward's gradients (no zero_grad(set_to_none=True)): torch semantics then are to ACCUMULATE into them.  The next
        backward therefore adds (atomic weight-gradient epilogues, scratch copies for the overwriting kernels) instead of zero-filling;
        after zero_grad(set_to_none=False) that adds into zeros, which is the same thing."""
        if not self._grads_live:
            return False
        for n, p, v in self._grad_views:
            if prefix is not None and not n.startswith(prefix):
                continue
            if exclude_prefix is not None and n.startswith(exclude_prefix):
                continue
            if p.grad is not None and p.grad.data_ptr() == v.data_ptr():
                return True
        return False

    def __getstate__(self):
        raise TypeError("VaultEngine holds device pointers and is not picklable; pickle / deepcopy the model instead (its engine is rebuilt lazily)")

    def backward(self, tape: Tape, dlhs: Optional[torch.Tensor], dpooled: Optional[torch.Tensor]):
        """Fills self.grad (fp32, flat) with dL/dparam for every trainable parameter; returns nothing."""
        for _ in self.backward_iter(tape, dlhs, dpooled, segments=False):
            pass

    def _first_off(self, prefix: str) -> Optional[int]:
        prefix = self._k(prefix)
        offs = [s.off for n, s in self.slots.items() if s.trainable and n.startswith(prefix)]
        return min(offs) if offs else None

    def backward_iter(self, tape: Tape, dlhs: Optional[torch.Tensor], dpooled: Optional[torch.Tensor], segments: bool = True):
        """Generator form of backward: yields the flat-gradient offset up to which gradients are FINAL each time a segment of
        the reverse-topological layout completes (upper ViLT half, all of ViLT, upper LM half), so a data-parallel caller
        can all-reduce that range while the rest of backward runs.  Exhausting it completes the backward."""
        if tape.done:
            raise RuntimeError("vault_b200: backward called twice on the same forward (activations already released)")
        self._lib, self._st = _abi.lib(), self._stream()
        lib, st = self._lib, self._st
        self._wgrad_group = None  # (a backward that raised mid-layer must not leave its queue behind)
        sv, mt = tape.t, tape.meta
        self._seed_buf = mt.get("seed_buf", self.seed_dev)
        B, T, S, pmax, gh, gw = mt["B"], mt["T"], mt["S"], mt["pmax"], mt["gh"], mt["gw"]
        H, M, Mt = self.H, B * S, B * T
        first = tape.gen >= self._gen
        pending = first and self.grads_pending(exclude_prefix="classifier.")
        if first:
            if not pending:
                if self.prezeroed:
                    self.prezeroed = False  # the caller zero-filled the accumulated ranges already (train.py: on the side stream, under the forward)
                else:
                    self.zero_accumulated_grads()
            self._gen = tape.gen + 1
        self._accumulate = (not first) or pending
        if dlhs is not None:
            g_lhs = dlhs.contiguous().float().clone().view(M, H)
        else:
            g_lhs = torch.zeros((M, H), device=self.device, dtype=torch.float32)
        if dpooled is not None and sv["pooled"] is not None:
            dpooled = dpooled.contiguous().float()
            pw, pb = self.g32("pooler.dense.weight"), self.g32("pooler.dense.bias")
            tmp = None
            if self._accumulate and pw:  # the kernel overwrites dW / db: route a later pass of the same generation through a scratch copy
                tmp = (torch.empty((H, H), device=self.device), torch.empty((H,), device=self.device))
                pw, pb = tmp[0].data_ptr(), tmp[1].data_ptr()
            _abi.check(lib.vault_small_linear_bwd(dpooled.data_ptr(), sv["pooled"].data_ptr(), sv["lhs"].data_ptr(), S * H, self.w32("pooler.dense.weight"),
                                                  g_lhs.data_ptr(), S * H, 1, pw or None, pb or None, B, H, H, 1, st), "pooler_bwd")
            if tmp is not None:
                self.grad_view("pooler.dense.weight").add_(tmp[0])
                self.grad_view("pooler.dense.bias").add_(tmp[1])
        g32, g16 = self.ln_bwd(g_lhs, None, sv["x_final"], sv["st_f"], M, "layernorm.weight", "layernorm.bias",
                               colsum_to=self._names("", self.L - 1, True)["b2"])
        for i in reversed(range(self.L)):
            g32, g16 = self.vilt_layer_bwd(i, g32, g16, M, B, S, sv["key_mask"], sv)
            if segments and i > 0 and i in (self.L * 2 // 3, self.L // 3):
                off = self._first_off(f"encoder.layer.{i - 1}.")
                if off:
                    self._cut()
                    yield off
        # ---- embeddings ----
        dtext_ln = self._new((Mt, H), torch.float32)
        if mt.get("embeds_mode"):
            d_img = self._new((B, pmax, H), torch.float32)  # gradient w.r.t. the caller's image_embeds (returned through autograd)
            _abi.check(lib.vault_vilt_assemble_embeds_bwd(g32.data_ptr(), dtext_ln.data_ptr(), d_img.data_ptr(),
                                                          self.g32("embeddings.token_type_embeddings.weight") or None, B, T, pmax, H, mt["img_type"], st),
                       "vilt_assemble_embeds_bwd")
            tape.meta["d_image_embeds"] = d_img
        else:
            gw_ptr = self.g32("embeddings.patch_embeddings.projection.weight")
            tma_wgrad = bool(gw_ptr) and sv["patches"] is None and sv.get("pixels") is not None
            dpatch = self._new((B * gh * gw, H), torch.bfloat16) if (sv["patches"] is not None) else None
            _abi.check(lib.vault_vilt_assemble_bwd(g32.data_ptr(), sv["hw"].data_ptr(), dtext_ln.data_ptr(), dpatch.data_ptr() if dpatch is not None else None,
                                                   self.g32("embeddings.cls_token") or None, self.g32("embeddings.position_embeddings") or None,
                                                   self.g32("embeddings.token_type_embeddings.weight") or None, B, T, pmax, gh, gw, self.grid, H,
                                                   mt["img_type"], st), "vilt_assemble_bwd")
            Kp = self.C * self.patch * self.patch
            if sv["patches"] is not None:
                self.linear_wgrad(dpatch, sv["patches"], B * gh * gw, "embeddings.patch_embeddings.projection.weight",
                                  "embeddings.patch_embeddings.projection.bias", H, Kp)
            elif tma_wgrad:
                # im2col-free: dW += dpatch^T * pixels with the pixel operand taken through the forward's 5-D TMA map (TF32 tcgen05, contraction
                # over the patches); fp32 patch-gradient rows + the bias gradient by one gather kernel.  Off the critical path: side stream.
                dp32 = self._new((B * gh * gw, H), torch.float32)
                wst = st
                if self.wgrad_side_stream:
                    ev = torch.cuda.Event()
                    ev.record(torch.cuda.current_stream(self.device))
                    self._side.wait_event(ev)
                    wst = self._side.cuda_stream
                    self._side_keep.append((dp32, sv["pixels"], g32))
                    self._side_dirty = True
                _abi.check(lib.vault_patch_grad_rows_f32(g32.data_ptr(), sv["hw"].data_ptr(), dp32.data_ptr(),
                                                         self.g32("embeddings.patch_embeddings.projection.bias") or None, B, T, pmax, gh, gw, H, wst),
                           "patch_grad_rows_f32")
                _abi.check(lib.vault_patch_embed_wgrad(sv["pixels"].data_ptr(), dp32.data_ptr(), gw_ptr, B, self.C, mt["Hi"], mt["Wi"], self.patch, H, wst),
                           "patch_embed_wgrad")
        dv_sum, _ = self.ln_bwd(dtext_ln, None, sv["v_sum"], sv["st_t"], Mt, "embeddings.text_embeddings.LayerNorm.weight",
                                "embeddings.text_embeddings.LayerNorm.bias", want16=False)
        tt_ptr = sv["tt"].data_ptr() if sv["tt"] is not None else None
        if self.lm is not None:
            gpos = self.g32("embeddings.text_embeddings.position_embeddings.weight") if self.use_text_pos() else 0
            _abi.check(lib.vault_vilt_text_embed_bwd(tt_ptr, dv_sum.data_ptr(), self.g32("embeddings.text_embeddings.token_type_embeddings.weight") or None,
                                                     gpos or None, B, T, H, st), "vilt_text_embed_bwd")
            if mt["lm_trains"]:
                if segments:
                    off = self._first_off("bert.")
                    if off:
                        self._cut()
                        yield off
                yield from self._lm_backward(dv_sum, sv, B, T, mt["training"], segments, meta=mt)
        elif mt.get("text_embeds_mode"):
            gpos = self.g32("embeddings.text_embeddings.position_embeddings.weight") if self.use_text_pos() else 0
            _abi.check(lib.vault_vilt_text_embed_bwd(tt_ptr, dv_sum.data_ptr(), self.g32("embeddings.text_embeddings.token_type_embeddings.weight") or None,
                                                     gpos or None, B, T, H, st), "vilt_text_embed_bwd")
            mt["d_inputs_embeds"] = dv_sum.view(B, T, H)  # gradient w.r.t. the caller's inputs_embeds (returned through autograd)
        else:
            _abi.check(lib.vault_lm_embed_bwd(sv["ids"].data_ptr(), tt_ptr, dv_sum.data_ptr(),
                                              self.g32("embeddings.text_embeddings.word_embeddings.weight") or None,
                                              self.g32("embeddings.text_embeddings.token_type_embeddings.weight") or None,
                                              self.g32("embeddings.text_embeddings.position_embeddings.weight") or None, B, T, H, -1,
                                              int(getattr(self.model.config, "pad_token_id", -1) if getattr(self.model.config, "pad_token_id", None) is not None else -1),
                                              st), "vilt_word_embed_bwd")
        self._join_side()
        tape.done = True
        tape.t = {}

    def _lm_embeds_pos_offset(self, T: int) -> int:
        """First position-table row used with text inputs_embeds: 0 for BERT, pad+1 for RoBERTa (create_position_ids_from_inputs_embeds)."""
        off = self.lm_roberta_pad + 1 if self.lm_roberta_pad >= 0 else 0
        rows = self.lm.embeddings.position_embeddings.weight.shape[0]
        if off + T > rows:
            raise ValueError(f"inputs_embeds: {T} tokens need position rows {off}..{off + T - 1}, the LM's table has {rows}")
        return off

    def _lm_backward(self, g32, sv, B, T, train, segments=False, meta=None):
        lib, st = self._lib, self._st
        Mt, H = B * T, self.H
        gx16 = None
        for i in reversed(range(self.lm_L)):
            g32, gx16 = self.lm_layer_bwd(i, g32, gx16, Mt, B, T, sv["lm.mask"], sv, train)
            if segments and i > 0 and i in (self.lm_L * 2 // 3, self.lm_L // 3):
                off = self._first_off(f"bert.encoder.layer.{i - 1}.")
                if off:
                    self._cut()
                    yield off
        if segments:
            off = self._first_off("bert.embeddings.")
            if off:
                self._cut()
                yield off
        p_emb = self.lm_p if train else 0.0
        dx_sum, _ = self.ln_bwd(g32, gx16, sv["lm.x_sum"], sv["lm.st0"], Mt, "bert.embeddings.LayerNorm.weight", "bert.embeddings.LayerNorm.bias",
                                want16=False, in_p=p_emb, in_site=self.SITE_LM_EMB)
        lm_tt = None if (self.lm_type_vocab < 2 or sv["tt"] is None) else sv["tt"].data_ptr()
        if meta is not None and meta.get("text_embeds_mode"):
            gpos = self.g32("bert.embeddings.position_embeddings.weight")
            _abi.check(lib.vault_vilt_text_embed_bwd(lm_tt, dx_sum.data_ptr(), self.g32("bert.embeddings.token_type_embeddings.weight") or None,
                                                     (gpos + 4 * H * self._lm_embeds_pos_offset(T)) if gpos else None, B, T, H, st), "lm_embeds_bwd")
            meta["d_inputs_embeds"] = dx_sum.view(B, T, H)  # gradient w.r.t. the caller's inputs_embeds (returned through autograd)
            return
        _abi.check(lib.vault_lm_embed_bwd(sv["ids"].data_ptr(), lm_tt, dx_sum.data_ptr(), self.g32("bert.embeddings.word_embeddings.weight") or None,
                                          self.g32("bert.embeddings.token_type_embeddings.weight") or None,
                                          self.g32("bert.embeddings.position_embeddings.weight") or None, B, T, H, self.lm_roberta_pad,
                                          self.lm_word_pad, st), "lm_embed_bwd")

    # ------------------------------------------------------------------------------------------------------------
    # fused optimizer (transformers==4.48.0 AdamW rule) over the flat trainable range
    # ------------------------------------------------------------------------------------------------------------
    def adamw_range(self, lo: int, hi: int, step: int, lr: float, beta1, beta2, eps, weight_decay, correct_bias, grad_scale, sched_dev, stream: int,
                    grad16: Optional[torch.Tensor] = None):
        """AdamW over the flat sub-range [lo, hi) on an explicit stream (per-segment updates overlapped with the rest of backward).
        grad16: bf16 copy of the flat gradient buffer to read instead of the fp32 one (data-parallel bf16 all-reduce)."""
        s = self.opt_state
        gptr = (grad16.data_ptr() + 2 * lo) if grad16 is not None else (self.grad.data_ptr() + 4 * lo)
        _abi.call("vault_adamw_step", self.master.data_ptr() + 4 * lo, gptr, int(grad16 is not None), s["m"].data_ptr() + 4 * lo,
                  s["v"].data_ptr() + 4 * lo, self.shadow.data_ptr() + 2 * lo, hi - lo, lr, beta1, beta2, eps, weight_decay, int(correct_bias),
                  max(1, step), grad_scale, sched_dev.data_ptr() if sched_dev is not None else None, stream)

    def init_opt_state(self):
        if self.opt_state is None:
            self.opt_state = dict(step=0, m=torch.zeros(self.n_train, device=self.device), v=torch.zeros(self.n_train, device=self.device))
        return self.opt_state

    def adamw_step(self, lr: float, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.0, correct_bias=False, grad_scale=1.0,
                   sched_dev: Optional[torch.Tensor] = None):
        """One fused launch over the whole trainable range.  With ``sched_dev`` (device float[2] = {step_size, lr*wd}) the
        scalars are read on the device (CUDA-graph friendly); otherwise they are computed here from lr / step."""
        s = self.init_opt_state()
        s["step"] += 1
        _abi.call("vault_adamw_step", self.master.data_ptr(), self.grad.data_ptr(), 0, s["m"].data_ptr(), s["v"].data_ptr(), self.shadow.data_ptr(),
                  self.n_train, lr, beta1, beta2, eps, weight_decay, int(correct_bias), max(1, s["step"]), grad_scale,
                  sched_dev.data_ptr() if sched_dev is not None else None, self._stream())
