"""Deterministic synthetic weights and inputs for the VAuLT hot path.

TEST INFRASTRUCTURE (oracle side).  Nothing under ``vault_b200/`` imports this
module; only ``tests/``, ``bench.py`` (``cpu_baseline`` / ``--impl reference``
legs and the synthetic-input generator) and ``__graft_entry__.smoke()`` do.

The reference (gchochla/VAuLT) has no tests and no golden vectors, and no
pretrained checkpoints are reachable from this image (no network), so every
parity case is built from

* ``Dims``          -- the shape of a (ViLT, LM) pair; ``Dims.base()`` is
                        ``ViltConfig()`` + ``BertConfig()`` i.e. vilt-b32 +
                        bert-base-uncased,
* ``make_state_dict`` -- weights with the *HuggingFace state_dict keys* the
                        reference classes carry, drawn
                        per key from a generator seeded by a hash of the key
                        so the result does not depend on iteration order,
* ``make_inputs``   -- ids / masks / pixels of the BASELINE shapes.

HF's random init leaves ``cls_token``, ``position_embeddings`` and every bias at
zero (HF:models/vilt/modeling_vilt.py:82,85) which would leave the pos-embed
interpolation and every bias epilogue untested, so here they are N(0, 0.02).
"""
from __future__ import annotations

import dataclasses
import hashlib
from typing import Dict, Optional

import torch


@dataclasses.dataclass(frozen=True)
class Dims:
    # ViLT trunk (HF ViltConfig defaults = dandelin/vilt-b32-*)
    hidden: int = 768
    layers: int = 12
    heads: int = 12
    inter: int = 3072
    patch: int = 32
    image_size: int = 384
    channels: int = 3
    vilt_vocab: int = 30522
    vilt_max_pos: int = 40
    vilt_type_vocab: int = 2
    modality_vocab: int = 2
    vilt_eps: float = 1e-12
    # language model (HF BertConfig defaults = bert-base-uncased); lm_layers=0 -> no LM attached
    lm_kind: str = "bert"  # "bert" | "roberta"
    lm_layers: int = 12
    lm_vocab: int = 30522
    lm_max_pos: int = 512
    lm_type_vocab: int = 2
    lm_pad_id: int = 0
    lm_eps: float = 1e-12
    lm_dropout: float = 0.1
    # head
    n_classes: int = 3
    head_dropout: float = 0.1

    @property
    def head_dim(self) -> int:
        return self.hidden // self.heads

    @property
    def grid(self) -> int:
        return self.image_size // self.patch

    @staticmethod
    def base(**kw) -> "Dims":
        """vilt-b32 + bert-base-uncased."""
        return Dims(**kw)

    @staticmethod
    def bertweet(**kw) -> "Dims":
        """vilt-b32 + BERTweet-shaped RoBERTa."""
        d = dict(lm_kind="roberta", lm_vocab=64001, lm_max_pos=130, lm_type_vocab=1, lm_pad_id=1, lm_eps=1e-5)
        d.update(kw)
        return Dims(**d)

    @staticmethod
    def tiny(**kw) -> "Dims":
        """2+2 layers, hidden 128 (2 heads of 64): seconds on a CPU."""
        d = dict(hidden=128, layers=2, heads=2, inter=512, vilt_vocab=512, lm_layers=2, lm_vocab=512, lm_max_pos=64)
        d.update(kw)
        return Dims(**d)


def _gen(seed: int, key: str) -> torch.Generator:
    h = hashlib.sha256(f"{seed}:{key}".encode()).digest()
    g = torch.Generator(device="cpu")
    g.manual_seed(int.from_bytes(h[:8], "little") & 0x7FFFFFFFFFFFFFFF)
    return g


def _encoder_keys(prefix: str, d_h: int, d_i: int, n_layers: int, style: str) -> Dict[str, tuple]:
    """style 'vilt' (pre-LN, HF:models/vilt/modeling_vilt.py:431-465) or 'bert' (post-LN, HF:models/bert/modeling_bert.py:359-421)."""
    out: Dict[str, tuple] = {}
    for i in range(n_layers):
        p = f"{prefix}encoder.layer.{i}."
        att = "attention.attention." if style == "vilt" else "attention.self."
        for n in ("query", "key", "value"):
            out[p + att + n + ".weight"] = (d_h, d_h)
            out[p + att + n + ".bias"] = (d_h,)
        out[p + "attention.output.dense.weight"] = (d_h, d_h)
        out[p + "attention.output.dense.bias"] = (d_h,)
        out[p + "intermediate.dense.weight"] = (d_i, d_h)
        out[p + "intermediate.dense.bias"] = (d_i,)
        out[p + "output.dense.weight"] = (d_h, d_i)
        out[p + "output.dense.bias"] = (d_h,)
        if style == "vilt":
            lns = ("layernorm_before", "layernorm_after")
        else:
            lns = ("attention.output.LayerNorm", "output.LayerNorm")
        for ln in lns:
            out[p + ln + ".weight"] = (d_h,)
            out[p + ln + ".bias"] = (d_h,)
    return out


def param_shapes(d: Dims, head: bool = True) -> Dict[str, tuple]:
    """HF-compatible state_dict keys -> shapes."""
    H = d.hidden
    s: Dict[str, tuple] = {
        "embeddings.cls_token": (1, 1, H),
        "embeddings.position_embeddings": (1, d.grid * d.grid + 1, H),
        "embeddings.text_embeddings.word_embeddings.weight": (d.vilt_vocab, H),
        "embeddings.text_embeddings.position_embeddings.weight": (d.vilt_max_pos, H),
        "embeddings.text_embeddings.token_type_embeddings.weight": (d.vilt_type_vocab, H),
        "embeddings.text_embeddings.LayerNorm.weight": (H,),
        "embeddings.text_embeddings.LayerNorm.bias": (H,),
        "embeddings.patch_embeddings.projection.weight": (H, d.channels, d.patch, d.patch),
        "embeddings.patch_embeddings.projection.bias": (H,),
        "embeddings.token_type_embeddings.weight": (d.modality_vocab, H),
    }
    s.update(_encoder_keys("", H, d.inter, d.layers, "vilt"))
    s["layernorm.weight"] = (H,)
    s["layernorm.bias"] = (H,)
    s["pooler.dense.weight"] = (H, H)
    s["pooler.dense.bias"] = (H,)
    if d.lm_layers > 0:
        s["bert.embeddings.word_embeddings.weight"] = (d.lm_vocab, H)
        s["bert.embeddings.position_embeddings.weight"] = (d.lm_max_pos, H)
        s["bert.embeddings.token_type_embeddings.weight"] = (d.lm_type_vocab, H)
        s["bert.embeddings.LayerNorm.weight"] = (H,)
        s["bert.embeddings.LayerNorm.bias"] = (H,)
        s.update(_encoder_keys("bert.", H, d.inter, d.lm_layers, "bert"))
    if head:
        s["classifier.1.weight"] = (d.n_classes, H)
        s["classifier.1.bias"] = (d.n_classes,)
    return s


def make_state_dict(d: Dims, seed: int = 0, head: bool = True) -> Dict[str, torch.Tensor]:
    sd: Dict[str, torch.Tensor] = {}
    for k, shp in param_shapes(d, head).items():
        g = _gen(seed, k)
        w = torch.randn(shp, generator=g, dtype=torch.float32) * 0.02
        if ("LayerNorm.weight" in k) or ("layernorm" in k and k.endswith(".weight")):
            w = 1.0 + 5.0 * w  # gamma = 1 + N(0, 0.1): exercises the scale path
        sd[k] = w
    return sd


def make_inputs(
    d: Dims,
    batch: int,
    text_len: int,
    image_hw=(384, 384),
    seed: int = 1,
    var_text: bool = False,
    mixed_images: bool = False,
    min_text: int = 8,
) -> Dict[str, torch.Tensor]:
    """Synthetic batch of the BASELINE shapes.

    ids ~ U{lo..vocab-1} with trailing pad; ``pixel_values ~ N(0,1)`` fp32 NCHW, zero where padded;
    ``pixel_mask`` int64 per-sample top-left rectangles (what ``safe_dict_concat`` produces,
    ref:vault/vl_utils/dataset_utils.py:21-36); labels ~ U{0..n_classes-1}.
    """
    g = _gen(seed, "inputs")
    Himg, Wimg = image_hw
    lo = min(1000, d.lm_vocab // 2) if d.lm_layers > 0 else min(1000, d.vilt_vocab // 2)
    vocab = d.lm_vocab if d.lm_layers > 0 else d.vilt_vocab
    pad = d.lm_pad_id if d.lm_layers > 0 else 0
    ids = torch.randint(lo, vocab, (batch, text_len), generator=g, dtype=torch.int64)
    if var_text:
        lens = torch.randint(min(min_text, text_len), text_len + 1, (batch,), generator=g)
        lens[0] = text_len
    else:
        lens = torch.full((batch,), text_len, dtype=torch.int64)
    ar = torch.arange(text_len)[None, :]
    attn = (ar < lens[:, None]).to(torch.int64)
    ids = torch.where(attn.bool(), ids, torch.full_like(ids, pad))
    tt = torch.zeros_like(ids)
    pix = torch.randn((batch, d.channels, Himg, Wimg), generator=g, dtype=torch.float32)
    pmask = torch.ones((batch, Himg, Wimg), dtype=torch.int64)
    if mixed_images:
        gh, gw = Himg // d.patch, Wimg // d.patch
        for b in range(1, batch):
            h = int(torch.randint(max(1, gh // 2), gh + 1, (1,), generator=g))
            w = int(torch.randint(max(1, gw // 2), gw + 1, (1,), generator=g))
            pmask[b] = 0
            pmask[b, : h * d.patch, : w * d.patch] = 1
        pix = pix * pmask[:, None].to(pix.dtype)
    labels = torch.randint(0, max(d.n_classes, 2), (batch,), generator=g, dtype=torch.int64)
    if d.n_classes == 1:
        labels = labels.clamp(max=1)
    else:
        labels = labels % d.n_classes
    return dict(
        input_ids=ids,
        attention_mask=attn,
        token_type_ids=tt,
        pixel_values=pix,
        pixel_mask=pmask,
        labels=labels,
    )


def fill_parameters(named_shapes: Dict[str, tuple], d: Dims, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Deterministic values for the parameters of a head wrapper (ViltForMaskedLM & co. around the VAuLT trunk): trunk / LM keys
    take the values ``make_state_dict`` gives the same tensor (with or without the wrapper's ``vilt.`` prefix), every other
    parameter (the head) is N(0, 0.05) from a key-hashed generator (LayerNorm scales 1 + N(0, 0.1)).  Used identically by
    ``oracle/make_golden_heads.py`` (reference side) and the GPU parity test (product side)."""
    base = make_state_dict(d, seed, head=False)
    out: Dict[str, torch.Tensor] = {}
    for k, shp in named_shapes.items():
        src = k[5:] if k.startswith("vilt.") else k
        if src in base and tuple(base[src].shape) == tuple(shp):
            out[k] = base[src]
            continue
        g = _gen(seed, "head:" + k)
        w = torch.randn(tuple(shp), generator=g, dtype=torch.float32) * 0.05
        if "LayerNorm.weight" in k or k.endswith("classifier.1.weight") and len(shp) == 1:
            w = 1.0 + 2.0 * w
        out[k] = w
    return out
