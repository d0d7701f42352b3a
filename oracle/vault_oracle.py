"""CPU fp32 restatement of the VAuLT hot path (oracle).

TEST INFRASTRUCTURE -- NOT PRODUCT CODE.  Only ``tests/``, ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs and ``__graft_entry__.smoke()``
may import this file; ``vault_b200/`` never does and has no CPU fallback.

Where the algorithm lives.  The reference's own code for this path is ~200 lines
of glue (ref:vault/models/vault/model.py:53-218, 512-570); the arithmetic is in
the un-vendored third-party dependency ``transformers`` **pinned ==4.48.0**
(ref:setup.py:11).  This file restates that published algorithm with plain
``torch`` fp32 ops on an HF-keyed state_dict, each function citing the lines it
follows (``HF:`` = transformers/...; line numbers of the installed 5.5.0 copy,
whose ViLT/BERT arithmetic is unchanged apart from the ``position_embedding_type``
gate that 5.x deleted and that is restored here -- SURVEY.md section 8c).

Pinning status.  The reference ships no tests, KATs or fixtures, so nothing of
its own pins this path.  The restatement is instead pinned against OUTPUTS OF
THE REFERENCE ITSELF, run in the build container by ``oracle/make_golden.py``
(reference ``model.py`` imported by path over the installed HF modules, 4.48.0
gate restored) and committed under ``tests/golden/``; ``tests/test_oracle.py``
checks the restatement against those fixtures on every run.

Deliberate difference from the reference: image tokens come out in raster order
(valid patches first, pad rows after) instead of the reference's random
permutation (HF:models/vilt/modeling_vilt.py:141-160).  Attention is
permutation-equivariant, so ``pooler_output``, the text rows and (after
un-permuting with the reference's ``patch_index``) the image rows are identical.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

from .synth import Dims

Tensor = torch.Tensor


def _ln(x: Tensor, sd: Dict[str, Tensor], name: str, eps: float) -> Tensor:
    return F.layer_norm(x, (x.shape[-1],), sd[name + ".weight"], sd[name + ".bias"], eps)


def _lin(x: Tensor, sd: Dict[str, Tensor], name: str) -> Tensor:
    return F.linear(x, sd[name + ".weight"], sd[name + ".bias"])


def _dropout(x: Tensor, p: float, train: bool) -> Tensor:
    return F.dropout(x, p, training=train) if (train and p > 0) else x


def _self_attention(x: Tensor, key_mask: Tensor, sd, pre: str, heads: int, p_drop: float, train: bool) -> Tensor:
    """softmax(Q K^T / sqrt(dh) + mask) V.

    HF:models/vilt/modeling_vilt.py:325-365 (scores / sqrt(dh), + additive mask, softmax, dropout, @V) and
    HF:models/bert/modeling_bert.py:115-140 (same with * scaling).  The additive mask is
    (1 - mask) * finfo.min of shape [B,1,1,S] (HF:modeling_utils.py:902-949).
    """
    B, S, H = x.shape
    dh = H // heads
    q = _lin(x, sd, pre + "query").view(B, S, heads, dh).transpose(1, 2)
    k = _lin(x, sd, pre + "key").view(B, S, heads, dh).transpose(1, 2)
    v = _lin(x, sd, pre + "value").view(B, S, heads, dh).transpose(1, 2)
    scores = torch.matmul(q, k.transpose(-1, -2)) / math.sqrt(dh)
    add = (1.0 - key_mask.to(scores.dtype))[:, None, None, :] * torch.finfo(scores.dtype).min
    probs = torch.softmax(scores + add, dim=-1)
    probs = _dropout(probs, p_drop, train)
    ctx = torch.matmul(probs, v).transpose(1, 2).reshape(B, S, H)
    return ctx


def roberta_position_ids(input_ids: Tensor, pad_id: int) -> Tensor:
    """HF:models/roberta/modeling_roberta.py:152-170: cumsum(ids != pad) * (ids != pad) + pad."""
    m = input_ids.ne(pad_id).to(torch.int64)
    return torch.cumsum(m, dim=1) * m + pad_id


def lm_forward(sd, d: Dims, input_ids: Optional[Tensor], attention_mask: Tensor, token_type_ids: Optional[Tensor], train: bool = False,
               inputs_embeds: Optional[Tensor] = None) -> Tensor:
    """BertModel / RobertaModel without pooler -> last_hidden_state.

    Embeddings HF:models/bert/modeling_bert.py:72-112 (RoBERTa: HF:models/roberta/modeling_roberta.py:79-126),
    post-LN layer HF:models/bert/modeling_bert.py:359-421, GELU = exact erf (BertIntermediate :330-342).
    ``inputs_embeds`` (instead of ``input_ids``) replaces the word-embedding lookup; RoBERTa then numbers the positions
    sequentially from pad+1 (create_position_ids_from_inputs_embeds: padding cannot be inferred from embeddings).
    """
    B, T = input_ids.shape if input_ids is not None else inputs_embeds.shape[:2]
    if d.lm_kind == "roberta":
        pos = roberta_position_ids(input_ids, d.lm_pad_id) if input_ids is not None else torch.arange(d.lm_pad_id + 1, T + d.lm_pad_id + 1)[None, :].expand(B, T)
    else:
        pos = torch.arange(T)[None, :].expand(B, T)
    if token_type_ids is None:
        token_type_ids = torch.zeros((B, T), dtype=torch.int64)
    p = d.lm_dropout
    x = (
        (sd["bert.embeddings.word_embeddings.weight"][input_ids] if inputs_embeds is None else inputs_embeds)
        + sd["bert.embeddings.token_type_embeddings.weight"][token_type_ids]
        + sd["bert.embeddings.position_embeddings.weight"][pos]
    )
    x = _dropout(_ln(x, sd, "bert.embeddings.LayerNorm", d.lm_eps), p, train)
    for i in range(d.lm_layers):
        pre = f"bert.encoder.layer.{i}."
        ctx = _self_attention(x, attention_mask, sd, pre + "attention.self.", d.heads, p, train)
        a = _ln(x + _dropout(_lin(ctx, sd, pre + "attention.output.dense"), p, train), sd, pre + "attention.output.LayerNorm", d.lm_eps)
        h = F.gelu(_lin(a, sd, pre + "intermediate.dense"))
        x = _ln(a + _dropout(_lin(h, sd, pre + "output.dense"), p, train), sd, pre + "output.LayerNorm", d.lm_eps)
    return x


def patch_grid_hw(pixel_mask: Tensor, patch: int) -> Tuple[Tensor, Tensor]:
    """Per-sample valid patch rows/cols.  HF:models/vilt/modeling_vilt.py:95-98: nearest down-sampling of the
    pixel mask (patch (i,j) valid iff pixel (patch*i, patch*j) is), x_h = valid rows in column 0, x_w = valid cols in row 0."""
    pm = pixel_mask[:, ::patch, ::patch].to(torch.int64)
    return pm[:, :, 0].sum(dim=1), pm[:, 0, :].sum(dim=1)


def resized_pos_embed(pos_table: Tensor, grid: int, h: int, w: int) -> Tensor:
    """Bilinear(align_corners=True) resize of the grid x grid spatial position table to (h, w) -> [h*w, H].
    HF:models/vilt/modeling_vilt.py:101-117."""
    H = pos_table.shape[-1]
    spatial = pos_table[:, 1:, :].transpose(1, 2).reshape(1, H, grid, grid)
    r = F.interpolate(spatial, size=(h, w), mode="bilinear", align_corners=True)
    return r.flatten(2).transpose(1, 2)[0]


def visual_embed_raster(sd, d: Dims, pixel_values: Tensor, pixel_mask: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
    """ViltEmbeddings.visual_embed (HF:models/vilt/modeling_vilt.py:91-177) with deterministic raster order.

    Returns (x [B,1+P,H], mask [B,1+P] int64, patch_index [B,P,2]) with P = max_b h_b*w_b (max_image_length=-1).
    Pad rows (mask 0) are zeros here; the reference fills them with randomly chosen padded patches, they are masked
    out of attention either way.
    """
    w_proj = sd["embeddings.patch_embeddings.projection.weight"]
    b_proj = sd["embeddings.patch_embeddings.projection.bias"]
    x = F.conv2d(pixel_values.to(w_proj.dtype), w_proj, b_proj, stride=d.patch)  # [B,H,gh,gw]
    B, Hd, gh, gw = x.shape
    xh, xw = patch_grid_hw(pixel_mask, d.patch)
    P = int((xh * xw).max())
    pos_table = sd["embeddings.position_embeddings"]
    rows, masks, pidx = [], [], []
    for b in range(B):
        h, w = int(xh[b]), int(xw[b])
        pe = resized_pos_embed(pos_table, d.grid, h, w)  # [h*w,H]
        xb = x[b, :, :h, :w].reshape(Hd, h * w).transpose(0, 1) + pe
        ii, jj = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
        idx = torch.stack([ii.reshape(-1), jj.reshape(-1)], dim=-1)
        n = h * w
        if n < P:
            xb = torch.cat([xb, xb.new_zeros(P - n, Hd)], dim=0)
            idx = torch.cat([idx, idx.new_full((P - n, 2), -1)], dim=0)
        rows.append(xb)
        pidx.append(idx)
        masks.append(torch.cat([torch.ones(1 + n, dtype=torch.int64), torch.zeros(P - n, dtype=torch.int64)]))
    xp = torch.stack(rows, dim=0)
    cls = (sd["embeddings.cls_token"] + pos_table[:, 0:1, :]).expand(B, -1, -1)
    return torch.cat([cls, xp], dim=1), torch.stack(masks, 0), torch.stack(pidx, 0)


def vilt_text_embed(sd, d: Dims, inputs_embeds: Tensor, token_type_ids: Tensor, use_pos: bool) -> Tensor:
    """TextEmbeddings.forward with inputs_embeds (HF:models/vilt/modeling_vilt.py:240-272) under transformers==4.48.0
    semantics: position embeddings are added only if position_embedding_type == "absolute"; VAuLT sets
    "NOT_absolute" whenever an LM is attached (ref:vault/models/vault/model.py:77-79, 112-116)."""
    T = inputs_embeds.shape[1]
    e = inputs_embeds + sd["embeddings.text_embeddings.token_type_embeddings.weight"][token_type_ids]
    if use_pos:
        e = e + sd["embeddings.text_embeddings.position_embeddings.weight"][:T][None]
    return _ln(e, sd, "embeddings.text_embeddings.LayerNorm", d.vilt_eps)


def vault_forward(
    sd: Dict[str, Tensor],
    d: Dims,
    input_ids: Tensor,
    attention_mask: Tensor,
    token_type_ids: Tensor,
    pixel_values: Tensor,
    pixel_mask: Tensor,
    train: bool = False,
    use_vilt_position_embeddings: bool = False,
    image_token_type_idx: int = 1,
    inputs_embeds: Optional[Tensor] = None,
    output_hidden_states: bool = False,
) -> Dict[str, Tensor]:
    """VaultMixin.forward (ref:vault/models/vault/model.py:151-218) -> ViltModel.forward
    (HF:models/vilt/modeling_vilt.py:550-660) -> ViltPooler (:663-675).  ``inputs_embeds`` with ``input_ids=None``: the text
    embeddings handed to the LM (ref :170-190) -- or, without an LM, to ViLT's TextEmbeddings -- in place of the word lookup."""
    if token_type_ids is None:  # HF: all-zero type ids
        token_type_ids = torch.zeros((input_ids if input_ids is not None else inputs_embeds).shape[:2], dtype=torch.int64)
    if d.lm_layers > 0:
        # ref:vault/models/vault/model.py:174-180: LM copy of the type ids is zeroed iff the LM type vocab < 2
        lm_tt = torch.zeros_like(token_type_ids) if d.lm_type_vocab < 2 else token_type_ids
        text_in = lm_forward(sd, d, input_ids, attention_mask, lm_tt, train, inputs_embeds=inputs_embeds)
        use_pos = use_vilt_position_embeddings
    else:
        text_in = sd["embeddings.text_embeddings.word_embeddings.weight"][input_ids] if inputs_embeds is None else inputs_embeds
        use_pos = True
    text = vilt_text_embed(sd, d, text_in, token_type_ids, use_pos)
    img, img_mask, patch_index = visual_embed_raster(sd, d, pixel_values, pixel_mask)
    mod = sd["embeddings.token_type_embeddings.weight"]
    text = text + mod[0]
    img = img + mod[image_token_type_idx]
    x = torch.cat([text, img], dim=1)
    mask = torch.cat([attention_mask.to(torch.int64), img_mask], dim=1)
    hidden = []  # ViltEncoder with output_hidden_states (HF:models/vilt/modeling_vilt.py:505-540): input of every layer, then the last output
    for i in range(d.layers):
        hidden.append(x)
        pre = f"encoder.layer.{i}."
        ctx = _self_attention(_ln(x, sd, pre + "layernorm_before", d.vilt_eps), mask, sd, pre + "attention.attention.", d.heads, 0.0, False)
        h = x + _lin(ctx, sd, pre + "attention.output.dense")
        m = F.gelu(_lin(_ln(h, sd, pre + "layernorm_after", d.vilt_eps), sd, pre + "intermediate.dense"))
        x = h + _lin(m, sd, pre + "output.dense")
    hidden.append(x)
    x = _ln(x, sd, "layernorm", d.vilt_eps)
    pooled = torch.tanh(_lin(x[:, 0], sd, "pooler.dense"))
    out = dict(last_hidden_state=x, pooler_output=pooled, mask=mask, patch_index=patch_index)
    if output_hidden_states:
        out["hidden_states"] = tuple(hidden)
    return out


def tmsc_logits(sd, d: Dims, pooled: Tensor, train: bool = False) -> Tensor:
    """VaultForTMSC head: Linear(Dropout(pooler_output)).squeeze(-1) (ref:vault/models/vault/model.py:547-550, 569)."""
    return _lin(_dropout(pooled, d.head_dropout, train), sd, "classifier.1").squeeze(-1)


def ce_loss(logits: Tensor, labels: Tensor) -> Tensor:
    """Twitter201XTrainer.calculate_loss: nn.CrossEntropyLoss(), mean over the batch
    (ref:vault/tmsc_utils/trainer.py:228-242)."""
    return F.cross_entropy(logits, labels)


def bce_loss(logits: Tensor, labels: Tensor) -> Tensor:
    """VaultTrainerForBloombergTwitterCorpus.calculate_loss: nn.BCEWithLogitsLoss(), mean (ref:vault/models/vault/trainer.py:42-56)."""
    return F.binary_cross_entropy_with_logits(logits, labels)


def mvsa_raw_loss(logits: Tensor, labels: Tensor) -> Tensor:
    """VaultTrainerForMVSA.calculate_loss on raw annotations: the logits' two halves against the text / image label columns,
    0.5 * (CE + CE) (ref:vault/models/vault/trainer.py:114-137)."""
    n = logits.shape[-1]
    return 0.5 * (F.cross_entropy(logits[..., : n // 2], labels[..., 0]) + F.cross_entropy(logits[..., n // 2:], labels[..., 1]))


def hf_adamw_step(p: Tensor, g: Tensor, m: Tensor, v: Tensor, step: int, lr: float, beta1=0.9, beta2=0.999, eps=1e-8,
                  weight_decay=0.0, correct_bias=False) -> None:
    """transformers.optimization.AdamW.step of transformers==4.48.0 (class removed in 5.x), restated:
    m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g^2 ; denom = sqrt(v) + eps ;
    step_size = lr * sqrt(1-b2^t)/(1-b1^t) if correct_bias else lr ; p -= step_size * m / denom ;
    then, if wd > 0, p -= lr * wd * p.  (trainer default correct_bias=False,
    ref:vault/tmsc_utils/trainer.py:69, 244-254.)"""
    m.mul_(beta1).add_(g, alpha=1.0 - beta1)
    v.mul_(beta2).addcmul_(g, g, value=1.0 - beta2)
    denom = v.sqrt().add_(eps)
    step_size = lr
    if correct_bias:
        step_size = lr * math.sqrt(1.0 - beta2 ** step) / (1.0 - beta1 ** step)
    p.addcdiv_(m, denom, value=-step_size)
    if weight_decay > 0.0:
        p.add_(p, alpha=-lr * weight_decay)


def linear_warmup_lr(step: int, total: int, base_lr: float, warmup_ratio: float = 0.1) -> float:
    """get_linear_schedule_with_warmup (HF:optimization.py:101-131) with warm = int(ratio * total)
    (ref:vault/tmsc_utils/trainer.py:256-280).  lr(0) = 0."""
    warm = int(warmup_ratio * total)
    if step < warm:
        return base_lr * float(step) / float(max(1, warm))
    return base_lr * max(0.0, float(total - step) / float(max(1, total - warm)))


def grads_never_set(d: Dims, use_vilt_position_embeddings: bool = False):
    """Parameters that get grad=None whenever an LM is attached (SURVEY.md section 8e traps)."""
    if d.lm_layers == 0:
        return set()
    s = {"embeddings.text_embeddings.word_embeddings.weight"}
    if not use_vilt_position_embeddings:
        s.add("embeddings.text_embeddings.position_embeddings.weight")
    return s


def train_step(sd: Dict[str, Tensor], d: Dims, batch: Dict[str, Tensor], lr: float, freeze_lm: bool = False,
               state: Optional[dict] = None, train_mode: bool = False, loss_fn=None) -> Dict[str, object]:
    """One fine-tuning step of Twitter201XTrainer.train (ref:vault/tmsc_utils/trainer.py:353-367): forward, CE loss,
    backward, HF-AdamW.  ``sd`` is updated in place.  Dropout only if ``train_mode`` (parity runs keep it off)."""
    params = {k: v.detach().clone().requires_grad_(not (freeze_lm and k.startswith("bert."))) for k, v in sd.items()}
    ctxmgr = torch.enable_grad()
    with ctxmgr:
        if freeze_lm:
            # ref:vault/models/vault/model.py:189: LM runs under set_grad_enabled(False)
            pass
        out = vault_forward(params, d, batch["input_ids"], batch["attention_mask"], batch["token_type_ids"],
                            batch["pixel_values"], batch["pixel_mask"], train=train_mode)
        logits = tmsc_logits(params, d, out["pooler_output"], train=train_mode)
        loss = (loss_fn or ce_loss)(logits, batch["labels"])  # bce_loss / mvsa_raw_loss for the other two trainers
    loss.backward()
    grads = {k: p.grad for k, p in params.items() if p.grad is not None}
    if state is None:
        state = {"step": 0, "m": {}, "v": {}}
    state["step"] += 1
    for k, g in grads.items():
        m = state["m"].setdefault(k, torch.zeros_like(g))
        v = state["v"].setdefault(k, torch.zeros_like(g))
        with torch.no_grad():
            hf_adamw_step(sd[k], g, m, v, state["step"], lr)
    return dict(loss=loss.detach(), logits=logits.detach(), pooler_output=out["pooler_output"].detach(), grads=grads, state=state)
