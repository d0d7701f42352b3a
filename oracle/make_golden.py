"""Generate tests/golden/*.pt by running the REAL reference (see oracle/ref_loader.py) in the build container.

    python -m oracle.make_golden            # writes tests/golden/<case>.pt and prints restatement-vs-reference errors

Each fixture holds only the case description (Dims kwargs, seeds, input kwargs) and reference OUTPUTS (pooler_output,
logits, loss, last_hidden_state text rows / un-permuted image rows for the small cases, per-parameter gradient norms
and leading slices).  Weights and inputs are regenerated from seeds by oracle/synth.py, so the files stay small.
"""
from __future__ import annotations

import dataclasses
import os
import sys

import torch

from . import ref_loader, synth, vault_oracle as O

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

CASES = {
    # name: (dims factory, dims kwargs, input kwargs, options)
    "tiny_bert_b3_t16": ("tiny", {}, dict(batch=3, text_len=16, image_hw=(384, 384), var_text=True), dict(grads=True, full_lhs=True)),
    "tiny_bert_mixed_b3_t24": ("tiny", {}, dict(batch=3, text_len=24, image_hw=(384, 640), var_text=True, mixed_images=True), dict(grads=True, full_lhs=True)),
    "tiny_roberta_b2_t20": ("tiny", dict(lm_kind="roberta", lm_type_vocab=1, lm_pad_id=1, lm_eps=1e-5), dict(batch=2, text_len=20, image_hw=(384, 384), var_text=True), dict(grads=True, full_lhs=True)),
    "tiny_frozen_b2_t16": ("tiny", {}, dict(batch=2, text_len=16, image_hw=(384, 384)), dict(grads=True, freeze_lm=True)),
    "tiny_nolm_b2_t16": ("tiny", dict(lm_layers=0), dict(batch=2, text_len=16, image_hw=(384, 384)), dict(grads=True)),
    "tiny_vpos_b2_t16": ("tiny", {}, dict(batch=2, text_len=16, image_hw=(384, 384)), dict(grads=True, use_vilt_pos=True)),
    "base_b1_t40": ("base", {}, dict(batch=1, text_len=40, image_hw=(384, 384)), dict(grads=False)),  # BASELINE config 1
    "base_b2_t40_train": ("base", {}, dict(batch=2, text_len=40, image_hw=(384, 384), var_text=True), dict(grads=True)),  # config 3 shape
    "base_b2_t64": ("base", {}, dict(batch=2, text_len=64, image_hw=(384, 384)), dict(grads=False)),  # config 2 shape
    "bertweet_b2_t128_mixed": ("bertweet", {}, dict(batch=2, text_len=128, image_hw=(384, 640), var_text=True, mixed_images=True), dict(grads=True)),  # config 4 shape
}


def _unpermute(lhs_img, patch_index, grid_w, n_valid):
    """Sort the reference's image rows of one sample into raster order using its patch_index."""
    key = patch_index[:, 0] * grid_w + patch_index[:, 1]
    order = torch.argsort(key[:n_valid])
    return lhs_img[:n_valid][order]


def run_case(name):
    fac, dkw, ikw, opt = CASES[name]
    d = getattr(synth.Dims, fac)(**dkw)
    sd = synth.make_state_dict(d, seed=0)
    inp = synth.make_inputs(d, seed=1, **ikw)
    use_vpos = opt.get("use_vilt_pos", False)
    freeze = opt.get("freeze_lm", False)
    model, mod = ref_loader.build_reference_tmsc(d, sd, use_vilt_position_embeddings=use_vpos)
    if freeze:
        model.freeze_lm = True
        if model.bert is not None:
            for p in model.bert.parameters():
                p.requires_grad_(False)
    model.eval()  # dropout off: parity is deterministic
    kw = {k: inp[k] for k in ("input_ids", "attention_mask", "token_type_ids", "pixel_values", "pixel_mask")}
    T = kw["input_ids"].shape[1]

    torch.manual_seed(1234)
    out = mod.VaultModel.forward(model, **kw)
    torch.manual_seed(1234)
    with torch.no_grad():
        _, img_mask, (patch_index, (gh, gw)) = model.embeddings.visual_embed(kw["pixel_values"], kw["pixel_mask"], -1)
    logits = model.classifier(out.pooler_output).squeeze(-1)
    loss = torch.nn.functional.cross_entropy(logits, inp["labels"])

    gold = dict(
        name=name, dims_factory=fac, dims_kwargs=dkw, input_kwargs=ikw, options=opt,
        pooler_output=out.pooler_output.detach().clone(), logits=logits.detach().clone(), loss=loss.detach().clone(),
        lhs_shape=tuple(out.last_hidden_state.shape), lhs_cls=out.last_hidden_state[:, 0].detach().clone(),
        lhs_image_cls=out.last_hidden_state[:, T].detach().clone(),
    )
    n_valid = (img_mask[:, 1:] > 0).sum(dim=1)
    lhs = out.last_hidden_state.detach()
    if opt.get("full_lhs"):
        gold["lhs_text"] = lhs[:, :T].clone()
        gold["lhs_image_raster"] = [
            _unpermute(lhs[b, T + 1:], patch_index[b], gw, int(n_valid[b])).clone() for b in range(lhs.shape[0])
        ]
    else:
        # first valid raster patch row and a checksum per sample
        gold["lhs_image_raster_first"] = torch.stack(
            [_unpermute(lhs[b, T + 1:], patch_index[b], gw, int(n_valid[b]))[0] for b in range(lhs.shape[0])]
        )
    gold["n_valid_patches"] = n_valid.clone()

    if opt.get("grads"):
        model.zero_grad()
        loss.backward()
        gn, gs = {}, {}
        for k, p in model.named_parameters():
            if p.grad is None:
                continue
            gn[k] = p.grad.norm().item()
            gs[k] = p.grad.flatten()[:32].clone()
        gold["grad_norm"] = gn
        gold["grad_head"] = gs
        gold["grad_none"] = sorted(k for k, p in model.named_parameters() if p.grad is None)
        # full gradients for a few structurally distinct parameters (small or tiny-model only)
        full = {}
        for k, p in model.named_parameters():
            if p.grad is None:
                continue
            pick = fac == "tiny" and (k.endswith("layer.0.attention.attention.query.weight")
                                      or k.endswith("layer.0.attention.self.query.weight")
                                      or k.endswith("layer.1.output.dense.weight"))
            if p.numel() <= 1024 or pick:
                full[k] = p.grad.detach().clone()
        gold["grad_full"] = full

    # --- check the restatement against the reference right here -----------------------------------------
    o = O.vault_forward(sd, d, use_vilt_position_embeddings=use_vpos, **kw)
    e_pool = (o["pooler_output"] - gold["pooler_output"]).abs().max().item()
    e_txt = (o["last_hidden_state"][:, :T] - lhs[:, :T]).abs().max().item()
    e_img = 0.0
    for b in range(lhs.shape[0]):
        nv = int(n_valid[b])
        ref_rows = _unpermute(lhs[b, T + 1:], patch_index[b], gw, nv)
        e_img = max(e_img, (o["last_hidden_state"][b, T + 1: T + 1 + nv] - ref_rows).abs().max().item())
    assert tuple(o["last_hidden_state"].shape) == gold["lhs_shape"], (o["last_hidden_state"].shape, gold["lhs_shape"])
    print(f"{name:28s} lhs{gold['lhs_shape']} restatement-vs-reference max-abs: pooler {e_pool:.2e} text {e_txt:.2e} image {e_img:.2e}")
    return gold


def main(argv):
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    names = argv or list(CASES)
    for n in names:
        g = run_case(n)
        torch.save(g, os.path.join(GOLDEN_DIR, n + ".pt"))
    tot = sum(os.path.getsize(os.path.join(GOLDEN_DIR, f)) for f in os.listdir(GOLDEN_DIR))
    print(f"golden dir: {tot/1e6:.2f} MB")


if __name__ == "__main__":
    torch.set_grad_enabled(True)
    main(sys.argv[1:])
