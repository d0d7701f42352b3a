"""Golden ViLT hidden states of the REAL reference (output_hidden_states=True through ref:vault/models/vault/model.py:207-218 ->
HF ViltModel.forward) for one tiny mixed-size case -> tests/golden/heads/hidden_states_tiny.pt.  TEST INFRASTRUCTURE; runs only where
/root/reference exists.  Stored per hidden state: the text rows, the image CLS row and the valid image rows sorted into raster order with the
reference's own patch_index (the reference permutes image tokens at random; attention is permutation-equivariant).

    python -m oracle.make_golden_hidden
"""
from __future__ import annotations

import os
import sys

import torch

from . import ref_loader, synth, vault_oracle as O
from .make_golden import GOLDEN_DIR, _unpermute

CASE = dict(dims_factory="tiny", input_kwargs=dict(batch=3, text_len=12, image_hw=(384, 640), var_text=True, mixed_images=True))


def main():
    d = getattr(synth.Dims, CASE["dims_factory"])()
    sd = synth.make_state_dict(d, seed=0)
    inp = synth.make_inputs(d, seed=21, **CASE["input_kwargs"])
    model, mod = ref_loader.build_reference_tmsc(d, sd)
    model.eval()
    kw = {k: inp[k] for k in ("input_ids", "attention_mask", "token_type_ids", "pixel_values", "pixel_mask")}
    T = kw["input_ids"].shape[1]
    torch.manual_seed(1234)
    with torch.no_grad():
        out = mod.VaultModel.forward(model, output_hidden_states=True, **kw)
    torch.manual_seed(1234)
    with torch.no_grad():
        _, img_mask, (patch_index, (gh, gw)) = model.embeddings.visual_embed(kw["pixel_values"], kw["pixel_mask"], -1)
    n_valid = (img_mask[:, 1:] > 0).sum(dim=1)
    hs = out.hidden_states
    gold = dict(case=CASE, seed=21, n_hidden=len(hs), n_valid_patches=n_valid.clone(), text=[h[:, :T].clone() for h in hs],
                image_cls=[h[:, T].clone() for h in hs],
                image_raster=[[_unpermute(h[b, T + 1:], patch_index[b], gw, int(n_valid[b])).clone() for b in range(h.shape[0])] for h in hs])
    o = O.vault_forward(sd, d, output_hidden_states=True, **kw)
    assert len(o["hidden_states"]) == len(hs)
    err = 0.0
    for k, h in enumerate(o["hidden_states"]):
        err = max(err, (h[:, :T] - gold["text"][k]).abs().max().item(), (h[:, T] - gold["image_cls"][k]).abs().max().item())
        for b in range(h.shape[0]):
            nv = int(n_valid[b])
            err = max(err, (h[b, T + 1:T + 1 + nv] - gold["image_raster"][k][b]).abs().max().item())
    print(f"hidden states: {len(hs)} tensors, restatement-vs-reference max-abs {err:.2e}")
    torch.save(gold, os.path.join(GOLDEN_DIR, "heads", "hidden_states_tiny.pt"))


if __name__ == "__main__":
    main()
