"""Golden outputs of the reference's head wrappers (ref:vault/models/vault/model.py:375-509) for tests/test_heads_gpu.py.

TEST INFRASTRUCTURE: runs only where /root/reference exists.  Each case builds the REAL reference class (VaultForMaskedLM,
VaultForQuestionAnswering, VaultForImageAndTextRetrieval, VaultForImagesAndTextClassification) from tiny configs, fills its parameters
with oracle.synth.fill_parameters, runs the reference forward in eval mode on oracle.synth.make_inputs batches and stores the
logits (+ loss where the head computes one).  Parameters and inputs are regenerated from seeds at test time; only outputs are stored.

    python -m oracle.make_golden_heads        # writes tests/golden/heads/*.pt
"""
from __future__ import annotations

import dataclasses
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import synth  # noqa: E402
from oracle.ref_loader import hf_configs, load_reference_module  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "heads")

# name -> (class name, Dims overrides, constructor kwargs, batch, text_len, n_images)
CASES = {
    "mlm": ("VaultForMaskedLM", {}, {}, 3, 16, 1),
    "vqa": ("VaultForQuestionAnswering", {}, {"n_classes": 13}, 3, 16, 1),
    "retrieval": ("VaultForImageAndTextRetrieval", {}, {}, 3, 16, 1),
    "nlvr2": ("VaultForImagesAndTextClassification", {"modality_vocab": 3}, {}, 2, 12, 2),
}


GRAD_KEYS = ("vilt.pooler.dense.bias", "vilt.layernorm.weight", "vilt.encoder.layer.1.output.dense.bias", "vilt.encoder.layer.0.attention.attention.value.bias",
             "vilt.embeddings.token_type_embeddings.weight", "vilt.embeddings.cls_token", "bert.encoder.layer.0.attention.self.value.bias",
             "bert.embeddings.LayerNorm.weight", "mlm_score.transform.dense.bias", "classifier.3.bias", "rank_output.bias")


def head_inputs(d, batch, text_len, n_images):
    """Shared with the GPU test: text from one seeded batch; `n_images` image sets from consecutive seeds (stacked on dim 1)."""
    inp = synth.make_inputs(d, batch=batch, text_len=text_len, seed=11, var_text=True)
    if n_images > 1:
        more = [synth.make_inputs(d, batch=batch, text_len=text_len, seed=11 + i, var_text=True) for i in range(n_images)]
        inp["pixel_values"] = torch.stack([m["pixel_values"] for m in more], dim=1)
        inp["pixel_mask"] = torch.stack([m["pixel_mask"] for m in more], dim=1)
    return inp


def head_labels(name, d, batch, text_len, vocab):
    g = torch.Generator().manual_seed(5)
    if name == "mlm":
        lab = torch.full((batch, text_len), -100, dtype=torch.long)
        lab[:, 1] = torch.randint(0, vocab, (batch,), generator=g)
        lab[0, 3] = 7
        return lab
    if name == "vqa":
        return torch.rand(batch, 13, generator=g)
    if name == "nlvr2":
        return torch.randint(0, 2, (batch,), generator=g)
    return None


def build(mod_or_pkg, name):
    cls_name, dkw, ckw, batch, text_len, n_images = CASES[name]
    d = synth.Dims.tiny(**dkw)
    vc, lc = hf_configs(d)
    cls = getattr(mod_or_pkg, cls_name)
    m = cls(vc, bert_config=lc, **ckw)
    trunk = m.vilt
    trunk.embeddings.text_embeddings.position_embedding_type = "NOT_absolute"
    shapes = {k: tuple(p.shape) for k, p in m.named_parameters()}
    missing, unexpected = m.load_state_dict(synth.fill_parameters(shapes, d, seed=0), strict=False)
    assert not unexpected, unexpected
    return m, d, (batch, text_len, n_images)


def embeds_case(pkg):
    """VaultModel fed pre-embedded image tokens (image_embeds + flat pixel_mask): the TomViLT-style call, ref:vault/models/tomvilt/model.py:281-287."""
    d = synth.Dims.tiny()
    vc, lc = hf_configs(d)
    m = pkg.VaultModel(vc, bert_config=lc)
    m.embeddings.text_embeddings.position_embedding_type = "NOT_absolute"
    shapes = {k: tuple(p.shape) for k, p in m.named_parameters()}
    missing, unexpected = m.load_state_dict(synth.fill_parameters(shapes, d, seed=0), strict=False)
    assert not unexpected
    inp = synth.make_inputs(d, batch=3, text_len=16, seed=21, var_text=True)
    g = torch.Generator().manual_seed(9)
    P = 10
    image_embeds = torch.randn(3, P, d.hidden, generator=g) * 0.5
    image_mask = torch.ones(3, P, dtype=torch.long)
    image_mask[1, 7:] = 0
    image_mask[2, 4] = 0
    w_pool = torch.randn(3, d.hidden, generator=g)
    w_lhs = torch.randn(3, 16 + P, d.hidden, generator=g) * 0.1
    return m, d, inp, image_embeds, image_mask, w_pool, w_lhs


TEXT_EMBEDS_KINDS = {"bert": {}, "roberta": dict(lm_kind="roberta", lm_type_vocab=1, lm_pad_id=1, lm_eps=1e-5), "nolm": dict(lm_layers=0)}
TEXT_EMBEDS_GRAD_KEYS = ("pooler.dense.bias", "layernorm.weight", "embeddings.text_embeddings.token_type_embeddings.weight",
                         "embeddings.text_embeddings.position_embeddings.weight", "embeddings.text_embeddings.LayerNorm.weight",
                         "bert.embeddings.position_embeddings.weight", "bert.embeddings.token_type_embeddings.weight",
                         "bert.embeddings.LayerNorm.bias", "bert.encoder.layer.1.output.dense.bias")


def text_embeds_case(pkg, kind):
    """VaultModel fed text ``inputs_embeds`` with ``input_ids=None`` (ref:vault/models/vault/model.py:170-200: they go to the LM in place
    of its word-embedding lookup; without an LM, to ViLT's TextEmbeddings)."""
    d = synth.Dims.tiny(**TEXT_EMBEDS_KINDS[kind])
    vc, lc = hf_configs(d)
    m = pkg.VaultModel(vc, bert_config=lc)
    m.embeddings.text_embeddings.position_embedding_type = "absolute" if lc is None else "NOT_absolute"
    shapes = {k: tuple(p.shape) for k, p in m.named_parameters()}
    missing, unexpected = m.load_state_dict(synth.fill_parameters(shapes, d, seed=0), strict=False)
    assert not unexpected
    B, T = 3, 16
    inp = synth.make_inputs(d, batch=B, text_len=T, seed=31, var_text=True)
    g = torch.Generator().manual_seed(13)
    text_embeds = torch.randn(B, T, d.hidden, generator=g) * 0.5
    w_pool = torch.randn(B, d.hidden, generator=g)
    w_text = torch.randn(B, T, d.hidden, generator=g) * 0.1
    return m, d, inp, text_embeds, w_pool, w_text


EMBEDS_GRAD_KEYS = ("pooler.dense.bias", "layernorm.weight", "embeddings.token_type_embeddings.weight", "encoder.layer.0.attention.attention.value.bias",
                    "bert.encoder.layer.1.output.dense.bias")


def main():
    os.makedirs(OUT, exist_ok=True)
    mod = load_reference_module()
    m, d, inp, image_embeds, image_mask, w_pool, w_lhs = embeds_case(mod)
    m.eval()
    ie = image_embeds.clone().requires_grad_(True)
    out = m(input_ids=inp["input_ids"], attention_mask=inp["attention_mask"], token_type_ids=inp["token_type_ids"], image_embeds=ie, pixel_mask=image_mask)
    loss = (out.pooler_output * w_pool).sum() + (out.last_hidden_state * w_lhs).sum()
    loss.backward()
    named = dict(m.named_parameters())
    torch.save(dict(case="image_embeds", pooler_output=out.pooler_output.detach().clone(), last_hidden_state=out.last_hidden_state.detach().clone(),
                    d_image_embeds=ie.grad.clone(), grads={k: named[k].grad.detach().clone() for k in EMBEDS_GRAD_KEYS}),
               os.path.join(OUT, "image_embeds.pt"))
    print("image_embeds", tuple(out.last_hidden_state.shape), float(loss))
    for kind in TEXT_EMBEDS_KINDS:
        m, d, inp, text_embeds, w_pool, w_text = text_embeds_case(mod, kind)
        m.eval()
        te = text_embeds.clone().requires_grad_(True)
        torch.manual_seed(1234)
        out = m(input_ids=None, attention_mask=inp["attention_mask"], token_type_ids=inp["token_type_ids"], pixel_values=inp["pixel_values"],
                pixel_mask=inp["pixel_mask"], inputs_embeds=te)
        T = text_embeds.shape[1]
        loss = (out.pooler_output * w_pool).sum() + (out.last_hidden_state[:, :T] * w_text).sum()
        loss.backward()
        named = dict(m.named_parameters())
        torch.save(dict(case="text_embeds_" + kind, pooler_output=out.pooler_output.detach().clone(), lhs_text=out.last_hidden_state[:, :T].detach().clone(),
                        lhs_shape=tuple(out.last_hidden_state.shape), d_inputs_embeds=te.grad.clone(),
                        grads={k: named[k].grad.detach().clone() for k in TEXT_EMBEDS_GRAD_KEYS if k in named and named[k].grad is not None}),
                   os.path.join(OUT, f"text_embeds_{kind}.pt"))
        # the restatement must agree with the reference right here
        from oracle import vault_oracle as O
        sd = {k: v.detach() for k, v in m.state_dict().items()}
        with torch.no_grad():
            o = O.vault_forward(sd, d, None, inp["attention_mask"], inp["token_type_ids"], inp["pixel_values"], inp["pixel_mask"], inputs_embeds=text_embeds)
        print("text_embeds_" + kind, tuple(out.last_hidden_state.shape), float(loss), "restatement-vs-reference max-abs: pooler",
              float((o["pooler_output"] - out.pooler_output).abs().max()), "text", float((o["last_hidden_state"][:, :T] - out.last_hidden_state[:, :T]).abs().max()))
    if "--text-embeds-only" in sys.argv:
        return
    for name in CASES:
        torch.manual_seed(0)
        m, d, (batch, text_len, n_images) = build(mod, name)
        m.eval()
        inp = head_inputs(d, batch, text_len, n_images)
        labels = head_labels(name, d, batch, text_len, d.vilt_vocab)
        kw = dict(input_ids=inp["input_ids"], attention_mask=inp["attention_mask"], token_type_ids=inp["token_type_ids"],
                  pixel_values=inp["pixel_values"], pixel_mask=inp["pixel_mask"])
        if labels is not None:
            kw["labels"] = labels
        out = m(**kw)  # eval mode (no dropout), autograd on: gradients of a few small parameters pin the backward through head + trunk + LM
        loss = out.loss if getattr(out, "loss", None) is not None else out.logits.sum()
        loss.backward()
        named = dict(m.named_parameters())
        rec = dict(case=name, logits=out.logits.detach().float().clone(),
                   loss=(out.loss.detach().float().clone() if getattr(out, "loss", None) is not None else None),
                   grads={k: named[k].grad.detach().float().clone() for k in GRAD_KEYS if k in named and named[k].grad is not None})
        torch.save(rec, os.path.join(OUT, name + ".pt"))
        print(name, tuple(rec["logits"].shape), None if rec["loss"] is None else float(rec["loss"]), sorted(rec["grads"]))


if __name__ == "__main__":
    main()
