"""GPU parity of the VaultFor* head wrappers against fixtures produced by the REAL reference classes
(oracle/make_golden_heads.py, ref:vault/models/vault/model.py:375-509): logits, loss and gradients through head + ViLT trunk + LM.

Tolerance: bf16 tensor-core trunk under fp32 heads on a 2+2-layer model -- logits max-abs / max|ref| <= 2e-2, loss |d| <= 2e-2,
per-parameter gradient cosine >= 0.99."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import make_golden_heads as G  # noqa: E402
from tests.golden_utils import cosine, rel_err  # noqa: E402

DEV = "cuda:0"
HEADS_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "heads")


@pytest.mark.parametrize("name", sorted(G.CASES))
def test_head_wrapper_matches_reference_fixture(name):
    import vault_b200.models.vault as pkg

    ref = torch.load(os.path.join(HEADS_DIR, name + ".pt"), weights_only=False)
    m, d, (batch, text_len, n_images) = G.build(pkg, name)
    m = m.to(DEV).eval()
    inp = G.head_inputs(d, batch, text_len, n_images)
    labels = G.head_labels(name, d, batch, text_len, d.vilt_vocab)
    kw = {k: inp[k].to(DEV) for k in ("input_ids", "attention_mask", "token_type_ids", "pixel_values", "pixel_mask")}
    if labels is not None:
        kw["labels"] = labels.to(DEV)
    out = m(**kw)
    assert rel_err(out.logits.detach().float().cpu(), ref["logits"]) <= 2e-2
    if ref["loss"] is not None:
        assert abs(out.loss.item() - ref["loss"].item()) <= 2e-2
    loss = out.loss if getattr(out, "loss", None) is not None else out.logits.sum()
    loss.backward()
    named = dict(m.named_parameters())
    assert ref["grads"], "fixture without gradients"
    for k, g_ref in ref["grads"].items():
        g = named[k].grad
        assert g is not None, k
        assert cosine(g.detach().float().cpu(), g_ref) >= 0.99, (k, cosine(g.detach().float().cpu(), g_ref))
        assert abs(g.float().norm().item() / max(g_ref.norm().item(), 1e-12) - 1.0) <= 5e-2, k
    # a second step must not accumulate into the first one's kernel-written gradients
    for p in m.parameters():
        p.grad = None
    out = m(**kw)
    (out.loss if getattr(out, "loss", None) is not None else out.logits.sum()).backward()
    k = "vilt.layernorm.weight"
    assert cosine(named[k].grad.float().cpu(), ref["grads"][k]) >= 0.99
    assert abs(named[k].grad.float().norm().item() / ref["grads"][k].norm().item() - 1.0) <= 5e-2


def test_image_embeds_path_matches_reference_fixture():
    """VaultModel(image_embeds=..., pixel_mask=flat mask) -- SURVEY.md section 8f rank 4 -- against the real reference: outputs on valid rows,
    gradient w.r.t. the caller's image_embeds and a few parameters."""
    import vault_b200.models.vault as pkg

    ref = torch.load(os.path.join(HEADS_DIR, "image_embeds.pt"), weights_only=False)
    m, d, inp, image_embeds, image_mask, w_pool, w_lhs = G.embeds_case(pkg)
    m = m.to(DEV).eval()
    ie = image_embeds.to(DEV).requires_grad_(True)
    out = m(input_ids=inp["input_ids"].to(DEV), attention_mask=inp["attention_mask"].to(DEV), token_type_ids=inp["token_type_ids"].to(DEV),
            image_embeds=ie, pixel_mask=image_mask.to(DEV))
    assert rel_err(out.pooler_output.detach().cpu(), ref["pooler_output"]) <= 2e-2
    valid = torch.cat([inp["attention_mask"], image_mask], dim=1).bool()
    assert rel_err(out.last_hidden_state.detach().cpu()[valid], ref["last_hidden_state"][valid]) <= 2e-2
    loss = (out.pooler_output * w_pool.to(DEV)).sum() + (out.last_hidden_state * w_lhs.to(DEV)).sum()
    loss.backward()
    assert cosine(ie.grad.cpu(), ref["d_image_embeds"]) >= 0.99
    named = dict(m.named_parameters())
    for k, g_ref in ref["grads"].items():
        assert cosine(named[k].grad.float().cpu(), g_ref) >= 0.99, k
    with torch.no_grad():
        again = m(input_ids=inp["input_ids"].to(DEV), attention_mask=inp["attention_mask"].to(DEV), token_type_ids=inp["token_type_ids"].to(DEV),
                  image_embeds=image_embeds.to(DEV), pixel_mask=image_mask.to(DEV))
    assert torch.equal(again.pooler_output, out.pooler_output.detach())


@pytest.mark.parametrize("vocab,tokens", [(510, 48), (30522, 80)])
def test_mlm_decoder_on_the_gemm_kernel(vocab, tokens):
    """``mlm_score.decoder`` of VaultForMaskedLM runs on the tcgen05 GEMM (forward, dgrad, wgrad, bias column sums), with the vocabulary
    padded to a multiple of 8 columns: against torch fp32 on the same bf16-rounded operands.  Tolerance: bf16 rounding of dlogits."""
    from vault_b200.model import _KernelDecoder

    torch.manual_seed(3)
    H = 128
    lin = torch.nn.Linear(H, vocab).to(DEV)
    with torch.no_grad():
        lin.weight.copy_(lin.weight.to(torch.bfloat16).float())
    ker = torch.nn.Linear(H, vocab).to(DEV)
    ker.load_state_dict(lin.state_dict())
    ker.__class__ = _KernelDecoder
    x = torch.randn(2, tokens // 2, H, device=DEV).to(torch.bfloat16).float()
    xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    labels = torch.randint(0, vocab, (tokens,), device=DEV)
    labels[::3] = -100
    ya, yb = lin(xa), ker(xb)
    assert yb.shape == ya.shape == (2, tokens // 2, vocab)
    assert rel_err(yb.detach().cpu(), ya.detach().cpu()) <= 1e-3
    la = torch.nn.functional.cross_entropy(ya.view(-1, vocab), labels)
    lb = torch.nn.functional.cross_entropy(yb.view(-1, vocab), labels)  # the HF head's own .view on the sliced logits
    assert abs(la.item() - lb.item()) <= 1e-3
    la.backward()
    lb.backward()
    for got, want in ((xb.grad, xa.grad), (ker.weight.grad, lin.weight.grad), (ker.bias.grad, lin.bias.grad)):
        assert got is not None and got.shape == want.shape
        assert cosine(got.float().cpu(), want.float().cpu()) >= 0.999
        assert rel_err(got.float().cpu(), want.float().cpu()) <= 2e-2


@pytest.mark.parametrize("kind", sorted(G.TEXT_EMBEDS_KINDS))
def test_text_inputs_embeds_path_matches_reference_fixture(kind):
    """VaultModel(input_ids=None, inputs_embeds=...) -- ref:vault/models/vault/model.py:170-200 -- against the REAL reference: BERT, RoBERTa
    (sequential position ids from pad+1) and no-LM variants; outputs, the gradient returned for the caller's embeddings, embedding-table
    gradients.  Tolerance as for the other fixtures of this file (bf16 trunk): 2e-2 relative, cosine >= 0.99."""
    import vault_b200.models.vault as pkg

    ref = torch.load(os.path.join(HEADS_DIR, f"text_embeds_{kind}.pt"), weights_only=False)
    m, d, inp, text_embeds, w_pool, w_text = G.text_embeds_case(pkg, kind)
    m = m.to(DEV).eval()
    T = text_embeds.shape[1]
    te = text_embeds.to(DEV).requires_grad_(True)
    kw = dict(attention_mask=inp["attention_mask"].to(DEV), token_type_ids=inp["token_type_ids"].to(DEV), pixel_values=inp["pixel_values"].to(DEV),
              pixel_mask=inp["pixel_mask"].to(DEV))
    out = m(input_ids=None, inputs_embeds=te, **kw)
    assert tuple(out.last_hidden_state.shape) == ref["lhs_shape"]
    assert rel_err(out.pooler_output.detach().cpu(), ref["pooler_output"]) <= 2e-2
    valid = inp["attention_mask"].bool()
    assert rel_err(out.last_hidden_state.detach().cpu()[:, :T][valid], ref["lhs_text"][valid]) <= 2e-2
    loss = (out.pooler_output * w_pool.to(DEV)).sum() + (out.last_hidden_state[:, :T] * w_text.to(DEV)).sum()
    loss.backward()
    assert te.grad is not None and cosine(te.grad.cpu(), ref["d_inputs_embeds"]) >= 0.99
    named = dict(m.named_parameters())
    for k, g_ref in ref["grads"].items():
        assert named[k].grad is not None, k
        assert cosine(named[k].grad.float().cpu(), g_ref) >= 0.99, k
    word = "bert.embeddings.word_embeddings.weight" if kind != "nolm" else "embeddings.text_embeddings.word_embeddings.weight"
    assert named[word].grad is None or float(named[word].grad.abs().max()) == 0.0  # the lookup never ran
    with torch.no_grad():
        again = m(input_ids=None, inputs_embeds=text_embeds.to(DEV), **kw)
    assert torch.equal(again.pooler_output, out.pooler_output.detach())
    with pytest.raises(ValueError):
        m(input_ids=inp["input_ids"].to(DEV), inputs_embeds=te, **kw)
