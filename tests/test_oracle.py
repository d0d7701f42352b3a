"""Pins the CPU restatement (oracle/vault_oracle.py) against fixtures produced by the REAL reference
(oracle/make_golden.py, run in the build container where /root/reference exists)."""
import pytest
import torch

from oracle import synth, vault_oracle as O
from tests.golden_utils import cosine, golden_names, load_case, rel_err

FWD_KEYS = ("input_ids", "attention_mask", "token_type_ids", "pixel_values", "pixel_mask")
NAMES = golden_names()


def test_golden_present():
    assert len(NAMES) >= 8


@pytest.mark.parametrize("name", NAMES)
def test_forward_matches_reference(name):
    g, d, sd, inp = load_case(name)
    opt = g["options"]
    with torch.no_grad():
        o = O.vault_forward(sd, d, use_vilt_position_embeddings=opt.get("use_vilt_pos", False), **{k: inp[k] for k in FWD_KEYS})
        logits = O.tmsc_logits(sd, d, o["pooler_output"])
        loss = O.ce_loss(logits, inp["labels"])
    T = inp["input_ids"].shape[1]
    assert tuple(o["last_hidden_state"].shape) == g["lhs_shape"]
    tol = 2e-5  # fp32 CPU vs fp32 CPU, different op order
    assert (o["pooler_output"] - g["pooler_output"]).abs().max() < tol
    assert (logits - g["logits"]).abs().max() < tol
    assert abs(loss.item() - g["loss"].item()) < tol
    assert (o["last_hidden_state"][:, 0] - g["lhs_cls"]).abs().max() < tol
    assert (o["last_hidden_state"][:, T] - g["lhs_image_cls"]).abs().max() < tol
    nv = g["n_valid_patches"]
    assert torch.equal(o["mask"][:, T + 1:].sum(1), nv)
    if "lhs_text" in g:
        assert (o["last_hidden_state"][:, :T] - g["lhs_text"]).abs().max() < tol
        for b, rows in enumerate(g["lhs_image_raster"]):
            assert (o["last_hidden_state"][b, T + 1: T + 1 + int(nv[b])] - rows).abs().max() < tol
    else:
        assert (o["last_hidden_state"][:, T + 1] - g["lhs_image_raster_first"]).abs().max() < tol


@pytest.mark.parametrize("name", [n for n in NAMES if n.startswith("tiny")] + ["base_b2_t40_train"])
def test_gradients_match_reference(name):
    g, d, sd, inp = load_case(name)
    opt = g["options"]
    if not opt.get("grads"):
        pytest.skip("no gradients in fixture")
    freeze = opt.get("freeze_lm", False)
    params = {k: v.clone().requires_grad_(not (freeze and k.startswith("bert."))) for k, v in sd.items()}
    o = O.vault_forward(params, d, use_vilt_position_embeddings=opt.get("use_vilt_pos", False), **{k: inp[k] for k in FWD_KEYS})
    loss = O.ce_loss(O.tmsc_logits(params, d, o["pooler_output"]), inp["labels"])
    loss.backward()
    none = sorted(k for k, p in params.items() if p.grad is None or (not p.requires_grad))
    assert none == g["grad_none"], (none, g["grad_none"])
    assert set(g["grad_none"]) >= O.grads_never_set(d, opt.get("use_vilt_pos", False))
    for k, n in g["grad_norm"].items():
        if k.endswith("key.bias"):  # identically-zero gradient (softmax shift invariance): rounding noise only
            assert n < 1e-6 and params[k].grad.norm().item() < 1e-6
            continue
        gn = params[k].grad.norm().item()
        assert abs(gn - n) <= 1e-4 * max(n, 1e-6) + 1e-9, (k, gn, n)
        assert (params[k].grad.flatten()[:32] - g["grad_head"][k]).abs().max() <= 1e-4 * max(n, 1e-6) + 1e-8, k
    for k, full in g["grad_full"].items():
        if k.endswith("key.bias") or full.norm() < 1e-9:  # key.bias: softmax is shift-invariant, its gradient is identically zero (rounding noise)
            continue
        assert cosine(params[k].grad, full) > 1 - 1e-6, k


def test_hf_adamw_rule():
    """transformers==4.48.0 AdamW with correct_bias=False: the first step moves every weight by lr*g/(|g|*sqrt(1-b2)... )."""
    p = torch.tensor([1.0, -2.0, 3.0]); g = torch.tensor([0.5, -0.25, 0.0])
    m = torch.zeros(3); v = torch.zeros(3)
    O.hf_adamw_step(p, g, m, v, step=1, lr=1e-2, beta1=0.9, beta2=0.999, eps=1e-8)
    exp_m = 0.1 * g
    exp_v = 0.001 * g * g
    exp_p = torch.tensor([1.0, -2.0, 3.0]) - 1e-2 * exp_m / (exp_v.sqrt() + 1e-8)
    assert torch.allclose(m, exp_m) and torch.allclose(v, exp_v) and torch.allclose(p, exp_p)
    # weight decay is applied AFTER the Adam update, on the updated weight
    p2 = torch.tensor([1.0]); O.hf_adamw_step(p2, torch.tensor([0.0]), torch.zeros(1), torch.zeros(1), 1, lr=0.1, weight_decay=0.5)
    assert torch.allclose(p2, torch.tensor([0.95]))
    # with bias correction the step is lr*sqrt(1-b2)/(1-b1) times larger at t=1
    p3 = torch.tensor([1.0]); O.hf_adamw_step(p3, torch.tensor([0.5]), torch.zeros(1), torch.zeros(1), 1, lr=1e-2, correct_bias=True)
    step = 1e-2 * (1 - 0.999) ** 0.5 / (1 - 0.9)
    assert torch.allclose(p3, 1.0 - step * 0.05 / (torch.tensor(0.001 * 0.25).sqrt() + 1e-8))


def test_linear_warmup_schedule():
    assert O.linear_warmup_lr(0, 100, 2e-5) == 0.0
    assert abs(O.linear_warmup_lr(5, 100, 2e-5) - 1e-5) < 1e-12
    assert abs(O.linear_warmup_lr(10, 100, 2e-5) - 2e-5) < 1e-12
    assert abs(O.linear_warmup_lr(55, 100, 2e-5) - 1e-5) < 1e-12
    assert O.linear_warmup_lr(100, 100, 2e-5) == 0.0


def test_roberta_position_ids():
    ids = torch.tensor([[5, 6, 7, 1, 1], [9, 1, 1, 1, 1]])
    assert O.roberta_position_ids(ids, 1).tolist() == [[2, 3, 4, 1, 1], [2, 1, 1, 1, 1]]


@pytest.mark.parametrize("kind", ["bert", "roberta", "nolm"])
def test_text_inputs_embeds_matches_reference(kind):
    """``inputs_embeds`` with ``input_ids=None`` (ref:vault/models/vault/model.py:170-200): restatement vs the REAL reference fixture --
    outputs, the gradient returned for the caller's embeddings and the LM / ViLT embedding-table gradients that replace the word lookup's."""
    import os

    from oracle import make_golden_heads as G
    from oracle.ref_loader import hf_configs
    from oracle import synth

    ref = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "heads", f"text_embeds_{kind}.pt"), weights_only=False)
    d = synth.Dims.tiny(**G.TEXT_EMBEDS_KINDS[kind])
    B, T = 3, 16
    inp = synth.make_inputs(d, batch=B, text_len=T, seed=31, var_text=True)
    g = torch.Generator().manual_seed(13)
    text_embeds = torch.randn(B, T, d.hidden, generator=g) * 0.5
    w_pool = torch.randn(B, d.hidden, generator=g)
    w_text = torch.randn(B, T, d.hidden, generator=g) * 0.1
    shapes = synth.param_shapes(d, head=False)
    sd = synth.fill_parameters(shapes, d, seed=0)
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    te = text_embeds.clone().requires_grad_(True)
    o = O.vault_forward(params, d, None, inp["attention_mask"], inp["token_type_ids"], inp["pixel_values"], inp["pixel_mask"], inputs_embeds=te)
    assert tuple(o["last_hidden_state"].shape) == ref["lhs_shape"]
    assert (o["pooler_output"] - ref["pooler_output"]).abs().max() < 2e-5
    assert (o["last_hidden_state"][:, :T] - ref["lhs_text"]).abs().max() < 2e-5
    ((o["pooler_output"] * w_pool).sum() + (o["last_hidden_state"][:, :T] * w_text).sum()).backward()
    assert cosine(te.grad, ref["d_inputs_embeds"]) > 1 - 1e-6
    for k, g_ref in ref["grads"].items():
        assert cosine(params[k].grad, g_ref) > 1 - 1e-6, k
    word = "bert.embeddings.word_embeddings.weight" if kind != "nolm" else "embeddings.text_embeddings.word_embeddings.weight"
    assert params[word].grad is None  # the lookup never ran


def test_hidden_states_match_reference():
    """output_hidden_states: the restatement's per-layer ViLT hidden states against the REAL reference's (fixture written by
    oracle/make_golden_hidden.py: text rows, image CLS row, valid image rows un-permuted into raster order)."""
    import os

    g = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "heads", "hidden_states_tiny.pt"), weights_only=False)
    d = getattr(synth.Dims, g["case"]["dims_factory"])()
    sd = synth.make_state_dict(d, seed=0)
    inp = synth.make_inputs(d, seed=g["seed"], **g["case"]["input_kwargs"])
    o = O.vault_forward(sd, d, output_hidden_states=True, **{k: inp[k] for k in ("input_ids", "attention_mask", "token_type_ids", "pixel_values", "pixel_mask")})
    T = inp["input_ids"].shape[1]
    assert len(o["hidden_states"]) == g["n_hidden"] == d.layers + 1
    for k, h in enumerate(o["hidden_states"]):
        assert (h[:, :T] - g["text"][k]).abs().max().item() < 1e-5 and (h[:, T] - g["image_cls"][k]).abs().max().item() < 1e-5
        for b in range(h.shape[0]):
            nv = int(g["n_valid_patches"][b])
            assert (h[b, T + 1:T + 1 + nv] - g["image_raster"][k][b]).abs().max().item() < 1e-5
