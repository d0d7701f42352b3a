"""GPU parity of the whole hot path (drop-in classes -> engine -> C ABI -> CUDA) against the fp32 CPU oracle and the committed
reference fixtures, plus size-independent properties at the BASELINE batch sizes.

Stated tolerance (bf16 tensor-core operands, fp32 accumulation / residual stream / statistics, 24 transformer layers):
  pooler_output : per-row relative L2 <= 1e-2 (the north-star's "within 1e-2 relative of the reference"; measured 5.5e-3 at the base
                  shapes)  and  max-abs / max|ref| <= 1.5e-2 -- the single worst ELEMENT of a 768-wide tanh output relative to the largest
                  one; measured 0.9e-2 .. 1.1e-2 at the base shapes, 4e-4 at the tiny ones.  It cannot be brought under 1e-2 without
                  leaving bf16: 24 layers of bf16-rounded GEMM operands put ~1e-2 on individual elements while the row as a whole (the
                  relative-L2 figure, what the target sentence bounds) stays at half of that
  loss          : |d| <= 5e-3 ;  gradients: per-parameter cosine >= 0.99, global cosine >= 0.999
"""
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import synth, vault_oracle as O  # noqa: E402
from oracle.ref_loader import hf_configs  # noqa: E402
from tests.golden_utils import cosine, golden_names, load_case  # noqa: E402

FWD = ("input_ids", "attention_mask", "token_type_ids", "pixel_values", "pixel_mask")
DEV = "cuda:0"


def build(d, sd, opt=None, n_classes=None):
    from vault_b200 import VaultForTMSC

    opt = opt or {}
    vc, lc = hf_configs(d)
    m = VaultForTMSC(vc, n_classes=d.n_classes, vilt_dropout_prob=d.head_dropout, bert_config=lc)
    m.embeddings.text_embeddings.position_embedding_type = "absolute" if (lc is None or opt.get("use_vilt_pos")) else "NOT_absolute"
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected and all(k.endswith("_ids") for k in missing)
    if opt.get("freeze_lm"):
        m.freeze_lm = True
        for p in m.bert.parameters():
            p.requires_grad_(False)
    return m.to(DEV).eval()


@pytest.mark.parametrize("name", golden_names())
def test_forward_vs_reference_fixture_and_oracle(name):
    g, d, sd, inp = load_case(name)
    m = build(d, sd, g["options"])
    cu = {k: inp[k].to(DEV) for k in FWD}
    T = inp["input_ids"].shape[1]
    with torch.no_grad():
        out = m.__class__.__mro__[1].forward(m, **cu)  # VaultModel.forward of a VaultForTMSC instance, as in the README usage
    lhs, pooled = out.last_hidden_state.float().cpu(), out.pooler_output.float().cpu()
    assert tuple(lhs.shape) == g["lhs_shape"]  # same dynamic sequence length as the reference (max valid patches in the batch)
    ref = g["pooler_output"]  # REAL reference output
    assert ((pooled - ref).norm(dim=1) / ref.norm(dim=1)).max() <= 1e-2
    assert (pooled - ref).abs().max() / ref.abs().max() <= 1.5e-2
    assert (lhs[:, 0] - g["lhs_cls"]).abs().max() <= 6e-2 and (lhs[:, T] - g["lhs_image_cls"]).abs().max() <= 6e-2
    with torch.no_grad():
        o = O.vault_forward(sd, d, use_vilt_position_embeddings=g["options"].get("use_vilt_pos", False), **{k: inp[k] for k in FWD})
    mask = o["mask"].bool()
    diff = (lhs - o["last_hidden_state"]).abs()
    assert diff[mask].max() <= 6e-2  # every valid row (text, image CLS, raster-ordered patches); |values| ~ 6
    assert torch.equal(out.last_hidden_state.new_tensor(mask.float()).cpu().bool(), mask)


@pytest.mark.parametrize("name", [n for n in golden_names() if n.startswith("tiny")] + ["base_b2_t40_train", "bertweet_b2_t128_mixed"])
def test_backward_vs_oracle(name):
    g, d, sd, inp = load_case(name)
    opt = g["options"]
    if not opt.get("grads"):
        pytest.skip("no gradients in this fixture")
    m = build(d, sd, opt)
    cu = {k: inp[k].to(DEV) for k in FWD}
    logits = m(**cu)
    loss = torch.nn.functional.cross_entropy(logits.float(), inp["labels"].to(DEV))
    loss.backward()
    assert abs(loss.item() - g["loss"].item()) <= 5e-3
    assert (logits.float().cpu() - g["logits"]).abs().max() <= 1e-2
    assert sorted(k for k, p in m.named_parameters() if p.grad is None) == g["grad_none"]
    freeze = opt.get("freeze_lm", False)
    params = {k: v.clone().requires_grad_(not (freeze and k.startswith("bert."))) for k, v in sd.items()}
    oo = O.vault_forward(params, d, use_vilt_position_embeddings=opt.get("use_vilt_pos", False), **{k: inp[k] for k in FWD})
    O.ce_loss(O.tmsc_logits(params, d, oo["pooler_output"]), inp["labels"]).backward()
    dot = n1 = n2 = 0.0
    for k, p in m.named_parameters():
        if p.grad is None:
            continue
        a, b = p.grad.float().cpu().double().flatten(), params[k].grad.double().flatten()
        if k.endswith("key.bias") or b.norm() < 1e-9:
            continue  # identically-zero gradient
        assert cosine(a, b) >= 0.99, (k, cosine(a, b))
        assert abs(a.norm() / b.norm() - 1) <= 0.05, (k, (a.norm() / b.norm()).item())
        assert abs(a.norm().item() - g["grad_norm"][k]) <= 0.05 * g["grad_norm"][k] + 1e-9, k  # vs the REAL reference
        dot += float(a @ b); n1 += float(a @ a); n2 += float(b @ b)
    assert dot / (n1 ** 0.5 * n2 ** 0.5) >= 0.999


def test_output_hidden_states_vs_reference_fixture():
    """model(..., output_hidden_states=True): the ViLT encoder's hidden states (embedding output + every layer's output, image rows in raster
    order) against the REAL reference's (tests/golden/heads/hidden_states_tiny.pt), tuple layout as HF's BaseModelOutputWithPooling."""
    import os

    g = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "heads", "hidden_states_tiny.pt"), weights_only=False)
    d = getattr(synth.Dims, g["case"]["dims_factory"])()
    sd = synth.make_state_dict(d, seed=0)
    inp = synth.make_inputs(d, seed=g["seed"], **g["case"]["input_kwargs"])
    m = build(d, sd)
    cu = {k: inp[k].to(DEV) for k in FWD}
    T = inp["input_ids"].shape[1]
    with torch.no_grad():
        out = m.__class__.__mro__[1].forward(m, output_hidden_states=True, **cu)
        tup = m.__class__.__mro__[1].forward(m, output_hidden_states=True, return_dict=False, **cu)
        plain = m.__class__.__mro__[1].forward(m, **cu)
    assert plain.hidden_states is None and len(tup) == 3 and len(tup[2]) == g["n_hidden"] == len(out.hidden_states)
    for k, h in enumerate(out.hidden_states):
        h = h.float().cpu()
        scale = max(1.0, g["text"][k].abs().max().item())
        assert (h[:, :T] - g["text"][k]).abs().max().item() <= 2e-2 * scale and (h[:, T] - g["image_cls"][k]).abs().max().item() <= 2e-2 * scale
        for b in range(h.shape[0]):
            nv = int(g["n_valid_patches"][b])
            assert (h[b, T + 1:T + 1 + nv] - g["image_raster"][k][b]).abs().max().item() <= 2e-2 * scale
    assert torch.equal(out.last_hidden_state, plain.last_hidden_state)


def test_train_step_matches_oracle_one_step():
    """VaultTrainStep (CUDA graph, fused CE head, fused AdamW) vs the oracle's train_step: loss and post-step weight delta."""
    from vault_b200 import VaultTrainStep

    d = synth.Dims.tiny()
    sd = synth.make_state_dict(d, seed=0)
    batch = synth.make_inputs(d, batch=4, text_len=16, seed=2, var_text=True, mixed_images=True, image_hw=(384, 640))
    m = build(d, sd).train()
    ts = VaultTrainStep(m, lr=1e-3, dropout=False, use_cuda_graph=True)
    before = {k: v.detach().float().cpu().clone() for k, v in m.state_dict().items() if k in sd}
    r1 = ts.step({k: v.pin_memory() for k, v in batch.items()})
    r2 = ts.step({k: v.pin_memory() for k, v in batch.items()})  # second buffer slot / second graph
    ref_sd = {k: v.clone() for k, v in sd.items()}
    o1 = O.train_step(ref_sd, d, batch, lr=1e-3)
    o2 = O.train_step(ref_sd, d, batch, lr=1e-3, state=o1["state"])
    assert abs(r1.loss() - o1["loss"].item()) <= 5e-3
    assert abs(r2.loss() - o2["loss"].item()) <= 2e-2
    torch.cuda.synchronize()
    after = {k: v.detach().float().cpu() for k, v in m.state_dict().items() if k in sd}
    # Adam's first steps move every weight by ~lr*sign(g): compare the update directions on the big tensors
    for k in ("encoder.layer.1.intermediate.dense.weight", "bert.encoder.layer.0.attention.self.value.weight", "classifier.1.weight",
              "embeddings.patch_embeddings.projection.weight"):
        mine, ref = after[k] - before[k], ref_sd[k] - sd[k]
        assert cosine(mine, ref) >= 0.9, (k, cosine(mine, ref))
    for k in O.grads_never_set(d):
        assert torch.equal(after[k], before[k])


def test_train_step_host_pixel_mask_reduced_on_host():
    """A HOST batch's pixel mask never crosses PCIe: step() reduces it to (patch rows, patch columns) per sample, exactly what
    vault_patch_grid computes from a device mask.  Same model, same batch: host path and device path give the same first loss bit for bit
    (the second differs in the last bits only: the gradient atomics are unordered)."""
    from vault_b200 import VaultTrainStep

    d = synth.Dims.tiny()
    sd = synth.make_state_dict(d, seed=0)
    batch = synth.make_inputs(d, batch=4, text_len=16, seed=5, var_text=True, mixed_images=True, image_hw=(384, 640))
    losses = []
    for on_host in (True, False):
        m = build(d, sd).train()
        ts = VaultTrainStep(m, lr=1e-3, dropout=False, use_cuda_graph=True)
        b = {k: (v.pin_memory() if on_host else v.to(DEV)) for k, v in batch.items()}
        losses.append([ts.step(b).loss() for _ in range(2)])
        if on_host:
            assert ts.last_h2d_bytes < sum(v.numel() * v.element_size() for v in batch.values()) - batch["pixel_mask"].numel() * 8 + 64
    assert losses[0][0] == losses[1][0] and abs(losses[0][1] - losses[1][1]) <= 1e-5, losses


@pytest.mark.parametrize("kind,n_classes", [("bce", 1), ("ce2", 6)])
def test_train_step_other_trainer_losses_match_oracle(kind, n_classes):
    """The Bloomberg (one logit, BCE-with-logits) and raw-MVSA (two label groups) losses through the fused step vs the oracle."""
    import dataclasses
    from vault_b200 import VaultForTMSC, VaultTrainStep

    d = dataclasses.replace(synth.Dims.tiny(), n_classes=n_classes)
    sd = synth.make_state_dict(d, seed=1)
    batch = synth.make_inputs(d, batch=4, text_len=16, seed=3, var_text=True)
    g = torch.Generator().manual_seed(0)
    batch["labels"] = torch.randint(0, 2, (4,), generator=g).float() if kind == "bce" else torch.randint(0, 3, (4, 2), generator=g)
    vc, lc = hf_configs(d)
    m = VaultForTMSC(vc, n_classes=n_classes, vilt_dropout_prob=d.head_dropout, bert_config=lc)
    m.embeddings.text_embeddings.position_embedding_type = "NOT_absolute"
    m.load_state_dict(sd, strict=False)
    m = m.to(DEV).train()
    ts = VaultTrainStep(m, lr=1e-3, dropout=False, loss=kind if kind != "bce" else "auto")
    before = m.classifier[1].weight.detach().float().cpu().clone()
    r = ts.step({k: v.pin_memory() for k, v in batch.items()})
    ref_sd = {k: v.clone() for k, v in sd.items()}
    o = O.train_step(ref_sd, d, batch, lr=1e-3, loss_fn=O.bce_loss if kind == "bce" else O.mvsa_raw_loss)
    assert abs(r.loss() - o["loss"].item()) <= 5e-3
    ts.synchronize()
    assert cosine(m.classifier[1].weight.detach().float().cpu() - before, ref_sd["classifier.1.weight"] - sd["classifier.1.weight"]) >= 0.9
    m.eval()
    with torch.no_grad():
        logits = m(**{k: batch[k].to(DEV) for k in FWD})
    assert tuple(logits.shape) == ((4,) if n_classes == 1 else (4, 6))  # ref:vault/models/vault/model.py:569 squeezes the single logit


def test_trainer_drop_in_runs_real_steps_and_evaluates():
    """vault_b200.trainer.VaultTrainerForTMSC end to end on a tiny model: pinned loader -> fused steps -> sharded evaluation."""
    from types import SimpleNamespace
    from vault_b200.trainer import VaultTrainerForTMSC

    d = synth.Dims.tiny()
    sd = synth.make_state_dict(d, seed=0)
    inp = synth.make_inputs(d, batch=24, text_len=16, seed=5, var_text=True)

    class DS(torch.utils.data.Dataset):
        name = "synthetic"

        def __len__(self):
            return 24

        def __getitem__(self, i):
            return (i, inp["input_ids"][i], inp["attention_mask"][i], inp["token_type_ids"][i], inp["pixel_values"][i], inp["pixel_mask"][i],
                    int(inp["labels"][i]))

        @staticmethod
        def collate_fn(items):
            cols = list(zip(*items))
            return [list(cols[0])] + [torch.stack(c) for c in cols[1:6]] + [torch.tensor(cols[6])]

    logged = []
    eh = SimpleNamespace(device=DEV, learning_rate=5e-4, adam_beta1=0.9, adam_beta2=0.999, adam_epsilon=1e-8, weight_decay=0.0, correct_bias=False,
                         train_batch_size=8, eval_batch_size=8, dataloader_num_workers=0, num_train_epochs=4, warmup_ratio=0.1, max_steps=-1,
                         eval_steps=None, disable_tqdm=True, early_stopping_patience=None, model_save=False, model_load_filename=None,
                         set_dict_metrics=lambda r, test=False: logged.append((test, dict(r))))
    m = build(d, sd)
    t = VaultTrainerForTMSC(m, DS(), eh, dev_dataset=DS(), test_dataset=DS())
    res = t.train()
    train_losses = [r["train_loss"] for test, r in logged if not test]
    assert len(train_losses) == 4 and all(torch.isfinite(torch.tensor(train_losses))) and train_losses[-1] < train_losses[0]
    assert set(res) >= {"eval_loss", "eval_accuracy", "macro_f1_score"} and 0.0 <= res["eval_accuracy"] <= 1.0
    assert t.train_step.total_steps == 12 and t.train_step.step_idx == 12  # len(loader) * epochs, the reference's schedule length


def test_graph_and_eager_steps_agree_bitwise_without_dropout():
    from vault_b200 import VaultTrainStep

    d = synth.Dims.tiny()
    sd = synth.make_state_dict(d, seed=0)
    batch = {k: v.to(DEV) for k, v in synth.make_inputs(d, batch=2, text_len=16, seed=4).items()}
    losses = []
    for graph in (True, False):
        m = build(d, sd).train()
        ts = VaultTrainStep(m, lr=1e-3, dropout=False, use_cuda_graph=graph)
        losses.append([ts.step(batch).loss() for _ in range(3)])
    assert losses[0][0] == losses[1][0]           # same kernels, same order
    assert abs(losses[0][2] - losses[1][2]) < 1e-3  # split-K atomics may reorder fp32 sums
    assert losses[0][2] < losses[0][0]            # it learns the repeated batch


def test_dropout_is_active_and_reseeded_in_training_mode():
    d = synth.Dims.tiny()
    sd = synth.make_state_dict(d, seed=0)
    inp = synth.make_inputs(d, batch=2, text_len=16, seed=4)
    cu = {k: inp[k].to(DEV) for k in FWD}
    m = build(d, sd)
    with torch.no_grad():
        e1, e2 = m(**cu), m(**cu)
    assert torch.equal(e1, e2)  # eval: deterministic
    m.train()
    a = m(**cu)
    b = m(**cu)
    assert not torch.equal(a, b)  # BERT-stack + head dropout, fresh masks per forward (ref: model.train() in the trainer)
    a.sum().backward()         # backward regenerates the masks of ITS forward... which was superseded -> must still run


def test_batch_rows_are_independent_at_baseline_size():
    """Size-independent property at the BASELINE shape (B=32, T=40, 384x384, bert-base + vilt-b32): each sample's output equals
    what it gets alone (padding and the other rows do not leak), and masked text padding can be extended freely."""
    d = synth.Dims.base()
    sd = synth.make_state_dict(d, seed=0)
    inp = synth.make_inputs(d, batch=32, text_len=40, seed=9, var_text=True)
    m = build(d, sd)
    cu = {k: inp[k].to(DEV) for k in FWD}
    with torch.no_grad():
        full = m._trunk(**cu)[1].float()
        solo = m._trunk(**{k: v[5:6] for k, v in cu.items()})[1].float()
        # extend the padding: T 40 -> 64 with masked pad tokens
        ext = {k: v[5:6] for k, v in cu.items()}
        pad = torch.zeros(1, 24, dtype=torch.long, device=DEV)
        for k in ("input_ids", "attention_mask", "token_type_ids"):
            ext[k] = torch.cat([ext[k], pad], dim=1)
        longer = m._trunk(**ext)[1].float()
    assert (full[5] - solo[0]).abs().max() <= 2e-3   # same arithmetic up to tile-boundary effects in bf16 GEMM inputs: none expected
    assert (longer[0] - solo[0]).abs().max() <= 1e-2  # key chunking of the online softmax shifts -> bf16-level differences over 24 layers
    assert torch.isfinite(full).all()


def test_inference_config2_batch64_matches_oracle_rows():
    """BASELINE config 2: VaultModel inference, batch 64, T=64, 384x384 (S=209), bf16 operands vs the fp32 oracle.  The oracle runs
    on a subset of rows (batch rows are independent -- checked separately), the CUDA path on the full batch."""
    d = synth.Dims.base()
    sd = synth.make_state_dict(d, seed=0)
    inp = synth.make_inputs(d, batch=64, text_len=64, seed=11)
    m = build(d, sd)
    cu = {k: inp[k].to(DEV) for k in FWD}
    with torch.no_grad():
        lhs, pooled, _ = m._trunk(**cu)
    assert tuple(lhs.shape) == (64, 209, 768) and torch.isfinite(pooled).all()
    rows = [0, 31, 63]
    sub = {k: inp[k][rows] for k in FWD}
    with torch.no_grad():
        o = O.vault_forward(sd, d, **sub)
    got, ref = pooled[rows].float().cpu(), o["pooler_output"]
    assert ((got - ref).norm(dim=1) / ref.norm(dim=1)).max() <= 1e-2
    assert (got - ref).abs().max() / ref.abs().max() <= 1.5e-2  # worst element (see the module docstring)
