"""Load the REAL reference (gchochla/VAuLT ``model.py``, and its trainer / EarlyStopping) on top of the installed HuggingFace ViLT/BERT code.

TEST INFRASTRUCTURE.  Source of the files: ``/root/reference`` where it exists (the build container), else the byte-identical copies
``oracle/build_ref.py`` put under ``oracle/_ref/`` (git-ignored build output that travels to the GPU box).  Used by
``oracle/make_golden.py`` to produce the committed fixtures, by the CPU tests that pin the trainer / checkpoint rows, and by
``bench.py --impl reference`` / its ``cpu_baseline`` leg.  Recipe recorded in SURVEY.md Appendix A:

1. stub ``vault.utils.set_parameter_requires_grad`` (ref:vault/utils.py:78-88) -- the real module imports ekphrasis/emoji;
2. exec ``ref:vault/models/vault/model.py`` by path, registered in ``sys.modules`` first;
3. restore the transformers==4.48.0 gate in ViLT ``TextEmbeddings.forward``: position embeddings are added only when
   ``position_embedding_type == "absolute"`` (5.x deleted the attribute, which silently turns the reference's
   ref:vault/models/vault/model.py:77-79 into a no-op).
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import torch

REF_ROOT = os.environ.get("VAULT_REFERENCE_ROOT", "/root/reference")
_REF_COPY = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
_MOD = None
_TRAINER = None


def _path(src: str, copy_name: str):
    """The reference file: in the reference tree if present, else its copy under oracle/_ref/ (None if neither exists)."""
    for cand in (os.path.join(REF_ROOT, src), os.path.join(_REF_COPY, copy_name)):
        if os.path.exists(cand):
            return cand
    return None


def available() -> bool:
    return _path("vault/models/vault/model.py", "vault_model.py") is not None


def trainer_available() -> bool:
    return _path("vault/tmsc_utils/trainer.py", "tmsc_trainer.py") is not None and _path("vault/train_utils.py", "train_utils.py") is not None


def _patched_text_embeddings_forward(self, input_ids=None, token_type_ids=None, position_ids=None, inputs_embeds=None):
    # transformers==4.48.0 semantics of HF:models/vilt/modeling_vilt.py:240-272
    if input_ids is not None:
        input_shape = input_ids.size()
    else:
        input_shape = inputs_embeds.size()[:-1]
    seq_length = input_shape[1]
    if token_type_ids is None:
        token_type_ids = torch.zeros(input_shape, dtype=torch.long, device=self.position_ids.device)
    if inputs_embeds is None:
        inputs_embeds = self.word_embeddings(input_ids)
    embeddings = inputs_embeds + self.token_type_embeddings(token_type_ids)
    if getattr(self, "position_embedding_type", "absolute") == "absolute":
        if position_ids is None:
            position_ids = self.position_ids[:, :seq_length]
        embeddings = embeddings + self.position_embeddings(position_ids)
    embeddings = self.LayerNorm(embeddings)
    return self.dropout(embeddings)


def load_reference_module():
    global _MOD
    if _MOD is not None:
        return _MOD
    if not available():
        raise RuntimeError(f"reference not present at {REF_ROOT} nor copied to {_REF_COPY} (python -m oracle.build_ref)")
    if "vault" not in sys.modules:
        vault = types.ModuleType("vault")
        vault.__path__ = []
        vu = types.ModuleType("vault.utils")

        def set_parameter_requires_grad(model, requires_grad=False):
            for p in model.parameters():
                p.requires_grad_(requires_grad)

        vu.set_parameter_requires_grad = set_parameter_requires_grad
        sys.modules["vault"], sys.modules["vault.utils"] = vault, vu
    spec = importlib.util.spec_from_file_location("ref_vault_model", _path("vault/models/vault/model.py", "vault_model.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["ref_vault_model"] = mod
    spec.loader.exec_module(mod)
    from transformers.models.vilt import modeling_vilt

    modeling_vilt.TextEmbeddings.forward = _patched_text_embeddings_forward
    _MOD = mod
    return mod


class HFAdamW(torch.optim.Optimizer):
    """``transformers.optimization.AdamW`` of the pinned 4.48.0 (the class was removed in 5.x), restated as a torch Optimizer so that the
    reference trainer -- which imports it by name, ref:vault/tmsc_utils/trainer.py:11,244-254 -- runs unchanged.  Rule per parameter with a
    gradient (HF:optimization.py AdamW.step): m, v moments; denom = sqrt(v) + eps; step_size = lr (x sqrt(1-b2^t)/(1-b1^t) if correct_bias);
    p -= step_size * m / denom; then p -= lr * wd * p (decoupled decay AFTER the update)."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-6, weight_decay=0.0, correct_bias=True, no_deprecation_warning=False):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, correct_bias=correct_bias))

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        for group in self.param_groups:
            for p in group["params"]:
                if p.grad is None:
                    continue
                state = self.state[p]
                if len(state) == 0:
                    state["step"] = 0
                    state["exp_avg"] = torch.zeros_like(p)
                    state["exp_avg_sq"] = torch.zeros_like(p)
                m, v = state["exp_avg"], state["exp_avg_sq"]
                b1, b2 = group["betas"]
                state["step"] += 1
                m.mul_(b1).add_(p.grad, alpha=1.0 - b1)
                v.mul_(b2).addcmul_(p.grad, p.grad, value=1.0 - b2)
                denom = v.sqrt().add_(group["eps"])
                step_size = group["lr"]
                if group["correct_bias"]:
                    step_size = step_size * (1.0 - b2 ** state["step"]) ** 0.5 / (1.0 - b1 ** state["step"])
                p.addcdiv_(m, denom, value=-step_size)
                if group["weight_decay"] > 0.0:
                    p.add_(p, alpha=-group["lr"] * group["weight_decay"])
        return loss


def load_reference_trainer():
    """(trainer module, train_utils module) of the reference: ``Twitter201XTrainer`` (ref:vault/tmsc_utils/trainer.py) and ``EarlyStopping``
    (ref:vault/train_utils.py:13-160), loaded by path.  Their un-importable neighbours are stubbed: ``vault.utils.flatten_list`` is restated
    (ref:vault/utils.py:91-117, pure list code), ``vault.logging_utils.ExperimentHandler`` (matplotlib / yaml; only a type annotation in
    the trainer) is an empty class, and ``transformers.optimization.AdamW`` (4.48.0; gone in 5.x) is HFAdamW above."""
    global _TRAINER
    if _TRAINER is not None:
        return _TRAINER
    if not trainer_available():
        raise RuntimeError("reference trainer files not present (python -m oracle.build_ref)")
    load_reference_module()  # installs the `vault` / `vault.utils` stubs
    vu = sys.modules["vault.utils"]
    if not hasattr(vu, "flatten_list"):
        def flatten_list(l, order=None):  # ref:vault/utils.py:91-117
            if not isinstance(l, list):
                l = list(l)
            if order is None:
                lc, order = l, 0
                while isinstance(lc, list) and lc:
                    lc = lc[0]
                    order += 1
            if order == 1:
                return l
            return [x for sub in l for x in flatten_list(sub, order - 1)]
        vu.flatten_list = flatten_list
    import transformers.optimization as topt

    if not hasattr(topt, "AdamW"):
        topt.AdamW = HFAdamW
    if "vault.logging_utils" not in sys.modules:
        lu = types.ModuleType("vault.logging_utils")
        lu.ExperimentHandler = type("ExperimentHandler", (), {})
        sys.modules["vault.logging_utils"] = lu
    spec = importlib.util.spec_from_file_location("vault.train_utils", _path("vault/train_utils.py", "train_utils.py"))
    tu = importlib.util.module_from_spec(spec)
    sys.modules["vault.train_utils"] = tu
    spec.loader.exec_module(tu)
    spec = importlib.util.spec_from_file_location("ref_tmsc_trainer", _path("vault/tmsc_utils/trainer.py", "tmsc_trainer.py"))
    tr = importlib.util.module_from_spec(spec)
    sys.modules["ref_tmsc_trainer"] = tr
    spec.loader.exec_module(tr)
    _TRAINER = (tr, tu)
    return _TRAINER


def hf_configs(d):
    """Dims -> (ViltConfig, LM config or None)."""
    from transformers import BertConfig, RobertaConfig, ViltConfig

    vc = ViltConfig(
        vocab_size=d.vilt_vocab, type_vocab_size=d.vilt_type_vocab, modality_type_vocab_size=d.modality_vocab,
        max_position_embeddings=d.vilt_max_pos, hidden_size=d.hidden, num_hidden_layers=d.layers,
        num_attention_heads=d.heads, intermediate_size=d.inter, image_size=d.image_size, patch_size=d.patch,
        num_channels=d.channels, layer_norm_eps=d.vilt_eps,
    )
    if d.lm_layers == 0:
        return vc, None
    common = dict(
        vocab_size=d.lm_vocab, hidden_size=d.hidden, num_hidden_layers=d.lm_layers, num_attention_heads=d.heads,
        intermediate_size=d.inter, max_position_embeddings=d.lm_max_pos, type_vocab_size=d.lm_type_vocab,
        layer_norm_eps=d.lm_eps, pad_token_id=d.lm_pad_id, hidden_dropout_prob=d.lm_dropout,
        attention_probs_dropout_prob=d.lm_dropout,
    )
    if d.lm_kind == "roberta":
        lc = RobertaConfig(bos_token_id=0, eos_token_id=2, **common)
    else:
        lc = BertConfig(**common)
    return vc, lc


def build_reference_tmsc(d, sd, use_vilt_position_embeddings: bool = False):
    """The reference ``VaultForTMSC`` built from configs, with the synthetic state_dict loaded."""
    mod = load_reference_module()
    vc, lc = hf_configs(d)
    m = mod.VaultForTMSC(vc, n_classes=d.n_classes, vilt_dropout_prob=d.head_dropout, bert_config=lc)
    # constructor path sets the flag on the config only (5.x ignores it): set it on the module, as from_pretrained does
    # (ref:vault/models/vault/model.py:112-116)
    m.embeddings.text_embeddings.position_embedding_type = (
        "NOT_absolute" if (lc is not None and not use_vilt_position_embeddings) else "absolute"
    )
    missing, unexpected = m.load_state_dict(sd, strict=False)
    missing = [k for k in missing if "position_ids" not in k and "token_type_ids" not in k]
    assert not missing and not unexpected, (missing, unexpected)
    return m, mod
