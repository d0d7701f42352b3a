"""Load the REAL reference (gchochla/VAuLT ``model.py``) on top of the installed HuggingFace ViLT/BERT code.

TEST INFRASTRUCTURE.  Works only where ``/root/reference`` exists (the build container); used by
``oracle/make_golden.py`` to produce the committed fixtures and by ``bench.py --impl reference`` when available.
Recipe recorded in SURVEY.md Appendix A:

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
_MOD = None


def available() -> bool:
    return os.path.exists(os.path.join(REF_ROOT, "vault/models/vault/model.py"))


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
        raise RuntimeError(f"reference not present at {REF_ROOT}")
    if "vault" not in sys.modules:
        vault = types.ModuleType("vault")
        vault.__path__ = []
        vu = types.ModuleType("vault.utils")

        def set_parameter_requires_grad(model, requires_grad=False):
            for p in model.parameters():
                p.requires_grad_(requires_grad)

        vu.set_parameter_requires_grad = set_parameter_requires_grad
        sys.modules["vault"], sys.modules["vault.utils"] = vault, vu
    spec = importlib.util.spec_from_file_location("ref_vault_model", os.path.join(REF_ROOT, "vault/models/vault/model.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["ref_vault_model"] = mod
    spec.loader.exec_module(mod)
    from transformers.models.vilt import modeling_vilt

    modeling_vilt.TextEmbeddings.forward = _patched_text_embeddings_forward
    _MOD = mod
    return mod


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
