"""GPU parity probe: vault_b200 (CUDA, bf16 operands / fp32 accumulate) vs the fp32 CPU oracle on the golden cases.

    python tools/parity_probe.py [case ...]
"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import vault_oracle as O  # noqa: E402
from oracle.ref_loader import hf_configs  # noqa: E402
from tests.golden_utils import cosine, golden_names, load_case  # noqa: E402
from vault_b200 import VaultForTMSC  # noqa: E402

FWD = ("input_ids", "attention_mask", "token_type_ids", "pixel_values", "pixel_mask")
dev = torch.device("cuda:0")


def build(d, sd, opt):
    vc, lc = hf_configs(d)
    m = VaultForTMSC(vc, n_classes=d.n_classes, vilt_dropout_prob=d.head_dropout, bert_config=lc)
    m.embeddings.text_embeddings.position_embedding_type = "absolute" if (lc is None or opt.get("use_vilt_pos")) else "NOT_absolute"
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected and all("_ids" in k for k in missing), (missing, unexpected)
    if opt.get("freeze_lm"):
        m.freeze_lm = True
        for p in m.bert.parameters():
            p.requires_grad_(False)
    return m.to(dev).eval()


def run(name, do_grads=True):
    g, d, sd, inp = load_case(name)
    opt = g["options"]
    t0 = time.time()
    m = build(d, sd, opt)
    cu = {k: inp[k].to(dev) for k in FWD}
    T = inp["input_ids"].shape[1]
    with torch.no_grad():
        out = m._trunk(**cu)
    lhs, pooled = out[0].float().cpu(), out[1].float().cpu()
    with torch.no_grad():
        o = O.vault_forward(sd, d, use_vilt_position_embeddings=opt.get("use_vilt_pos", False), **{k: inp[k] for k in FWD})
    ref_pool = g["pooler_output"]
    res = dict(case=name, lhs_shape=list(lhs.shape), ref_shape=list(g["lhs_shape"]))
    res["pooler_rel_inf"] = ((pooled - ref_pool).abs().max() / ref_pool.abs().max()).item()
    res["pooler_rel_l2_row_max"] = ((pooled - ref_pool).norm(dim=1) / ref_pool.norm(dim=1)).max().item()
    mask = o["mask"].bool()
    diff = (lhs - o["last_hidden_state"]).abs()
    res["lhs_text_max_abs"] = diff[:, :T][mask[:, :T]].max().item()
    res["lhs_img_max_abs"] = diff[:, T:][mask[:, T:]].max().item()
    res["lhs_ref_absmax"] = o["last_hidden_state"].abs().max().item()
    if do_grads and opt.get("grads"):
        m.zero_grad()
        logits = m(**cu)
        loss = torch.nn.functional.cross_entropy(logits.float(), inp["labels"].to(dev))
        loss.backward()
        res["loss"] = loss.item()
        res["loss_ref"] = g["loss"].item()
        res["logits_max_abs"] = (logits.float().cpu() - g["logits"]).abs().max().item()
        # oracle grads
        freeze = opt.get("freeze_lm", False)
        params = {k: v.clone().requires_grad_(not (freeze and k.startswith("bert."))) for k, v in sd.items()}
        oo = O.vault_forward(params, d, use_vilt_position_embeddings=opt.get("use_vilt_pos", False), **{k: inp[k] for k in FWD})
        O.ce_loss(O.tmsc_logits(params, d, oo["pooler_output"]), inp["labels"]).backward()
        worst, dots, n1, n2 = [], 0.0, 0.0, 0.0
        none_mine = sorted(k for k, p in m.named_parameters() if p.grad is None)
        none_ref = sorted(k for k, p in params.items() if p.grad is None)
        res["grad_none_match"] = none_mine == none_ref
        if not res["grad_none_match"]:
            res["grad_none_diff"] = sorted(set(none_mine) ^ set(none_ref))[:8]
        for k, p in m.named_parameters():
            if p.grad is None or params[k].grad is None:
                continue
            a, b = p.grad.float().cpu(), params[k].grad
            if k.endswith("key.bias") or b.norm() < 1e-9:
                continue
            c = cosine(a, b)
            worst.append((c, k, (a.norm() / b.norm()).item()))
            dots += (a.double().flatten() @ b.double().flatten()).item()
            n1 += a.double().norm().item() ** 2
            n2 += b.double().norm().item() ** 2
        worst.sort()
        res["grad_cos_global"] = dots / (n1 ** 0.5 * n2 ** 0.5)
        res["grad_cos_min"] = [(round(c, 5), k, round(r, 4)) for c, k, r in worst[:6]]
    res["seconds"] = time.time() - t0
    print(json.dumps(res), flush=True)


if __name__ == "__main__":
    names = sys.argv[1:] or golden_names()
    for n in names:
        try:
            run(n)
        except Exception as e:  # noqa: BLE001
            import traceback

            traceback.print_exc()
            print(json.dumps(dict(case=n, exception=repr(e))), flush=True)
            if "CUDA" in repr(e) or "launch" in repr(e):
                break
