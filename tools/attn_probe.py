"""GPU probe: fused attention fwd/bwd vs torch fp32 on the same bf16-rounded inputs (+ timing)."""
import ctypes as C, json, math, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vault_b200 import _abi

dev = torch.device("cuda:0")
lib = _abi.lib()
st = lambda: torch.cuda.current_stream().cuda_stream


def run(B, S, heads, p=0.0, seed=0, mask_mode="mid"):
    H = heads * 64
    torch.manual_seed(S * 7 + B)
    qkv = (torch.randn(B * S, 3 * H, device=dev) * 0.7).to(torch.bfloat16)
    mask = torch.ones(B, S, dtype=torch.uint8, device=dev)
    if mask_mode == "mid":
        for b in range(B):
            a = 3 + (b * 5) % max(1, S // 3); z = min(S - 1, a + (b * 3) % max(1, S // 4))
            mask[b, a:z] = 0
            if b % 2: mask[b, S - (b % 7) - 1:] = 0
    ctx = torch.empty(B * S, H, device=dev, dtype=torch.bfloat16); lse = torch.empty(B, heads, S, device=dev)
    _abi.check(lib.vault_attn_fwd(qkv.data_ptr(), mask.data_ptr(), ctx.data_ptr(), lse.data_ptr(), B, S, heads, p, seed, None, 5, st()), "fwd")
    x = qkv.float().view(B, S, 3, heads, 64).permute(2, 0, 3, 1, 4).clone().requires_grad_(True)
    q, k, v = x[0], x[1], x[2]
    sc = q @ k.transpose(-1, -2) / 8.0
    sc = sc.masked_fill(mask[:, None, None, :] == 0, float("-inf"))
    pr = torch.softmax(sc, -1)
    ref = (pr @ v).permute(0, 2, 1, 3).reshape(B * S, H)
    ref_lse = torch.logsumexp(sc, -1)
    dctx = (torch.randn(B * S, H, device=dev) * 0.5).to(torch.bfloat16)
    res = dict(case=f"attn B{B} S{S} h{heads} p{p}")
    if p == 0.0:
        res["ctx_err"] = (ctx.float() - ref).abs().max().item()
        res["lse_err"] = (lse - ref_lse).abs().max().item()
        ref.backward(dctx.float())
        dref = x.grad.permute(1, 3, 0, 2, 4).reshape(B * S, 3 * H)
        dqkv = torch.empty_like(qkv); delta = torch.empty(B, heads, S, device=dev)
        _abi.check(lib.vault_attn_bwd(qkv.data_ptr(), mask.data_ptr(), ctx.data_ptr(), dctx.data_ptr(), lse.data_ptr(), delta.data_ptr(), dqkv.data_ptr(),
                                      B, S, heads, 0.0, 0, None, 5, st()), "bwd")
        d = (dqkv.float() - dref).abs()
        scale = dref.abs().max().item()
        res.update(dq_err=d[:, :H].max().item() / scale, dk_err=d[:, H:2 * H].max().item() / scale, dv_err=d[:, 2 * H:].max().item() / scale)
        res["ok"] = bool(res["ctx_err"] < 2e-2 and res["lse_err"] < 1e-2 and max(res["dq_err"], res["dk_err"], res["dv_err"]) < 2e-2)
    else:
        # dropout: E[ctx] == ref over seeds, fwd deterministic, and finite-difference-free bwd check through linearity in V:
        acc = torch.zeros_like(ref); n = 24
        for s in range(n):
            _abi.check(lib.vault_attn_fwd(qkv.data_ptr(), mask.data_ptr(), ctx.data_ptr(), lse.data_ptr(), B, S, heads, p, 1000 + s, None, 5, st()), "fwd")
            acc += ctx.float()
        res["mean_err"] = ((acc / n - ref).abs().mean() / ref.abs().mean()).item()
        c1 = torch.empty_like(ctx); c2 = torch.empty_like(ctx)
        lib.vault_attn_fwd(qkv.data_ptr(), mask.data_ptr(), c1.data_ptr(), lse.data_ptr(), B, S, heads, p, 77, None, 5, st())
        lib.vault_attn_fwd(qkv.data_ptr(), mask.data_ptr(), c2.data_ptr(), lse.data_ptr(), B, S, heads, p, 77, None, 5, st())
        res["deterministic"] = bool(torch.equal(c1, c2))
        # backward consistency: <dctx, ctx> is linear in V  =>  <dV, V> == <dctx, ctx> when the same mask is regenerated
        dqkv = torch.empty_like(qkv); delta = torch.empty(B, heads, S, device=dev)
        _abi.check(lib.vault_attn_bwd(qkv.data_ptr(), mask.data_ptr(), c1.data_ptr(), dctx.data_ptr(), lse.data_ptr(), delta.data_ptr(), dqkv.data_ptr(),
                                      B, S, heads, p, 77, None, 5, st()), "bwd")
        lhs = (dqkv[:, 2 * H:].float() * qkv[:, 2 * H:].float()).sum().item(); rhs = (dctx.float() * c1.float()).sum().item()
        # normalised by the norms, not by |rhs|: rhs is a random-sign sum and can be small by chance (the exact-mask check lives in
        # tests/test_kernels_gpu.py::test_attention_dropout_exact_mask)
        res["dv_linearity_rel"] = abs(lhs - rhs) / max((dctx.float().norm() * c1.float().norm()).item(), 1e-6)
        # softmax rows: sum_k dS = 0  =>  <dQ, Q> == <dK, K>
        a = (dqkv[:, :H].float() * qkv[:, :H].float()).sum().item(); b = (dqkv[:, H:2 * H].float() * qkv[:, H:2 * H].float()).sum().item()
        res["dq_dk_balance_rel"] = abs(a - b) / max(abs(a), abs(b), 1e-6)
        res["ok"] = bool(res["mean_err"] < 0.08 and res["deterministic"] and res["dv_linearity_rel"] < 2e-3 and res["dq_dk_balance_rel"] < 5e-2)
    print(json.dumps(res), flush=True)


for args in [(2, 64, 2), (3, 185, 12), (2, 369, 12), (2, 40, 12), (1, 17, 2), (4, 128, 12), (2, 209, 12)]:
    run(*args)
run(2, 40, 12, p=0.1); run(2, 128, 12, p=0.1)
# timing at the BASELINE shapes
for B, S, heads in ((32, 185, 12), (32, 369, 12), (32, 40, 12)):
    H = heads * 64
    qkv = torch.randn(B * S, 3 * H, device=dev).to(torch.bfloat16); mask = torch.ones(B, S, dtype=torch.uint8, device=dev)
    ctx = torch.empty(B * S, H, device=dev, dtype=torch.bfloat16); lse = torch.empty(B, heads, S, device=dev)
    dctx = torch.randn(B * S, H, device=dev).to(torch.bfloat16); dqkv = torch.empty_like(qkv); delta = torch.empty(B, heads, S, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f = lambda: lib.vault_attn_fwd(qkv.data_ptr(), mask.data_ptr(), ctx.data_ptr(), lse.data_ptr(), B, S, heads, 0.0, 0, None, 0, st())
    g = lambda: lib.vault_attn_bwd(qkv.data_ptr(), mask.data_ptr(), ctx.data_ptr(), dctx.data_ptr(), lse.data_ptr(), delta.data_ptr(), dqkv.data_ptr(), B, S, heads, 0.0, 0, None, 0, st())
    out = {}
    for nm, fn, mult in (("fwd", f, 4), ("bwd", g, 14)):
        for _ in range(3): fn()
        torch.cuda.synchronize(); e0.record()
        for _ in range(20): fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        out[nm + "_ms"] = ms; out[nm + "_tflops"] = mult * B * heads * S * S * 64 / ms / 1e9
    print(json.dumps(dict(case=f"attn time B{B} S{S}", **out)), flush=True)
