"""GPU probe: tcgen05 attention (attention_tc.cu) vs the mma.sync kernels vs torch fp32, plus timing.  usage: attn_tc_probe.py fwd|bwd"""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vault_b200 import _abi

mode = sys.argv[1] if len(sys.argv) > 1 else "fwd"
dev = torch.device("cuda:0")
lib = _abi.lib()
st = lambda: torch.cuda.current_stream().cuda_stream


def mk(B, S, heads):
    H = heads * 64
    torch.manual_seed(S * 7 + B)
    qkv = (torch.randn(B * S, 3 * H, device=dev) * 0.7).to(torch.bfloat16)
    mask = torch.ones(B, S, dtype=torch.uint8, device=dev)
    for b in range(B):
        a = 3 + (b * 5) % max(1, S // 3); z = min(S - 1, a + (b * 3) % max(1, S // 4))
        mask[b, a:z] = 0
        if b % 2: mask[b, S - (b % 7) - 1:] = 0
    dctx = (torch.randn(B * S, H, device=dev) * 0.5).to(torch.bfloat16)
    return qkv, mask, dctx


def ref(qkv, mask, dctx, B, S, heads):
    H = heads * 64
    x = qkv.float().view(B, S, 3, heads, 64).permute(2, 0, 3, 1, 4).clone().requires_grad_(True)
    sc = x[0] @ x[1].transpose(-1, -2) / 8.0
    sc = sc.masked_fill(mask[:, None, None, :] == 0, float("-inf"))
    out = (torch.softmax(sc, -1) @ x[2]).permute(0, 2, 1, 3).reshape(B * S, H)
    lse = torch.logsumexp(sc, -1)
    out.backward(dctx.float())
    return out.detach(), lse.detach(), x.grad.permute(1, 3, 0, 2, 4).reshape(B * S, 3 * H)


def fwd(impl, qkv, mask, B, S, heads):
    _abi.set_attn_impl(impl)
    H = heads * 64
    ctx = torch.full((B * S, H), float("nan"), device=dev, dtype=torch.bfloat16); lse = torch.full((B, heads, S), float("nan"), device=dev)
    _abi.check(lib.vault_attn_fwd(qkv.data_ptr(), mask.data_ptr(), ctx.data_ptr(), lse.data_ptr(), B, S, heads, 0.0, 0, None, 0, st()), "fwd")
    torch.cuda.synchronize()
    return ctx, lse


def bwd(impl, qkv, mask, ctx, dctx, lse, B, S, heads):
    _abi.set_attn_impl(impl)
    dqkv = torch.full_like(qkv, float("nan")); delta = torch.empty(B, heads, S, device=dev)
    _abi.check(lib.vault_attn_bwd(qkv.data_ptr(), mask.data_ptr(), ctx.data_ptr(), dctx.data_ptr(), lse.data_ptr(), delta.data_ptr(), dqkv.data_ptr(),
                                  B, S, heads, 0.0, 0, None, 0, st()), "bwd")
    torch.cuda.synchronize()
    return dqkv


def timeit(f, n=30):
    for _ in range(5): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


shapes = [(3, 185, 12), (2, 128, 2), (2, 192, 4), (3, 100, 2), (2, 129, 2), (2, 65, 2)] + ([(2, 256, 2), (1, 225, 2), (2, 209, 12)] if mode == "fwd" else [])
for B, S, heads in shapes:
    H = heads * 64
    qkv, mask, dctx = mk(B, S, heads)
    r_out, r_lse, r_d = ref(qkv, mask, dctx, B, S, heads)
    res = dict(case=f"{mode} B{B} S{S} h{heads}")
    c1, l1 = fwd(1, qkv, mask, B, S, heads)
    if mode == "fwd":
        c2, l2 = fwd(0, qkv, mask, B, S, heads)
        res.update(legacy_ctx_err=(c1.float() - r_out).abs().max().item(), tc_ctx_err=(c2.float() - r_out).abs().max().item(),
                   tc_lse_err=(l2 - r_lse).abs().max().item(), tc_vs_legacy=(c2.float() - c1.float()).abs().max().item(),
                   nan=int(torch.isnan(c2.float()).sum().item()))
    else:
        d1 = bwd(1, qkv, mask, c1, dctx, l1, B, S, heads)
        d2 = bwd(0, qkv, mask, c1, dctx, l1, B, S, heads)
        sc = r_d.abs().max().item()
        for nm, lo in (("dq", 0), ("dk", H), ("dv", 2 * H)):
            res[f"legacy_{nm}"] = (d1[:, lo:lo + H].float() - r_d[:, lo:lo + H]).abs().max().item() / sc
            res[f"tc_{nm}"] = (d2[:, lo:lo + H].float() - r_d[:, lo:lo + H]).abs().max().item() / sc
        res["nan"] = int(torch.isnan(d2.float()).sum().item())
    print(json.dumps(res), flush=True)

B, S, heads = 32, 185, 12
qkv, mask, dctx = mk(B, S, heads)
H = heads * 64
ctx = torch.empty(B * S, H, device=dev, dtype=torch.bfloat16); lse = torch.empty(B, heads, S, device=dev)
dqkv = torch.empty_like(qkv); delta = torch.empty(B, heads, S, device=dev)
f = lambda: lib.vault_attn_fwd(qkv.data_ptr(), mask.data_ptr(), ctx.data_ptr(), lse.data_ptr(), B, S, heads, 0.0, 0, None, 0, st())
g = lambda: lib.vault_attn_bwd(qkv.data_ptr(), mask.data_ptr(), ctx.data_ptr(), dctx.data_ptr(), lse.data_ptr(), delta.data_ptr(), dqkv.data_ptr(), B, S, heads, 0.0, 0, None, 0, st())
out = dict(case=f"timing B{B} S{S} h{heads} ({mode})")
for impl, nm in ((1, "legacy"), (0, "tc")):
    _abi.set_attn_impl(impl)
    f(); torch.cuda.synchronize()
    out[f"{nm}_us"] = round(timeit(f if mode == "fwd" else g), 2)
flops = (4 if mode == "fwd" else 10) * S * S * 64 * B * heads
out["tc_tflops_algorithmic"] = round(flops / out["tc_us"] / 1e6, 1)
print(json.dumps(out), flush=True)
