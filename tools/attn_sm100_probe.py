"""GPU probe: pipelined tcgen05 attention (attention_sm100.cu, impl 0) vs mma.sync (impl 1) / first-generation tcgen05 (impl 2) vs torch fp32,
correctness first (small + ragged shapes), then timing at the bench shapes.  usage: attn_sm100_probe.py [check|time|all]"""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vault_b200 import _abi

what = sys.argv[1] if len(sys.argv) > 1 else "all"
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


def fwd(impl, qkv, mask, B, S, heads, sync=True):
    _abi.set_attn_impl(impl)
    H = heads * 64
    ctx = torch.full((B * S, H), float("nan"), device=dev, dtype=torch.bfloat16); lse = torch.full((B, heads, S), float("nan"), device=dev)
    _abi.check(lib.vault_attn_fwd(qkv.data_ptr(), mask.data_ptr(), ctx.data_ptr(), lse.data_ptr(), B, S, heads, 0.0, 0, None, 0, st()), "fwd")
    if sync: torch.cuda.synchronize()
    return ctx, lse


def bwd(impl, qkv, mask, ctx, dctx, lse, B, S, heads, sync=True):
    _abi.set_attn_impl(impl)
    dqkv = torch.full_like(qkv, float("nan")); delta = torch.empty(B, heads, S, device=dev)
    _abi.check(lib.vault_attn_bwd(qkv.data_ptr(), mask.data_ptr(), ctx.data_ptr(), dctx.data_ptr(), lse.data_ptr(), delta.data_ptr(), dqkv.data_ptr(),
                                  B, S, heads, 0.0, 0, None, 0, st()), "bwd")
    if sync: torch.cuda.synchronize()
    return dqkv


def rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-12)).item()


def check():
    ok = True
    for (B, S, heads) in [(1, 17, 2), (2, 40, 2), (2, 64, 2), (2, 128, 2), (3, 129, 2), (3, 185, 12), (2, 256, 2), (2, 257, 2), (2, 369, 12), (1, 384, 2),
                          (40, 369, 12)]:
        qkv, mask, dctx = mk(B, S, heads)
        r_out, r_lse, r_d = ref(qkv, mask, dctx, B, S, heads)
        try:
            ctx, lse = fwd(0, qkv, mask, B, S, heads)
            e_f = (ctx.float() - r_out).abs().max().item(); e_l = (lse - r_lse).abs().max().item()
            dq = bwd(0, qkv, mask, ctx, dctx, lse, B, S, heads)
            H = heads * 64
            parts = [rel(dq[:, i * H:(i + 1) * H], r_d[:, i * H:(i + 1) * H]) for i in range(3)]
            nan = bool(torch.isnan(ctx.float()).any() or torch.isnan(dq.float()).any() or torch.isnan(lse).any())
            good = e_f < 2e-2 and e_l < 1e-3 and max(parts) < 2e-2 and not nan
            ok &= good
            print(json.dumps(dict(B=B, S=S, heads=heads, fwd_maxabs=e_f, lse_maxabs=e_l, rel_dq=parts[0], rel_dk=parts[1], rel_dv=parts[2], nan=nan, ok=good)), flush=True)
        except Exception as e:  # noqa
            ok = False
            print(json.dumps(dict(B=B, S=S, heads=heads, error=str(e)[:300])), flush=True)
            break
    print("CHECK", "PASS" if ok else "FAIL", flush=True)
    return ok


def timeit(f, n=20):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def time_all():
    for (B, S, heads) in [(32, 369, 12), (32, 185, 12), (32, 128, 12), (32, 40, 12), (64, 209, 12)]:
        qkv, mask, dctx = mk(B, S, heads)
        fl_f = 4.0 * S * S * 64 * B * heads
        for impl in (0, 1, 2):
            try:
                ctx, lse = fwd(impl, qkv, mask, B, S, heads)
                dq = torch.empty_like(qkv); delta = torch.empty(B, heads, S, device=dev); c2 = torch.empty_like(ctx); l2 = torch.empty_like(lse)
                s_ = st()
                tf = timeit(lambda: lib.vault_attn_fwd(qkv.data_ptr(), mask.data_ptr(), c2.data_ptr(), l2.data_ptr(), B, S, heads, 0.0, 0, None, 0, s_))
                tb = timeit(lambda: lib.vault_attn_bwd(qkv.data_ptr(), mask.data_ptr(), ctx.data_ptr(), dctx.data_ptr(), lse.data_ptr(), delta.data_ptr(),
                                                        dq.data_ptr(), B, S, heads, 0.0, 0, None, 0, s_))
                print(json.dumps(dict(B=B, S=S, impl=impl, fwd_us=tf, bwd_us=tb, fwd_tflops=fl_f / tf / 1e6, bwd_tflops=2.5 * fl_f / tb / 1e6)), flush=True)
            except Exception as e:  # noqa
                print(json.dumps(dict(B=B, S=S, impl=impl, error=str(e)[:200])), flush=True)
    _abi.set_attn_impl(0)


if what in ("check", "all"):
    good = check()
    if not good and what == "all":
        sys.exit(1)
if what in ("time", "all"):
    time_all()
