"""GPU parity tests of the individual sm_100a kernels, called through the C ABI, against plain torch fp32 on the same
(bf16-rounded) inputs.  Tolerances are stated per test."""
import ctypes as C
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import vault_oracle as O  # noqa: E402  (checker only)


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def ops():
    from vault_b200 import ops as o

    return o


def _rnd(dev, *shape, scale=1.0):
    return (torch.randn(*shape, device=dev) * scale).to(torch.bfloat16)


def _rel(got, ref):
    return ((got.float() - ref).abs().max() / ref.abs().max().clamp_min(1e-6)).item()


# ---------------------------------------------------------------- GEMM ----------------------------------------------------------------
@pytest.mark.parametrize("M,N,K,a_mn,b_mn,bn", [
    (128, 128, 64, False, False, 128), (128, 64, 64, False, False, 64), (128, 256, 768, False, False, 256),
    (1000, 768, 768, False, False, 0), (5920, 2304, 768, False, False, 0), (185, 128, 128, False, False, 0),
    (128, 256, 64, False, True, 256), (5920, 768, 2304, False, True, 0), (130, 3072, 768, False, True, 0),       # dgrad layout
    (128, 128, 64, True, True, 128), (256, 256, 1000, True, True, 0), (2304, 768, 5920, True, True, 128),          # wgrad layout
    (128, 128, 64, True, False, 128),
    (4096, 768, 768, False, False, 192), (4096, 768, 3072, False, True, 192), (304, 384, 200, True, True, 192), (4096, 768, 2304, False, True, 0),   # 128x192 tiles
])
def test_gemm_layouts(dev, ops, M, N, K, a_mn, b_mn, bn):
    """bit-level claim: bf16 x bf16 products accumulated in fp32 -> relative error <= 1e-4 of the fp32 torch result."""
    torch.manual_seed(M + N + K)
    a, b = _rnd(dev, M, K, scale=0.5), _rnd(dev, N, K, scale=0.5)
    ref = a.float() @ b.float().t()
    A = a.t().contiguous() if a_mn else a
    Bm = b.t().contiguous() if b_mn else b
    got = ops.gemm(A, Bm, ops.EPI_STORE_F32, a_mn=a_mn, b_mn=b_mn, block_n=bn)
    assert _rel(got, ref) < 1e-4


@pytest.mark.parametrize("n_out,k_in,tokens,split,bn", [
    (2304, 768, 11808, 2, 192), (3072, 768, 4096, 2, 256), (768, 768, 5920, 8, 256), (768, 3072, 1000, 2, 256), (136, 128, 70, 1, 128),
    (384, 384, 1280, 1, 128),   # several tiles per m block and (grid < tiles) several tiles per CTA
])
def test_gemm_wgrad_fused_bias_gradient(dev, ops, n_out, k_in, tokens, split, bn):
    """a_colsum: db = column sums of dy taken from the A tiles of the weight-gradient launch itself == torch sum over tokens (fp32 sum of
    the bf16 values), and dW is unchanged by the extra readers of the smem ring."""
    torch.manual_seed(n_out + k_in + tokens)
    dy, x = _rnd(dev, tokens, n_out, scale=0.5), _rnd(dev, tokens, k_in, scale=0.5)
    ref_w = dy.float().t() @ x.float()
    ref_b = dy.float().sum(0)
    for max_ctas in (0, 3):
        gw = torch.zeros(n_out, k_in, device=dev)
        gb = torch.zeros(n_out, device=dev)
        ops.gemm(dy, x, ops.EPI_ATOMIC_F32 if split > 1 else ops.EPI_STORE_F32, a_mn=True, b_mn=True, split_k=split, out=gw, block_n=bn,
                 a_colsum=gb, max_ctas=max_ctas)
        assert _rel(gw, ref_w) < 1e-4
        assert _rel(gb, ref_b) < 1e-5, (max_ctas, (gb - ref_b).abs().max().item())


@pytest.mark.parametrize("tokens,max_ctas", [(11808, 0), (4096, 0), (1000, 0), (200, 5), (64, 0)])
def test_gemm_wgrad_grouped(dev, ops, tokens, max_ctas):
    """The four weight gradients of a layer as ONE persistent launch over their pooled tiles == the four fp32 products, with the fused bias
    gradients of the QKV / MLP-1 problems; ragged token counts, few CTAs (many tiles per CTA across problem boundaries), a lone k-block."""
    torch.manual_seed(tokens)
    H, I = 768, 3072
    shapes = [(3 * H, H, True), (H, H, False), (I, H, True), (H, I, False)]  # (n_out, k_in, bias gradient?)
    probs, refs = [], []
    for n_out, k_in, bias in shapes:
        dy, x = _rnd(dev, tokens, n_out, scale=0.5), _rnd(dev, tokens, k_in, scale=0.5)
        dw = torch.zeros(n_out, k_in, device=dev)
        db = torch.zeros(n_out, device=dev) if bias else None
        probs.append((dy, x, dw, 2, db))
        refs.append((dy.float().t() @ x.float(), dy.float().sum(0)))
    ops.gemm_wgrad_grouped(probs, max_ctas=max_ctas)
    for (dy, x, dw, _, db), (rw, rb) in zip(probs, refs):
        assert _rel(dw, rw) < 1e-4
        if db is not None:
            assert _rel(db, rb) < 1e-5
    ops.gemm_wgrad_grouped(probs[2:3], max_ctas=max_ctas)  # a group of one; accumulates on top
    assert _rel(probs[2][2], 2 * refs[2][0]) < 1e-4


@pytest.mark.parametrize("split", [2, 4, 8])
def test_gemm_split_k_atomic(dev, ops, split):
    a, b = _rnd(dev, 768, 5920, scale=0.5), _rnd(dev, 768, 5920, scale=0.5)
    ref = a.float() @ b.float().t()
    out = torch.zeros(768, 768, device=dev)
    ops.gemm(a.t().contiguous(), b.t().contiguous(), ops.EPI_ATOMIC_F32, a_mn=True, b_mn=True, split_k=split, out=out, block_n=128)
    assert _rel(out, ref) < 1e-4


def test_gemm_epilogues(dev, ops):
    M, N, K = 777, 768, 768
    a, b = _rnd(dev, M, K, scale=0.5), _rnd(dev, N, K, scale=0.5)
    bias = torch.randn(N, device=dev)
    acc = a.float() @ b.float().t()
    tol = 1e-2  # outputs rounded to bf16 (2^-9 relative) on values up to ~max|acc|
    assert _rel(ops.gemm(a, b, ops.EPI_BIAS_BF16, bias=bias), acc + bias) < tol
    pre = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    act = ops.gemm(a, b, ops.EPI_BIAS_GELU_BF16, bias=bias, out2=pre)
    assert _rel(act, torch.nn.functional.gelu(acc + bias)) < tol and _rel(pre, acc + bias) < tol
    resid = torch.randn(M, N, device=dev)
    assert _rel(ops.gemm(a, b, ops.EPI_BIAS_RESID_F32, bias=bias, resid=resid), acc + bias + resid) < 1e-4
    assert _rel(ops.gemm(a, b, ops.EPI_PLAIN_BF16), acc) < tol
    aux = _rnd(dev, M, N)
    x = aux.float()
    gp = 0.5 * (1 + torch.erf(x / 2 ** 0.5)) + x * torch.exp(-0.5 * x * x) / (2 * math.pi) ** 0.5
    assert _rel(ops.gemm(a, b, ops.EPI_DGELU_BF16, aux=aux), acc * gp) < tol
    assert _rel(ops.gemm(a, b, ops.EPI_BIAS_F32, bias=bias), acc + bias) < 1e-4
    # training pair: the forward saves gelu'(x) next to gelu(x), the backward epilogue multiplies by it
    xr = (acc + bias).clone().requires_grad_(True)
    torch.nn.functional.gelu(xr).sum().backward()
    dact = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    act2 = ops.gemm(a, b, ops.EPI_BIAS_GELU_GRAD_BF16, bias=bias, out2=dact)
    assert _rel(act2, torch.nn.functional.gelu(acc + bias)) < tol and torch.equal(act2, act)
    assert (dact.float() - xr.grad).abs().max().item() < 1e-2  # gelu' in [-0.13, 1.13], bf16 output
    assert _rel(ops.gemm(a, b, ops.EPI_MUL_AUX_BF16, aux=aux), acc * aux.float()) < tol


def test_gemm_dropout_epilogue_statistics(dev, ops):
    M, N, K = 1024, 768, 128
    a, b = _rnd(dev, M, K), _rnd(dev, N, K)
    acc = a.float() @ b.float().t()
    z = torch.zeros(M, N, device=dev)
    y = ops.gemm(a, b, ops.EPI_BIAS_RESID_F32, resid=z, dropout_p=0.1, seed=11, site=3)
    y2 = ops.gemm(a, b, ops.EPI_BIAS_RESID_F32, resid=z, dropout_p=0.1, seed=11, site=3)
    y3 = ops.gemm(a, b, ops.EPI_BIAS_RESID_F32, resid=z, dropout_p=0.1, seed=12, site=3)
    keep = (y != 0)
    assert torch.equal(y, y2) and not torch.equal(y, y3)
    assert abs(keep.float().mean().item() - 0.9) < 5e-3
    assert torch.allclose(y[keep], acc[keep] / 0.9, rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("split", [1, 4])
def test_gemm_atomic_bias_dropout_epilogue(dev, ops, split):
    """Split-K form of the residual epilogue: out (holding the residual) += dropout(acc + bias), the same Philox mask as BIAS_RESID."""
    M, N, K = 1280, 768, 3072
    a, b = _rnd(dev, M, K, scale=0.3), _rnd(dev, N, K, scale=0.3)
    bias = torch.randn(N, device=dev)
    resid = torch.randn(M, N, device=dev)
    for p in (0.0, 0.1):
        want = ops.gemm(a, b, ops.EPI_BIAS_RESID_F32, bias=bias, resid=resid, dropout_p=p, seed=5, site=9)
        out = resid.clone()
        ops.gemm(a, b, ops.EPI_ATOMIC_BIAS_DROP_F32, bias=bias, out=out, dropout_p=p, seed=5, site=9, split_k=split, block_n=256)
        assert _rel(out, want) < 1e-5
    assert _rel(ops.gemm(a, b, ops.EPI_BIAS_RESID_F32, bias=bias, resid=resid), a.float() @ b.float().t() + bias + resid) < 1e-4


def test_gemm_rejects_bad_arguments(dev, ops):
    a, b = _rnd(dev, 128, 64), _rnd(dev, 100, 64)  # N not a multiple of 8
    with pytest.raises(RuntimeError, match="multiple of 8"):
        ops.gemm(a, b, ops.EPI_STORE_F32)
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        ops.gemm(a.cpu(), b.cpu(), ops.EPI_STORE_F32)


# ---------------------------------------------------------------- LayerNorm ----------------------------------------------------------------
@pytest.mark.parametrize("rows,cols", [(5920, 768), (37, 128), (1, 768), (1280, 1024)])
def test_layernorm_fwd_bwd(dev, ops, rows, cols):
    torch.manual_seed(rows)
    x = torch.randn(rows, cols, device=dev) * 2 + 0.5
    g = 1 + 0.1 * torch.randn(cols, device=dev)
    b = 0.1 * torch.randn(cols, device=dev)
    y16, y32, mean, rstd = ops.layernorm_fwd(x, g, b, 1e-12, want_bf16=True, want_f32=True)
    xr, gr, br = x.clone().requires_grad_(True), g.clone().requires_grad_(True), b.clone().requires_grad_(True)
    ref = torch.nn.functional.layer_norm(xr, (cols,), gr, br, 1e-12)
    assert (y32 - ref).abs().max() < 1e-4               # fp32 path
    assert (y16.float() - ref).abs().max() < 4e-2       # bf16 rounding of O(4) values
    dy16 = torch.randn(rows, cols, device=dev).to(torch.bfloat16)
    dres = torch.randn(rows, cols, device=dev)
    ref.backward(dy16.float())
    dg, db = torch.zeros(cols, device=dev), torch.zeros(cols, device=dev)
    dx32, dx16 = ops.layernorm_bwd(None, dy16, x, mean, rstd, g, dres, dg, db)
    assert (dx32 - (xr.grad + dres)).abs().max() < 1e-3
    assert _rel(dg, gr.grad) < 1e-3 and _rel(db, br.grad) < 1e-3
    assert (dx16.float() - dx32).abs().max() < 4e-2


def test_layernorm_dropout_roundtrip(dev, ops):
    """Output dropout: the backward regenerates the forward's mask (same seed/site) -- checked through the kept pattern."""
    rows, cols = 512, 768
    x = torch.randn(rows, cols, device=dev)
    g, b = torch.ones(cols, device=dev), torch.zeros(cols, device=dev)
    _, y, mean, rstd = ops.layernorm_fwd(x, g, b, 1e-5, want_bf16=False, want_f32=True, dropout_p=0.1, seed=5, site=9)
    _, y0, _, _ = ops.layernorm_fwd(x, g, b, 1e-5, want_bf16=False, want_f32=True)
    keep = y != 0
    assert abs(keep.float().mean().item() - 0.9) < 5e-3
    assert torch.allclose(y[keep], y0[keep] / 0.9, rtol=1e-5, atol=1e-5)
    dy = torch.randn(rows, cols, device=dev)
    dx, _ = ops.layernorm_bwd(dy, None, x, mean, rstd, g, None, None, None, want_bf16=False, in_p=0.1, in_site=9, seed=5)
    xr = x.clone().requires_grad_(True)
    torch.nn.functional.layer_norm(xr, (cols,), g, b, 1e-5).backward(dy * keep / 0.9)
    assert (dx - xr.grad).abs().max() < 1e-3


# ---------------------------------------------------------------- attention ----------------------------------------------------------------
def _attn_ref(qkv, mask, B, S, heads, drop_mask=None, p=0.0):
    H = heads * 64
    x = qkv.float().view(B, S, 3, heads, 64).permute(2, 0, 3, 1, 4).clone().requires_grad_(True)
    sc = x[0] @ x[1].transpose(-1, -2) / 8.0
    sc = sc.masked_fill(mask[:, None, None, :] == 0, float("-inf"))
    pr = torch.softmax(sc, -1)
    lse = torch.logsumexp(sc, -1)
    if drop_mask is not None:
        pr = pr * drop_mask / (1 - p)
    return x, (pr @ x[2]).permute(0, 2, 1, 3).reshape(B * S, H), lse


def _mid_mask(dev, B, S):
    mask = torch.ones(B, S, dtype=torch.uint8, device=dev)
    for b in range(B):
        a = 3 + (b * 5) % max(1, S // 3)
        z = min(S - 1, a + (b * 3) % max(1, S // 4))
        mask[b, a:z] = 0  # padding in the MIDDLE of the sequence (text pad before the image tokens)
        if b % 2:
            mask[b, S - (b % 7) - 1:] = 0
    return mask


@pytest.mark.parametrize("B,S,heads", [(2, 64, 2), (3, 185, 12), (2, 369, 12), (2, 40, 12), (1, 17, 2), (2, 209, 12), (1, 512, 2), (2, 65, 2),
                                       (2, 128, 4), (3, 129, 2), (2, 192, 4), (2, 193, 2), (1, 256, 2), (2, 257, 2), (1, 384, 2), (5, 300, 12),
                                       (13, 369, 12)])
def test_attention_fwd_bwd(dev, B, S, heads):
    """vault_attn_fwd / vault_attn_bwd against torch fp32 (HF:models/vilt/modeling_vilt.py:306-365 semantics: scores / 8, key mask, softmax, @V) at
    every kernel family's shapes: mma.sync (<= 64, > 384), whole-row tcgen05 (65..192), pipelined tcgen05 (193..384; 13 x 12 = 156 (sample, head)
    items = more than one item per persistent CTA)."""
    from vault_b200 import _abi

    lib, st = _abi.lib(), torch.cuda.current_stream().cuda_stream
    H = heads * 64
    torch.manual_seed(S)
    qkv = _rnd(dev, B * S, 3 * H, scale=0.7)
    mask = _mid_mask(dev, B, S)
    ctx = torch.empty(B * S, H, device=dev, dtype=torch.bfloat16)
    lse = torch.empty(B, heads, S, device=dev)
    _abi.check(lib.vault_attn_fwd(qkv.data_ptr(), mask.data_ptr(), ctx.data_ptr(), lse.data_ptr(), B, S, heads, 0.0, 0, None, 0, st))
    x, ref, ref_lse = _attn_ref(qkv, mask, B, S, heads)
    assert (ctx.float() - ref).abs().max() < 2e-2  # bf16 P and bf16 output on O(1) values
    assert (lse - ref_lse).abs().max() < 1e-3
    dctx = _rnd(dev, B * S, H, scale=0.5)
    ref.backward(dctx.float())
    dref = x.grad.permute(1, 3, 0, 2, 4).reshape(B * S, 3 * H)
    dqkv = torch.empty_like(qkv)
    delta = torch.empty(B, heads, S, device=dev)
    _abi.check(lib.vault_attn_bwd(qkv.data_ptr(), mask.data_ptr(), ctx.data_ptr(), dctx.data_ptr(), lse.data_ptr(), delta.data_ptr(), dqkv.data_ptr(), B, S,
                                  heads, 0.0, 0, None, 0, st))
    assert _rel(dqkv, dref) < 2e-2


@pytest.mark.parametrize("B,S,heads,tc_impl", [(3, 185, 12, 0), (2, 100, 2, 0), (2, 192, 2, 0), (1, 241, 2, 0), (2, 369, 12, 0),
                                               (1, 17, 2, 3), (2, 40, 12, 3), (2, 128, 2, 3), (3, 185, 12, 3), (2, 256, 2, 3), (25, 128, 12, 3),
                                               (2, 369, 12, 4), (27, 128, 12, 4), (13, 300, 12, 4), (3, 185, 2, 4)])
def test_attention_tcgen05_matches_mma_sync(dev, B, S, heads, tc_impl):
    """The tensor-memory kernels (attention_tc.cu, attention_sm100.cu; tc_impl 3 forces the pipelined kernels at every length, 4 additionally
    its one-tile-per-warpgroup forward; the multi-item cases put several (sample, head) items on one persistent CTA) and the
    mma.sync kernels implement the same contract: same LSE convention (either forward feeds either backward), outputs equal to bf16
    rounding, every output row written, nothing written past a sample's rows."""
    from vault_b200 import _abi

    lib, st = _abi.lib(), torch.cuda.current_stream().cuda_stream
    H = heads * 64
    torch.manual_seed(S + 1)
    qkv = _rnd(dev, B * S, 3 * H, scale=0.7)
    mask = _mid_mask(dev, B, S)
    dctx = _rnd(dev, B * S, H, scale=0.5)
    out = {}
    try:
        for impl in (1, tc_impl):
            _abi.set_attn_impl(impl)
            ctx = torch.full((B * S + 64, H), float("nan"), device=dev, dtype=torch.bfloat16)
            ctx[B * S:] = 7.0  # 64 guard rows behind the last sample
            lse = torch.full((B, heads, S), float("nan"), device=dev)
            _abi.check(lib.vault_attn_fwd(qkv.data_ptr(), mask.data_ptr(), ctx.data_ptr(), lse.data_ptr(), B, S, heads, 0.0, 0, None, 0, st))
            assert torch.all(ctx[B * S:] == 7.0) and not torch.isnan(ctx.float()).any() and not torch.isnan(lse).any()
            out[impl] = (ctx[:B * S].clone(), lse)
        (c1, l1), (c0, l0) = out[1], out[tc_impl]
        assert (c0.float() - c1.float()).abs().max() <= 2 ** -8 * max(1.0, c1.float().abs().max().item())
        assert (l0 - l1).abs().max() < 1e-5
        if True:
            grads = {}
            for impl in (1, tc_impl):
                _abi.set_attn_impl(impl)
                dqkv = torch.full((B * S + 64, 3 * H), float("nan"), device=dev, dtype=torch.bfloat16)
                dqkv[B * S:] = 7.0
                delta = torch.empty(B, heads, S, device=dev)
                # forward of the OTHER family feeds this backward
                cf, lf = out[tc_impl if impl == 1 else 1]
                _abi.check(lib.vault_attn_bwd(qkv.data_ptr(), mask.data_ptr(), cf.data_ptr(), dctx.data_ptr(), lf.data_ptr(), delta.data_ptr(),
                                              dqkv.data_ptr(), B, S, heads, 0.0, 0, None, 0, st))
                assert torch.all(dqkv[B * S:] == 7.0) and not torch.isnan(dqkv.float()).any()
                grads[impl] = dqkv[:B * S].float()
            assert _rel(grads[tc_impl], grads[1]) < 1e-2
    finally:
        _abi.set_attn_impl(0)


@pytest.mark.parametrize("S", [40, 64, 100, 128, 185, 257, 369])
def test_attention_dropout_exact_mask(dev, S):
    """With V = identity over one 64-key chunk the context IS that chunk of the dropped probability matrix: recover the Philox mask chunk
    by chunk (the mask depends only on seed / site / (b, h, q, k), never on the data), then check forward and backward against torch
    with that exact mask -- several key chunks, ragged tails and mid-sequence padding included."""
    from vault_b200 import _abi

    lib, st = _abi.lib(), torch.cuda.current_stream().cuda_stream
    B, heads, p, seed, site = 2, 2, 0.1, 99, 4
    H = heads * 64
    torch.manual_seed(7)
    qkv = _rnd(dev, B * S, 3 * H, scale=0.7)
    mask = _mid_mask(dev, B, S) if S > 64 else torch.ones(B, S, dtype=torch.uint8, device=dev)
    ones = torch.ones(B, S, dtype=torch.uint8, device=dev)
    ctx = torch.empty(B * S, H, device=dev, dtype=torch.bfloat16)
    lse = torch.empty(B, heads, S, device=dev)
    drop_mask = torch.zeros(B, heads, S, S, device=dev)
    for c0 in range(0, S, 64):
        probe = qkv.clone().view(B, S, 3, heads, 64)
        probe[:, :, 2] = 0
        for k in range(c0, min(S, c0 + 64)):
            probe[:, k, 2, :, k - c0] = 1.0
        probe = probe.view(B * S, 3 * H).contiguous()
        _abi.check(lib.vault_attn_fwd(probe.data_ptr(), ones.data_ptr(), ctx.data_ptr(), lse.data_ptr(), B, S, heads, p, seed, None, site, st))
        n = min(S, c0 + 64) - c0
        pd = ctx.float().view(B, S, heads, 64)[..., :n].permute(0, 2, 1, 3)  # [B,h,q,k in chunk] dropped probabilities
        drop_mask[..., c0:c0 + n] = (pd != 0).float()
    assert abs(drop_mask.mean().item() - 0.9) < 0.02
    _abi.check(lib.vault_attn_fwd(qkv.data_ptr(), mask.data_ptr(), ctx.data_ptr(), lse.data_ptr(), B, S, heads, p, seed, None, site, st))
    x, ref, _ = _attn_ref(qkv, mask, B, S, heads, drop_mask, p)
    assert (ctx.float() - ref).abs().max() < 2e-2
    dctx = _rnd(dev, B * S, H, scale=0.5)
    ref.backward(dctx.float())
    dref = x.grad.permute(1, 3, 0, 2, 4).reshape(B * S, 3 * H)
    dqkv = torch.empty_like(qkv)
    delta = torch.empty(B, heads, S, device=dev)
    _abi.check(lib.vault_attn_bwd(qkv.data_ptr(), mask.data_ptr(), ctx.data_ptr(), dctx.data_ptr(), lse.data_ptr(), delta.data_ptr(), dqkv.data_ptr(), B, S,
                                  heads, p, seed, None, site, st))
    assert _rel(dqkv, dref) < 2e-2


# ---------------------------------------------------------------- optimizer / loss / reductions ----------------------------------------------------------------
@pytest.mark.parametrize("correct_bias,wd", [(False, 0.0), (True, 0.01)])
def test_adamw_matches_hf_rule(dev, correct_bias, wd):
    """fp32 elementwise rule: bit-comparable up to fused-multiply-add rounding -> 1e-6 relative."""
    from vault_b200 import _abi

    n = 100003
    torch.manual_seed(1)
    p = torch.randn(n); g = torch.randn(n) * 1e-2
    m, v = torch.zeros(n), torch.zeros(n)
    pc, mc, vc = p.to(dev), m.to(dev), v.to(dev)
    n_al = (n + 63) // 64 * 64
    pd, md, vd, gd = (torch.zeros(n_al, device=dev) for _ in range(4))
    pd[:n], gd[:n] = pc, g.to(dev)
    sh = torch.zeros(n_al, device=dev, dtype=torch.bfloat16)
    for step in (1, 2, 3):
        O.hf_adamw_step(p, g, m, v, step, lr=1e-3, weight_decay=wd, correct_bias=correct_bias)
        _abi.call("vault_adamw_step", pd.data_ptr(), gd.data_ptr(), 0, md.data_ptr(), vd.data_ptr(), sh.data_ptr(), n, 1e-3, 0.9, 0.999, 1e-8, wd,
                  int(correct_bias), step, 1.0, None, torch.cuda.current_stream().cuda_stream)
    assert torch.allclose(pd[:n].cpu(), p, rtol=2e-6, atol=1e-7)
    assert torch.allclose(md[:n].cpu(), m, rtol=2e-6, atol=1e-9) and torch.allclose(vd[:n].cpu(), v, rtol=2e-6, atol=1e-12)
    assert torch.equal(sh[:n].cpu(), pd[:n].cpu().to(torch.bfloat16))


@pytest.mark.parametrize("B,T,H,n_types", [(32, 128, 768, 2), (3, 17, 128, 1), (8, 40, 768, 4), (2, 9, 1152, 2)])
def test_embedding_backward_scatter_adds(dev, B, T, H, n_types):
    """lm_embed_bwd / vilt_text_embed_bwd: word / token-type / position table gradients == torch index_add_ of the row gradients
    (token types 0 and 1 are summed per CTA in registers + shared memory, other types and rows wider than 1,024 take direct atomics)."""
    from vault_b200 import _abi
    torch.manual_seed(B * T + H)
    V = 97
    ids = torch.randint(1, V, (B, T), device=dev)
    ids[:, -2:] = 0  # padding id: no word gradient
    tt = torch.randint(0, n_types, (B, T), device=dev)
    dx = torch.randn(B * T, H, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    dword, dtype, dpos = torch.zeros(V, H, device=dev), torch.zeros(n_types, H, device=dev), torch.zeros(T, H, device=dev)
    _abi.call("vault_lm_embed_bwd", ids.data_ptr(), tt.data_ptr(), dx.data_ptr(), dword.data_ptr(), dtype.data_ptr(), dpos.data_ptr(), B, T, H, -1, 0, st)
    rw = torch.zeros(V, H, device=dev).index_add_(0, ids.flatten(), dx)
    rw[0] = 0
    rt = torch.zeros(n_types, H, device=dev).index_add_(0, tt.flatten(), dx)
    rp = torch.zeros(T, H, device=dev).index_add_(0, torch.arange(T, device=dev).repeat(B), dx)
    for got, ref in ((dword, rw), (dtype, rt), (dpos, rp)):
        assert (got - ref).abs().max().item() <= 1e-4 * max(1.0, ref.abs().max().item())
    dtype2, dpos2 = torch.zeros(n_types, H, device=dev), torch.zeros(T, H, device=dev)
    _abi.call("vault_vilt_text_embed_bwd", tt.data_ptr(), dx.data_ptr(), dtype2.data_ptr(), dpos2.data_ptr(), B, T, H, st)
    for got, ref in ((dtype2, rt), (dpos2, rp)):
        assert (got - ref).abs().max().item() <= 1e-4 * max(1.0, ref.abs().max().item())
    dtype3 = torch.zeros(n_types, H, device=dev)
    _abi.call("vault_vilt_text_embed_bwd", None, dx.data_ptr(), dtype3.data_ptr(), None, B, T, H, st)  # no token types: everything is type 0
    assert (dtype3[0] - dx.sum(0)).abs().max().item() <= 1e-4 * max(1.0, dx.sum(0).abs().max().item())


def test_ce_loss_and_colsum(dev):
    from vault_b200 import _abi

    st = torch.cuda.current_stream().cuda_stream
    logits = torch.randn(32, 3, device=dev)
    labels = torch.randint(0, 3, (32,), device=dev)
    loss = torch.zeros(1, device=dev)
    dl = torch.empty_like(logits)
    _abi.call("vault_ce_loss", logits.data_ptr(), labels.data_ptr(), loss.data_ptr(), dl.data_ptr(), 32, 3, 1.0, st)
    lr = logits.clone().requires_grad_(True)
    ref = torch.nn.functional.cross_entropy(lr, labels)
    ref.backward()
    assert abs(loss.item() - ref.item()) < 1e-6 and (dl - lr.grad).abs().max() < 1e-6
    x = _rnd(dev, 5920, 2304)
    out = torch.zeros(2304, device=dev)
    _abi.call("vault_colsum_bf16", x.data_ptr(), 2304, out.data_ptr(), 5920, 2304, st)
    assert _rel(out, x.float().sum(0)) < 1e-4


def test_head_loss_bce_and_two_group_ce(dev):
    """The Bloomberg (BCE-with-logits, one logit) and raw-MVSA (two label groups) losses against the oracle's restatement."""
    from oracle import vault_oracle as O
    from vault_b200 import _abi

    st = torch.cuda.current_stream().cuda_stream
    torch.manual_seed(3)
    for kind, logits, labels, ref_fn in (
            (1, torch.randn(37, device=dev) * 3, (torch.rand(37, device=dev) > 0.5).float(), O.bce_loss),
            (2, torch.randn(37, 6, device=dev), torch.randint(0, 3, (37, 2), device=dev), O.mvsa_raw_loss),
            (0, torch.randn(37, 3, device=dev), torch.randint(0, 3, (37,), device=dev), O.ce_loss)):
        n = 1 if logits.dim() == 1 else logits.shape[1]
        loss = torch.zeros(1, device=dev)
        dl = torch.empty_like(logits)
        _abi.call("vault_head_loss", logits.data_ptr(), labels.data_ptr(), loss.data_ptr(), dl.data_ptr(), 37, n, kind, 1.0, st)
        lr = logits.clone().requires_grad_(True)
        ref = ref_fn(lr, labels)
        ref.backward()
        assert abs(loss.item() - ref.item()) < 2e-6 and (dl - lr.grad).abs().max() < 1e-6


# ---------------------------------------------------------------- im2col-free patch embedding ----------------------------------------------------------------
@pytest.mark.parametrize("B,Hi,Wi,N", [(2, 384, 384, 768), (5, 384, 640, 768), (3, 640, 384, 128), (33, 384, 384, 256), (1, 32, 32, 128)])
def test_patch_embed_tma_tf32(dev, B, Hi, Wi, N):
    """TF32 operands (10-bit mantissa), fp32 accumulate over K=3072: relative error <= 2e-3 of conv2d in fp32."""
    from vault_b200 import _abi

    torch.manual_seed(B)
    px = torch.randn(B, 3, Hi, Wi, device=dev)
    w = torch.randn(N, 3, 32, 32, device=dev) * 0.02
    b = torch.randn(N, device=dev)
    out = torch.full((B * (Hi // 32) * (Wi // 32), N), float("nan"), device=dev)
    _abi.call("vault_patch_embed_fwd", px.data_ptr(), w.data_ptr(), b.data_ptr(), out.data_ptr(), B, 3, Hi, Wi, 32, N, torch.cuda.current_stream().cuda_stream)
    ref = torch.nn.functional.conv2d(px, w, b, stride=32).flatten(2).transpose(1, 2).reshape(-1, N)
    assert torch.isfinite(out).all()  # every output row written exactly once
    assert _rel(out, ref) < 2e-3


@pytest.mark.parametrize("B,Hi,Wi,N,T", [(2, 384, 384, 768, 8), (5, 384, 640, 768, 40), (3, 640, 384, 128, 5), (33, 384, 384, 256, 16), (32, 384, 640, 768, 128),
                                         (2, 512, 512, 128, 3), (3, 416, 608, 128, 7), (2, 32, 32, 128, 2)])
def test_patch_embed_wgrad_tma_tf32(dev, B, Hi, Wi, N, T):
    """im2col-free weight gradient of the patch projection (vault_patch_grad_rows_f32 + vault_patch_embed_wgrad) against torch's conv2d
    backward (HF:models/vilt/modeling_vilt.py:293-303): dW, and the bias gradient, with mixed per-sample valid patch rectangles (rows of
    invalid patches are zero); TF32 operands -> 2e-3 relative.  The result is ACCUMULATED into dW (a second call doubles it)."""
    from vault_b200 import _abi

    lib, st = _abi.lib(), torch.cuda.current_stream().cuda_stream
    gh, gw = Hi // 32, Wi // 32
    assert lib.vault_patch_embed_wgrad_ok(3, Hi, Wi, 32, N) == 1
    torch.manual_seed(B + Hi)
    px = torch.randn(B, 3, Hi, Wi, device=dev)
    hw = torch.tensor([[gh, gw]] + [[max(1, gh - (b * 3) % gh), max(1, gw - (b * 5) % gw)] for b in range(1, B)], dtype=torch.int32, device=dev)
    pmax = int((hw[:, 0] * hw[:, 1]).max())
    S = T + 1 + pmax
    dX = torch.randn(B, S, N, device=dev)
    # reference: dpatch rows in the full-grid layout, zero outside each sample's rectangle
    dpatch_ref = torch.zeros(B, gh, gw, N, device=dev)
    for b in range(B):
        h, w = int(hw[b, 0]), int(hw[b, 1])
        dpatch_ref[b, :h, :w] = dX[b, T + 1:T + 1 + h * w].view(h, w, N)
    wt = torch.zeros(N, 3, 32, 32, device=dev, requires_grad=True)
    bias = torch.zeros(N, device=dev, requires_grad=True)
    y = torch.nn.functional.conv2d(px, wt, bias, stride=32)  # [B, N, gh, gw]
    y.backward(dpatch_ref.permute(0, 3, 1, 2).contiguous())
    dp32 = torch.full((B * gh * gw, N), float("nan"), device=dev)
    db = torch.zeros(N, device=dev)
    dW = torch.zeros(N, 3 * 32 * 32, device=dev)
    for rep in (1, 2):
        _abi.check(lib.vault_patch_grad_rows_f32(dX.data_ptr(), hw.data_ptr(), dp32.data_ptr(), db.data_ptr(), B, T, pmax, gh, gw, N, st))
        _abi.check(lib.vault_patch_embed_wgrad(px.data_ptr(), dp32.data_ptr(), dW.data_ptr(), B, 3, Hi, Wi, 32, N, st))
        assert torch.equal(dp32.view(B, gh, gw, N), dpatch_ref)
        assert _rel(dW, rep * wt.grad.view(N, -1)) < 2e-3
        assert _rel(db, rep * bias.grad) < 1e-5


def test_patch_embed_wgrad_rejects_grids_without_a_k_block(dev):
    from vault_b200 import _abi

    lib = _abi.lib()
    assert lib.vault_patch_embed_wgrad_ok(3, 384, 416, 32, 768) == 1    # 13 patches per row: k-blocks of 4 rows = 52 patches, zero-padded to 56
    assert lib.vault_patch_embed_wgrad_ok(3, 32, 32, 32, 768) == 1      # one patch per image: k-block of 1, padded to 8
    assert lib.vault_patch_embed_wgrad_ok(3, 64, 32 * 65, 32, 768) == 0  # a patch row of 65 patches does not fit a k-block
    assert lib.vault_patch_embed_wgrad_ok(3, 384, 384, 16, 768) == 0    # patch size
    assert lib.vault_patch_embed_wgrad_ok(3, 384, 384, 32, 100) == 0    # N % 128


@pytest.mark.parametrize("cl,M,N,K,a_mn,b_mn,bn", [
    (1, 5920, 2304, 768, False, False, 256), (1, 5920, 768, 2304, False, True, 256), (1, 300, 256, 192, False, False, 128),
    (2, 1280, 768, 768, False, False, 64), (2, 2304, 768, 1280, True, True, 128), (2, 128, 192, 64, False, False, 64), (1, 128, 256, 64, True, True, 128),
])
def test_gemm_cta_pair_multicast(dev, ops, cl, M, N, K, a_mn, b_mn, bn):
    """CTA-pair TMA multicast (incl. odd tile counts -> phantom tiles) must be bit-identical to the unclustered kernel."""
    torch.manual_seed(cl * 1000 + M)
    a, b = _rnd(dev, M, K, scale=0.5), _rnd(dev, N, K, scale=0.5)
    A = a.t().contiguous() if a_mn else a
    Bm = b.t().contiguous() if b_mn else b
    ref = ops.gemm(A, Bm, ops.EPI_STORE_F32, a_mn=a_mn, b_mn=b_mn, block_n=bn)
    got = ops.gemm(A, Bm, ops.EPI_STORE_F32, a_mn=a_mn, b_mn=b_mn, block_n=bn, cluster=cl)
    assert torch.equal(got, ref)
    assert _rel(got, a.float() @ b.float().t()) < 1e-4


@pytest.mark.parametrize("M,N,K,a_mn,b_mn,bn,split", [(5920, 2304, 768, False, False, 0, 1), (130, 128, 64, False, False, 0, 1), (768, 768, 5920, True, True, 128, 4),
                                                    (1280, 3072, 768, False, True, 0, 1)])
def test_gemm_dynamic_tile_scheduler(dev, ops, M, N, K, a_mn, b_mn, bn, split):
    """Dynamic (atomic-counter) tile claiming gives the same result as the static round-robin, and re-arms its counters."""
    torch.manual_seed(M)
    a, b = _rnd(dev, M, K, scale=0.5), _rnd(dev, N, K, scale=0.5)
    A = a.t().contiguous() if a_mn else a
    Bm = b.t().contiguous() if b_mn else b
    epi = ops.EPI_ATOMIC_F32 if split > 1 else ops.EPI_STORE_F32
    sched = torch.zeros(2, device=dev, dtype=torch.int32)
    ref = a.float() @ b.float().t()
    for _ in range(3):  # repeated launches reuse the same counters
        out = torch.zeros(M, N, device=dev)
        ops.gemm(A, Bm, epi, a_mn=a_mn, b_mn=b_mn, block_n=bn, split_k=split, out=out, sched=sched)
        assert _rel(out, ref) < 1e-4
        assert sched.tolist() == [0, 0]
