"""Small invocations of every mbarrier / TMEM / TMA kernel family, for compute-sanitizer (racecheck, synccheck, memcheck):
    compute-sanitizer --tool racecheck python tools/sanitize_cases.py [gemm|attn|attn_tc|patch|ln|all]
Shapes are tiny (sanitizer slow-down ~100x) but hit every code path: multi-tile persistent loops, split-K, ragged tails, multi-item CTAs."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vault_b200 import _abi, ops

what = sys.argv[1] if len(sys.argv) > 1 else "all"
dev = torch.device("cuda:0"); lib = _abi.lib(); st = torch.cuda.current_stream().cuda_stream
torch.manual_seed(0)
rnd = lambda *s: (torch.randn(*s, device=dev) * 0.5).to(torch.bfloat16)

if what in ("gemm", "all"):
    for (M, N, K, a_mn, b_mn, bn, split, epi) in [(300, 256, 192, False, False, 0, 1, ops.EPI_STORE_F32), (520, 768, 256, False, True, 192, 1, ops.EPI_PLAIN_BF16),
                                                   (256, 256, 1000, True, True, 256, 4, ops.EPI_ATOMIC_F32), (260, 128, 64, False, False, 64, 1, ops.EPI_BIAS_GELU_BF16)]:
        a, b = rnd(M, K), rnd(N, K)
        A = a.t().contiguous() if a_mn else a
        Bm = b.t().contiguous() if b_mn else b
        kw = dict(bias=torch.zeros(N, device=dev)) if epi == ops.EPI_BIAS_GELU_BF16 else {}
        out = torch.zeros(M, N, device=dev) if epi == ops.EPI_ATOMIC_F32 else None
        ops.gemm(A, Bm, epi, a_mn=a_mn, b_mn=b_mn, block_n=bn, split_k=split, out=out, **kw)
    torch.cuda.synchronize(); print("gemm ok")

def attn(B, S, heads, impl, p=0.0):
    H = heads * 64
    _abi.set_attn_impl(impl)
    qkv = rnd(B * S, 3 * H); mask = torch.ones(B, S, dtype=torch.uint8, device=dev); mask[:, 3:5] = 0
    ctx = torch.empty(B * S, H, device=dev, dtype=torch.bfloat16); lse = torch.empty(B, heads, S, device=dev)
    _abi.check(lib.vault_attn_fwd(qkv.data_ptr(), mask.data_ptr(), ctx.data_ptr(), lse.data_ptr(), B, S, heads, p, 1, None, 3, st), "fwd")
    dctx = rnd(B * S, H); dq = torch.empty_like(qkv); delta = torch.empty(B, heads, S, device=dev)
    _abi.check(lib.vault_attn_bwd(qkv.data_ptr(), mask.data_ptr(), ctx.data_ptr(), dctx.data_ptr(), lse.data_ptr(), delta.data_ptr(), dq.data_ptr(), B, S, heads, p, 1, None, 3, st), "bwd")
    torch.cuda.synchronize()

if what in ("attn", "all"):  # pipelined tcgen05 kernels (attention_sm100.cu): 1, 2, 3 key blocks; 150 items -> two items on some CTAs
    for (B, S, heads) in [(1, 40, 2), (1, 200, 2), (2, 369, 2), (75, 130, 2)]:
        attn(B, S, heads, 3)
    print("attn_sm100 ok")
if what in ("attn_tc", "all"):  # whole-row tcgen05 kernels (attention_tc.cu) and the mma.sync kernels with dropout
    attn(2, 185, 2, 2); attn(2, 100, 2, 2); attn(2, 72, 2, 1, p=0.1); attn(1, 400, 2, 1)
    _abi.set_attn_impl(0)
    print("attn_tc / mma.sync ok")
if what in ("patch", "all"):
    B, Hi, Wi, N = 2, 96, 160, 128
    px = torch.randn(B, 3, Hi, Wi, device=dev); w = torch.randn(N, 3 * 32 * 32, device=dev) * 0.02; bias = torch.zeros(N, device=dev)
    out = torch.empty(B * 15, N, device=dev)
    _abi.call("vault_patch_embed_fwd", px.data_ptr(), w.data_ptr(), bias.data_ptr(), out.data_ptr(), B, 3, Hi, Wi, 32, N, st)
    hw = torch.tensor([[3, 5], [2, 4]], dtype=torch.int32, device=dev); dX = torch.randn(B, 4 + 1 + 15, N, device=dev)
    dp = torch.empty(B * 15, N, device=dev); db = torch.zeros(N, device=dev); dW = torch.zeros(N, 3072, device=dev)
    _abi.call("vault_patch_grad_rows_f32", dX.data_ptr(), hw.data_ptr(), dp.data_ptr(), db.data_ptr(), B, 4, 15, 3, 5, N, st)
    _abi.call("vault_patch_embed_wgrad", px.data_ptr(), dp.data_ptr(), dW.data_ptr(), B, 3, Hi, Wi, 32, N, st)
    torch.cuda.synchronize(); print("patch ok")
if what in ("ln", "all"):
    x = torch.randn(300, 768, device=dev); g = torch.ones(768, device=dev); b = torch.zeros(768, device=dev)
    y16, _, mean, rstd = ops.layernorm_fwd(x, g, b, 1e-12)
    dg, dbb = torch.zeros(768, device=dev), torch.zeros(768, device=dev)
    ops.layernorm_bwd(None, y16, x, mean, rstd, g, torch.randn(300, 768, device=dev), dg, dbb)
    torch.cuda.synchronize(); print("ln ok")
