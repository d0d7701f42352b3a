"""GPU bring-up probe for the tcgen05 GEMM: simplest case first, prints one line per case, keeps going after numeric
mismatches (a device trap ends the process -- run groups in separate processes).

    python tools/gemm_probe.py [group ...]      groups: kk kmn mnmn epi time
"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vault_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")


def rnd(*shape, scale=1.0):
    return (torch.randn(*shape, device=dev) * scale).to(torch.bfloat16)


def report(name, got, ref, tol=2e-2):
    got = got.float()
    err = (got - ref).abs()
    denom = ref.abs().max().clamp_min(1e-6)
    rel = (err.max() / denom).item()
    bad = (err > tol * denom)
    line = dict(case=name, rel=rel, bad_frac=bad.float().mean().item(), ok=bool(rel < tol))
    if not line["ok"]:
        rows = bad.any(dim=1).nonzero().flatten()
        cols = bad.any(dim=0).nonzero().flatten()
        line["bad_rows"] = [int(rows.min()), int(rows.max()), int(rows.numel())] if rows.numel() else []
        line["bad_cols"] = [int(cols.min()), int(cols.max()), int(cols.numel())] if cols.numel() else []
        line["got00"] = got[:2, :4].tolist()
        line["ref00"] = ref[:2, :4].tolist()
    print(json.dumps(line), flush=True)
    return line["ok"]


def case(name, M, N, K, a_mn=False, b_mn=False, bn=0, epi=ops.EPI_STORE_F32, split_k=1):
    a = rnd(M, K, scale=0.5)
    b = rnd(N, K, scale=0.5)
    ref = a.float() @ b.float().t()
    A = a.t().contiguous() if a_mn else a
    B = b.t().contiguous() if b_mn else b
    out = None
    if epi == ops.EPI_ATOMIC_F32:
        out = torch.zeros(M, N, device=dev)
    got = ops.gemm(A, B, epi, a_mn=a_mn, b_mn=b_mn, block_n=bn, split_k=split_k, out=out)
    torch.cuda.synchronize()
    return report(name, got, ref)


def group_kk():
    for bn in (128, 64, 256):
        case(f"kk 128x{bn}x64 bn{bn}", 128, bn, 64, bn=bn)
    for bn in (128, 64, 256):
        case(f"kk 128x{bn}x768 bn{bn}", 128, bn, 768, bn=bn)
    case("kk 256x256x768 bn128", 256, 256, 768, bn=128)
    case("kk 1000x768x768 auto", 1000, 768, 768)
    case("kk 5920x2304x768 auto", 5920, 2304, 768)
    case("kk 5920x768x3072 auto", 5920, 768, 3072)
    case("kk 1280x3072x768 auto", 1280, 3072, 768)
    case("kk 185x128x128 auto (tiny)", 185, 128, 128)


def group_kmn():
    for bn in (128, 64, 256):
        case(f"k,mn 128x{bn}x64 bn{bn}", 128, bn, 64, b_mn=True, bn=bn)
    case("k,mn 128x256x768", 128, 256, 768, b_mn=True)
    case("k,mn dgrad 5920x768x2304", 5920, 768, 2304, b_mn=True)
    case("k,mn dgrad 5920x3072x768", 5920, 3072, 768, b_mn=True)


def group_mnmn():
    for bn in (128, 64, 256):
        case(f"mn,mn 128x{bn}x64 bn{bn}", 128, bn, 64, a_mn=True, b_mn=True, bn=bn)
    case("mn,k 128x128x64", 128, 128, 64, a_mn=True, b_mn=False, bn=128)
    case("mn,mn 256x256x1000", 256, 256, 1000, a_mn=True, b_mn=True)
    case("mn,mn wgrad 2304x768x5920 store", 2304, 768, 5920, a_mn=True, b_mn=True)
    case("mn,mn wgrad 768x768x5920 splitk4", 768, 768, 5920, a_mn=True, b_mn=True, epi=ops.EPI_ATOMIC_F32, split_k=4)
    case("mn,mn wgrad 768x3072x5920 splitk2", 768, 3072, 5920, a_mn=True, b_mn=True, epi=ops.EPI_ATOMIC_F32, split_k=2)


def group_epi():
    M, N, K = 777, 768, 768
    a, b = rnd(M, K, scale=0.5), rnd(N, K, scale=0.5)
    bias = torch.randn(N, device=dev)
    acc = a.float() @ b.float().t()
    report("epi bias_bf16", ops.gemm(a, b, ops.EPI_BIAS_BF16, bias=bias), acc + bias)
    pre = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    act = ops.gemm(a, b, ops.EPI_BIAS_GELU_BF16, bias=bias, out2=pre)
    report("epi gelu act", act, torch.nn.functional.gelu(acc + bias))
    report("epi gelu pre", pre, acc + bias)
    resid = torch.randn(M, N, device=dev)
    report("epi bias_resid_f32", ops.gemm(a, b, ops.EPI_BIAS_RESID_F32, bias=bias, resid=resid), acc + bias + resid, tol=1e-2)
    report("epi plain_bf16", ops.gemm(a, b, ops.EPI_PLAIN_BF16), acc)
    aux = rnd(M, N)
    x = aux.float()
    gp = 0.5 * (1 + torch.erf(x / 2 ** 0.5)) + x * torch.exp(-0.5 * x * x) / (2 * torch.pi) ** 0.5
    report("epi dgelu", ops.gemm(a, b, ops.EPI_DGELU_BF16, aux=aux), acc * gp)
    report("epi bias_f32", ops.gemm(a, b, ops.EPI_BIAS_F32, bias=bias), acc + bias, tol=1e-2)
    # dropout: keep-rate and scale
    p = 0.1
    y = ops.gemm(a, b, ops.EPI_BIAS_RESID_F32, bias=bias, resid=torch.zeros(M, N, device=dev), dropout_p=p, seed=123, site=7)
    keep = (y != 0).float().mean().item()
    kept = y != 0
    ratio = (y[kept] / (acc + bias)[kept]).median().item()
    y2 = ops.gemm(a, b, ops.EPI_BIAS_RESID_F32, bias=bias, resid=torch.zeros(M, N, device=dev), dropout_p=p, seed=123, site=7)
    print(json.dumps(dict(case="epi dropout", keep=keep, scale=ratio, deterministic=bool(torch.equal(y, y2)),
                          ok=bool(abs(keep - 0.9) < 0.01 and abs(ratio - 1 / 0.9) < 0.02))), flush=True)


def group_time():
    shapes = [("qkv", 5920, 2304, 768), ("attn_out", 5920, 768, 768), ("mlp1", 5920, 3072, 768), ("mlp2", 5920, 768, 3072),
              ("lm_qkv", 1280, 2304, 768), ("lm_mlp1", 1280, 3072, 768), ("big", 11808, 3072, 768), ("sq8k", 8192, 8192, 8192)]
    for name, M, N, K in shapes:
        a, b = rnd(M, K), rnd(N, K)
        for bn in (0, 128, 256):
            out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
            for _ in range(3):
                ops.gemm(a, b, ops.EPI_PLAIN_BF16, out=out, block_n=bn)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 20
            e0.record()
            for _ in range(reps):
                ops.gemm(a, b, ops.EPI_PLAIN_BF16, out=out, block_n=bn)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            tf = 2.0 * M * N * K / (ms * 1e-3) / 1e12
            # cuBLAS reference
            for _ in range(3):
                torch.matmul(a, b.t())
            torch.cuda.synchronize()
            e0.record()
            for _ in range(reps):
                torch.matmul(a, b.t())
            e1.record()
            torch.cuda.synchronize()
            ms_ref = e0.elapsed_time(e1) / reps
            print(json.dumps(dict(case=f"time {name} {M}x{N}x{K} bn{bn}", ms=ms, tflops=tf, cublas_ms=ms_ref,
                                  cublas_tflops=2.0 * M * N * K / (ms_ref * 1e-3) / 1e12)), flush=True)


if __name__ == "__main__":
    torch.manual_seed(0)
    groups = sys.argv[1:] or ["kk", "kmn", "mnmn", "epi", "time"]
    for g in groups:
        t0 = time.time()
        try:
            globals()["group_" + g]()
        except Exception as e:  # noqa: BLE001
            print(json.dumps(dict(group=g, exception=repr(e))), flush=True)
        print(json.dumps(dict(group=g, seconds=time.time() - t0)), flush=True)
