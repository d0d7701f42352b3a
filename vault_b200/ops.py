"""Thin torch-tensor wrappers over the C ABI (include/vault_b200.h): pull data_ptr / shapes / the current CUDA stream and
call the kernel.  No arithmetic happens here and nothing falls back to torch ops."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _abi
from ._abi import (EPI_ATOMIC_BIAS_DROP_F32, EPI_ATOMIC_F32, EPI_BIAS_BF16, EPI_BIAS_F32, EPI_BIAS_GELU_BF16, EPI_BIAS_GELU_GRAD_BF16, EPI_BIAS_RESID_F32,
                   EPI_DGELU_BF16, EPI_MUL_AUX_BF16, EPI_PLAIN_BF16, EPI_STORE_F32, GemmArgs)

_OUT_F32 = {EPI_BIAS_RESID_F32, EPI_ATOMIC_F32, EPI_BIAS_F32, EPI_STORE_F32, EPI_ATOMIC_BIAS_DROP_F32}


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("vault_b200 kernels take CUDA tensors only (there is no CPU path)")


def gemm(a: torch.Tensor, b: torch.Tensor, epilogue: int, *, a_mn: bool = False, b_mn: bool = False, bias=None, resid=None,
         aux=None, out=None, out2=None, dropout_p: float = 0.0, seed: int = 0, site: int = 0, split_k: int = 1,
         block_n: int = 0, max_ctas: int = 0, cluster: int = 0, sched=None, a_colsum=None) -> torch.Tensor:
    """C[M,N] = sum_k A[m,k] B[n,k] with a fused epilogue.  a: [M,K] (or [K,M] if a_mn), b: [N,K] (or [K,N] if b_mn), bf16,
    last dim contiguous."""
    _cuda(a, b, bias, resid, aux, out, out2)
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16 and a.dim() == 2 and b.dim() == 2
    assert a.stride(1) == 1 and b.stride(1) == 1
    M, K = (a.shape[1], a.shape[0]) if a_mn else (a.shape[0], a.shape[1])
    N, Kb = (b.shape[1], b.shape[0]) if b_mn else (b.shape[0], b.shape[1])
    assert K == Kb, (a.shape, b.shape, a_mn, b_mn)
    if out is None:
        out = torch.empty((M, N), device=a.device, dtype=torch.float32 if epilogue in _OUT_F32 else torch.bfloat16)
    assert out.stride(1) == 1 and out.shape == (M, N)
    g = GemmArgs()
    g.M, g.N, g.K = M, N, K
    g.A, g.lda, g.a_mn = a.data_ptr(), a.stride(0), int(a_mn)
    g.B, g.ldb, g.b_mn = b.data_ptr(), b.stride(0), int(b_mn)
    g.epilogue = epilogue
    g.bias = _ptr(bias)
    g.resid, g.ldr = _ptr(resid), (resid.stride(0) if resid is not None else 0)
    g.aux, g.ldaux = _ptr(aux), (aux.stride(0) if aux is not None else 0)
    g.out, g.ldo = out.data_ptr(), out.stride(0)
    g.out2, g.ldo2 = _ptr(out2), (out2.stride(0) if out2 is not None else 0)
    g.dropout_p, g.seed, g.site = dropout_p, seed, site
    g.split_k, g.block_n, g.max_ctas, g.cluster = split_k, block_n, max_ctas, cluster
    g.sched = _ptr(sched)
    g.a_colsum = _ptr(a_colsum)
    _abi.call("vault_gemm_bf16", C.byref(g), _stream())
    return out


def layernorm_fwd(x, gamma, beta, eps, *, want_bf16=True, want_f32=False, dropout_p=0.0, seed=0, site=0):
    _cuda(x, gamma, beta)
    assert x.dtype == torch.float32 and x.is_contiguous()
    rows, cols = x.numel() // x.shape[-1], x.shape[-1]
    y16 = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16) if want_bf16 else None
    y32 = torch.empty_like(x) if want_f32 else None
    mean = torch.empty(rows, device=x.device, dtype=torch.float32)
    rstd = torch.empty(rows, device=x.device, dtype=torch.float32)
    _abi.call("vault_layernorm_fwd_drop", x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), _ptr(y16), _ptr(y32), mean.data_ptr(),
              rstd.data_ptr(), rows, cols, eps, dropout_p, seed, None, site, _stream())
    return y16, y32, mean, rstd


def layernorm_bwd(dy_f32, dy_bf16, x, mean, rstd, gamma, dres, dgamma, dbeta, *, want_bf16=True, in_p=0.0, in_site=0, out_p=0.0,
                  out_site=0, seed=0, dcolsum=None):
    _cuda(dy_f32, dy_bf16, x, mean, rstd, gamma, dres, dgamma, dbeta)
    rows, cols = x.numel() // x.shape[-1], x.shape[-1]
    dx32 = torch.empty_like(x)
    dx16 = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16) if want_bf16 else None
    _abi.call("vault_layernorm_bwd_drop", _ptr(dy_f32), _ptr(dy_bf16), x.data_ptr(), mean.data_ptr(), rstd.data_ptr(), gamma.data_ptr(),
              _ptr(dres), dx32.data_ptr(), _ptr(dx16), _ptr(dgamma), _ptr(dbeta), _ptr(dcolsum), rows, cols, in_p, in_site, out_p, out_site, seed, None,
              _stream())
    return dx32, dx16


def gemm_wgrad_grouped(problems, max_ctas: int = 0):
    """problems: list of (dy [tokens, n_out] bf16, x [tokens, k_in] bf16, dW [n_out, k_in] fp32 (accumulated into), split_k, db [n_out] fp32 or None).
    One persistent launch over all their 128x256 tiles (include/vault_b200.h: vault_gemm_wgrad_grouped)."""
    arr = (GemmArgs * len(problems))()
    for g, (dy, x, dw, split, db) in zip(arr, problems):
        _cuda(dy, x, dw, db)
        assert dy.dtype == torch.bfloat16 and x.dtype == torch.bfloat16 and dw.dtype == torch.float32 and dy.shape[0] == x.shape[0]
        g.M, g.N, g.K = dy.shape[1], x.shape[1], dy.shape[0]
        g.A, g.lda, g.a_mn = dy.data_ptr(), dy.stride(0), 1
        g.B, g.ldb, g.b_mn = x.data_ptr(), x.stride(0), 1
        g.epilogue = EPI_ATOMIC_F32
        g.out, g.ldo = dw.data_ptr(), dw.stride(0)
        g.split_k, g.max_ctas = split, max_ctas
        g.a_colsum = _ptr(db)
    _abi.call("vault_gemm_wgrad_grouped", arr, len(problems), _stream())
