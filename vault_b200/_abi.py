"""ctypes binding of include/vault_b200.h (the C-ABI of the sm_100a kernels).

The shared library is built in-tree by ``python -m vault_b200.build`` (also run by ``__graft_entry__.build()``).  There is
no fallback: if the library is missing, or a call returns non-zero, this raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libvault_b200.so")

c_i32, c_i64, c_u32, c_u64, c_f32, c_p = C.c_int32, C.c_int64, C.c_uint32, C.c_uint64, C.c_float, C.c_void_p

EPI_BIAS_BF16, EPI_BIAS_GELU_BF16, EPI_BIAS_RESID_F32, EPI_PLAIN_BF16 = 0, 1, 2, 3
EPI_DGELU_BF16, EPI_ATOMIC_F32, EPI_BIAS_F32, EPI_STORE_F32 = 4, 5, 6, 7
EPI_BIAS_GELU_GRAD_BF16, EPI_MUL_AUX_BF16 = 9, 10
EPI_ATOMIC_BIAS_DROP_F32 = 8


class GemmArgs(C.Structure):
    _fields_ = [
        ("M", c_i32), ("N", c_i32), ("K", c_i32),
        ("A", c_p), ("lda", c_i64), ("a_mn", c_i32),
        ("B", c_p), ("ldb", c_i64), ("b_mn", c_i32),
        ("epilogue", c_i32),
        ("bias", c_p), ("resid", c_p), ("ldr", c_i64),
        ("aux", c_p), ("ldaux", c_i64),
        ("out", c_p), ("ldo", c_i64), ("out2", c_p), ("ldo2", c_i64),
        ("dropout_p", c_f32), ("seed", c_u64), ("seed_dev", c_p), ("site", c_u32),
        ("split_k", c_i32), ("block_n", c_i32), ("max_ctas", c_i32), ("sched", c_p), ("cluster", c_i32),
        ("a_colsum", c_p),
    ]


# name -> argtypes (restype is int unless listed in _RESTYPES); mirrors include/vault_b200.h declaration by declaration
SIGNATURES = {
    "vault_version": [],
    "vault_last_error": [C.c_char_p, C.c_size_t],
    "vault_check_device": [c_i32],
    "vault_gemm_bf16": [C.POINTER(GemmArgs), c_p],
    "vault_gemm_wgrad_grouped": [C.POINTER(GemmArgs), c_i32, c_p],
    "vault_patch_embed_fwd": [c_p, c_p, c_p, c_p, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_p],
    "vault_patch_embed_wgrad_ok": [c_i32, c_i32, c_i32, c_i32, c_i32],
    "vault_patch_embed_wgrad": [c_p, c_p, c_p, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_p],
    "vault_patch_grad_rows_f32": [c_p, c_p, c_p, c_p, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_p],
    "vault_layernorm_fwd": [c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_i64, c_i32, c_f32, c_p],
    "vault_layernorm_bwd": [c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_i64, c_i32, c_p],
    "vault_layernorm_fwd_drop": [c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_i64, c_i32, c_f32, c_f32, c_u64, c_p, c_u32, c_p],
    "vault_layernorm_bwd_drop": [c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_i64, c_i32, c_f32, c_u32, c_f32, c_u32, c_u64, c_p, c_p],
    "vault_attn_fwd": [c_p, c_p, c_p, c_p, c_i32, c_i32, c_i32, c_f32, c_u64, c_p, c_u32, c_p],
    "vault_attn_bwd": [c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_i32, c_i32, c_i32, c_f32, c_u64, c_p, c_u32, c_p],
    "vault_attn_set_impl": [c_i32],
    "vault_lm_embed_fwd": [c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_i32, c_i32, c_i32, c_i32, c_p],
    "vault_lm_embed_bwd": [c_p, c_p, c_p, c_p, c_p, c_p, c_i32, c_i32, c_i32, c_i32, c_i32, c_p],
    "vault_vilt_text_embed_fwd": [c_p, c_p, c_p, c_p, c_p, c_i32, c_i32, c_i32, c_p],
    "vault_vilt_text_embed_bwd": [c_p, c_p, c_p, c_p, c_i32, c_i32, c_i32, c_p],
    "vault_patch_grid": [c_p, c_i32, c_p, c_i32, c_i32, c_i32, c_i32, c_p],
    "vault_vilt_assemble_fwd": [c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_p],
    "vault_vilt_assemble_bwd": [c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_p],
    "vault_vilt_assemble_embeds_fwd": [c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_i32, c_i32, c_i32, c_i32, c_i32, c_p],
    "vault_vilt_assemble_embeds_bwd": [c_p, c_p, c_p, c_p, c_i32, c_i32, c_i32, c_i32, c_i32, c_p],
    "vault_image_preprocess": [c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_i32, c_i32, c_i32, c_i32, c_i32, c_p],
    "vault_patchify_bf16": [c_p, c_p, c_i32, c_i32, c_i32, c_i32, c_i32, c_p],
    "vault_small_linear_fwd": [c_p, c_i64, c_p, c_p, c_p, c_i32, c_i32, c_i32, c_i32, c_p],
    "vault_small_linear_bwd": [c_p, c_p, c_p, c_i64, c_p, c_p, c_i64, c_i32, c_p, c_p, c_i32, c_i32, c_i32, c_i32, c_p],
    "vault_dropout_f32": [c_p, c_p, c_i64, c_f32, c_u64, c_p, c_u32, c_p],
    "vault_ce_loss": [c_p, c_p, c_p, c_p, c_i32, c_i32, c_f32, c_p],
    "vault_head_loss": [c_p, c_p, c_p, c_p, c_i32, c_i32, c_i32, c_f32, c_p],
    "vault_colsum_bf16": [c_p, c_i64, c_p, c_i64, c_i32, c_p],
    "vault_adamw_step": [c_p, c_p, c_i32, c_p, c_p, c_p, c_i64, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, c_i32, c_i32, c_f32, c_p, c_p],
    "vault_mc_adamw_step": [c_p, c_p, c_p, c_i64, c_p, c_i32, c_p, c_p, c_p, c_i64, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, c_i32, c_i32, c_f32,
                            c_p, c_i32, c_p],
    "vault_mc_broadcast_f32": [c_p, c_p, c_i64, c_i32, c_p],
    "vault_cast_f32_bf16": [c_p, c_p, c_i64, c_p],
}
_RESTYPES = {"vault_last_error": C.c_size_t}

_lib = None


def lib() -> C.CDLL:
    """The loaded C-ABI library.  Raises if it has not been built -- there is no other implementation to fall back to."""
    global _lib
    if _lib is None:
        path = os.environ.get("VAULT_B200_LIB", LIB_PATH)  # override: A/B builds of the same ABI (tools/)
        if not os.path.exists(path):
            raise RuntimeError(
                f"{path} is missing: build it with `python -m vault_b200.build` (nvcc, sm_100a). "
                "vault_b200 has no CPU or eager fallback."
            )
        l = C.CDLL(path)
        partial = os.environ.get("VAULT_B200_ALLOW_PARTIAL") == "1"  # kernel bring-up probes only
        for name, argtypes in SIGNATURES.items():
            try:
                fn = getattr(l, name)  # AttributeError here = header / library mismatch: fail loudly
            except AttributeError:
                if partial:
                    continue
                raise
            fn.argtypes = argtypes
            fn.restype = _RESTYPES.get(name, C.c_int)
        _lib = l
    return _lib


def last_error() -> str:
    buf = C.create_string_buffer(512)
    lib().vault_last_error(buf, 512)
    return buf.value.decode(errors="replace")


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        raise RuntimeError(f"vault_b200 {what} failed (code {rc}): {last_error()}")


def call(name: str, *args) -> None:
    check(getattr(lib(), name)(*args), name)


ATTN_IMPL = 0  # mirror of vault_attn_set_impl (set through set_attn_impl below)


def set_attn_impl(impl: int) -> None:
    """0 = automatic routing, 1 = mma.sync kernels only, 2 = whole-row tcgen05 kernels where allowed, 3 = pipelined tcgen05 kernels for every
    S <= 384, 4 = 3 with the one-tile-per-warpgroup forward (include/vault_b200.h)."""
    global ATTN_IMPL
    call("vault_attn_set_impl", int(impl))
    ATTN_IMPL = int(impl)


# kernels launched per ABI call (for launch accounting in bench.py; torch memsets / fills are not counted)
KERNELS_PER_CALL = {"vault_attn_bwd": 2, "vault_vilt_assemble_bwd": 2, "vault_small_linear_bwd": 2}


class CountingLib:
    """Proxy around the loaded library that counts kernel launches and records GEMM problems (bench.py instrumentation)."""

    def __init__(self, inner=None):
        self._inner = inner or lib()
        self.launches = 0
        self.calls = {}
        self.gemms = []  # GemmArgs copies, in launch order
        self.trace = []  # (name, args) of every call, GemmArgs copied -- replayable with replay_trace()

    def __getattr__(self, name):
        fn = getattr(self._inner, name)
        if not name.startswith("vault_") or name in ("vault_version", "vault_last_error", "vault_check_device", "vault_attn_set_impl",
                                                     "vault_patch_embed_wgrad_ok"):
            return fn

        def wrapped(*args):
            n = KERNELS_PER_CALL.get(name, 1)
            if name == "vault_attn_bwd" and ATTN_IMPL != 1 and float(args[10]) == 0.0 and 64 < int(args[8]) <= 192:
                n = 1  # fused tcgen05 backward (attention_tc.cu) instead of the dQ + dK/dV pair
            elif name == "vault_attn_bwd" and ATTN_IMPL in (0, 3, 4) and ((float(args[10]) == 0.0 and 192 < int(args[8]) <= 384)
                                                                        or (float(args[10]) > 0.0 and 64 < int(args[8]) <= 384)):
                n = 3  # delta + dQ + dK/dV kernels of attention_sm100.cu
            self.launches += n
            self.calls[name] = self.calls.get(name, 0) + 1
            if name == "vault_gemm_bf16":
                src = args[0]._obj
                cp = GemmArgs()
                C.memmove(C.byref(cp), C.byref(src), C.sizeof(GemmArgs))
                self.gemms.append(cp)
                self.trace.append((name, (cp,)))
            elif name == "vault_gemm_wgrad_grouped":
                n_prob = int(args[1])
                cp = (GemmArgs * n_prob)()
                C.memmove(cp, args[0], C.sizeof(GemmArgs) * n_prob)
                self.gemms.append(cp)  # a ctypes ARRAY of GemmArgs = one grouped launch (replay_gemm() tells them apart)
                self.trace.append((name, (cp, n_prob)))
            else:
                self.trace.append((name, args[:-1]))
            return fn(*args)

        return wrapped


def install_counter():
    """Route every subsequent ABI call through a CountingLib; returns it.  uninstall_counter() restores the plain library."""
    global _lib
    c = CountingLib(lib() if not isinstance(_lib, CountingLib) else _lib._inner)
    _lib = c
    return c


def uninstall_counter():
    global _lib
    if isinstance(_lib, CountingLib):
        _lib = _lib._inner


def replay_gemm(g, stream: int):
    """Re-issue one recorded GEMM launch: a GemmArgs (vault_gemm_bf16) or a GemmArgs array (vault_gemm_wgrad_grouped)."""
    l = lib()
    inner = l._inner if isinstance(l, CountingLib) else l
    if isinstance(g, GemmArgs):
        return inner.vault_gemm_bf16(C.byref(g), stream)
    return inner.vault_gemm_wgrad_grouped(g, len(g), stream)


def gemm_flops(g) -> float:
    return 2.0 * g.M * g.N * g.K if isinstance(g, GemmArgs) else sum(2.0 * x.M * x.N * x.K for x in g)


def replay_trace(trace, stream: int, only=None):
    """Re-issue recorded ABI calls (same pointers) on `stream`; `only` filters by entry-point name."""
    l = lib()
    inner = l._inner if isinstance(l, CountingLib) else l
    for name, args in trace:
        if only is not None and not only(name):
            continue
        if name == "vault_gemm_bf16":
            inner.vault_gemm_bf16(C.byref(args[0]), stream)
        else:
            getattr(inner, name)(*args, stream)
