"""CPU checks of the drop-in boundary: the C-ABI library builds for sm_100a, loads, and exports every symbol that
include/vault_b200.h declares; the ctypes table mirrors the header; no compute is called (no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "vault_b200.h")


def header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(?:int|size_t)\s+(vault_\w+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib_path():
    from vault_b200 import build

    return build.build()


def test_header_declares_the_expected_families():
    fns = header_functions()
    for must in ("vault_gemm_bf16", "vault_layernorm_fwd", "vault_layernorm_bwd", "vault_attn_fwd", "vault_attn_bwd", "vault_lm_embed_fwd",
                 "vault_vilt_assemble_fwd", "vault_ce_loss", "vault_adamw_step", "vault_last_error", "vault_version"):
        assert must in fns


def test_library_exports_every_declared_symbol(lib_path):
    l = ctypes.CDLL(lib_path)
    for fn in header_functions():
        assert hasattr(l, fn), f"{fn} declared in include/vault_b200.h but not exported by {lib_path}"


def test_ctypes_table_matches_header(lib_path):
    from vault_b200 import _abi

    assert sorted(_abi.SIGNATURES) == header_functions()
    src = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    for name, argtypes in _abi.SIGNATURES.items():
        m = re.search(r"\b" + name + r"\s*\((.*?)\)\s*;", src, flags=re.S)
        assert m, name
        params = [p for p in m.group(1).split(",") if p.strip() and p.strip() != "void"]
        assert len(params) == len(argtypes), (name, len(params), len(argtypes))
    lib = _abi.lib()
    assert lib.vault_version() >= 100


def test_gemm_args_struct_layout_matches_header():
    """Field order of the ctypes Structure == field order of `struct vault_gemm_args`."""
    from vault_b200 import _abi

    src = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    body = re.search(r"typedef struct vault_gemm_args \{(.*?)\} vault_gemm_args;", src, flags=re.S).group(1)
    names = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        for part in decl.split(","):
            names.append(re.findall(r"(\w+)\s*$", part.strip())[0])
    assert names == [f[0] for f in _abi.GemmArgs._fields_]


def test_no_gpu_means_loud_failure():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from vault_b200 import _abi

    rc = _abi.lib().vault_check_device(0)
    assert rc != 0 and "no CUDA device" in _abi.last_error()


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under vault_b200/ may import it."""
    pkg = os.path.join(ROOT, "vault_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
                assert "vault_oracle" not in txt and "ref_loader" not in txt, f
