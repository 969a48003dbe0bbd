"""The C-ABI library must load on a CPU-only box and export every symbol that
include/itjl_b200.h declares (no compute calls here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "itjl_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(itjl_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_expected_entry_points():
    syms = declared_symbols()
    for must in ["itjl_ctx_create", "itjl_model_create", "itjl_abc_distances", "itjl_abc_generation",
                 "itjl_abc_accept", "itjl_generate_ou", "itjl_generate_ou_osc", "itjl_acf",
                 "itjl_acf_missing", "itjl_psd_hamming", "itjl_acw", "itjl_acw_acf", "itjl_fit_expdecay", "itjl_fit_expdecay_3_parameters",
                 "itjl_last_error", "itjl_fp32_peak"]:
        assert must in syms


def test_library_exports_every_declared_symbol(pkg):
    path = pkg.library_path()
    assert os.path.exists(path), "libitjl_b200.so not built (python __graft_entry__.py)"
    lib = ctypes.CDLL(path)
    for s in declared_symbols():
        assert hasattr(lib, s), f"{s} declared in include/itjl_b200.h but not exported"
    assert set(declared_symbols()) == set(pkg._lib.SIGNATURES), "ctypes table out of sync with the header"
    assert b"sm_100a" in pkg._lib.lib().itjl_version()


def test_no_device_is_a_loud_error_not_a_fallback(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(pkg.ItjlError) as e:
        pkg.Context(0)
    assert e.value.code == -3


def test_product_never_imports_the_oracle():
    pkg_dir = os.path.join(ROOT, "intrinsictimescales.jl_b200")
    for dirpath, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "itjl_oracle" not in src, f


def _declarations():
    """name -> list of C parameter types, parsed from the header (comments stripped)."""
    text = open(os.path.join(ROOT, "include", "itjl_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(itjl_[a-z0-9_]+)\s*\(([^()]*)\)\s*;", text):
        params = [p.strip() for p in m.group(2).split(",")]
        if params == ["void"] or params == [""]:
            params = []
        out[m.group(1)] = params
    return out


def test_ctypes_signatures_agree_with_the_header(pkg):
    """Arity and type class of every parameter: a pointer in the header is a void pointer in the table, an integer /
    floating type the ctypes type of the same width."""
    decl = _declarations()
    assert set(decl) == set(pkg._lib.SIGNATURES)
    scalar = {"int": ctypes.c_int, "int32_t": ctypes.c_int32, "int64_t": ctypes.c_int64, "uint32_t": ctypes.c_uint32,
              "uint64_t": ctypes.c_uint64, "double": ctypes.c_double}
    for name, params in decl.items():
        _, argtypes = pkg._lib.SIGNATURES[name]
        assert len(argtypes) == len(params), f"{name}: {len(params)} parameters in the header, {len(argtypes)} in the ctypes table"
        for p, a in zip(params, argtypes):
            if "*" in p:
                assert a in (ctypes.c_void_p, ctypes.c_char_p) or issubclass(a, ctypes._Pointer), f"{name}: `{p}` is a pointer"
            else:
                ctype = re.sub(r"\bconst\b", "", p).split()[0]
                assert ctype in scalar, f"{name}: unrecognised scalar type in `{p}`"
                assert ctypes.sizeof(a) == ctypes.sizeof(scalar[ctype]) and (a is ctypes.c_double) == (ctype == "double"), \
                    f"{name}: `{p}` bound as {a.__name__}"


def test_documents_name_only_symbols_the_header_declares():
    """INTEGRATION.md (the Julia-side binding), DESIGN.md and README.md may only cite itjl_* names that exist."""
    header = open(os.path.join(ROOT, "include", "itjl_b200.h")).read()
    known = set(re.findall(r"\bitjl_[a-z0-9_]+", header)) | {"itjl_oracle", "itjl_b200"}
    for doc in ("INTEGRATION.md", "DESIGN.md", "README.md"):
        cited = set(re.findall(r"\bitjl_[a-z0-9_]+", open(os.path.join(ROOT, doc)).read()))
        assert cited <= known, f"{doc} cites names the header does not declare: {sorted(cited - known)}"
