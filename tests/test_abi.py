"""The C-ABI library loads and exports every symbol include/sisgpu.h declares (no compute
calls: the build container has no GPU), and refuses to run without a device."""
import ctypes as C
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    txt = (ROOT / "include" / "sisgpu.h").read_text()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(sisgpu_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_exported():
    from sis_b200.api import library
    lib = library()
    syms = declared_symbols()
    assert len(syms) >= 40
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from sis_b200._capi import SisError
    from sis_b200.api import SisGpu
    with pytest.raises(SisError, match="no CUDA device"):
        SisGpu(0)


def test_product_does_not_use_oracle():
    """The product package may mention the oracle in prose, never import / link / load it."""
    pat = re.compile(r"(^\s*(from|import)\s+oracle\b)|sis_oracle|liboracle|oracle_[a-z]+\s*\(", re.M)
    for p in (ROOT / "sis_b200").rglob("*"):
        if p.suffix in (".py", ".cu", ".cuh", ".h", ".F90"):
            assert not pat.search(p.read_text()), p


def header_prototypes():
    """name -> number of arguments, parsed from include/sisgpu.h"""
    txt = (ROOT / "include" / "sisgpu.h").read_text()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(sisgpu_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", txt, flags=re.S):
        args = m.group(2).strip()
        out[m.group(1)] = 0 if args in ("", "void") else args.count(",") + 1
    return out


def test_fortran_shim_matches_header():
    """Every bind(C) interface in sis_b200/fortran/sisgpu_mod.F90 names a function the header
    declares, with the same number of arguments (no Fortran compiler exists in this image)."""
    protos = header_prototypes()
    src = (ROOT / "sis_b200" / "fortran" / "sisgpu_mod.F90").read_text()
    src = re.sub(r"&\s*\n\s*&?", " ", src)     # join continuation lines
    src = re.sub(r"!.*", "", src)
    found = {}
    for m in re.finditer(r"function\s+(sisgpu_\w+)\s*\(([^)]*)\)\s*bind\(C,\s*name='(\w+)'\)", src):
        assert m.group(1) == m.group(3)
        args = m.group(2).strip()
        found[m.group(1)] = 0 if not args else args.count(",") + 1
    assert len(found) >= 35
    for name, n in found.items():
        assert name in protos, f"{name} not declared in sisgpu.h"
        assert protos[name] == n, f"{name}: shim has {n} arguments, header {protos[name]}"
