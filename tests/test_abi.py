"""CPU-only: the C-ABI library loads and exports every symbol include/knockoffs_b200.h declares;
the ctypes table mirrors the header; the product refuses to run without a GPU (no CPU fallback)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "knockoffs_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(kb_[a-z0-9_]+)\s*\(", src)) - {"kb_ctx"})


def test_library_exports_every_declared_symbol(kb):
    import ctypes
    lib = ctypes.CDLL(kb.LIB_PATH)
    names = _header_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/knockoffs_b200.h but not exported"


def test_ctypes_table_matches_header(kb):
    assert sorted(kb.SIGNATURES) == _header_functions()


def test_struct_layout_matches_header(kb):
    """kb_solve_opts / kb_solve_info field order as in the header."""
    src = open(os.path.join(ROOT, "include", "knockoffs_b200.h")).read()
    opts = re.search(r"typedef struct \{((?:(?!typedef struct).)*?)\} kb_solve_opts;", src, re.S).group(1)
    info = re.search(r"typedef struct \{((?:(?!typedef struct).)*?)\} kb_solve_info;", src, re.S).group(1)
    strip = lambda b: re.findall(r"\b(?:int32_t|double)\s+([a-z_]+)", re.sub(r"/\*.*?\*/", "", b, flags=re.S))
    from importlib import import_module
    _lib = import_module("knockoffs_b200._lib")
    assert strip(opts) == [f[0] for f in _lib.SolveOpts._fields_]
    assert strip(info) == [f[0] for f in _lib.SolveInfo._fields_]


def test_no_cpu_fallback(kb):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(kb.KnockoffsError):
        kb.Context(0)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "knockoffs.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in text.replace("CPU oracle lives in oracle/", ""), f"{f} mentions oracle/"
