"""The C-ABI library loads and exports every symbol include/*.h declares (no GPU needed)."""
import ctypes
import os
import re

from conftest import ROOT


def declared_symbols():
    names = set()
    for header in ("gfb200.h", "gfb_rays.h", "graph_c_binding.h"):
        text = open(os.path.join(ROOT, "include", header)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        text = re.sub(r"//[^\n]*", "", text)
        for m in re.finditer(r"\b((?:gfb|graph)_[a-zA-Z0-9_]+)\s*\(", text):
            names.add(m.group(1))
    names.discard("graph_c_context")
    return sorted(names)


def test_every_declared_symbol_is_exported(lib):
    so = ctypes.CDLL(os.path.join(ROOT, "graph_framework_b200", "libgfb200.so"))
    syms = declared_symbols()
    assert len(syms) > 80
    missing = [s for s in syms if not hasattr(so, s)]
    assert not missing, missing


def test_python_binding_covers_every_declared_symbol(lib):
    bound = set(lib._gfb_signatures)
    assert set(declared_symbols()) - bound <= {"gfb_set_last_error"}


def test_no_device_fails_loudly(lib):
    """No CPU fallback: without a CUDA device creating a context or a tracer is an error."""
    if lib.gfb_device_count() > 0:
        return
    assert not lib.gfb_ctx_create(0)
    assert b"no CUDA device" in lib.gfb_last_error()
    assert not lib.gfb_rays_create(b"simple", b"slab", b"", b"rk4", 4, 0.1, 0, None)
    assert b"no CPU fallback" in lib.gfb_last_error()


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under graph_framework_b200/ may import, include,
    link or execute anything under oracle/."""
    pkg = os.path.join(ROOT, "graph_framework_b200")
    pattern = re.compile(r"(from\s+oracle|import\s+oracle|oracle/|oracle\.|#include\s+[\"<][^\n]*oracle)")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cpp", ".hpp", ".cu", ".cuh", ".h", ".inc")):
                text = open(os.path.join(base, f), errors="replace").read()
                assert not pattern.search(text), os.path.join(base, f)


def test_build_entry_nvrtc_smoke(lib):
    """The hand-written traits struct that __graft_entry__.build() compiles must follow the skeleton."""
    import sys
    sys.path.insert(0, ROOT)
    import __graft_entry__
    __graft_entry__.nvrtc_smoke()
