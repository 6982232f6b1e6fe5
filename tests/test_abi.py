"""CPU tests of the drop-in boundary: the product library loads and exports
every symbol the headers declare (no compute calls: there is no GPU here)."""
import os
import re

import pytest

from msolve_b200 import backend

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    syms = set()
    for h in ("f4la_b200.h", "f4la_b200_neogb.h"):
        src = open(os.path.join(ROOT, "include", h)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        syms |= set(re.findall(r"\b(f4la_b200_\w+)\s*\(", src))
    return syms


def test_library_exports_every_declared_symbol():
    L = backend.load_library()
    declared = declared_symbols()
    assert declared == set(backend.EXPORTED_SYMBOLS)
    for s in declared:
        assert hasattr(L, s), f"{s} declared in include/ but not exported"


def test_version_and_device_count_callable_without_gpu():
    L = backend.load_library()
    assert b"f4la_b200" in L.f4la_b200_version()
    assert L.f4la_b200_device_count() >= 0


def test_no_cpu_fallback_when_no_device():
    L = backend.load_library()
    if L.f4la_b200_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(backend.F4LAError):
        backend.Backend(0)


def test_product_does_not_reference_the_oracle():
    """Nothing under msolve_b200/ may import, link or dlopen oracle/."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "msolve_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", "Makefile")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                code = "\n".join(l for l in txt.splitlines() if "oracle" in l and not l.strip().startswith(("#", "*", "/*", "//", '"')))
                assert "import oracle" not in code and "from oracle" not in code and "oracle/" not in code.replace("oracle/`", ""), (f, code)


def test_layout_descriptor_matches_reference_when_available():
    from oracle import refharness as rh
    if not rh.available():
        pytest.skip("reference harness not built")
    import ctypes as C
    buf = (C.c_uint32 * 64)()
    rh.lib().refh_fill_layout(buf)
    L = backend.load_library()
    assert L.f4la_b200_neogb_configure(buf) == 0
    assert buf[0] == 44 * 4          # sizeof(f4la_neogb_layout): 44 uint32 fields
    assert list(buf[36:43]) == [6, 5, 4, 3, 2, 1, 0]
