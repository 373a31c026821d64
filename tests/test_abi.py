"""CPU tests of the boundary: the C-ABI library loads, exports every symbol its header
declares, and fails loudly (no CPU fallback) when there is no GPU."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT


def declared_symbols(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(ptmb_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from rti_b200 import _lib
    lib = _lib.load()
    names = declared_symbols("ptm_b200.h")
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), "libptm_b200.so does not export %s" % n
    # and the Python binding covers the whole header
    assert sorted(_lib.EXPORTS) == names


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from rti_b200 import _lib
    from rti_b200.device import Context
    with pytest.raises(_lib.PtmError):
        Context(0)


def test_product_does_not_import_oracle():
    # the oracle is test infrastructure; nothing under rti_b200/ may reach into it
    for dirpath, _, files in os.walk(os.path.join(ROOT, "rti_b200")):
        for f in files:
            if f.endswith((".py", ".c", ".cu", ".h", ".cuh", "Makefile")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "import oracle" not in txt and "from oracle" not in txt, f
                assert "liboracle" not in txt and "libptmref" not in txt, f
