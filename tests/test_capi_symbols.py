"""CPU: the C-ABI library loads and exports every symbol include/auncel_b200.h declares."""
import os
import re

from auncel_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_exported():
    hdr = open(os.path.join(ROOT, "include", "auncel_b200.h")).read()
    declared = set(re.findall(r"\b(auncel_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations found"
    L = _lib.lib()
    for name in declared:
        assert hasattr(L, name), f"{name} not exported"
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
