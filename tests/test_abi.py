"""CPU: the C-ABI library builds, loads and exports every symbol include/*.h declares."""
import ctypes
import glob
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = []
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        text = re.sub(r"/\*.*?\*/", "", open(h).read(), flags=re.S)
        names += re.findall(r"\b(aadg_[a-z0-9_]+)\s*\(", text)
    return sorted(set(names))


@pytest.fixture(scope="module")
def built():
    from aadg_b200 import build
    return build.build()


def test_header_declares_something():
    assert "aadg_version" in declared_symbols()
    assert len(declared_symbols()) >= 5


def test_library_exports_every_declared_symbol(built):
    lib = ctypes.CDLL(built)
    missing = [n for n in declared_symbols() if not hasattr(lib, n)]
    assert not missing, missing


def test_binding_table_matches_header(built):
    from aadg_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_symbols()
    l = _lib.lib()
    assert l.aadg_version() == 1
    assert l.aadg_u8_workspace_bytes(0, 1, 8, 8) == 0
    assert l.aadg_u8_workspace_bytes(4, 2, 64, 64) > 4 * 3 * 64 * 64 * 3


def test_row_struct_size_matches_header():
    from aadg_b200.data.decisions import ROW_DTYPE
    text = open(os.path.join(ROOT, "include", "aadg_b200.h")).read()
    assert "AADG_MAX_OPS 4" in text
    assert ROW_DTYPE.itemsize == 4 * (2 + 4 + 4 + 24 + 6)


def test_ops_refuse_cpu_tensors():
    import numpy as np
    import torch
    from aadg_b200.ops import u8
    from aadg_b200.data.decisions import ROW_DTYPE
    with pytest.raises(RuntimeError):
        u8.apply_policy(torch.zeros(1, 8, 8, 3, dtype=torch.uint8), None, np.zeros(1, ROW_DTYPE))
