"""CPU: the C-ABI library builds, loads and exports every symbol include/*.h declares, and its entry points reject bad
arguments (host-side checks, no device needed)."""
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


def test_entry_points_reject_bad_arguments_before_touching_the_device(built):
    """error behaviour of the boundary (include/aadg_b200.h: 0 / negative AADG_E* + aadg_last_error()): argument checks
    run on the host before any CUDA call, so they can be exercised without a GPU.  The fake pointers are 16-byte
    aligned and never dereferenced."""
    from aadg_b200 import _lib
    l = _lib.lib()
    EINVAL = -1
    P = 0x100000

    def err():
        return l.aadg_last_error().decode()
    # depthwise: channels must be a multiple of 8
    assert l.aadg_dwconv3x3_strided(P, 1, 8, 8, 12, 16, P, 1, 2, 0, P, 4, 4, 16, None) == EINVAL
    assert "multiple of 8" in err()
    # ... and the output size is fixed by the geometry (MobileNetV2's padding = dilation convention)
    assert l.aadg_dwconv3x3_strided(P, 1, 8, 8, 16, 16, P, 1, 2, 0, P, 5, 4, 16, None) == EINVAL
    assert "output size mismatch" in err()
    assert l.aadg_dwconv3x3_strided_wgrad(P, 1, 9, 9, 16, 16, P, 4, 5, 16, 1, 2, P, None) == EINVAL
    # channel strides that break the 16-byte access rule
    assert l.aadg_dwconv3x3_strided(P, 1, 8, 8, 16, 12, P, 1, 2, 0, P, 4, 4, 16, None) == EINVAL
    assert "16-byte" in err()
    # tensor-core convolution: channels in multiples of 8, sane geometry
    assert l.aadg_conv_fprop_bf16(P, 1, 8, 8, 12, 16, P, 64, 3, 3, 1, 1, 1, P, 8, 8, 64, 0, 0, None) < 0
    assert l.aadg_conv_fprop_bf16(P, 0, 8, 8, 16, 16, P, 64, 3, 3, 1, 1, 1, P, 8, 8, 64, 0, 0, None) < 0
    # stem im2col: patch length must hold R*S*3 values in multiples of 8
    assert l.aadg_im2col_stem(P, 1, 32, 32, 7, 7, 2, 3, 100, P, None) == EINVAL
    assert "kp" in err()
    # a successful query leaves the codes alone
    assert l.aadg_version() == 1
