"""The C-ABI library builds for sm_100a, loads without a GPU and exports every symbol that
include/lcgan_b200.h declares (no compute calls here)."""
import ctypes
import os
import re

from lcgan_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "lcgan_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lcgan_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    so = _lib.build()
    assert os.path.exists(so)
    lib = ctypes.CDLL(so)
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/lcgan_b200.h but not exported"
    assert set(names) == set(_lib.EXPORTS), set(names) ^ set(_lib.EXPORTS)
    assert _lib.lib().lcgan_version() >= 1


def test_descriptor_struct_matches_header_size():
    # 7 int32 (+pad) + 8 int64 + 3 int32 + 2 + 3 + 1 + 1 int32 + 27 int32 (+pad) + int64 + 3 float (+pad)
    d = _lib.TapConvDesc()
    assert ctypes.sizeof(d) % 8 == 0
    assert _lib.TapConvDesc.xs_n.offset == 32 and _lib.TapConvDesc.w_ld.offset % 8 == 0


def test_ops_fail_loudly_without_cuda():
    import pytest
    import torch
    from lcgan_b200 import ops
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.Box3.apply(torch.zeros(1, 4, 4, 4))
