"""The built objects really contain the Blackwell instructions the design relies on (cuobjdump -sass of the sm_100a
objects __graft_entry__.build() produces): tcgen05.mma (UTCHMMA), TMA loads (UTMALDG), TMEM loads (LDTM) in the conv
kernels; TMA window loads and the mixed-precision FMA (FHFMA) in the flow-warp kernels; TMA windows in the box filter."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "lcgan_b200", "csrc", "_build")


def _sass(obj):
    path = os.path.join(BUILD, obj)
    if shutil.which("cuobjdump") is None or not os.path.exists(path):
        pytest.skip("needs cuobjdump and the built objects (python -c 'import __graft_entry__ as e; e.build()')")
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    assert "sm_100a" in out, "objects must be compiled for sm_100a"
    return out


def _count(sass, mnemonic):
    return len(re.findall(r"\b" + re.escape(mnemonic), sass))


def test_conv_kernels_use_tcgen05_tmem_and_tma():
    s = _sass("conv_tc.o")
    assert _count(s, "UTCHMMA") > 300          # tcgen05.mma in the forward (2 instantiations) and three wgrad kernels
    assert _count(s, "UTMALDG") > 100          # cp.async.bulk.tensor
    assert _count(s, "LDTM") >= 8              # tcgen05.ld
    assert _count(s, "UTCBAR") >= 8            # tcgen05.commit -> mbarrier
    assert "HMMA.16816" not in s and "WGMMA" not in s.upper().replace("UTCHMMA", "")   # no legacy tensor-core paths


def test_warp_and_box_kernels_use_tma_windows_and_mixed_precision_fma():
    w = _sass("warp.o")
    assert _count(w, "FHFMA") > 500            # fma.rn.f32.bf16: the 16-tap gathers without unpack instructions
    assert _count(w, "UTMALDG") >= 5           # one bulk tensor copy per window
    b = _sass("resample.o")
    assert _count(b, "UTMALDG") >= 3
