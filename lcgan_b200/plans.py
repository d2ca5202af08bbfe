"""Tap-convolution plans: how each conv / transposed conv / linear of the reference, and each of
their data- and weight-gradients, maps onto the one contraction the CUDA kernels implement
(include/lcgan_b200.h `struct lcgan_tapconv`).

A *launch* covers a lattice (m, n), m < MH, n < MW:
    output pixel  (m*os + py, n*os + px)
    input pixel of tap t  (m*is + dy_t, n*is + dx_t)      (out of range -> zero)
    weight slice of tap t  W[..., wtap_t]                  (wtap = kh*k + kw of the 4-D weight)

  conv k3 s1 p1   (custom_layers.py:41,43,83)   1 launch, 9 taps, os=1, is=1, d = k-1
  conv k3 s2 p1   (custom_layers.py:41,43)      1 launch, 9 taps, os=1, is=2, d = k-1
  conv k1         (custom_layers.py:41,83,25)   1 launch, 1 tap
  conv_transpose k3 s2 p1 op1 (custom_layers.py:78; out[2i-1+k] += x[i] w[k])
                                                4 phase launches (1,2,2,4 taps), os=2, is=1
The adjoint (data gradient) of a plan is again a plan (`adjoint`), so forward, dgrad and the
double-backward all run through the same kernels.  Pure Python, no torch: testable on CPU.
"""
from __future__ import annotations

from dataclasses import dataclass
from functools import lru_cache
from typing import Tuple


@dataclass(frozen=True)
class Launch:
    os: int
    py: int
    px: int
    is_: int
    MH: int
    MW: int
    taps: Tuple[Tuple[int, int, int], ...]   # (dy, dx, wtap)


@dataclass(frozen=True)
class Plan:
    IH: int
    IW: int
    OH: int
    OW: int
    k: int                       # kernel size of the 4-D weight (wtap = kh*k + kw)
    launches: Tuple[Launch, ...]


@lru_cache(maxsize=None)
def conv(k: int, stride: int, H: int, W: int) -> Plan:
    """F.conv2d(kernel k, stride, padding k//2)."""
    assert k in (1, 3) and stride in (1, 2)
    pad = k // 2
    OH = (H + 2 * pad - k) // stride + 1
    OW = (W + 2 * pad - k) // stride + 1
    taps = tuple((kh - pad, kw - pad, kh * k + kw) for kh in range(k) for kw in range(k))
    return Plan(H, W, OH, OW, k, (Launch(1, 0, 0, stride, OH, OW, taps),))


@lru_cache(maxsize=None)
def conv_transpose_up2(k: int, H: int, W: int) -> Plan:
    """F.conv_transpose2d(stride 2, padding (k-1)//2, output_padding 1) with the [O,I,k,k] weight
    used unflipped: out[2i - pad + ki, 2j - pad + kj] += x[i, j] * w[ki, kj]."""
    assert k == 3, "every x2 layer of the reference is 3x3 (k=1 would leave 3 of 4 phases unwritten)"
    pad = (k - 1) // 2
    OH, OW = 2 * H, 2 * W        # (H-1)*2 - 2*pad + k + 1
    launches = []
    for py in (0, 1):
        for px in (0, 1):
            taps = []
            for ki in range(k):
                if (py + pad - ki) % 2:
                    continue
                for kj in range(k):
                    if (px + pad - kj) % 2:
                        continue
                    taps.append(((py + pad - ki) // 2, (px + pad - kj) // 2, ki * k + kj))
            if taps:
                launches.append(Launch(2, py, px, 1, H, W, tuple(taps)))
    return Plan(H, W, OH, OW, k, tuple(launches))


@lru_cache(maxsize=None)
def adjoint(p: Plan) -> Plan:
    """Plan of the transpose map (dY -> dX): input/output roles swap, taps keep their wtap."""
    launches = []
    if all(l.is_ == 1 for l in p.launches):
        # i = m + d  ->  dX[i] = sum_t dY[(i - d_t)*os + p] W_t : one launch over i, stride os
        os_ = p.launches[0].os
        assert all(l.os == os_ for l in p.launches)
        taps = []
        for l in p.launches:
            taps += [(l.py - dy * os_, l.px - dx * os_, wt) for dy, dx, wt in l.taps]
        launches.append(Launch(1, 0, 0, os_, p.IH, p.IW, tuple(taps)))
    else:
        # os == 1, is == s: i = s*m + d -> per phase q of i: m = m' + (q - d)/s
        assert len(p.launches) == 1 and p.launches[0].os == 1
        l = p.launches[0]
        s = l.is_
        assert p.IH % s == 0 and p.IW % s == 0
        for qy in range(s):
            for qx in range(s):
                taps = [((qy - dy) // s, (qx - dx) // s, wt) for dy, dx, wt in l.taps
                        if (qy - dy) % s == 0 and (qx - dx) % s == 0]
                if taps:
                    launches.append(Launch(s, qy, qx, 1, p.IH // s, p.IW // s, tuple(taps)))
    assert all(len(l.taps) <= 9 for l in launches)
    return Plan(p.OH, p.OW, p.IH, p.IW, p.k, tuple(launches))


def linear() -> Plan:
    """F.linear as a 1x1 conv on a 1x1 image (custom_layers.py:25)."""
    return conv(1, 1, 1, 1)


def macs(p: Plan, N: int, Cin: int, Cout: int) -> int:
    """Multiply-accumulates of one application (algorithmic: zero-inserted taps not counted)."""
    return sum(l.MH * l.MW * len(l.taps) for l in p.launches) * N * Cin * Cout
