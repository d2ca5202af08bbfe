"""Helpers of the GPU parity tests.

Leaky-relu masks.  A gradient through lrelu(z) depends on sign(z); two correct implementations disagree on
the sign of a pre-activation that lies within rounding noise of zero, and ONE such flip among 1e5 elements
reads as ~2e-3 rel-L2 on a block's gradients (numerics).  `masked_oracle` makes that effect
measurable instead of hiding it behind a loose bound: the oracle is run twice on the same inputs,
  * free   - with its own masks: forward parity, and the number of masks that differ from ours is counted;
  * shared - with OUR masks injected into its leaky-relus: every remaining gradient difference is arithmetic
             (operand rounding, summation order), and is held to north_star's per-layer bound.
Measured values are appended to gpurun_out/parity_report.jsonl when that directory exists (profiles/ keeps
a copy of a B200 run).
"""
import contextlib
import json
import os

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def report(**row):
    d = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(d):
        with open(os.path.join(d, "parity_report.jsonl"), "a") as f:
            f.write(json.dumps(row) + "\n")


@contextlib.contextmanager
def masked_oracle(O, masks=None):
    """Patch the oracle's leaky-relu (oracle.lcgan_oracle.lrelu).  Yields the list of masks the oracle itself
    would have used, in call order.  masks: list of bool tensors (logical NCHW) to use instead, in call order."""
    own, it = [], iter(masks) if masks is not None else None
    orig = O.lrelu

    def lrelu(x, gain=1.0):
        own.append((x > 0).detach())
        if it is None:
            return orig(x, gain)
        m = next(it).to(x.device)
        assert m.shape == x.shape, (m.shape, x.shape)
        return x * torch.where(m, 1.0, 0.2).to(x.dtype) * gain

    O.lrelu = lrelu
    try:
        yield own
    finally:
        O.lrelu = orig


class MaskRecorder:
    """Collects the sign masks of our fused activations, in forward order.  Every leaky-relu of the CUDA path lives in
    a kernel epilogue, so the recorder wraps the two entry points that apply one: ops.tapconv (conv epilogue, recorded
    when the call has slope != 1) and ops.Box3Act / ops.Box3ActMod (box filter + lrelu [* style])."""

    def __init__(self, modules=(), box3act=True):
        from lcgan_b200 import ops
        self.masks, self._ops = [], ops
        self._tapconv, self._box, self._boxmod = ops.tapconv, ops.Box3Act.apply, ops.Box3ActMod.apply

        def tapconv(*a, **k):
            out = self._tapconv(*a, **k)
            slope = k.get("slope", a[7] if len(a) > 7 else 1.0)
            if slope != 1.0:
                self.masks.append((a[2] > 0).detach())
            return out

        def box(*a):
            out = self._box(*a)
            self.masks.append((out > 0).detach())
            return out

        def boxmod(x, s, *a):                 # output = lrelu(box(x)) * gain * s: the mask is the sign of output * s
            out = self._boxmod(x, s, *a)
            self.masks.append(((out.float() * s[:, :, None, None]) > 0).detach())
            return out
        ops.tapconv, ops.Box3Act.apply, ops.Box3ActMod.apply = tapconv, box, boxmod

    def close(self):
        self._ops.tapconv, self._ops.Box3Act.apply, self._ops.Box3ActMod.apply = self._tapconv, self._box, self._boxmod

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def flip_fraction(ours, theirs):
    """Largest fraction of differing mask bits over the activations of a block."""
    assert len(ours) == len(theirs), (len(ours), len(theirs))
    return max(float((a.to(b.device) != b).float().mean()) for a, b in zip(ours, theirs)) if ours else 0.0
