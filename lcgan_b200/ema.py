"""EMA of generator weights with the surface of the reference's ema.py (ema.py:4-32); the update
is one multi-tensor kernel launch instead of one lerp + one copy per tensor."""
import torch

from . import ops


class Ema(object):
    def __init__(self, source, target, decay=0.9999, start_iter=0):
        self.source = source
        self.target = target
        self.decay = decay
        self.start_iter = start_iter
        with torch.no_grad():
            for p_ema, p in zip(self.target.parameters(), self.source.parameters()):
                p_ema.copy_(p)
            for b_ema, b in zip(self.target.buffers(), self.source.buffers()):
                b_ema.copy_(b)

    def decay_at(self, iter):
        """ema.py:19-23: plain copy (decay 0) before start_iter."""
        return 0.0 if (iter >= 0 and iter < self.start_iter) else self.decay

    def update(self, iter=None, decay_dev=None):
        """decay_dev: optional device scalar holding decay_at(iter) (CUDA-graph replays)."""
        decay = self.decay_at(iter)
        with torch.no_grad():
            dst, src = [], []
            for p_ema, p in zip(self.target.parameters(), self.source.parameters()):
                dst.append(p_ema); src.append(p)
            for (name, b_ema), (_, b) in zip(self.target.named_buffers(), self.source.named_buffers()):
                if "num_batches_tracked" in name:
                    b_ema.copy_(b)
                else:
                    dst.append(b_ema); src.append(b)
            ops.ema_lerp_(dst, src, float(decay), decay_dev)
