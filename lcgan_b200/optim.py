"""Fused multi-tensor Adam (reference worker.py:98-110 builds
torch.optim.Adam(lr, betas=(0, 0.99), eps=1e-8) over 165 generator / 60 discriminator tensors).

One kernel launch per 48 tensors instead of torch's per-dtype foreach chains; the tensor pointers travel
by value in the kernel arguments, so `step()` is CUDA-graph capturable without a device pointer table.
Semantics are torch.optim.Adam's (amsgrad=False, weight_decay=0, maximize=False): parameters whose
gradient is None are skipped and keep their own step count, bias corrections use each tensor's count.
With beta1 == 0 (the reference's setting) the first moment equals the gradient and is not stored.
"""
import ctypes as C

import torch

from . import _lib, ops


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        if lr < 0 or eps < 0 or not (0 <= betas[0] < 1) or not (0 <= betas[1] < 1):
            raise ValueError("FusedAdam: bad hyper-parameters")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))

    def _init_state(self, group):
        """Per-parameter state: 'step' (0-dim f32 device tensor, a view into one buffer per group),
        'exp_avg_sq', and 'exp_avg' unless beta1 == 0."""
        todo = [p for p in group["params"] if p not in self.state or len(self.state[p]) == 0]
        if not todo:
            return
        dev = todo[0].device
        steps = torch.zeros(len(todo), dtype=torch.float32, device=dev)
        for i, p in enumerate(todo):
            if p.dtype != torch.float32 or not p.is_contiguous():
                raise TypeError("FusedAdam: parameters must be contiguous float32")
            st = self.state[p]
            st["step"] = steps[i]
            st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
            if group["betas"][0] != 0:
                st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            self._init_state(group)
            b1, b2 = group["betas"]
            live = [p for p in group["params"] if p.grad is not None]
            if not live:
                continue
            ops._need_cuda(live[0])
            stream = ops._stream(live[0])
            for i in range(0, len(live), _lib.MT_MAX):
                part = live[i:i + _lib.MT_MAX]
                ch = _lib.AdamChunk()
                keep = []
                for k, p in enumerate(part):
                    g = p.grad
                    if g.dtype != torch.float32 or not g.is_contiguous():
                        g = g.float().contiguous()
                        keep.append(g)
                    st = self.state[p]
                    ch.p[k], ch.g[k], ch.v[k] = p.data_ptr(), g.data_ptr(), st["exp_avg_sq"].data_ptr()
                    ch.m[k] = st["exp_avg"].data_ptr() if "exp_avg" in st else None
                    ch.step[k], ch.numel[k] = st["step"].data_ptr(), p.numel()
                ch.count = len(part)
                _lib.call("lcgan_adam_step", C.byref(ch), C.c_float(group["lr"]), C.c_float(b1), C.c_float(b2),
                          C.c_float(group["eps"]), stream, tag="adam",
                          nbytes=sum(p.numel() for p in part) * 4 * (5 if b1 == 0 else 7))
        ops.bump_generation()          # parameters changed behind autograd's version counters
        return loss
