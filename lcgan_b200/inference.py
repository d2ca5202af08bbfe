"""Generator-only inference (BASELINE config 5): the `generator_ema(geo, app, w_psi)` forward
that the reference's fake_image_generation / demo_generation loops call under no_grad (worker.py:427-441, 447-485),
with the truncation lerp of cnn.py:99-101.

`GeneratorRunner` keeps, per batch size, one captured CUDA graph of the whole forward: latents in a static buffer,
weights packed to bf16 once (the packs are held by the runner), no autograd graph and therefore no saved
activations - intermediates are freed as the forward proceeds, so batch 64 at 1024x1024 fits easily.  The
reference's `((x + 1) / 2).clamp(0, 1)` (worker.py:436) and the uint8 conversion torchvision's save_image does
before encoding are part of the graph; JPEG / mp4 encoding stays on the host as in the reference.
"""
from __future__ import annotations

import torch

from . import _lib, ops


class GeneratorRunner:
    def __init__(self, generator, w_psi: float = 1.0, batch_sizes=(1, 2, 4, 8, 16, 32, 64)):
        g = generator.module if hasattr(generator, "module") else generator
        self.G = g.eval()
        self.w_psi = float(w_psi)
        assert self.w_psi > 0, "inference uses the truncation branch (cnn.py:99-101); w_psi <= 0 updates avg_latent"
        self.device = next(g.parameters()).device
        if self.device.type != "cuda":
            raise RuntimeError("lcgan_b200 inference needs CUDA tensors (there is no CPU fallback)")
        self.batch_sizes = tuple(sorted(batch_sizes))
        self._graphs = {}
        self._packs = []
        self.launches = {}

    # ---- eager path (any batch size) ---------------------------------------------------------------
    @torch.no_grad()
    def forward_eager(self, z_geo, z_app):
        return self.G(z_geo, z_app, self.w_psi)

    @staticmethod
    def postprocess(x):
        """worker.py:436: ((x + 1) / 2).clamp(0, 1)"""
        return ((x + 1) / 2).clamp_(0.0, 1.0)

    @staticmethod
    def to_uint8(x01):
        """what torchvision.utils.save_image does to a [0,1] image before encoding"""
        return x01.mul(255).add_(0.5).clamp_(0, 255).to(torch.uint8)

    # ---- captured path -------------------------------------------------------------------------------
    def _capture(self, b):
        g = self.G
        dev = self.device
        st = {"z_geo": torch.zeros(b, g.geo_noise_dim, device=dev), "z_app": torch.zeros(b, g.app_noise_dim, device=dev)}
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side), torch.no_grad():
            ops.prepack(g)
            for _ in range(2):                          # lazy inits; fills the weight-pack cache (packed ONCE)
                self.forward_eager(st["z_geo"], st["z_app"])
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        # the graph reads the packs through raw pointers: keep them alive here, whatever happens to the cache
        ids = {id(p) for p in g.parameters()}
        self._packs.append([t for (wid, _kind), (_tag, t) in list(ops._pack_cache.items()) if wid in ids])
        graph = torch.cuda.CUDAGraph()
        n0 = _lib.launches
        with torch.cuda.graph(graph), torch.no_grad():
            img = self.forward_eager(st["z_geo"], st["z_app"])
            st["image"] = img
            st["image01"] = self.postprocess(img.clone())
            st["uint8"] = self.to_uint8(st["image01"].clone())
        self.launches[b] = _lib.launches - n0
        st["graph"] = graph
        self._graphs[b] = st
        return st

    def refresh(self):
        """Call after the generator's weights changed (load_state_dict, EMA update): drops the graphs, which hold
        packed copies of the old weights."""
        self._graphs.clear()
        self._packs.clear()

    def static_inputs(self, b):
        st = self._graphs.get(b) or self._capture(b)
        return st["z_geo"], st["z_app"]

    def replay(self, b):
        """Run the captured forward on whatever is in the static latent buffers; returns the static outputs
        (raw image [-1,1] fp32 NCHW, [0,1] image, uint8 image) - overwritten by the next replay."""
        st = self._graphs.get(b) or self._capture(b)
        st["graph"].replay()
        return st["image"], st["image01"], st["uint8"]

    def __call__(self, z_geo, z_app):
        """generator_ema(z_geo, z_app, w_psi) -> [b,3,R,R] fp32 in [-1,1] (a fresh tensor).  Batch sizes in
        `batch_sizes` replay a graph; others are padded up to the next captured size, or run eagerly if larger."""
        b = z_geo.shape[0]
        cap = next((n for n in self.batch_sizes if n >= b), None)
        if cap is None:
            return self.forward_eager(z_geo.to(self.device), z_app.to(self.device))
        zg, za = self.static_inputs(cap)
        zg[:b].copy_(z_geo, non_blocking=True)
        za[:b].copy_(z_app, non_blocking=True)
        img, _, _ = self.replay(cap)
        return img[:b].clone()
