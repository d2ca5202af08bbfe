"""The reference's OWN modules on the host cores  --  TEST / BASELINE INFRASTRUCTURE, NOT PRODUCT CODE.

bench.py's `--impl reference` arm and its `cpu_baseline` leg time the unmodified reference implementation of the
hot path: `cnn.Generator`, `cnn.Discriminator`, `loss.*` and `ema.Ema` imported from the reference checkout
(`baseline/_ref`, git-ignored, shipped to the GPU box; `/root/reference` in the build container), stepped by a
restatement of `loader.py:44-54` + `worker.py:137-217` with the CUDA-only lines removed (worker.py itself cannot
be imported on a CPU: it hard-codes `.cuda()`, DDP(device_ids=...) and needs albumentations / av).  Nothing of
lcgan_b200 is on this path.  If no checkout is present the callers fall back to the oracle port
(oracle/lcgan_oracle.py) and say so (`kind: "port"`).
"""
from __future__ import annotations

import copy
import importlib
import os
import sys
import time
import types

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CANDIDATES = (os.path.join(ROOT, "baseline", "_ref"), "/root/reference")


def find_checkout():
    for d in CANDIDATES:
        if os.path.exists(os.path.join(d, "cnn.py")) and os.path.exists(os.path.join(d, "loss.py")):
            return d
    return None


def load_modules(ref_dir):
    """Import the reference's cnn / loss / ema (and custom_layers through cnn) without leaving them in
    sys.modules or on sys.path, so they can never shadow - or be shadowed by - the drop-in modules."""
    names = ("custom_layers", "cnn", "loss", "ema")
    saved = {n: sys.modules.pop(n, None) for n in names}
    sys.path.insert(0, ref_dir)
    try:
        mods = {n: importlib.import_module(n) for n in names}
        for n, m in mods.items():
            assert os.path.dirname(os.path.abspath(m.__file__)) == os.path.abspath(ref_dir), (n, m.__file__)
    finally:
        sys.path.remove(ref_dir)
        for n in names:
            sys.modules.pop(n, None)
            if saved[n] is not None:
                sys.modules[n] = saved[n]
    return types.SimpleNamespace(**mods)


class ReferenceTrainer:
    """WORKER.set_cnn_models / train_generator / train_discriminator / ema_update (worker.py:75-112, 137-217) and the
    requires_grad / freezeD toggling of loader.py:44-53, on the reference's modules, device-agnostic, without DDP."""

    def __init__(self, ref, args, batch, device="cpu", seed=0):
        self.ref, self.args, self.b, self.dev = ref, args, batch, torch.device(device)
        torch.manual_seed(seed)
        self.G = ref.cnn.Generator(args).to(self.dev)
        self.D = ref.cnn.Discriminator(args).to(self.dev)
        self.g_opt = torch.optim.Adam(list(self.G.parameters()), lr=args.g_lr, betas=(args.beta1, args.beta2), eps=1e-8)
        self.d_opt = torch.optim.Adam(list(self.D.parameters()), lr=args.d_lr, betas=(args.beta1, args.beta2), eps=1e-8)
        self.G_ema = copy.deepcopy(self.G)
        self.ema = ref.ema.Ema(self.G, self.G_ema, args.g_ema_decay, args.g_ema_start)
        self.gen = torch.Generator().manual_seed(1000 + seed)

    @staticmethod
    def _requires_grad(model, flag):
        for p in model.parameters():
            p.requires_grad = flag

    def _randn(self, d):
        return torch.randn(self.b, d, generator=self.gen).to(self.dev)

    def _images(self):
        r = self.args.img_resolution
        return [(torch.rand(self.b, 3, r, r, generator=self.gen) * 2 - 1).to(self.dev) for _ in range(3)]

    def g_step(self, epoch):
        a, L = self.args, self.ref.loss
        self._requires_grad(self.G, True); self._requires_grad(self.D, False)
        self.g_opt.zero_grad()
        rand1, rand2 = self._randn(a.geo_noise_dim), self._randn(a.app_noise_dim)
        resample1, resample2 = self._randn(a.geo_noise_dim), self._randn(a.app_noise_dim)
        ones = torch.ones(self.b, 1, device=self.dev)
        if epoch % 2 == 1:
            logit, _, _ = self.D(self.G(rand1, rand2), False)
            g_loss = F.binary_cross_entropy_with_logits(logit, ones)
        else:
            anchor, re_geo, re_app = self.G(rand1, rand2), self.G(resample1, rand2), self.G(rand1, resample2)
            logit, gf, af = self.D(anchor, True)
            _, gp, an = self.D(re_geo, True)
            _, gn, ap = self.D(re_app, True)
            aug = (L.contrastive_loss(gf, gp, gn, a.tau) + L.contrastive_loss(af, ap, an, a.tau)) * a.l_aux
            d1 = self.G.geometry_mapping.diagonal_params.view(-1)
            d2 = self.G.appearance_mapping.diagonal_params.view(-1)
            g_loss = F.binary_cross_entropy_with_logits(logit, ones) + aug + torch.norm(torch.cat([d1, d2]), p=1) * a.l_s
        g_loss.backward()
        self.g_opt.step()
        self.ema.update(epoch)
        return g_loss.item()

    def d_step(self, epoch):
        a, L = self.args, self.ref.loss
        self._requires_grad(self.G, False); self._requires_grad(self.D, True)
        if epoch >= a.freezeD_start:
            for i, (_n, layer) in enumerate(self.D.shared_model.named_children()):
                if i < a.freezeD_layer + 2:
                    for p in layer.parameters():
                        p.requires_grad = False
        self.d_opt.zero_grad()
        image, geo, app = self._images()
        rand1, rand2 = self._randn(a.geo_noise_dim), self._randn(a.app_noise_dim)
        fake_logit, _, _ = self.D(self.G(rand1, rand2), False)
        ones, zeros = torch.ones(self.b, 1, device=self.dev), torch.zeros(self.b, 1, device=self.dev)
        if epoch % 2 == 1:
            image.requires_grad_(True)
            real_logit, _, _ = self.D(image, False)
            d_loss = F.binary_cross_entropy_with_logits(real_logit, ones) + F.binary_cross_entropy_with_logits(fake_logit, zeros)
            if epoch % 8 == 1:
                d_loss = d_loss + L.cal_r1_reg(real_logit, image, self.dev) * a.l_r1
        else:
            real_logit, gf, af = self.D(image, True)
            _, gp, an = self.D(geo, True)
            _, gn, ap = self.D(app, True)
            d_loss = F.binary_cross_entropy_with_logits(real_logit, ones) + F.binary_cross_entropy_with_logits(fake_logit, zeros) \
                + (L.contrastive_loss(gf, gp, gn, a.tau) + L.contrastive_loss(af, ap, an, a.tau)) * a.l_aux
        d_loss.backward()
        self.d_opt.step()
        return d_loss.item()

    def half_step(self, k):
        """k-th half-iteration of the reference schedule: k = 2*epoch (G step + EMA) or 2*epoch + 1 (D step)."""
        return self.g_step(k // 2) if k % 2 == 0 else self.d_step(k // 2)


def make_args(res, lr, freeze_d_start=10 ** 9, freeze_d_layer=5):
    """main.py:12-61 defaults with the README recipe's learning rate."""
    return types.SimpleNamespace(
        img_resolution=res, geo_noise_dim=64, app_noise_dim=64, geo_projection_dim=256, app_projection_dim=256,
        geo_latent_dim=64, app_latent_dim=512, max_flow_scale=0.1, tau=0.05, l_aux=0.5, l_r1=10.0, l_s=1e-7,
        g_lr=lr, d_lr=lr, beta1=0.0, beta2=0.99, g_ema_decay=0.9999, g_ema_start=0,
        freezeD_start=freeze_d_start, freezeD_layer=freeze_d_layer)


class PortTrainer:
    """Same interface over the oracle port, used when no reference checkout is available."""

    def __init__(self, res, lr, batch, seed=0):
        from oracle import lcgan_oracle as O
        self.O, self.b = O, batch
        self.cfg, hp = O.Config(img_resolution=res), O.Hyper(lr=lr)
        self.tr = O.OracleTrainer(self.cfg, hp, O.make_generator_state(self.cfg, seed), O.make_discriminator_state(self.cfg, seed + 1))
        self.gen = torch.Generator().manual_seed(1000 + seed)

    def half_step(self, k):
        O, it = self.O, k // 2
        z = O.synthetic_latents(self.b, self.cfg, self.gen)
        if k % 2 == 0:
            loss = self.tr.g_step(it, z)
            self.tr.ema_step(it)
            return loss
        return self.tr.d_step(it, z, O.synthetic_data(self.b, self.cfg, self.gen))


def make_trainer(res, lr, batch, threads=None):
    """(trainer, kind, description of what runs)"""
    torch.set_num_threads(threads or os.cpu_count() or 1)
    ref_dir = find_checkout()
    if ref_dir is not None:
        ref = load_modules(ref_dir)
        return (ReferenceTrainer(ref, make_args(res, lr), batch), "reference",
                f"unmodified reference modules ({os.path.relpath(ref_dir, ROOT) if ref_dir.startswith(ROOT) else ref_dir}: "
                "cnn.py, custom_layers.py, loss.py, ema.py) stepped by the restated worker.py:137-217 schedule")
    return PortTrainer(res, lr, batch), "port", "oracle port (oracle/lcgan_oracle.py; no reference checkout on this box)"


def time_half_steps(trainer, first, count):
    """Wall-clock seconds of half-steps first .. first+count-1 (each ends with .item(), like worker.py:177,214)."""
    out = []
    for k in range(first, first + count):
        t0 = time.perf_counter()
        trainer.half_step(k)
        out.append(time.perf_counter() - t0)
    return out
