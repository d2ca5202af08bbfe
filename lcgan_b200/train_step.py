"""The reference training iteration (loader.py:44-54, worker.py:127-217) restated around
nn.Modules, without the dataset / checkpoint / logging code: G step, EMA, D step, with the loss
schedule the reference runs (aux on even iterations, R1 on iteration % 8 == 1, l_s on even G
steps).  bench.py and the parity tests drive the hot path through this; the reference's own
worker.py drives it the same way through the drop-in modules.
"""
from __future__ import annotations

import copy

import torch
import torch.nn.functional as F

from . import loss as L
from .ema import Ema


def requires_grad(model, flag=True):
    """worker.py:133-135"""
    for p in model.parameters():
        p.requires_grad = flag


def _bare(m):
    return m.module if hasattr(m, "module") else m


def freeze_discriminator(discriminator, freeze_up_to_index=5):
    """worker.py:127-131: first freeze_up_to_index+2 children of D.shared_model."""
    for i, (_name, layer) in enumerate(_bare(discriminator).shared_model.named_children()):
        if i < freeze_up_to_index + 2:
            for p in layer.parameters():
                p.requires_grad = False


def generator_loss(G, D, hp, it, z):
    """worker.py:187-210.  z: dict rand1, rand2, resample1, resample2."""
    b = z["rand1"].shape[0]
    ones = torch.ones(b, 1, device=z["rand1"].device)
    if it % 2 == 1:
        logit, _, _ = D(G(z["rand1"], z["rand2"]), False)
        return F.binary_cross_entropy_with_logits(logit, ones)
    anchor = G(z["rand1"], z["rand2"])
    re_geo = G(z["resample1"], z["rand2"])
    re_app = G(z["rand1"], z["resample2"])
    logit, gf, af = D(anchor, True)
    _, gp, an = D(re_geo, True)
    _, gn, ap = D(re_app, True)
    aux = (L.contrastive_loss(gf, gp, gn, hp.tau) + L.contrastive_loss(af, ap, an, hp.tau)) * hp.l_aux
    d1 = _bare(G).geometry_mapping.diagonal_params.view(-1)
    d2 = _bare(G).appearance_mapping.diagonal_params.view(-1)
    sparsity = torch.norm(torch.cat([d1, d2]), p=1) * hp.l_s
    return F.binary_cross_entropy_with_logits(logit, ones) + aux + sparsity


def discriminator_loss(G, D, hp, it, z, data):
    """worker.py:145-173.  data: dict image, geometry_change, appearance_change."""
    b = z["rand1"].shape[0]
    dev = z["rand1"].device
    ones, zeros = torch.ones(b, 1, device=dev), torch.zeros(b, 1, device=dev)
    fake_logit, _, _ = D(G(z["rand1"], z["rand2"]), False)
    if it % 2 == 1:
        image = data["image"].detach().requires_grad_(True)
        real_logit, _, _ = D(image, False)
        loss = F.binary_cross_entropy_with_logits(real_logit, ones) \
            + F.binary_cross_entropy_with_logits(fake_logit, zeros)
        if it % 8 == 1:
            loss = loss + L.cal_r1_reg(real_logit, image, dev) * hp.l_r1
        return loss
    real_logit, gf, af = D(data["image"], True)
    _, gp, an = D(data["geometry_change"], True)
    _, gn, ap = D(data["appearance_change"], True)
    aux = (L.contrastive_loss(gf, gp, gn, hp.tau) + L.contrastive_loss(af, ap, an, hp.tau)) * hp.l_aux
    return F.binary_cross_entropy_with_logits(real_logit, ones) \
        + F.binary_cross_entropy_with_logits(fake_logit, zeros) + aux


class Trainer:
    """G, D (optionally DDP-wrapped), EMA copy and the two Adam optimizers (worker.py:75-112)."""

    def __init__(self, G, D, hp, ema_decay=0.9999, ema_start=0, freeze_d_start=10 ** 9, freeze_d_layer=5):
        self.G, self.D, self.hp = G, D, hp
        self.g_opt = torch.optim.Adam(list(G.parameters()), lr=hp.lr, betas=(hp.beta1, hp.beta2), eps=1e-8)
        self.d_opt = torch.optim.Adam(list(D.parameters()), lr=hp.lr, betas=(hp.beta1, hp.beta2), eps=1e-8)
        self.G_ema = copy.deepcopy(_bare(G))
        self.ema = Ema(_bare(G), self.G_ema, ema_decay, ema_start)
        self.freeze_d_start, self.freeze_d_layer = freeze_d_start, freeze_d_layer

    def g_step(self, it, z):
        requires_grad(self.G, True); requires_grad(self.D, False)
        self.g_opt.zero_grad()
        loss = generator_loss(self.G, self.D, self.hp, it, z)
        loss.backward()
        self.g_opt.step()
        return loss

    def d_step(self, it, z, data):
        requires_grad(self.G, False); requires_grad(self.D, True)
        if it >= self.freeze_d_start:
            freeze_discriminator(self.D, self.freeze_d_layer)
        self.d_opt.zero_grad()
        loss = discriminator_loss(self.G, self.D, self.hp, it, z, data)
        loss.backward()
        self.d_opt.step()
        return loss

    def iteration(self, it, zg, zd, data):
        """One reference iteration; returns (g_loss, d_loss) as python floats (the reference
        calls .item() on both every iteration: worker.py:177,214)."""
        g = self.g_step(it, zg).item()
        self.ema.update(it)
        d = self.d_step(it, zd, data).item()
        return g, d
