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
from . import ops
from .ema import Ema
from .optim import FusedAdam


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

    def __init__(self, G, D, hp, ema_decay=0.9999, ema_start=0, freeze_d_start=10 ** 9, freeze_d_layer=5,
                 fused_adam=True):
        self.G, self.D, self.hp = G, D, hp
        # worker.py:98-110; fused_adam=False keeps torch.optim.Adam (what the reference's worker.py builds)
        adam = FusedAdam if fused_adam else torch.optim.Adam
        self.g_opt = adam(list(G.parameters()), lr=hp.lr, betas=(hp.beta1, hp.beta2), eps=1e-8)
        self.d_opt = adam(list(D.parameters()), lr=hp.lr, betas=(hp.beta1, hp.beta2), eps=1e-8)
        self.G_ema = copy.deepcopy(_bare(G))
        self.ema = Ema(_bare(G), self.G_ema, ema_decay, ema_start)
        self.freeze_d_start, self.freeze_d_layer = freeze_d_start, freeze_d_layer

    def g_step(self, it, z):
        requires_grad(self.G, True); requires_grad(self.D, False)
        ops.prepack(self.G, self.D)                    # stale weight packs / demod tables, in bulk
        self.g_opt.zero_grad()
        loss = generator_loss(self.G, self.D, self.hp, it, z)
        loss.backward()
        self.g_opt.step()
        return loss

    def d_step(self, it, z, data):
        requires_grad(self.G, False); requires_grad(self.D, True)
        if it >= self.freeze_d_start:
            freeze_discriminator(self.D, self.freeze_d_layer)
        ops.prepack(self.G, self.D)
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


class GraphedTrainer(Trainer):
    """The same iteration with every step variant captured once into a CUDA graph and replayed:
    the ~1900 kernel launches of an iteration are then issued by the driver, not by Python
    (B200 guide: "capture launch-bound inner loops in CUDA graphs").  Five graphs cover the
    reference schedule: G even / G odd (+EMA), D even / D odd / D odd with R1.  Inputs are static
    device buffers that the caller fills before each replay; losses are left in device buffers.
    requires_grad / freezeD flags are frozen at capture time; the EMA decay is read from a device
    scalar, so the ema.py:19-23 start_iter schedule is followed across replays.

    world > 1: data parallel without the DDP wrapper (DDP's reducer is not capturable).  Rank 0's
    initial weights are broadcast (what DDP's constructor does, worker.py:88-96); every parameter's
    post-accumulate-grad hook copies the gradient into its slot of a flat bucket, and the moment a
    bucket is complete its NCCL all-reduce (average) is issued on a side stream, so the exchange
    overlaps the rest of backward (worker.py:88-96: DDP's bucketed overlap) - all captured in the graphs.
    Parameters whose gradient stays None (unused heads on odd iterations, frozen layers) are skipped
    by Adam exactly as under DDP(find_unused_parameters=True)  (dist_utils.GradExchange).
    """

    VARIANTS = ("g_even", "g_odd", "d_even", "d_odd", "d_r1")
    BUCKET_BYTES = 48 << 20

    def __init__(self, G, D, hp, batch, device, world=1, **kw):
        super().__init__(G, D, hp, **kw)
        self.world = world
        if world > 1:
            import torch.distributed as dist
            from .dist_utils import GradExchange
            for t in list(G.parameters()) + list(G.buffers()) + list(D.parameters()) + list(D.buffers()):
                dist.broadcast(t.data, 0)
            self.G_ema.load_state_dict(_bare(G).state_dict())
            self.exchange = GradExchange({"g": G, "d": D}, device, world, self.BUCKET_BYTES)
        for opt in (self.g_opt, self.d_opt):          # torch Adam: state on device so step() is capturable
            if not isinstance(opt, FusedAdam):
                for grp in opt.param_groups:
                    grp["capturable"] = True
        g = _bare(G)
        res = g.img_resolution
        dims = {"rand1": g.geo_noise_dim, "rand2": g.app_noise_dim, "resample1": g.geo_noise_dim,
                "resample2": g.app_noise_dim}
        self.z = {k: torch.zeros(batch, d, device=device) for k, d in dims.items()}
        self.zd = {k: torch.zeros(batch, dims[k], device=device) for k in ("rand1", "rand2")}
        self.data = {k: torch.zeros(batch, 3, res, res, device=device)
                     for k in ("image", "geometry_change", "appearance_change")}
        self.g_loss = torch.zeros((), device=device)
        self.d_loss = torch.zeros((), device=device)
        self.ema_decay = torch.full((), float(self.ema.decay_at(0)), device=device)
        self.graphs = {}
        self.launches = {}

    @staticmethod
    def variant(it, which):
        if which == "g":
            return "g_even" if it % 2 == 0 else "g_odd"
        return "d_even" if it % 2 == 0 else ("d_r1" if it % 8 == 1 else "d_odd")

    def g_step(self, it, z):
        if self.world == 1:
            return super().g_step(it, z)
        requires_grad(self.G, True); requires_grad(self.D, False)
        ops.prepack(self.G, self.D)                    # stale weight packs / demod tables, in bulk
        self.g_opt.zero_grad()
        loss = generator_loss(self.G, self.D, self.hp, it, z)
        self.exchange.backward(loss, self.variant(it, "g"), "g")
        self.g_opt.step()
        return loss

    def d_step(self, it, z, data):
        if self.world == 1:
            return super().d_step(it, z, data)
        requires_grad(self.G, False); requires_grad(self.D, True)
        if it >= self.freeze_d_start:
            freeze_discriminator(self.D, self.freeze_d_layer)
        ops.prepack(self.G, self.D)
        self.d_opt.zero_grad()
        loss = discriminator_loss(self.G, self.D, self.hp, it, z, data)
        frozen = "_frozen" if it >= self.freeze_d_start else ""      # a different set of gradients
        self.exchange.backward(loss, self.variant(it, "d") + frozen, "d")
        self.d_opt.step()
        return loss

    # ---- capture / replay -------------------------------------------------------------------------
    _VARIANT_IT = {"g_even": 0, "g_odd": 3, "d_even": 0, "d_odd": 3, "d_r1": 1}

    def _run(self, name):
        it = self._VARIANT_IT[name]
        if name.startswith("g"):
            loss = self.g_step(it, self.z)
            self.ema.update(it, self.ema_decay)
            self.g_loss.copy_(loss.detach())
        else:
            loss = self.d_step(it, self.zd, self.data)
            self.d_loss.copy_(loss.detach())

    def capture(self, warmup=3):
        from . import _lib
        dev = self.g_loss.device
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):                   # eager warm-up: lazy inits, Adam state, bucket plans
            for _ in range(max(warmup, 2)):
                for name in self.VARIANTS:
                    self._run(name)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        pool = None
        for name in self.VARIANTS:
            ops.clear_pack_cache()
            g = torch.cuda.CUDAGraph()
            n0 = _lib.launches
            with torch.cuda.graph(g, pool=pool):
                self._run(name)
            self.launches[name] = _lib.launches - n0
            pool = g.pool()
            self.graphs[name] = g
        ops.clear_pack_cache()
        torch.cuda.synchronize(dev)

    def reset_optimizer_state(self):
        """Zero Adam moments / step counters in place (the graphs hold their addresses)."""
        for opt in (self.g_opt, self.d_opt):
            for st in opt.state.values():
                for v in st.values():
                    if torch.is_tensor(v):
                        v.zero_()

    def iteration_graphed(self, it):
        """One iteration from the static input buffers; returns the number of our kernels replayed."""
        return self.replay_g(it) + self.replay_d(it)

    def replay_g(self, it):
        v = self.variant(it, "g")
        self.ema_decay.fill_(float(self.ema.decay_at(it)))
        self.graphs[v].replay()
        ops.bump_generation()          # the replay stepped the optimizer and the EMA behind autograd's back
        return self.launches[v]

    def replay_d(self, it):
        v = self.variant(it, "d")
        self.graphs[v].replay()
        ops.bump_generation()
        return self.launches[v]
