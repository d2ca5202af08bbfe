"""CPU oracle for the LC-GAN training hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this file.  The product path (``lcgan_b200``) never
routes through it.

What it is: a *functional* restatement, in plain PyTorch ops on whatever device/dtype the
parameters live on (fp32 CPU by default, fp64 for a tighter truth), of the arithmetic in the
reference's ``cnn.py`` / ``custom_layers.py`` / ``loss.py`` and of the step schedule in
``worker.py:137-214``.  It works on a flat ``state_dict`` (the reference's own key names), so the
same dict can be loaded into the reference modules (to pin this oracle, see
``oracle/make_golden.py``) and into the ``lcgan_b200`` modules (to check the CUDA path).

Pinning: the reference ships no tests or golden vectors, so this oracle is
pinned against outputs of the unmodified reference modules run in the build container:
``oracle/make_golden.py`` imports ``/root/reference`` and stores its outputs for seeded inputs in
``tests/golden/*.pt``; ``tests/test_oracle_golden.py`` replays them through this file.

The modulated convolution is restated in its shared-weight form (scale activations by the style,
convolve with the shared weight, scale by the demodulation coefficient) instead of the reference's
per-sample grouped convolution; the golden vectors prove the two agree.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
State = Dict[str, Tensor]

SQRT2 = math.sqrt(2.0)
SQRT_HALF = math.sqrt(0.5)


# --------------------------------------------------------------------------------------
# configuration / channel schedule   (cnn.py:10-17, 49-54, 79-84)
# --------------------------------------------------------------------------------------
class Config:
    """Hyper-parameters read from ``args`` by the reference constructors (main.py:25-31)."""

    def __init__(self, img_resolution=256, geo_noise_dim=64, app_noise_dim=64, geo_latent_dim=64,
                 app_latent_dim=512, geo_projection_dim=256, app_projection_dim=256,
                 max_flow_scale=0.1):
        self.img_resolution = img_resolution
        self.geo_noise_dim = geo_noise_dim
        self.app_noise_dim = app_noise_dim
        self.geo_latent_dim = geo_latent_dim
        self.app_latent_dim = app_latent_dim
        self.geo_projection_dim = geo_projection_dim
        self.app_projection_dim = app_projection_dim
        self.max_flow_scale = max_flow_scale

    @property
    def num_blocks(self):                    # cnn.py:13, 52
        return int(math.log2(self.img_resolution)) - 2

    @property
    def base_nf(self):                       # cnn.py:17, 54
        return {1024: 32, 512: 64}.get(self.img_resolution, 128)

    def g_channels(self):
        """[(in, out, out_resolution)] per synthesis block (cnn.py:79-84)."""
        chans, cin = [], 512
        for i in range(self.num_blocks):
            cout = min(self.base_nf * 2 ** (self.num_blocks - i - 1), 512)
            chans.append((cin, cout, 2 ** (3 + i)))
            cin = cout
        return chans

    def d_channels(self):
        """[(in, out)] per discriminator block (cnn.py:22-25)."""
        return [(min(self.base_nf * 2 ** i, 512), min(self.base_nf * 2 ** (i + 1), 512))
                for i in range(self.num_blocks)]

    def namespace(self):
        import types
        return types.SimpleNamespace(**self.__dict__)


# --------------------------------------------------------------------------------------
# deterministic parameter construction (same distributions as the reference constructors)
# --------------------------------------------------------------------------------------
def _randn(gen, *shape):
    return torch.randn(*shape, generator=gen, dtype=torch.float32)


def _add_linear(sd, gen, key, fin, fout, bias=0.0, lr_mul=1.0):
    # EqualizedLinear / EqualizedWeight init: custom_layers.py:11, 21
    sd[key + ".bias"] = torch.full((fout,), float(bias))
    sd[key + ".weight.weight"] = _randn(gen, fout, fin) / lr_mul


def _add_conv(sd, gen, key, cin, cout, k, bias=True):
    # EqualizedConv2d / ModulatedConv2d init: custom_layers.py:32-35, 56-57
    if bias:
        sd[key + ".bias"] = torch.zeros(cout)
    sd[key + ".weight.weight"] = _randn(gen, cout, cin, k, k)


def _add_synth_layer(sd, gen, key, cin, cout, latent, k=3):
    # SynthesisLayer: custom_layers.py:96-97 (use_noise is False everywhere, cnn.py:83,87)
    _add_linear(sd, gen, key + ".linear", latent, cin, bias=1.0)
    _add_conv(sd, gen, key + ".modulated_conv", cin, cout, k)


def make_generator_state(cfg: Config, seed: int = 0) -> State:
    """Random-init generator parameters + buffers under the reference's state_dict keys."""
    gen = torch.Generator().manual_seed(seed)
    sd: State = {}
    sd["const"] = _randn(gen, 512, 4, 4)                                    # cnn.py:76
    sd["avg_latent1"] = torch.zeros(cfg.geo_latent_dim)                    # cnn.py:63-64
    sd["avg_latent2"] = torch.zeros(cfg.app_latent_dim)
    g, a = cfg.geo_latent_dim, cfg.app_latent_dim
    geo = [cfg.geo_noise_dim] + [g] * 12                                    # cnn.py:66-68
    app = [cfg.app_noise_dim, a // 4, a // 2] + [a] * 10                    # cnn.py:70-72
    for name, chans in (("geometry_mapping", geo), ("appearance_mapping", app)):
        m = chans[0]
        sd[name + ".diagonal_params"] = _randn(gen, m)                      # custom_layers.py:264-265
        sd[name + ".basis_params"] = _randn(gen, m, m)
        for i in range(12):
            _add_linear(sd, gen, f"{name}.mlp.{i}", chans[i], chans[i + 1], lr_mul=0.01)
    for i, (cin, cout, _res) in enumerate(cfg.g_channels()):
        p = f"model.{i}"
        _add_synth_layer(sd, gen, p + ".modulated_conv0", cin, cout, a)
        _add_synth_layer(sd, gen, p + ".modulated_conv1", cout, cout, a)
        _add_conv(sd, gen, p + ".skip_layer", cin, cout, 1, bias=False)
        _add_synth_layer(sd, gen, p + ".flow_layer", cin, 2, g)
    c = cfg.g_channels()[-1][1]
    _add_synth_layer(sd, gen, "rgb_layer.modulated_conv0", c, c, a)
    _add_synth_layer(sd, gen, "rgb_layer.modulated_conv1", c, 3, a, k=1)
    return sd


def make_discriminator_state(cfg: Config, seed: int = 1) -> State:
    gen = torch.Generator().manual_seed(seed)
    sd: State = {}
    _add_conv(sd, gen, "shared_model.0", 3, cfg.base_nf, 1)                 # cnn.py:20
    for i, (cin, cout) in enumerate(cfg.d_channels()):
        p = f"shared_model.{i + 2}"
        _add_conv(sd, gen, p + ".conv0", cin, cin, 3)
        _add_conv(sd, gen, p + ".conv1", cin, cout, 3)
        _add_conv(sd, gen, p + ".skip_layer", cin, cout, 1, bias=False)
    c = cfg.d_channels()[-1][1]
    _add_conv(sd, gen, "discriminator_epilogue.conv", c + 1, c, 3)
    _add_linear(sd, gen, "discriminator_epilogue.linear", c * 16, c, lr_mul=0.01)
    _add_linear(sd, gen, "logit_mapper.mlp.0", c, 1, lr_mul=0.01)
    for h, dim in (("projection_header1", cfg.geo_projection_dim),
                   ("projection_header2", cfg.app_projection_dim)):
        for j, (fi, fo) in enumerate(((c * 16, c * 4), (c * 4, c), (c, dim))):
            _add_linear(sd, gen, f"{h}.mlp.{2 * j}", fi, fo, lr_mul=0.01)
    return sd


# --------------------------------------------------------------------------------------
# layer arithmetic
# --------------------------------------------------------------------------------------
def _scaled_weight(w: Tensor, lr_mul: float) -> Tensor:
    """EqualizedWeight.forward (custom_layers.py:10,14): W * lr_mul / sqrt(fan_in)."""
    return w * (lr_mul / math.sqrt(w[0].numel()))


def eq_linear(sd: State, key: str, x: Tensor, lr_mul: float = 1.0) -> Tensor:
    """EqualizedLinear.forward (custom_layers.py:24-25)."""
    return x @ _scaled_weight(sd[key + ".weight.weight"], lr_mul).t() + sd[key + ".bias"] * lr_mul


def eq_conv(sd: State, key: str, x: Tensor, stride: int = 1) -> Tensor:
    """EqualizedConv2d.forward (custom_layers.py:39-44); lr_mul is 1.0 at every call site."""
    w = _scaled_weight(sd[key + ".weight.weight"], 1.0)
    return F.conv2d(x, w, sd.get(key + ".bias"), stride=stride, padding=w.shape[-1] // 2)


def box3(x: Tensor) -> Tensor:
    """box_filter (custom_layers.py:136-138, 196-198): 3x3 mean, zero padded, always /9."""
    c = x.shape[1]
    k = torch.full((c, 1, 3, 3), 1.0 / 9.0, dtype=x.dtype, device=x.device)
    return F.conv2d(x, k, padding=1, groups=c)


def lrelu(x: Tensor, gain: float = 1.0) -> Tensor:
    return torch.where(x > 0, x, x * 0.2) * gain


def mod_conv(sd: State, key: str, x: Tensor, s: Tensor, up: int = 1, eps: float = 1e-8) -> Tensor:
    """ModulatedConv2d.forward (custom_layers.py:60-86) in shared-weight form.

    y[b,o] = d[b,o] * sum_{c,k} w[o,c,k] * (s[b,c] x[b,c,p+k]) + bias[o],
    d[b,o] = rsqrt(sum_{c,k} (w[o,c,k] s[b,c])^2 + eps).
    up=2 is conv_transpose2d(stride 2, padding 1, output_padding 1) with the *unflipped* weight.
    """
    w = _scaled_weight(sd[key + ".weight.weight"], 1.0)           # [O, I, k, k]
    wsq = w.square().sum(dim=(2, 3))                               # [O, I]
    d = torch.rsqrt(s.square() @ wsq.t() + eps)                    # [b, O]
    xs = x * s[:, :, None, None]
    pad = (w.shape[-1] - 1) // 2
    if up > 1:
        y = F.conv_transpose2d(xs, w.transpose(0, 1), stride=up, padding=pad, output_padding=1)
    else:
        y = F.conv2d(xs, w, padding=pad)
    return y * d[:, :, None, None] + sd[key + ".bias"][None, :, None, None]


def synth_layer(sd: State, key: str, x: Tensor, latent: Tensor, up: int = 1) -> Tensor:
    """SynthesisLayer.forward (custom_layers.py:103-111), use_noise=False."""
    s = eq_linear(sd, key + ".linear", latent)
    return mod_conv(sd, key + ".modulated_conv", x, s, up=up)


def _cubic_w(t: Tensor):
    """Bicubic convolution coefficients, A=-0.75 (ATen UpSample.h get_cubic_upsample_coefficients)."""
    A = -0.75
    def near(u):  # |u| <= 1
        return ((A + 2) * u - (A + 3)) * u * u + 1
    def far(u):   # 1 < |u| < 2
        return ((A * u - 5 * A) * u + 8 * A) * u - 4 * A
    return far(t + 1), near(t), near(1 - t), far(2 - t)


def bicubic_warp(x: Tensor, grid: Tensor) -> Tensor:
    """F.grid_sample(x, grid, mode='bicubic') with the defaults the reference relies on
    (custom_layers.py:165): zeros padding, align_corners=False.  Written out explicitly so the
    oracle documents the exact sampling rule the CUDA kernel must follow.
    x [b,C,H,W]; grid [b,H,W,2] (x,y) in [-1,1]."""
    b, c, h, w = x.shape
    ix = ((grid[..., 0] + 1) * w - 1) / 2
    iy = ((grid[..., 1] + 1) * h - 1) / 2
    x0, y0 = torch.floor(ix), torch.floor(iy)
    wx, wy = _cubic_w(ix - x0), _cubic_w(iy - y0)
    flat = x.reshape(b, c, h * w)
    out = torch.zeros(b, c, grid.shape[1], grid.shape[2], dtype=x.dtype, device=x.device)
    for j in range(4):
        yy = (y0 + (j - 1)).long()
        for i in range(4):
            xx = (x0 + (i - 1)).long()
            ok = ((xx >= 0) & (xx < w) & (yy >= 0) & (yy < h)).to(x.dtype)
            idx = (yy.clamp(0, h - 1) * w + xx.clamp(0, w - 1)).reshape(b, 1, -1).expand(b, c, -1)
            v = flat.gather(2, idx).reshape(out.shape)
            out = out + v * (wx[i] * wy[j] * ok)[:, None]
    return out


def base_coordinates(h: int, w: int, like: Tensor) -> Tensor:
    """get_coordinates (custom_layers.py:127-134): linspace(-1,1) grid, [1,2,h,w] (x then y)."""
    ys = 2 * torch.arange(h, dtype=like.dtype, device=like.device) / (h - 1) - 1
    xs = 2 * torch.arange(w, dtype=like.dtype, device=like.device) / (w - 1) - 1
    gy, gx = torch.meshgrid(ys, xs, indexing="ij")
    return torch.stack((gx, gy))[None]


def synthesis_block(sd: State, key: str, x: Tensor, g_lat: Tensor, a_lat: Tensor,
                    max_flow_scale: float, explicit_warp: bool = False,
                    a_lat1: Optional[Tensor] = None) -> Tensor:
    """SynthesisBlock.forward (custom_layers.py:140-166).  a_lat1: latent of the second
    modulated conv when it differs from the first (the Generator passes the same code twice)."""
    a_lat1 = a_lat if a_lat1 is None else a_lat1
    skip = eq_conv(sd, key + ".skip_layer", x) * SQRT_HALF
    skip = box3(F.interpolate(skip, scale_factor=2, mode="nearest"))
    flow = torch.tanh(box3(synth_layer(sd, key + ".flow_layer", x, g_lat, up=2)))
    y = lrelu(box3(synth_layer(sd, key + ".modulated_conv0", x, a_lat, up=2)), SQRT2)
    y = lrelu(synth_layer(sd, key + ".modulated_conv1", y, a_lat1))
    y = skip + y
    grid = (base_coordinates(y.shape[2], y.shape[3], y) + flow * max_flow_scale).permute(0, 2, 3, 1)
    if explicit_warp:
        return bicubic_warp(y, grid)
    return F.grid_sample(y, grid, mode="bicubic", padding_mode="zeros", align_corners=False)


def mapping_matrix(sd: State, key: str) -> Tensor:
    """L = Q(tanh(basis)) . diag(|d| + 1e-6)   (custom_layers.py:274-281)."""
    q, _ = torch.linalg.qr(torch.tanh(sd[key + ".basis_params"]))
    return q * (sd[key + ".diagonal_params"].abs() + 1e-6)[None, :]


def mapping_network(sd: State, key: str, z: Tensor) -> Tensor:
    """MappingNetwork.forward (custom_layers.py:278-287): x = L z, then 12 linear layers."""
    x = z @ mapping_matrix(sd, key).t()
    for i in range(12):
        x = eq_linear(sd, f"{key}.mlp.{i}", x, lr_mul=0.01)
    return x


def generator_forward(sd: State, cfg: Config, z_geo: Tensor, z_app: Tensor, w_psi: float = -1.0,
                      update_avg: bool = True, explicit_warp: bool = False) -> Tensor:
    """Generator.forward (cnn.py:89-114).  Mutates sd['avg_latent*'] like the reference."""
    g = mapping_network(sd, "geometry_mapping", z_geo)
    a = mapping_network(sd, "appearance_mapping", z_app)
    if w_psi <= 0 and update_avg:                                           # cnn.py:95-97
        sd["avg_latent1"] = g.detach().mean(0).lerp(sd["avg_latent1"], 0.998)
        sd["avg_latent2"] = a.detach().mean(0).lerp(sd["avg_latent2"], 0.998)
    if w_psi > 0:                                                           # cnn.py:99-101
        g = sd["avg_latent1"].lerp(g, w_psi)
        a = sd["avg_latent2"].lerp(a, w_psi)
    x = sd["const"][None].expand(z_geo.shape[0], -1, -1, -1)
    for i in range(cfg.num_blocks):
        x = synthesis_block(sd, f"model.{i}", x, g, a, cfg.max_flow_scale, explicit_warp)
    x = lrelu(synth_layer(sd, "rgb_layer.modulated_conv0", x, a))          # custom_layers.py:177-182
    return synth_layer(sd, "rgb_layer.modulated_conv1", x, a)


def minibatch_std(x: Tensor, group_size: int = 8) -> Tensor:
    """MinibatchStdLayer.forward (custom_layers.py:243-256), num_channels=1."""
    n, c, h, w = x.shape
    g = min(group_size, n)
    y = x.reshape(g, n // g, c, h, w)
    y = (y - y.mean(0)).square().mean(0)
    y = (y + 1e-8).sqrt().mean(dim=(1, 2, 3))                               # [n/g]
    y = y.reshape(-1, 1, 1, 1).repeat(g, 1, h, w)
    return torch.cat([x, y], dim=1)


def discriminator_block(sd: State, key: str, x: Tensor) -> Tensor:
    """DiscriminatorBlock.forward, skip=True (custom_layers.py:200-209)."""
    skip = eq_conv(sd, key + ".skip_layer", F.avg_pool2d(x, 2)) * SQRT_HALF
    y = box3(lrelu(eq_conv(sd, key + ".conv0", x), SQRT2))
    y = lrelu(eq_conv(sd, key + ".conv1", y, stride=2))
    return skip + y


def projection_head(sd: State, key: str, x: Tensor, n_layers: int) -> Tensor:
    """ProjectionHead.forward (custom_layers.py:290-306): LeakyReLU between, not after."""
    for j in range(n_layers):
        x = eq_linear(sd, f"{key}.mlp.{2 * j}", x, lr_mul=0.01)
        if j < n_layers - 1:
            x = lrelu(x)
    return x


def discriminator_forward(sd: State, cfg: Config, image: Tensor, get_embedding_features=False):
    """Discriminator.forward (cnn.py:33-43) -> (logit, geo_emb|None, app_emb|None)."""
    h = lrelu(eq_conv(sd, "shared_model.0", image))
    for i in range(cfg.num_blocks):
        h = discriminator_block(sd, f"shared_model.{i + 2}", h)
    e = lrelu(eq_conv(sd, "discriminator_epilogue.conv", minibatch_std(h, 8)))   # custom_layers.py:228-234
    e = lrelu(eq_linear(sd, "discriminator_epilogue.linear", e.flatten(1), lr_mul=0.01))
    logit = projection_head(sd, "logit_mapper", e, 1)
    geo = app = None
    if get_embedding_features:
        f = h.flatten(1)
        geo = F.normalize(projection_head(sd, "projection_header1", f, 3))
        app = F.normalize(projection_head(sd, "projection_header2", f, 3))
    return logit, geo, app


# --------------------------------------------------------------------------------------
# losses (loss.py, worker.py:151-173, 187-210)
# --------------------------------------------------------------------------------------
def contrastive_loss(anchor: Tensor, pos: Tensor, neg: Tensor, tau: float) -> Tensor:
    """loss.py:9-15.  -log(e^{ap/t} / (e^{ap/t}+e^{an/t})) == softplus((an-ap)/t)."""
    ap = (anchor * pos).sum(1) / tau
    an = (anchor * neg).sum(1) / tau
    return F.softplus(an - ap).mean()


def r1_penalty(real_logit: Tensor, images: Tensor) -> Tensor:
    """cal_r1_reg (loss.py:18-24): 0.5 * mean_b sum_chw (d sum(D(x)) / dx)^2, graph kept."""
    (grad,) = torch.autograd.grad(real_logit.sum(), images, create_graph=True)
    return 0.5 * grad.square().flatten(1).sum(1).mean()


def bce_logits(logit: Tensor, target: float) -> Tensor:
    """F.binary_cross_entropy_with_logits against a constant label (worker.py:156-157)."""
    return F.softplus(-logit).mean() if target == 1.0 else F.softplus(logit).mean()


class Hyper:
    """README.md:29/45/49 recipes."""
    def __init__(self, tau=0.05, l_aux=0.5, l_r1=10.0, l_s=1e-7, lr=2e-3, beta1=0.0, beta2=0.99):
        self.tau, self.l_aux, self.l_r1, self.l_s = tau, l_aux, l_r1, l_s
        self.lr, self.beta1, self.beta2 = lr, beta1, beta2


def _params(sd: State, requires_grad: bool):
    for k, v in sd.items():
        if not k.startswith("avg_latent"):
            v.requires_grad_(requires_grad)


def generator_loss(gsd: State, dsd: State, cfg: Config, hp: Hyper, it: int, z: Dict[str, Tensor]):
    """train_generator's loss (worker.py:187-210). z: rand1, rand2, resample1, resample2."""
    if it % 2 == 1:
        img = generator_forward(gsd, cfg, z["rand1"], z["rand2"])
        logit, _, _ = discriminator_forward(dsd, cfg, img, False)
        return bce_logits(logit, 1.0)
    img = generator_forward(gsd, cfg, z["rand1"], z["rand2"])
    img_g = generator_forward(gsd, cfg, z["resample1"], z["rand2"])
    img_a = generator_forward(gsd, cfg, z["rand1"], z["resample2"])
    logit, gf, af = discriminator_forward(dsd, cfg, img, True)
    _, gp, an = discriminator_forward(dsd, cfg, img_g, True)
    _, gn, ap = discriminator_forward(dsd, cfg, img_a, True)
    aux = (contrastive_loss(gf, gp, gn, hp.tau) + contrastive_loss(af, ap, an, hp.tau)) * hp.l_aux
    l1 = torch.cat([gsd["geometry_mapping.diagonal_params"],
                    gsd["appearance_mapping.diagonal_params"]]).abs().sum() * hp.l_s
    return bce_logits(logit, 1.0) + aux + l1


def discriminator_loss(gsd: State, dsd: State, cfg: Config, hp: Hyper, it: int,
                       z: Dict[str, Tensor], data: Dict[str, Tensor]):
    """train_discriminator's loss (worker.py:145-173). data: image, geometry_change, appearance_change."""
    fake = generator_forward(gsd, cfg, z["rand1"], z["rand2"])
    fake_logit, _, _ = discriminator_forward(dsd, cfg, fake, False)
    if it % 2 == 1:
        image = data["image"].detach().requires_grad_(True)
        real_logit, _, _ = discriminator_forward(dsd, cfg, image, False)
        loss = bce_logits(real_logit, 1.0) + bce_logits(fake_logit, 0.0)
        if it % 8 == 1:
            loss = loss + r1_penalty(real_logit, image) * hp.l_r1
        return loss
    real_logit, gf, af = discriminator_forward(dsd, cfg, data["image"], True)
    _, gp, an = discriminator_forward(dsd, cfg, data["geometry_change"], True)
    _, gn, ap = discriminator_forward(dsd, cfg, data["appearance_change"], True)
    aux = (contrastive_loss(gf, gp, gn, hp.tau) + contrastive_loss(af, ap, an, hp.tau)) * hp.l_aux
    return bce_logits(real_logit, 1.0) + bce_logits(fake_logit, 0.0) + aux


class OracleTrainer:
    """The reference iteration (loader.py:44-54 + worker.py:137-217) on oracle state dicts:
    G step, EMA, D step, with torch.optim.Adam(betas=(0,0.99), eps=1e-8) (worker.py:98-110)."""

    def __init__(self, cfg: Config, hp: Hyper, gsd: State, dsd: State, ema_decay=0.9999,
                 ema_start=0, freeze_d_start=10 ** 9, freeze_d_layer=3):
        self.cfg, self.hp, self.gsd, self.dsd = cfg, hp, gsd, dsd
        self.g_keys = [k for k in gsd if not k.startswith("avg_latent")]
        self.d_keys = list(dsd)
        self.g_opt = torch.optim.Adam([gsd[k] for k in self.g_keys], lr=hp.lr,
                                      betas=(hp.beta1, hp.beta2), eps=1e-8)
        self.d_opt = torch.optim.Adam([dsd[k] for k in self.d_keys], lr=hp.lr,
                                      betas=(hp.beta1, hp.beta2), eps=1e-8)
        self.ema = {k: v.detach().clone() for k, v in gsd.items()}
        self.ema_decay, self.ema_start = ema_decay, ema_start
        self.freeze_d_start, self.freeze_d_layer = freeze_d_start, freeze_d_layer

    def _frozen_d_prefixes(self):
        # worker.py:127-131: first freezeD_layer+2 children of shared_model
        return tuple(f"shared_model.{i}." for i in range(self.freeze_d_layer + 2))

    def g_step(self, it: int, z):
        _params(self.gsd, True); _params(self.dsd, False)
        self.g_opt.zero_grad(set_to_none=True)
        loss = generator_loss(self.gsd, self.dsd, self.cfg, self.hp, it, z)
        loss.backward()
        self.g_opt.step()
        return float(loss)

    def ema_step(self, it: int):
        decay = 0.0 if 0 <= it < self.ema_start else self.ema_decay        # ema.py:19-23
        with torch.no_grad():
            for k, v in self.gsd.items():
                self.ema[k] = v.detach().lerp(self.ema[k], decay)

    def d_step(self, it: int, z, data):
        _params(self.gsd, False); _params(self.dsd, True)
        if it >= self.freeze_d_start:
            for k in self.d_keys:
                if k.startswith(self._frozen_d_prefixes()):
                    self.dsd[k].requires_grad_(False)
        self.d_opt.zero_grad(set_to_none=True)
        loss = discriminator_loss(self.gsd, self.dsd, self.cfg, self.hp, it, z, data)
        loss.backward()
        self.d_opt.step()
        return float(loss)

    def iteration(self, it: int, zg, zd, data):
        g = self.g_step(it, zg)
        self.ema_step(it)
        d = self.d_step(it, zd, data)
        return g, d


def synthetic_latents(b: int, cfg: Config, gen: torch.Generator, device="cpu"):
    keys = ("rand1", "rand2", "resample1", "resample2")
    dims = (cfg.geo_noise_dim, cfg.app_noise_dim, cfg.geo_noise_dim, cfg.app_noise_dim)
    return {k: torch.randn(b, d, generator=gen).to(device) for k, d in zip(keys, dims)}


def synthetic_data(b: int, cfg: Config, gen: torch.Generator, device="cpu"):
    r = cfg.img_resolution
    return {k: (torch.rand(b, 3, r, r, generator=gen) * 2 - 1).to(device)
            for k in ("image", "geometry_change", "appearance_change")}
