"""Generate tests/golden/*.pt from the UNMODIFIED reference modules (/root/reference).

Run in the build container only (the reference is not present on the GPU box):

    python oracle/make_golden.py

For each case the script builds deterministic parameters with ``oracle.lcgan_oracle.make_*_state``
(plain seeded torch.randn, same distributions as the reference constructors), loads them into the
reference's ``cnn.Generator`` / ``cnn.Discriminator`` with ``load_state_dict``, runs the
reference's own forward / losses / backward, and stores inputs-by-seed plus the reference outputs.
The fixtures are what pins the oracle: tests/test_oracle_golden.py replays them through
oracle/lcgan_oracle.py; tests then compare the CUDA path against the oracle.
"""
import os
import sys
import warnings

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference"

from oracle import lcgan_oracle as O  # noqa: E402


def import_reference():
    """Import the reference's cnn/loss under their own names, isolated from our package."""
    saved = {k: sys.modules.pop(k) for k in ("cnn", "custom_layers", "loss", "ema") if k in sys.modules}
    sys.path.insert(0, REF)
    try:
        import cnn as ref_cnn
        import loss as ref_loss
    finally:
        sys.path.remove(REF)
        for k in ("cnn", "custom_layers", "loss", "ema"):
            sys.modules.pop(k, None)
        sys.modules.update(saved)
    return ref_cnn, ref_loss


def grad_summary(named):
    """Per-parameter gradient L2 norm + signed checksum (small, order-sensitive enough)."""
    out = {}
    for k, g in named:
        if g is None:
            out[k] = None
        else:
            g = g.detach().double()
            ramp = torch.linspace(0.5, 1.5, g.numel(), dtype=torch.float64).reshape(g.shape)
            out[k] = (float(g.norm()), float((g * ramp).sum()))
    return out


def case(res, b, seed):
    ref_cnn, ref_loss = import_reference()
    cfg = O.Config(img_resolution=res)
    hp = O.Hyper()
    gsd = O.make_generator_state(cfg, seed=seed)
    dsd = O.make_discriminator_state(cfg, seed=seed + 1)
    G = ref_cnn.Generator(cfg.namespace())
    D = ref_cnn.Discriminator(cfg.namespace())
    G.load_state_dict(gsd)
    D.load_state_dict(dsd)
    gen = torch.Generator().manual_seed(1000 + seed)
    z = O.synthetic_latents(b, cfg, gen)
    data = O.synthetic_data(b, cfg, gen)
    out = {"res": res, "b": b, "seed": seed}

    # ---- plain forwards -----------------------------------------------------------------
    with torch.no_grad():
        img = G(z["rand1"], z["rand2"])
        out["g_image"] = img.clone()
        out["avg_latent1"] = G.avg_latent1.clone()
        out["avg_latent2"] = G.avg_latent2.clone()
        out["g_image_psi07"] = G(z["rand1"], z["rand2"], 0.7).clone()
        logit, ge, ae = D(data["image"], True)
        out["d_logit"], out["d_geo"], out["d_app"] = logit.clone(), ge.clone(), ae.clone()
    G.load_state_dict(gsd)  # reset avg_latent buffers

    # ---- generator losses (worker.py:187-210) + parameter gradients -----------------------
    for it in (0, 1):
        G.load_state_dict(gsd); G.zero_grad(); D.zero_grad()
        for p in G.parameters(): p.requires_grad_(True)
        for p in D.parameters(): p.requires_grad_(False)
        ones = torch.ones(b, 1)
        if it % 2 == 1:
            logit, _, _ = D(G(z["rand1"], z["rand2"]), False)
            loss = F.binary_cross_entropy_with_logits(logit, ones)
        else:
            a_img = G(z["rand1"], z["rand2"])
            g_img = G(z["resample1"], z["rand2"])
            p_img = G(z["rand1"], z["resample2"])
            logit, gf, af = D(a_img, True)
            _, gp, an = D(g_img, True)
            _, gn, ap = D(p_img, True)
            aux = (ref_loss.contrastive_loss(gf, gp, gn, hp.tau)
                   + ref_loss.contrastive_loss(af, ap, an, hp.tau)) * hp.l_aux
            d1 = G.geometry_mapping.diagonal_params.view(-1)
            d2 = G.appearance_mapping.diagonal_params.view(-1)
            loss = F.binary_cross_entropy_with_logits(logit, ones) + aux \
                + torch.norm(torch.cat([d1, d2]), p=1) * hp.l_s
        loss.backward()
        out[f"g_loss_it{it}"] = float(loss)
        out[f"g_grads_it{it}"] = grad_summary((k, p.grad) for k, p in G.named_parameters())

    # ---- discriminator losses (worker.py:145-173), incl. R1 double backward ---------------
    for it in (0, 1, 3):
        G.load_state_dict(gsd); G.zero_grad(); D.zero_grad()
        for p in G.parameters(): p.requires_grad_(False)
        for p in D.parameters(): p.requires_grad_(True)
        ones, zeros = torch.ones(b, 1), torch.zeros(b, 1)
        fake_logit, _, _ = D(G(z["rand1"], z["rand2"]), False)
        if it % 2 == 1:
            image = data["image"].clone().requires_grad_(True)
            real_logit, _, _ = D(image, False)
            loss = F.binary_cross_entropy_with_logits(real_logit, ones) \
                + F.binary_cross_entropy_with_logits(fake_logit, zeros)
            if it % 8 == 1:
                r1 = ref_loss.cal_r1_reg(real_logit, image, "cpu")
                out["r1_value"] = float(r1)
                loss = loss + r1 * hp.l_r1
        else:
            real_logit, gf, af = D(data["image"], True)
            _, gp, an = D(data["geometry_change"], True)
            _, gn, ap = D(data["appearance_change"], True)
            aux = (ref_loss.contrastive_loss(gf, gp, gn, hp.tau)
                   + ref_loss.contrastive_loss(af, ap, an, hp.tau)) * hp.l_aux
            loss = F.binary_cross_entropy_with_logits(real_logit, ones) \
                + F.binary_cross_entropy_with_logits(fake_logit, zeros) + aux
        loss.backward()
        out[f"d_loss_it{it}"] = float(loss)
        out[f"d_grads_it{it}"] = grad_summary((k, p.grad) for k, p in D.named_parameters())
    return out


def layer_cases():
    """Isolated-layer vectors from the reference's custom_layers classes (small shapes)."""
    ref_cnn, _ = import_reference()
    cl = sys.modules.get("custom_layers")
    saved = {k: sys.modules.pop(k) for k in ("cnn", "custom_layers") if k in sys.modules}
    sys.path.insert(0, REF)
    import custom_layers as cl  # noqa: F811
    sys.path.remove(REF)
    for k in ("cnn", "custom_layers"):
        sys.modules.pop(k, None)
    sys.modules.update(saved)
    gen = torch.Generator().manual_seed(7)
    out = {}
    # modulated conv, up=1 and up=2, odd channel counts
    for up in (1, 2):
        m = cl.ModulatedConv2d(6, 5, 3, up=up)
        w = torch.randn(5, 6, 3, 3, generator=gen); bias = torch.randn(5, generator=gen)
        m.load_state_dict({"weight.weight": w, "bias": bias})
        x = torch.randn(3, 6, 7, 9, generator=gen); s = torch.randn(3, 6, generator=gen)
        out[f"modconv_up{up}"] = {"w": w, "bias": bias, "x": x, "s": s, "y": m(x, s).detach()}
    # synthesis block
    blk = cl.SynthesisBlock(8, 6, 4, 5, 16, 0.1)
    sd = {k: torch.randn(v.shape, generator=gen) for k, v in blk.state_dict().items()}
    blk.load_state_dict(sd)
    x = torch.randn(2, 8, 8, 8, generator=gen)
    gl = torch.randn(2, 1, 4, generator=gen); al = torch.randn(2, 2, 5, generator=gen)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out["synthesis_block"] = {"sd": sd, "x": x, "g": gl, "a": al, "y": blk(x, gl, al).detach()}
    # discriminator block
    blk = cl.DiscriminatorBlock(6, 10, skip=True)
    sd = {k: torch.randn(v.shape, generator=gen) for k, v in blk.state_dict().items()}
    blk.load_state_dict(sd)
    x = torch.randn(2, 6, 8, 8, generator=gen)
    out["discriminator_block"] = {"sd": sd, "x": x, "y": blk(x).detach()}
    # minibatch std at b=4 (one group) and b=16 (two strided groups of 8)
    for n in (4, 16):
        x = torch.randn(n, 5, 4, 4, generator=gen)
        out[f"mbstd_b{n}"] = {"x": x, "y": cl.MinibatchStdLayer(8)(x).detach()}
    # mapping network
    mp = cl.MappingNetwork([8, 16, 16])
    sd = {k: torch.randn(v.shape, generator=gen) for k, v in mp.state_dict().items()}
    mp.load_state_dict(sd)
    z = torch.randn(3, 8, generator=gen)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out["mapping"] = {"sd": sd, "z": z, "y": mp(z).detach()}
    return out


if __name__ == "__main__":
    torch.set_num_threads(8)
    gold = os.path.join(ROOT, "tests", "golden")
    os.makedirs(gold, exist_ok=True)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        torch.save(layer_cases(), os.path.join(gold, "layers.pt"))
        for res, b, seed in ((16, 4, 0), (32, 2, 3)):
            torch.save(case(res, b, seed), os.path.join(gold, f"model_r{res}_b{b}.pt"))
            print("wrote", res, b)
