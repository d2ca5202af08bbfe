"""Pin the oracle (oracle/lcgan_oracle.py) against outputs of the unmodified reference
(tests/golden/*.pt, written by oracle/make_golden.py).  CPU only."""
import os

import pytest
import torch

from oracle import lcgan_oracle as O
from conftest import rel_l2

TOL = 2e-5          # fp32 CPU, different op order (shared-weight modconv) -> ~1e-6 observed
GRAD_TOL = 2e-3     # gradient norms/checksums after ~30 layers of fp32


def _load(golden_dir, name):
    return torch.load(os.path.join(golden_dir, name), weights_only=False)


def test_layer_vectors(golden_dir):
    L = _load(golden_dir, "layers.pt")
    for up in (1, 2):
        c = L[f"modconv_up{up}"]
        sd = {"m.weight.weight": c["w"], "m.bias": c["bias"]}
        assert rel_l2(O.mod_conv(sd, "m", c["x"], c["s"], up=up), c["y"]) < TOL
    c = L["synthesis_block"]
    sd = {"blk." + k: v for k, v in c["sd"].items()}
    for explicit in (False, True):
        y = O.synthesis_block(sd, "blk", c["x"], c["g"][:, 0], c["a"][:, 0], 0.1, explicit_warp=explicit,
                              a_lat1=c["a"][:, 1])
        assert rel_l2(y, c["y"]) < TOL
    c = L["discriminator_block"]
    sd = {"blk." + k: v for k, v in c["sd"].items()}
    assert rel_l2(O.discriminator_block(sd, "blk", c["x"]), c["y"]) < TOL
    for n in (4, 16):
        c = L[f"mbstd_b{n}"]
        assert rel_l2(O.minibatch_std(c["x"], 8), c["y"]) < TOL


def test_mapping_vector(golden_dir):
    c = _load(golden_dir, "layers.pt")["mapping"]
    sd = {"m." + k: v for k, v in c["sd"].items()}
    x = c["z"] @ O.mapping_matrix(sd, "m").t()
    for i in range(2):
        x = O.eq_linear(sd, f"m.mlp.{i}", x, lr_mul=0.01)
    assert rel_l2(x, c["y"]) < TOL


def _grad_summary(sd, keys):
    out = {}
    for k in keys:
        g = sd[k].grad
        if g is None:
            out[k] = None
            continue
        g = g.detach().double()
        ramp = torch.linspace(0.5, 1.5, g.numel(), dtype=torch.float64).reshape(g.shape)
        out[k] = (float(g.norm()), float((g * ramp).sum()))
    return out


def _check_grads(mine, gold):
    worst = 0.0
    for k, ref in gold.items():
        if ref is None:
            assert mine[k] is None or mine[k][0] == 0.0, k
            continue
        assert mine[k] is not None, k
        err = abs(mine[k][0] - ref[0]) / max(ref[0], 1e-12)
        worst = max(worst, err)
        assert err < GRAD_TOL, (k, mine[k], ref)
    return worst


@pytest.mark.parametrize("name", ["model_r16_b4.pt", "model_r32_b2.pt"])
def test_model_forward_losses_grads(golden_dir, name):
    g = _load(golden_dir, name)
    cfg = O.Config(img_resolution=g["res"]); hp = O.Hyper()
    gen = torch.Generator().manual_seed(1000 + g["seed"])
    z = O.synthetic_latents(g["b"], cfg, gen)
    data = O.synthetic_data(g["b"], cfg, gen)

    def fresh():
        return (O.make_generator_state(cfg, g["seed"]), O.make_discriminator_state(cfg, g["seed"] + 1))

    gsd, dsd = fresh()
    with torch.no_grad():
        img = O.generator_forward(gsd, cfg, z["rand1"], z["rand2"])
        assert rel_l2(img, g["g_image"]) < TOL
        assert rel_l2(gsd["avg_latent1"], g["avg_latent1"]) < TOL
        assert rel_l2(gsd["avg_latent2"], g["avg_latent2"]) < TOL
        assert rel_l2(O.generator_forward(gsd, cfg, z["rand1"], z["rand2"], 0.7), g["g_image_psi07"]) < TOL
        logit, ge, ae = O.discriminator_forward(dsd, cfg, data["image"], True)
        assert rel_l2(logit, g["d_logit"]) < TOL
        assert rel_l2(ge, g["d_geo"]) < TOL and rel_l2(ae, g["d_app"]) < TOL

    for it in (0, 1):
        gsd, dsd = fresh()
        O._params(gsd, True); O._params(dsd, False)
        loss = O.generator_loss(gsd, dsd, cfg, hp, it, z)
        loss.backward()
        assert abs(float(loss) - g[f"g_loss_it{it}"]) < 1e-4 * max(1.0, abs(g[f"g_loss_it{it}"]))
        keys = [k for k in gsd if not k.startswith("avg_latent")]
        _check_grads(_grad_summary(gsd, keys), g[f"g_grads_it{it}"])

    for it in (0, 1, 3):
        gsd, dsd = fresh()
        O._params(gsd, False); O._params(dsd, True)
        loss = O.discriminator_loss(gsd, dsd, cfg, hp, it, z, data)
        loss.backward()
        assert abs(float(loss) - g[f"d_loss_it{it}"]) < 1e-4 * max(1.0, abs(g[f"d_loss_it{it}"]))
        _check_grads(_grad_summary(dsd, list(dsd)), g[f"d_grads_it{it}"])
