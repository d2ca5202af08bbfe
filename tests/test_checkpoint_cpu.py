"""Checkpoint interchange with the reference (CPU; the reference modules are used when the checkout is
mounted, otherwise the oracle's state dicts, which carry the same keys)."""
import os
import sys

import pytest
import torch

from lcgan_b200 import checkpoint, cnn
from oracle import lcgan_oracle as O

REF = "/root/reference"


def _models(res=32):
    cfg = O.Config(img_resolution=res)
    torch.manual_seed(0)
    return cfg, cnn.Generator(cfg.namespace()), cnn.Discriminator(cfg.namespace())


def test_round_trip_through_the_reference_layout(tmp_path):
    cfg, G, D = _models()
    gsd, dsd = O.make_generator_state(cfg, 3), O.make_discriminator_state(cfg, 4)
    for m, sd, name in ((G, gsd, "gen_model.ckpt"), (D, dsd, "disc_model.ckpt")):
        path = str(tmp_path / name)
        torch.save(checkpoint.add_module_prefix(sd), path)            # what worker.py:save_model writes
        checkpoint.load_reference_checkpoint(m, path)
        own = m.state_dict()
        assert set(own) == set(sd)
        assert all(torch.equal(own[k], sd[k]) for k in sd)
        out = checkpoint.reference_state_dict(m)
        assert all(k.startswith("module.") for k in out)
        assert all(torch.equal(out["module." + k], sd[k]) and out["module." + k].dtype == sd[k].dtype for k in sd)
        checkpoint.save_reference_checkpoint(m, path)
        again = torch.load(path, map_location="cpu")
        assert set(again) == set(out)


def test_mismatches_are_named():
    cfg, G, _ = _models()
    sd = O.make_generator_state(cfg, 0)
    bad = dict(sd); bad.pop(next(iter(bad)))
    with pytest.raises(KeyError, match="missing"):
        checkpoint.load_reference_checkpoint(G, bad)
    bad = dict(sd); k = next(k for k, v in sd.items() if v.dim() >= 2); bad[k] = bad[k][..., :1]
    with pytest.raises(ValueError, match=k.replace(".", r"\.")):
        checkpoint.load_reference_checkpoint(G, bad)


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not mounted")
def test_state_dicts_of_the_unmodified_reference_modules_load():
    """Keys, shapes and dtypes of the reference's own Generator / Discriminator (DDP-prefixed, as its
    trainer saves them) load strictly into the drop-in modules and come back identical."""
    saved = {k: sys.modules.pop(k) for k in ("cnn", "custom_layers", "loss", "ema") if k in sys.modules}
    sys.path.insert(0, REF)
    try:
        import cnn as ref_cnn                                           # the reference's module
        cfg, G, D = _models(64)
        torch.manual_seed(1)
        RG, RD = ref_cnn.Generator(cfg.namespace()), ref_cnn.Discriminator(cfg.namespace())
        for ours, ref in ((G, RG), (D, RD)):
            state = {"module." + k: v.clone() for k, v in ref.state_dict().items()}
            checkpoint.load_reference_checkpoint(ours, state)
            back = checkpoint.reference_state_dict(ours)
            assert set(back) == set(state)
            assert all(torch.equal(back[k], state[k]) and back[k].dtype == state[k].dtype for k in state)
            ref.load_state_dict(checkpoint.strip_module_prefix(back), strict=True)   # and the other way
    finally:
        sys.path.remove(REF)
        for k in ("cnn", "custom_layers", "loss", "ema"):
            sys.modules.pop(k, None)
        sys.modules.update(saved)
