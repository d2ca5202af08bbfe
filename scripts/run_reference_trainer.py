#!/usr/bin/env python
"""Run the reference's UNMODIFIED main.py / loader.py / worker.py on the lcgan_b200 drop-in modules.

    python scripts/run_reference_trainer.py --ref /path/to/lcgan -- --model_name /tmp/run \
        --img_resolution 64 --batch_size 8 --epoch 16 --print_interval 4

How: `lcgan_b200/dropin/` (same-named shims `cnn.py`, `custom_layers.py`, `loss.py`,
`ema.py`) is put ahead of the reference checkout on sys.path, so `worker.py`'s `import cnn`, `import
loss`, `from ema import Ema` resolve to the B200-native modules while `main.py`, `loader.py`,
`worker.py`, `eval/` come from the reference.  Where `albumentations` / `av` are not installed (this
image) they are stubbed, and with `--synthetic` the dataset is replaced by a synthetic
`custom_dataset.Dataset_` that yields (image, geometry_change, appearance_change) triples in [-1,1].
mp.spawn children inherit sys.path, so every rank sees the same modules.
"""
import argparse
import os
import runpy
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default=os.path.join(ROOT, "baseline", "_ref"), help="reference checkout")
    ap.add_argument("--synthetic", action="store_true", default=True)
    ap.add_argument("rest", nargs=argparse.REMAINDER)
    a = ap.parse_args()
    rest = [x for x in a.rest if x != "--"]
    stub_dir = os.path.join(ROOT, "scripts", "_stubs")
    sys.path[:0] = [os.path.join(ROOT, "lcgan_b200", "dropin"), ROOT, stub_dir, a.ref]
    os.environ["PYTHONPATH"] = os.pathsep.join(sys.path[:4] + [os.environ.get("PYTHONPATH", "")])
    sys.argv = [os.path.join(a.ref, "main.py")] + rest
    runpy.run_path(sys.argv[0], run_name="__main__")


if __name__ == "__main__":
    main()
