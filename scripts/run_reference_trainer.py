#!/usr/bin/env python
"""Run the reference's UNMODIFIED main.py / loader.py / worker.py on the lcgan_b200 drop-in modules.

    python scripts/run_reference_trainer.py [--ref DIR] [--modules dropin|reference] [--seed S]
        [--precision bf16|fp32] -- --model_name /tmp/run --img_resolution 64 --batch_size 8 --epoch 16 ...

How: `lcgan_b200/dropin/` (same-named shims `cnn.py`, `custom_layers.py`, `loss.py`,
`ema.py`) is put ahead of the reference checkout on sys.path, so `worker.py`'s `import cnn`, `import
loss`, `from ema import Ema` resolve to the B200-native modules while `main.py`, `loader.py`,
`worker.py`, `eval/` come from the reference.  `--modules reference` leaves the drop-ins out: the same
launcher then runs the reference end to end on its own torch/cuDNN modules (the oracle arm of
tests/test_gpu_reference_trainer.py).  Where `albumentations` / `av` are not installed (this image)
they are stubbed, and the dataset is a synthetic `custom_dataset.Dataset_` that yields (image,
geometry_change, appearance_change) triples in [-1,1].  mp.spawn children inherit PYTHONPATH and the
environment, so every rank sees the same modules, seed and precision.
"""
import argparse
import os
import runpy
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default=os.path.join(ROOT, "baseline", "_ref"), help="reference checkout")
    ap.add_argument("--modules", default="dropin", choices=["dropin", "reference"])
    ap.add_argument("--real-dataset", action="store_true",
                    help="use the reference's own custom_dataset.py (needs PIL + albumentations and an ImageFolder)")
    ap.add_argument("--seed", type=int, default=None, help="torch.manual_seed in every rank before the models are built")
    ap.add_argument("--precision", default=None, choices=["bf16", "fp32"],
                    help="drop-in modules: activation/compute mode; fp32 also switches TF32 off for torch's own ops")
    ap.add_argument("rest", nargs=argparse.REMAINDER)
    a = ap.parse_args()
    rest = [x for x in a.rest if x != "--"]
    paths = []
    if a.modules == "dropin":
        paths += [os.path.join(ROOT, "lcgan_b200", "dropin"), ROOT]
    if not a.real_dataset:
        paths.append(os.path.join(ROOT, "scripts", "_stubs"))
    paths.append(a.ref)
    sys.path[:0] = paths
    os.environ["PYTHONPATH"] = os.pathsep.join(paths + [os.environ.get("PYTHONPATH", "")])
    if a.seed is not None:
        os.environ["LCGAN_SEED"] = str(a.seed)
    if a.precision:
        os.environ["LCGAN_PRECISION"] = a.precision
        if a.precision == "fp32":
            os.environ["LCGAN_NO_TF32"] = "1"
    sys.argv = [os.path.join(a.ref, "main.py")] + rest
    runpy.run_path(sys.argv[0], run_name="__main__")


if __name__ == "__main__":
    main()
