"""Noise floor of the 100-step loss trajectory  --  TEST INFRASTRUCTURE (generates tests/golden/trajectory_noise_floor.json).

north_star asks for loss trajectories "within 1% over 100 steps".  The LC-GAN iteration with Adam(beta1=0) is a
chaotic map: any two correct implementations drift apart exponentially from rounding differences alone.  This
script measures that floor with the oracle against ITSELF - the same 100 iterations (res 32, batch 8, reference
hyper-parameters, seeded weights / latents / images) in
    fp64 (first 12 iterations here: CPU fp64 convolutions are slow; the GPU test runs the full fp64 oracle on
          the B200 and checks it against these 12),
    fp32, 8 threads and fp32, 1 thread   (two legitimate fp32 evaluations: only the summation order differs),
so tests/test_gpu_trajectory.py can hold the CUDA path to "no further from the fp64 trajectory than the oracle's
own fp32 runs are" instead of to a bound no fp32 implementation can meet.

    python oracle/make_noise_floor.py            (about 20 minutes on 8 cores)
"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import lcgan_oracle as O  # noqa: E402

RES, BATCH, STEPS, DATA_SEED = 32, 8, 100, 5


def run(dtype, threads, steps=STEPS):
    torch.set_num_threads(threads)
    cfg, hp = O.Config(img_resolution=RES), O.Hyper()
    gen = torch.Generator().manual_seed(DATA_SEED)
    gsd = {k: v.to(dtype) for k, v in O.make_generator_state(cfg, 0).items()}
    dsd = {k: v.to(dtype) for k, v in O.make_discriminator_state(cfg, 1).items()}
    tr = O.OracleTrainer(cfg, hp, gsd, dsd)
    out = []
    for it in range(steps):
        zg, zd = O.synthetic_latents(BATCH, cfg, gen), O.synthetic_latents(BATCH, cfg, gen)
        data = O.synthetic_data(BATCH, cfg, gen)
        cast = lambda d: {k: v.to(dtype) for k, v in d.items()}
        out.append(tr.iteration(it, cast(zg), cast(zd), cast(data)))
    return out


if __name__ == "__main__":
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else STEPS
    res = {"res": RES, "batch": BATCH, "steps": steps, "data_seed": DATA_SEED, "weights_seeds": [0, 1],
           "torch": torch.__version__}
    for name, dtype, threads, n in (("fp64", torch.float64, 8, min(steps, 12)), ("fp32_t8", torch.float32, 8, steps),
                                    ("fp32_t1", torch.float32, 1, steps)):
        t0 = time.time()
        res[name] = run(dtype, threads, n)
        print(name, f"{time.time() - t0:.0f}s", res[name][:2], flush=True)
    with open(os.path.join(ROOT, "tests", "golden", "trajectory_noise_floor.json"), "w") as f:
        json.dump(res, f)
