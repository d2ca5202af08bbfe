"""100-step loss trajectories against a MEASURED noise floor (north_star: "loss trajectories over 100 steps within 1%").

The LC-GAN iteration with Adam(beta1 = 0) is a chaotic map: rounding differences between two correct evaluations
grow exponentially until the trajectories decorrelate.  So the floor is measured, not assumed:
  * tests/golden/trajectory_noise_floor.json (oracle/make_noise_floor.py) holds the oracle's own CPU runs of these
    100 iterations: fp32 with 8 threads, fp32 with 1 thread (summation order only), and fp64 for the first 12;
  * here the oracle runs on the B200 in fp64 (the truth; pinned to the committed CPU fp64 steps) and in fp32 (cuDNN,
    TF32 off);
  * divergence(t) of a run = max over {g_loss, d_loss} of |run - fp64| / max(|fp64|, 1e-6); the noise floor F(t) is
    the largest divergence among the three oracle fp32 runs, and its running maximum defines the window in which
    the oracle agrees with itself to 1%.
Measured (profiles/r02_trajectory_curves.json): the oracle's fp32 runs are 2e-7 from fp64 at iteration 0, 2.5e-4 at
iteration 1 (the first Adam step with beta1 = 0 is lr * sign(g): every sign disagreement on a near-zero gradient
moves a weight by 2 lr), 3e-3 at iteration 4, 1.4e-2 at iteration 9, O(0.1-1) from iteration 11 on.
What is asserted for our CUDA path (fp32 mode run deterministically, and bf16 mode):
  1. within 1% of the truth - north_star's criterion - for every iteration before the oracle's own fp32 runs leave
     a third of that budget (the "window", about 5 iterations);
  2. over all 100 steps our fp32 divergence never runs ahead of the floor by more than a fixed factor
     (running maxima: ours <= K * floor + 1e-6) - i.e. we diverge like the oracle diverges from itself;
  3. bf16 starts three orders of magnitude above the fp32 floor (operand rounding: 6e-4 at iteration 0): within 1%
     for the first 3 iterations, never ahead of the floor by more than K_BF, and - like fp32 - its 100-step mean
     losses stay within the spread of the oracle's own decorrelated runs.
The curves are written to gpurun_out/trajectory_curves.json (profiles/ keeps a B200 copy).
"""
import json
import os
import statistics

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
K_FLOOR = 20.0          # criterion 2: allowed lead of fp32 over the oracle's self-divergence (running maxima)
K_BF = 50.0             # criterion 3: the same for bf16 (measured lead: 3x at iteration 1, 10x at iteration 4)


def _run(mode, steps, res, b, data_seed):
    from lcgan_b200 import cnn, ops, train_step as T
    from oracle import lcgan_oracle as O
    dev = "cuda"
    cfg, hp = O.Config(img_resolution=res), O.Hyper()
    gen = torch.Generator().manual_seed(data_seed)
    gsd, dsd = O.make_generator_state(cfg, 0), O.make_discriminator_state(cfg, 1)
    if mode.startswith("oracle"):
        dt = torch.float64 if mode == "oracle_fp64" else torch.float32
        tr = O.OracleTrainer(cfg, hp, {k: v.to(dev, dt) for k, v in gsd.items()}, {k: v.to(dev, dt) for k, v in dsd.items()})
    else:
        dt = torch.float32
        ops.set_precision(mode)
        ops.set_deterministic(True)
        G, D = cnn.Generator(cfg.namespace()), cnn.Discriminator(cfg.namespace())
        G.load_state_dict(gsd); D.load_state_dict(dsd)
        tr = T.Trainer(G.to(dev), D.to(dev), hp)
    out = []
    for it in range(steps):
        zg, zd = O.synthetic_latents(b, cfg, gen, dev), O.synthetic_latents(b, cfg, gen, dev)
        data = O.synthetic_data(b, cfg, gen, dev)
        cast = lambda d: {k: v.to(dt) for k, v in d.items()}
        g, d = tr.iteration(it, cast(zg), cast(zd), cast(data))
        out.append((float(g), float(d)))
    return out


def _div(run, truth):
    return [max(abs(a - t) / max(abs(t), 1e-6) for a, t in zip(r, tr)) for r, tr in zip(run, truth)]


def _cummax(xs):
    out, m = [], 0.0
    for x in xs:
        m = max(m, x)
        out.append(m)
    return out


def test_loss_trajectories_100_steps():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    from lcgan_b200 import ops
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "trajectory_noise_floor.json")))
    steps, res, b, seed = gold["steps"], gold["res"], gold["batch"], gold["data_seed"]
    try:
        truth = _run("oracle_fp64", steps, res, b, seed)
        runs = {m: _run(m, steps, res, b, seed) for m in ("oracle_fp32", "fp32", "bf16")}
        rerun = _run("fp32", 12, res, b, seed)
    finally:
        ops.set_precision("bf16")
        ops.set_deterministic(False)
    runs["oracle_cpu_fp32_t8"], runs["oracle_cpu_fp32_t1"] = gold["fp32_t8"], gold["fp32_t1"]

    # the truth is pinned: the B200 fp64 oracle reproduces the committed CPU fp64 iterations
    n64 = len(gold["fp64"])
    d64 = _div(truth[:n64], gold["fp64"])
    assert max(d64) < 1e-6, d64
    # deterministic mode: the fp32 run repeats bit for bit
    assert rerun == runs["fp32"][:12], "fp32 deterministic mode is not run-to-run reproducible"

    div = {k: _div(v, truth) for k, v in runs.items()}
    floor = [max(div[k][t] for k in ("oracle_fp32", "oracle_cpu_fp32_t8", "oracle_cpu_fp32_t1")) for t in range(steps)]
    cfloor, cours, cbf = _cummax(floor), _cummax(div["fp32"]), _cummax(div["bf16"])
    window = next((t for t, v in enumerate(cfloor) if v > 1e-2 / 3), steps)  # oracle within a third of 1% before this
    curves = {"steps": steps, "window": window, "divergence": div, "floor": floor, "truth": truth,
              "ours_fp32": runs["fp32"], "ours_bf16": runs["bf16"], "K": K_FLOOR, "K_BF": K_BF}
    out_dir = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out_dir):
        json.dump(curves, open(os.path.join(out_dir, "trajectory_curves.json"), "w"))

    assert all(all(x == x and abs(x) < 1e4 for x in r) for m in ("fp32", "bf16") for r in runs[m])
    # 1. within 1% wherever the oracle's own fp32 runs are within a third of that
    assert window >= 3, ("noise floor exceeds 0.33% almost immediately", cfloor[:6])
    assert all(div["fp32"][t] < 1e-2 for t in range(window)), ("fp32", window, div["fp32"][:window])
    # 2. never ahead of the floor by more than K (running maxima), over all 100 steps
    lead = max(cours[t] / (K_FLOOR * cfloor[t] + 1e-6) for t in range(steps))
    assert lead <= 1.0, ("fp32 diverges faster than the oracle from itself", lead, cours[:12], cfloor[:12])
    # 3. bf16
    assert all(div["bf16"][t] < 1e-2 for t in range(3)), ("bf16", div["bf16"][:4])
    lead_bf = max(cbf[t] / (K_BF * cfloor[t] + 1e-3) for t in range(steps))
    assert lead_bf <= 1.0, ("bf16 diverges faster than operand rounding explains", lead_bf, cbf[:12], cfloor[:12])
    for j in range(2):
        means = [statistics.mean(x[j] for x in runs[k]) for k in ("oracle_fp32", "oracle_cpu_fp32_t8", "oracle_cpu_fp32_t1")]
        means.append(statistics.mean(x[j] for x in truth))
        lo, hi = min(means), max(means)
        slack = 0.5 * (hi - lo) + 0.25 * abs(statistics.mean(means))
        for m in ("fp32", "bf16"):
            mine = statistics.mean(x[j] for x in runs[m])
            assert lo - slack <= mine <= hi + slack, (m, j, mine, means)
