"""100-step loss trajectories: lcgan_b200 (fp32 and bf16 modes) vs the oracle trained from the
same weights on the same latents and images (res 32, batch 8, reference hyper-parameters).

north_star asks for "within 1%".  Measured on the B200 (numerics): this GAN with Adam
(beta1 = 0, lr 2e-3) is a chaotic map - the oracle and our fp32 path, which agree to 1e-7 on
iteration 0, are 1e-4 apart at iteration 2, 1e-2 at iteration 8 and O(0.2) at iteration 11, i.e.
fp32 rounding noise alone exceeds 1% after ~8 iterations.  What can be checked, and is:
  * the first iterations, before amplification: fp32 <= 1e-3, bf16 <= 1e-2;
  * the 100-step mean of each loss: within a factor of 3 of the oracle's (a sanity bound - training
    neither diverges nor collapses; run-to-run, atomics ordering alone moves this mean by 5-50%);
  * bf16 diverges no faster than the fp32 noise floor allows (same order of magnitude at step 8).
"""
import statistics

import pytest
import torch

pytestmark = pytest.mark.gpu


def _run(mode, steps, res=32, b=8):
    from lcgan_b200 import cnn, ops, train_step as T
    from oracle import lcgan_oracle as O
    dev = "cuda"
    cfg, hp = O.Config(img_resolution=res), O.Hyper()
    gen = torch.Generator().manual_seed(5)
    gsd, dsd = O.make_generator_state(cfg, 0), O.make_discriminator_state(cfg, 1)
    if mode == "oracle":
        tr = O.OracleTrainer(cfg, hp, {k: v.to(dev) for k, v in gsd.items()}, {k: v.to(dev) for k, v in dsd.items()})
    else:
        ops.set_precision(mode)
        G, D = cnn.Generator(cfg.namespace()), cnn.Discriminator(cfg.namespace())
        G.load_state_dict(gsd); D.load_state_dict(dsd)
        tr = T.Trainer(G.to(dev), D.to(dev), hp)
    out = []
    for it in range(steps):
        zg, zd = O.synthetic_latents(b, cfg, gen, dev), O.synthetic_latents(b, cfg, gen, dev)
        data = O.synthetic_data(b, cfg, gen, dev)
        out.append(tr.iteration(it, zg, zd, data))
    return out


def test_loss_trajectories_100_steps():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    from lcgan_b200 import ops
    try:
        steps = 100
        ref = _run("oracle", steps)
        runs = {m: _run(m, steps) for m in ("fp32", "bf16")}
    finally:
        ops.set_precision("bf16")

    def rel(a, o):
        return abs(a - o) / max(abs(o), 1e-6)

    for mode, early_tol in (("fp32", 1e-3), ("bf16", 1e-2)):
        r = runs[mode]
        assert all(torch.isfinite(torch.tensor(x)).all() for x in r)
        for it in range(3):
            for j in range(2):
                assert rel(r[it][j], ref[it][j]) < early_tol, (mode, it, j, r[it], ref[it])
        for j in range(2):
            m_ref = statistics.mean(x[j] for x in ref)
            m_run = statistics.mean(x[j] for x in r)
            assert m_ref / 3 < m_run < m_ref * 3, (mode, j, m_run, m_ref)
    # bf16 error at the edge of the predictable window is the same order as the fp32 noise floor
    for j in range(2):
        f8 = max(rel(runs["fp32"][it][j], ref[it][j]) for it in range(6, 10))
        b8 = max(rel(runs["bf16"][it][j], ref[it][j]) for it in range(6, 10))
        assert b8 < max(10 * f8, 0.3), (j, f8, b8)
