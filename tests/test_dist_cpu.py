"""world_size-2 gloo tests (CPU) of the N>1 host logic: batch sharding, max-over-ranks timing, and
the nn.Module obligations DDP puts on our drop-in modules (worker.py:40,88-96): construction
broadcast from rank 0, 'module.'-prefixed state_dict, deepcopy through DDP's pickle path, gradient
averaging with requires_grad toggled per half-iteration.  The CUDA forward is replaced by a stub
that touches the parameters (no kernels run on CPU; the product has no CPU fallback)."""
import copy
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from lcgan_b200 import cnn, dist_utils
        from oracle.lcgan_oracle import Config
        res = {}
        res["local_batch"] = dist_utils.local_batch(32, world)
        res["seed"] = dist_utils.rank_seed(rank)
        res["max_ms"] = dist_utils.max_over_ranks(10.0 + rank)

        torch.manual_seed(100 + rank)                      # different init per rank on purpose
        cfg = Config(img_resolution=16)
        G = cnn.Generator(cfg.namespace())
        # stub forward: every parameter contributes, scaled differently per rank
        G.forward = lambda scale: sum((p * p).sum() for p in G.parameters() if p.requires_grad) * scale
        ddp = dist_utils.wrap_ddp(G)
        # rank 0's parameters were broadcast at construction
        probe = ddp.module.const.detach().clone()
        gathered = [torch.zeros_like(probe) for _ in range(world)]
        dist.all_gather(gathered, probe)
        res["broadcast_ok"] = all(torch.equal(gathered[0], g) for g in gathered)
        res["keys_prefixed"] = all(k.startswith("module.") for k in ddp.state_dict())
        ema = copy.deepcopy(ddp)                           # worker.py:40 deep-copies the DDP wrapper
        res["deepcopy_ok"] = torch.equal(ema.module.const, ddp.module.const) and ema.module.const is not ddp.module.const

        # half-iteration 1: G trainable; gradient = mean over ranks of 2*p*scale_r
        for p in ddp.parameters():
            p.requires_grad = True
        loss = ddp(float(rank + 1))
        loss.backward()
        p = ddp.module.const
        expect = 2 * p.detach() * (sum(range(1, world + 1)) / world)
        res["grad_avg_ok"] = torch.allclose(p.grad, expect, rtol=1e-5, atol=1e-6)
        # half-iteration 2: everything frozen except the mapping diagonals (unused-parameter path)
        for n, q in ddp.named_parameters():
            q.requires_grad = n.endswith("diagonal_params")
            q.grad = None
        loss = ddp(1.0)
        loss.backward()
        d = ddp.module.geometry_mapping.diagonal_params
        res["frozen_ok"] = ddp.module.const.grad is None and torch.allclose(d.grad, 2 * d.detach(), rtol=1e-5)
        out[rank] = res
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_gloo_host_logic():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert len(out) == world
    for rank in range(world):
        r = out[rank]
        assert r["local_batch"] == 16 and r["seed"] == 1000 + rank
        assert r["max_ms"] == 11.0                       # max over ranks of 10, 11
        assert r["broadcast_ok"] and r["keys_prefixed"] and r["deepcopy_ok"]
        assert r["grad_avg_ok"] and r["frozen_ok"]


def test_local_batch_rejects_bad_split():
    from lcgan_b200 import dist_utils
    with pytest.raises(ValueError):
        dist_utils.local_batch(4, 8)
    assert dist_utils.local_batch(32, 8) == 4
