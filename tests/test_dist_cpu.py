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
        from lcgan_b200.config import Config
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


def _exchange_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from lcgan_b200.dist_utils import GradExchange
        torch.manual_seed(0)                               # same weights on every rank
        net = torch.nn.Sequential(torch.nn.Linear(24, 40), torch.nn.Tanh(), torch.nn.Linear(40, 40), torch.nn.Tanh(),
                                  torch.nn.Linear(40, 8))
        unused = torch.nn.Linear(8, 8)                     # never touched by variant "a"
        net.add_module("unused", unused)
        fwd = lambda x: net[4](net[3](net[2](net[1](net[0](x)))))
        ex = GradExchange({"n": net}, "cpu", world, bucket_bytes=4096)     # several buckets
        torch.manual_seed(10 + rank)                       # different data per rank
        xs = [torch.randn(6, 24) for _ in range(3)]
        ok = True
        for step, x in enumerate(xs):                      # step 0 records the order, 1..2 use the buckets
            for variant in ("a", "b"):
                net.zero_grad()
                y = fwd(x)
                loss = (y * y).mean() if variant == "a" else (unused(y) ** 2).mean()
                ex.backward(loss, variant, "n")
                mine = {k: (None if p.grad is None else p.grad.clone()) for k, p in net.named_parameters()}
                # reference: plain backward + all-reduce of every gradient
                net.zero_grad()
                y = fwd(x)
                loss = (y * y).mean() if variant == "a" else (unused(y) ** 2).mean()
                loss.backward()
                for k, p in net.named_parameters():
                    if p.grad is None:
                        ok &= mine[k] is None
                        continue
                    dist.all_reduce(p.grad); p.grad /= world
                    ok &= mine[k] is not None and torch.allclose(mine[k], p.grad, rtol=1e-6, atol=1e-7)
        out[rank] = {"ok": bool(ok), "buckets": len(ex.plans["a"].flats), "unused_none": mine is not None}
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_grad_exchange_matches_mean_of_rank_grads():
    """dist_utils.GradExchange (the graph-capturable replacement of DDP's reducer): N-rank gradients ==
    mean of the single-rank gradients, unused parameters keep grad None, several buckets per variant."""
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_exchange_worker, args=(world, port, out), nprocs=world, join=True)
    for rank in range(world):
        assert out[rank]["ok"], out[rank]
        assert out[rank]["buckets"] >= 2


def test_local_batch_rejects_bad_split():
    from lcgan_b200 import dist_utils
    with pytest.raises(ValueError):
        dist_utils.local_batch(4, 8)
    assert dist_utils.local_batch(32, 8) == 4
