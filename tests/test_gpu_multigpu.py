"""Data parallel on real GPUs (needs >= 2 visible devices; run with `gpurun --gpus 2 -- python -m pytest
tests/test_gpu_multigpu.py -m gpu`).  SURVEY section 4 item 4 / section 8e: N-GPU gradients == mean of the N
single-rank gradients at the same local batch - checked on the path bench.py times: GraphedTrainer(world=N), i.e.
post-accumulate-grad hooks, flat buckets, NCCL all-reduce on a side stream, all captured in CUDA graphs."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import json, os, sys
import torch, torch.distributed as dist
sys.path.insert(0, os.environ["LCGAN_ROOT"])
from lcgan_b200 import cnn, ops, train_step as T
from lcgan_b200.config import Config, Hyper

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
ops.set_precision("bf16")
ops.set_deterministic(True)
cfg, b = Config(img_resolution=32), 4
hp = Hyper(lr=0.0)                                   # lr 0: the optimizer steps inside the graphs leave the weights alone

def build():
    torch.manual_seed(0)
    return cnn.Generator(cfg.namespace()).to(dev), cnn.Discriminator(cfg.namespace()).to(dev)

def inputs(r):
    g = torch.Generator().manual_seed(500 + r)
    z = {k: torch.randn(b, 64, generator=g).to(dev) for k in ("rand1", "rand2", "resample1", "resample2")}
    data = {k: (torch.rand(b, 3, 32, 32, generator=g) * 2 - 1).to(dev) for k in ("image", "geometry_change", "appearance_change")}
    return z, data

def single_rank_grads(r):
    """plain eager backward of every variant on rank r's inputs (no exchange)"""
    G, D = build()
    z, data = inputs(r)
    out = {}
    for name, it in T.GraphedTrainer._VARIANT_IT.items():
        G.zero_grad(); D.zero_grad()
        T.requires_grad(G, name[0] == "g"); T.requires_grad(D, name[0] == "d")
        loss = T.generator_loss(G, D, hp, it, z) if name[0] == "g" else T.discriminator_loss(G, D, hp, it, {k: z[k] for k in ("rand1", "rand2")}, data)
        loss.backward()
        net = G if name[0] == "g" else D
        out[name] = {k: p.grad.double().clone() for k, p in net.named_parameters() if p.grad is not None}
    return out

mean = None
for r in range(world):
    gr = single_rank_grads(r)
    if mean is None:
        mean = gr
    else:
        for v in gr:
            assert gr[v].keys() == mean[v].keys()
            for k in gr[v]:
                mean[v][k] += gr[v][k]
for v in mean:
    for k in mean[v]:
        mean[v][k] /= world

G, D = build()
tr = T.GraphedTrainer(G, D, hp, b, dev, world=world)
z, data = inputs(rank)
for k in tr.z: tr.z[k].copy_(z[k])
for k in tr.zd: tr.zd[k].copy_(z[k])
for k in tr.data: tr.data[k].copy_(data[k])
tr.capture(warmup=2)
worst = {}
for name, it in T.GraphedTrainer._VARIANT_IT.items():
    tr.graphs[name].replay()
    torch.cuda.synchronize()
    net = G if name[0] == "g" else D
    plan = tr.exchange.plans[name]
    got = {}
    for k, p in net.named_parameters():
        if p in plan.slot:
            bi, off = plan.slot[p]
            got[k] = plan.flats[bi][off:off + p.numel()].view_as(p).double()
    assert got.keys() == mean[name].keys(), (name, set(got) ^ set(mean[name]))
    w = 0.0
    for k in got:
        e = float((got[k] - mean[name][k]).norm() / mean[name][k].norm().clamp_min(1e-30))
        w = max(w, e)
    worst[name] = (w, len(plan.flats))
if rank == 0:
    print("RESULT " + json.dumps(worst), flush=True)
# a captured graph that still holds NCCL kernels makes destroy_process_group() wait forever: drop the graphs first,
# and leave without the teardown
tr.graphs.clear()
torch.cuda.synchronize()
dist.barrier()
sys.stdout.flush()
os._exit(0)
'''


def test_graph_captured_nccl_gradients_equal_mean_of_rank_gradients(tmp_path):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, LCGAN_ROOT=ROOT)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29731", str(script)],
                       cwd=ROOT, env=env, capture_output=True, text=True, timeout=420)
    assert r.returncode == 0, (r.stdout[-3000:], r.stderr[-5000:])
    line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")][-1]
    worst = json.loads(line[7:])
    out_dir = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out_dir):
        json.dump(worst, open(os.path.join(out_dir, "multigpu_grad_parity.json"), "w"))
    for name, (err, nbuckets) in worst.items():
        # deterministic kernels on both sides: the only difference is fp32 summation order of the all-reduce
        assert err < 1e-5, (name, err)
    assert worst["d_even"][1] >= 2, "the discriminator's 260 MB of gradients should span several buckets"
