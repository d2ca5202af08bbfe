"""Graph-replay kernel list of one even + one odd(R1) iteration at a small local batch (the 8-GPU
local batch of the 1024x1024 recipe is 4): run under
  ncu --metrics gpu__time_duration.sum --clock-control none --graph-profiling node --profile-from-start off
to see what the batch-independent part of the iteration is made of."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lcgan_b200 import cnn, train_step as T   # noqa: E402
from lcgan_b200.config import Config, Hyper   # noqa: E402

res = int(os.environ.get("RES", "1024"))
b = int(os.environ.get("B", "4"))
iters = int(os.environ.get("ITERS", "2"))
dev = torch.device("cuda")
cfg, hp = Config(img_resolution=res), Hyper(lr=1e-3)
torch.manual_seed(0)
G, D = cnn.Generator(cfg.namespace()).to(dev), cnn.Discriminator(cfg.namespace()).to(dev)
tr = T.GraphedTrainer(G, D, hp, b, dev)
tr.capture(warmup=2)
for k in tr.data:
    tr.data[k].uniform_(-1, 1)
for t in list(tr.z.values()) + list(tr.zd.values()):
    t.normal_()
for it in range(8):
    tr.iteration_graphed(it)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for it in range(8, 24):
    tr.iteration_graphed(it)
e1.record()
torch.cuda.synchronize()
print(f"res {res} b {b}: {e0.elapsed_time(e1) / 16:.2f} ms / iteration (graph replay), launches/iter "
      f"{sum(tr.launches.values())}", flush=True)
torch.cuda.profiler.start()
for it in range(iters):
    tr.iteration_graphed(it)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
