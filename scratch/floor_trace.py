"""Kernel table of graph replays at a small local batch via the torch profiler (CUPTI activity records of the
replayed kernels; seconds instead of the ~10 minutes `ncu --graph-profiling node` needs for 16 000 launches)."""
import collections, os, re, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lcgan_b200 import cnn, train_step as T
from lcgan_b200.config import Config, Hyper

res, b = int(os.environ.get("RES", "1024")), int(os.environ.get("B", "4"))
dev = torch.device("cuda")
torch.manual_seed(0)
cfg = Config(img_resolution=res)
G, D = cnn.Generator(cfg.namespace()).to(dev), cnn.Discriminator(cfg.namespace()).to(dev)
tr = T.GraphedTrainer(G, D, Hyper(lr=1e-3), b, dev)
tr.capture(warmup=2)
for k in tr.data: tr.data[k].uniform_(-1, 1)
for t in list(tr.z.values()) + list(tr.zd.values()): t.normal_()
for it in range(8): tr.iteration_graphed(it)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for it in range(8, 24): tr.iteration_graphed(it)
e1.record(); torch.cuda.synchronize()
print(f"res {res} b {b}: {e0.elapsed_time(e1) / 16:.2f} ms / iteration (graph replay), our launches/iter {sum(tr.launches.values()) / 2.5:.0f}")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for it in range(8): tr.iteration_graphed(it)
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for ev in prof.events():
    if ev.device_type != torch.autograd.DeviceType.CUDA: continue
    n = ev.name.replace("(anonymous namespace)::", "").replace("<unnamed>::", "").replace("void ", "")
    m = re.match(r"(at::native::)?(\w+)(<[^(]{0,60})?", n)
    key = n[:70] if n.startswith("at::") else (m.group(2) if m else n[:40])
    agg[key][0] += 1; agg[key][1] += ev.device_time_total if hasattr(ev, "device_time_total") else ev.cuda_time_total
tot = sum(v[1] for v in agg.values())
print(f"kernels/iter {sum(v[0] for v in agg.values()) / 8:.0f}  summed kernel time {tot / 8e3:.2f} ms/iter")
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    print(f"{t / 8e3:8.3f} ms/iter {c / 8:7.0f}/iter avg {t / c:8.1f} us  {k}")
