"""2-rank debug of dist_utils.GradExchange: eager first, then inside a CUDA graph; dumps stacks if it hangs."""
import faulthandler, os, sys, time
faulthandler.dump_traceback_later(int(os.environ.get("DUMP_AFTER", "70")), exit=True)
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lcgan_b200.dist_utils import GradExchange

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
mode = sys.argv[1]
torch.manual_seed(0)
net = torch.nn.Sequential(torch.nn.Linear(256, 512), torch.nn.Tanh(), torch.nn.Linear(512, 512), torch.nn.Tanh(),
                          torch.nn.Linear(512, 64)).to(dev)
ex = GradExchange({"n": net}, dev, world, bucket_bytes=256 << 10)
x = torch.randn(32, 256, device=dev) * (rank + 1)

def step():
    net.zero_grad()
    loss = net(x).square().mean()
    ex.backward(loss, "a", "n")
    return loss

def log(*a):
    print(f"[rank {rank}]", *a, flush=True)

s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    step(); log("recording pass done")
    step(); log("bucketed eager pass done")
    torch.cuda.synchronize()
    g1 = [p.grad.clone() for p in net.parameters()]
torch.cuda.current_stream().wait_stream(s)
torch.cuda.synchronize()
log("eager ok, buckets:", len(ex.plans["a"].flats))
if mode == "graph":
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        step()
    log("captured")
    g.replay(); torch.cuda.synchronize()
    log("replayed")
    g2 = [p.grad.clone() for p in net.parameters()]
    log("max diff", max(float((a - b).abs().max()) for a, b in zip(g1, g2)))
dist.barrier()
log("done")
os._exit(0)
