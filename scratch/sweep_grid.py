import os, sys, torch
cap = sys.argv[1]
os.environ["LCGAN_GRID_CAP"] = cap
sys.path.insert(0, '.')
from lcgan_b200 import ops
ops.set_precision("bf16")
dev = 'cuda'
def cl(x): return x.contiguous(memory_format=torch.channels_last)
def timeit(fn, n=6):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
N, C, R = 32, 32, 1024
x = cl(torch.randn(N, C, R, R, device=dev).bfloat16()); nb = x.numel() * 2
s = torch.randn(N, C, device=dev)
lo = cl(torch.randn(N, C, R // 2, R // 2, device=dev).bfloat16())
r = []
ms = timeit(lambda: ops._modulate_raw(x, s)); r.append(f"modulate {ms:.3f} ms {2*nb/ms/1e6:.0f} GB/s")
ms = timeit(lambda: ops.Up2BoxAdd.apply(lo, x)); r.append(f"up2box_add {ms:.3f} ms {2.25*nb/ms/1e6:.0f} GB/s")
ms = timeit(lambda: ops.Up2.apply(lo, 1.0)); r.append(f"up2 {ms:.3f} ms {1.25*nb/ms/1e6:.0f} GB/s")
ms = timeit(lambda: ops.Pool2.apply(x, 0.25)); r.append(f"pool2 {ms:.3f} ms {1.25*nb/ms/1e6:.0f} GB/s")
print(f"cap {cap:>3s}: " + "   ".join(r), flush=True)
