import os, sys, torch
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
for (N, C, R) in ((32, 32, 1024), (32, 64, 512), (32, 256, 128)):
    x = cl(torch.randn(N, C, R, R, device=dev).bfloat16()); g = cl(torch.randn(N, C, R, R, device=dev).bfloat16())
    nb = x.numel() * 2
    from lcgan_b200 import _lib as L
    sc = torch.randn(N, C, device=dev); ds = torch.zeros(N, C, device=dev); dx = torch.empty_like(x)
    for bps in ("1", "2", "3", "4", "6"):
        os.environ["LCGAN_ACT_STAGES"] = "4"; os.environ["LCGAN_ACT_BPS"] = bps
        ms = timeit(lambda: ops._act_bwd_raw(g, x, None, 0.2, 1.4, True, False))
        ms2 = timeit(lambda: L.call("lcgan_modulate_bwd", ops._ptr(x), ops._ptr(g), ops._ptr(sc), ops._ptr(dx), ops._ptr(ds), ops._dt(x), N, R * R, C, ops._stream(x)))
        print(f"C{C} R{R} bps {bps:>2s}: act_bwd {ms:7.3f} ms {3*nb/ms/1e6:7.0f} GB/s   modulate_bwd {ms2:7.3f} ms {3*nb/ms2/1e6:7.0f} GB/s", flush=True)
    del x, g
