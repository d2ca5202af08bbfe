import sys, torch
sys.path.insert(0, '.')
from lcgan_b200 import ops, plans
ops.set_precision("bf16")
dev = 'cuda'
def cl(x): return x.contiguous(memory_format=torch.channels_last)
def timeit(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for (N, Cin, Cout, R) in ((32, 64, 32, 512), (32, 128, 64, 256), (32, 256, 128, 128)):
    x = cl(torch.randn(N, Cin, R, R, device=dev).bfloat16())
    w2 = torch.randn(Cout, 9 * Cin, device=dev).bfloat16(); plan = plans.conv_transpose_up2(3, R, R)
    y = ops.empty_cl(N, Cout, 2 * R, 2 * R, torch.bfloat16, dev); bias = torch.randn(Cout, device=dev)
    rs = torch.rand(N, Cout, device=dev)
    nb = x.numel() * 2 + y.numel() * 2
    r = []
    for fused in (False, True):
        ops.set_up2_fused(fused and Cin <= 128)
        ms = timeit(lambda: ops.tapconv(x, w2, y, plan, rs, bias, None, slope=0.2, gain=1.4))
        r.append(f"{'fused' if fused else '4-phase'} {ms:.3f} ms {nb/ms/1e6:.0f} GB/s")
    print(f"C{Cin}->{Cout} {R}->{2*R}: " + "   ".join(r), flush=True)
    del x, y
