"""Micro-benchmark of individual launches at real 1024^2 shapes (CUDA events, L2 flushed by size)."""
import os, sys, torch
sys.path.insert(0, '.')
from lcgan_b200 import ops, plans, _lib
ops.set_precision("bf16")
dev = 'cuda'
def cl(x): return x.contiguous(memory_format=torch.channels_last)
def timeit(fn, n=5):
    fn(); torch.cuda.synchronize()
    if os.environ.get("PROF"):
        from torch.profiler import profile, ProfilerActivity
        with profile(activities=[ProfilerActivity.CUDA]) as pr:
            for _ in range(n): fn()
            torch.cuda.synchronize()
        for e in sorted(pr.key_averages(), key=lambda e: -e.device_time_total)[:8]:
            print(f"      {e.device_time_total / n / 1e3:8.3f} ms  x{e.count // n:3d}  {e.key[:90]}")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
which = sys.argv[1:] or ["fwd32", "wg32", "fwd64", "wg64", "fwd128", "box", "actbwd", "warpf", "warpb"]
N = 32
import os
FS = float(os.environ.get('FLOW_STD', '0.1'))   # tanh(0.1)*0.1*512 = 5 px std, smooth (64-px correlation)
for w in which:
    if w.startswith("fwd") or w.startswith("wg"):
        C = int(w[3:] if w.startswith("fwd") else w[2:]); R = {32: 1024, 64: 512, 128: 256, 256: 128}[C]
        x = cl(torch.randn(N, C, R, R, device=dev).bfloat16()); g = cl(torch.randn(N, C, R, R, device=dev).bfloat16())
        w2 = torch.randn(C, 9 * C, device=dev).bfloat16(); plan = plans.conv(3, 1, R, R)
        y = ops.empty_cl(N, C, R, R, torch.bfloat16, dev); bias = torch.randn(C, device=dev)
        flops = 2.0 * N * R * R * 9 * C * C; nbytes = 2 * x.numel() * 2
        if w.startswith("fwd"):
            ms = timeit(lambda: ops.tapconv(x, w2, y, plan, None, bias, None, slope=0.2, gain=1.4))
        else:
            ms = timeit(lambda: ops.tapconv_wgrad(x, g, plan, C, C))
        print(f"{w:8s} {ms:8.3f} ms  {flops/ms/1e9:8.1f} TF/s  {nbytes/ms/1e6:8.0f} GB/s (algorithmic)")
    elif w == "up2box":
        sk = cl(torch.randn(N, 32, 512, 512, device=dev).bfloat16()); t = cl(torch.randn(N, 32, 1024, 1024, device=dev).bfloat16())
        ms = timeit(lambda: ops.Up2BoxAdd.apply(sk, t)); tr = (sk.numel() + 2 * t.numel()) * 2
        print(f"{w:8s} {ms:8.3f} ms  {tr/ms/1e6:8.0f} GB/s (algorithmic)")
    elif w == "skinny":
        x = torch.randn(32, 512, device=dev); wt = torch.randn(512, 512, device=dev); y = torch.empty(32, 512, device=dev); bias = torch.randn(512, device=dev)
        x4, y4 = x.view(32, 512, 1, 1), y.view(32, 512, 1, 1)
        ms = timeit(lambda: ops.tapconv(x4, wt, y4, plans.conv(1, 1, 1, 1), None, bias, None, slope=0.2, gain=1.4), n=50)
        print(f"{w:8s} {ms*1e3:8.1f} us per launch (512 -> 512, 32 rows, fp32)")
        g = torch.randn(32, 512, device=dev).view(32, 512, 1, 1)
        ms = timeit(lambda: ops.tapconv_wgrad(x4, g, plans.conv(1, 1, 1, 1), 512, 512), n=50)
        print(f"{w:8s} {ms*1e3:8.1f} us per launch (wgrad, incl. the zero fill)")
    elif w == "s2_32":
        x = cl(torch.randn(N, 32, 1024, 1024, device=dev).bfloat16()); w2 = torch.randn(64, 9 * 32, device=dev).bfloat16()
        y = ops.empty_cl(N, 64, 512, 512, torch.bfloat16, dev); bias = torch.randn(64, device=dev)
        ms = timeit(lambda: ops.tapconv(x, w2, y, plans.conv(3, 2, 1024, 1024), None, bias, None, slope=0.2, gain=1.4)); tr = x.numel() * 2 + y.numel() * 2
        print(f"{w:8s} {ms:8.3f} ms  {tr/ms/1e6:8.0f} GB/s (algorithmic)")
    elif w == "res32":
        x = cl(torch.randn(N, 32, 512, 512, device=dev).bfloat16()); w2 = torch.randn(64, 32, device=dev).bfloat16()
        y = ops.empty_cl(N, 64, 512, 512, torch.bfloat16, dev); res = cl(torch.randn(N, 64, 512, 512, device=dev).bfloat16())
        ms = timeit(lambda: ops.tapconv(x, w2, y, plans.conv(1, 1, 512, 512), None, None, res, gain=0.7)); tr = x.numel() * 2 + 2 * y.numel() * 2
        print(f"{w:8s} {ms:8.3f} ms  {tr/ms/1e6:8.0f} GB/s (algorithmic)")
    elif w == "up64":
        x = cl(torch.randn(N, 64, 512, 512, device=dev).bfloat16()); wp = torch.nn.Parameter(torch.randn(32, 64, 3, 3, device=dev))
        w2 = ops.pack_weight(wp, False, torch.bfloat16); wf = ops._derive(wp, ("up2f", torch.bfloat16)); plan = plans.conv_transpose_up2(3, 512, 512)
        y = ops.empty_cl(N, 32, 1024, 1024, torch.bfloat16, dev); bias = torch.randn(32, device=dev); rs = torch.rand(N, 32, device=dev)
        ms = timeit(lambda: ops.tapconv(x, w2, y, plan, rs, bias, None, up2f=wf)); tr = x.numel() * 2 + y.numel() * 2
        print(f"{w:8s} {ms:8.3f} ms  {tr/ms/1e6:8.0f} GB/s (algorithmic)")
    elif w in ("rgbout", "rgbin"):
        R = 1024
        if w == "rgbout":
            x = cl(torch.randn(N, 32, R, R, device=dev).bfloat16()); w2 = torch.randn(3, 32, device=dev).bfloat16()
            y = torch.empty(N, 3, R, R, device=dev); rs = torch.rand(N, 3, device=dev) + 0.5; bias = torch.randn(3, device=dev)
            ms = timeit(lambda: ops.tapconv(x, w2, y, plans.conv(1, 1, R, R), rs, bias, None)); tr = x.numel() * 2 + y.numel() * 4
        else:
            x = torch.randn(N, 3, R, R, device=dev); w2 = torch.randn(32, 3, device=dev)
            y = ops.empty_cl(N, 32, R, R, torch.bfloat16, dev); bias = torch.randn(32, device=dev)
            ms = timeit(lambda: ops.tapconv(x, w2, y, plans.conv(1, 1, R, R), None, bias, None, slope=0.2)); tr = x.numel() * 4 + y.numel() * 2
        print(f"{w:8s} {ms:8.3f} ms  {tr/ms/1e6:8.0f} GB/s (algorithmic)")
    elif w in ("flowf", "flowd", "floww"):
        C, R = 64, 512
        x = cl(torch.randn(N, C, R, R, device=dev).bfloat16())
        wt = torch.randn(2, C, 3, 3, device=dev); plan = plans.conv_transpose_up2(3, R, R)
        if w == "flowf":
            w2 = ops.pack_weight(wt, False, torch.bfloat16)
            y = ops.empty_cl(N, 2, 2 * R, 2 * R, torch.float32, dev)
            wp = torch.nn.Parameter(wt)
            wf = ops._derive(wp, ("up2f", torch.bfloat16)) if os.environ.get("LCGAN_NO_FLOW_TC") != "1" else None
            ms = timeit(lambda: ops.tapconv(x, w2, y, plan, None, None, None, up2f=wf)); tr = x.numel() * 2 + y.numel() * 4
        elif w == "flowd":
            g = cl(torch.randn(N, 2, 2 * R, 2 * R, device=dev)); w2 = ops.pack_weight(wt, True, torch.bfloat16)
            y = ops.empty_cl(N, C, R, R, torch.bfloat16, dev)
            ms = timeit(lambda: ops.tapconv(g, w2, y, plans.adjoint(plan))); tr = y.numel() * 2 + g.numel() * 4
        else:
            g = cl(torch.randn(N, 2, 2 * R, 2 * R, device=dev))
            ops._FLOW_TC = os.environ.get("LCGAN_NO_FLOW_TC") != "1"
            ms = timeit(lambda: ops.tapconv_wgrad(x, g, plan, C, 2)); tr = x.numel() * 2 + g.numel() * 4
        print(f"{w:8s} {ms:8.3f} ms  {tr/ms/1e6:8.0f} GB/s (algorithmic)")
    else:
        C, R = 32, 1024
        x = cl(torch.randn(N, C, R, R, device=dev).bfloat16()); g = cl(torch.randn(N, C, R, R, device=dev).bfloat16())
        nb = x.numel() * 2
        if w == "box": ms = timeit(lambda: ops.Box3.apply(x)); tr = 2 * nb
        elif w == "modb":
            sc = torch.randn(N, C, device=dev); ds = torch.zeros(N, C, device=dev); dx = torch.empty_like(x)
            from lcgan_b200 import _lib as L
            ms = timeit(lambda: L.call("lcgan_modulate_bwd", ops._ptr(x), ops._ptr(g), ops._ptr(sc), ops._ptr(dx), ops._ptr(ds), ops._dt(x), N, R * R, C, ops._stream(x))); tr = 3 * nb
        elif w == "actbwd": ms = timeit(lambda: ops._act_bwd_raw(g, x, None, 0.2, 1.4, True, False)); tr = 3 * nb
        elif w == "warpf":
            flow = cl(torch.nn.functional.interpolate(torch.randn(N, 2, R // 64, R // 64, device=dev) * FS, size=(R, R), mode='bilinear')); ms = timeit(lambda: ops.Warp.apply(x, flow, 0.1)); tr = 2 * nb
        elif w == "warpb":
            flow = cl(torch.nn.functional.interpolate(torch.randn(N, 2, R // 64, R // 64, device=dev) * FS, size=(R, R), mode='bilinear')); xr = x.clone().requires_grad_(); fr = flow.clone().requires_grad_()
            out = ops.Warp.apply(xr, fr, 0.1)
            ms = timeit(lambda: torch.autograd.grad(out, (xr, fr), g, retain_graph=True)); tr = 4 * nb
        print(f"{w:8s} {ms:8.3f} ms  {tr/ms/1e6:8.0f} GB/s (algorithmic)")
