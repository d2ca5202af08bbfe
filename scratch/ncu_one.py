"""One launch of each hot kernel at the 1024^2 shapes (for `ncu --set full`)."""
import sys, torch
sys.path.insert(0, '.')
from lcgan_b200 import ops, plans
ops.set_precision("bf16")
dev = 'cuda'
def cl(x): return x.contiguous(memory_format=torch.channels_last)
which = sys.argv[1:] or ["fwd32", "box", "warpf", "warpb"]
N, C, R = 32, 32, 1024
x = cl(torch.randn(N, C, R, R, device=dev).bfloat16()); g = cl(torch.randn(N, C, R, R, device=dev).bfloat16())
reps = 1
for w in which:
    for _ in range(reps):
        if w == "fwd32":
            w2 = torch.randn(C, 9 * C, device=dev).bfloat16(); plan = plans.conv(3, 1, R, R)
            y = ops.empty_cl(N, C, R, R, torch.bfloat16, dev); bias = torch.randn(C, device=dev)
            ops.tapconv(x, w2, y, plan, None, bias, None, slope=0.2, gain=1.4)
        elif w == "fwd64":
            x64 = cl(torch.randn(N, 64, 512, 512, device=dev).bfloat16())
            w2 = torch.randn(64, 9 * 64, device=dev).bfloat16(); plan = plans.conv(3, 1, 512, 512)
            y = ops.empty_cl(N, 64, 512, 512, torch.bfloat16, dev); bias = torch.randn(64, device=dev)
            ops.tapconv(x64, w2, y, plan, None, bias, None, slope=0.2, gain=1.4)
            del x64, y
        elif w == "fwd128":
            x128 = cl(torch.randn(N, 128, 256, 256, device=dev).bfloat16())
            w2 = torch.randn(128, 9 * 128, device=dev).bfloat16(); plan = plans.conv(3, 1, 256, 256)
            y = ops.empty_cl(N, 128, 256, 256, torch.bfloat16, dev); bias = torch.randn(128, device=dev)
            ops.tapconv(x128, w2, y, plan, None, bias, None, slope=0.2, gain=1.4)
            del x128, y
        elif w == "up64":
            x64 = cl(torch.randn(N, 64, 512, 512, device=dev).bfloat16())
            w2 = torch.randn(32, 9 * 64, device=dev).bfloat16(); plan = plans.conv_transpose_up2(3, 512, 512)
            y = ops.empty_cl(N, 32, 1024, 1024, torch.bfloat16, dev); bias = torch.randn(32, device=dev)
            ops.tapconv(x64, w2, y, plan, None, bias, None)
            del x64, y
        elif w in ("wg32", "wg64"):
            Cw = int(w[2:]); Rw = {32: 1024, 64: 512}[Cw]
            xw = cl(torch.randn(N, Cw, Rw, Rw, device=dev).bfloat16()); gw = cl(torch.randn(N, Cw, Rw, Rw, device=dev).bfloat16())
            ops.tapconv_wgrad(xw, gw, plans.conv(3, 1, Rw, Rw), Cw, Cw)
            del xw, gw
        elif w == "box":
            ops.Box3.apply(x)
        elif w == "actbwd":
            ops._act_bwd_raw(g, x, None, 0.2, 1.4, True, False)
        elif w == "flowf":
            x64 = cl(torch.randn(N, 64, 512, 512, device=dev).bfloat16())
            wt = torch.randn(2, 64, 3, 3, device=dev); plan = plans.conv_transpose_up2(3, 512, 512)
            w2 = ops.pack_weight(wt, False, torch.bfloat16)
            y = ops.empty_cl(N, 2, 1024, 1024, torch.float32, dev)
            ops.tapconv(x64, w2, y, plan, None, None, None)
        elif w == "warpf":
            flow = cl(torch.nn.functional.interpolate(torch.randn(N, 2, R // 64, R // 64, device=dev) * 0.1, size=(R, R), mode='bilinear')); ops.Warp.apply(x, flow, 0.1)
        elif w == "warpb":
            flow = cl(torch.nn.functional.interpolate(torch.randn(N, 2, R // 64, R // 64, device=dev) * 0.1, size=(R, R), mode='bilinear')); xr = x.clone().requires_grad_(); fr = flow.clone().requires_grad_()
            out = ops.Warp.apply(xr, fr, 0.1)
            torch.autograd.grad(out, (xr, fr), g)
        torch.cuda.synchronize()
print("done")
