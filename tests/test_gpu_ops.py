"""Kernel-level parity on the B200: every CUDA op against the plain PyTorch fp32 op it replaces
(same seeded inputs).  Tolerances: fp32 path rel-L2 <= 1e-5 (north_star asks <= 1e-4);
bf16 path <= 1e-2."""
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_l2

pytestmark = pytest.mark.gpu

FP32_TOL = 1e-5
BF16_TOL = 1e-2


@pytest.fixture(autouse=True)
def _setup():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)
    from lcgan_b200 import ops
    ops.set_precision("fp32")
    yield
    ops.set_precision("bf16")


def _cl(x):
    return x.contiguous(memory_format=torch.channels_last)


def _ops():
    from lcgan_b200 import ops, plans
    return ops, plans


CONV_CASES = [  # (k, stride, up, N, Cin, Cout, H, W)
    (3, 1, 1, 2, 16, 24, 8, 8), (3, 2, 1, 2, 16, 24, 8, 8), (1, 1, 1, 3, 8, 5, 4, 6),
    (3, 1, 2, 2, 16, 24, 8, 8), (3, 1, 1, 1, 513, 64, 4, 4), (1, 1, 1, 2, 3, 32, 16, 16),
    (3, 1, 2, 2, 64, 2, 8, 8), (3, 1, 1, 4, 64, 128, 16, 16), (3, 2, 1, 4, 64, 128, 16, 16),
    (3, 1, 2, 2, 128, 2, 16, 8), (3, 1, 2, 1, 512, 2, 4, 4), (3, 1, 2, 3, 32, 2, 5, 7),   # flow layers: fused x2 thin path
]


def _torch_conv(x, w, stride, up):
    if up == 2:
        return F.conv_transpose2d(x, w.transpose(0, 1), stride=2, padding=1, output_padding=1)
    return F.conv2d(x, w, stride=stride, padding=w.shape[-1] // 2)


@pytest.mark.parametrize("case", CONV_CASES)
@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_conv_act_forward_backward(case, mode):
    ops, plans = _ops()
    ops.set_precision(mode)
    k, stride, up, N, Cin, Cout, H, W = case
    dt = torch.float32 if mode == "fp32" else torch.bfloat16
    tol = FP32_TOL if mode == "fp32" else BF16_TOL
    dev = "cuda"
    x = torch.randn(N, Cin, H, W, device=dev)
    w = torch.randn(Cout, Cin, k, k, device=dev)
    bias = torch.randn(Cout, device=dev)
    rs = torch.rand(N, Cout, device=dev) + 0.5
    wscale = 1.0 / (Cin * k * k) ** 0.5
    plan = plans.conv_transpose_up2(k, H, W) if up == 2 else plans.conv(k, stride, H, W)

    xq = x.to(dt).float()        # both sides see the same (possibly bf16-rounded) operands:
    xr, wr, br, rr = (t.clone().requires_grad_() for t in (xq, w, bias, rs))
    wq = wr + (wr.to(dt).float() - wr).detach()   # parameter values rounded, straight-through grad
    ref = F.leaky_relu(_torch_conv(xr, wq * wscale, stride, up) * rr[:, :, None, None] + br[None, :, None, None] * 0.5, 0.2) * 1.4
    g = torch.randn_like(ref)
    gq = g.to(dt).float()
    ref.backward(gq)

    xm = _cl(xq.to(dt)).requires_grad_()
    wm, bm, rm = (t.clone().requires_grad_() for t in (w, bias, rs))
    y = ops.conv_act(xm, wm, bm, rm, None, wscale=wscale, plan=plan, slope=0.2, gain=1.4, bias_scale=0.5)
    assert y.dtype == dt and y.shape == ref.shape
    y.backward(_cl(gq.to(dt)))
    assert rel_l2(y.float(), ref) < tol
    assert rel_l2(xm.grad.float(), xr.grad) < tol * 2
    assert rel_l2(wm.grad, wr.grad) < tol * 2
    assert rel_l2(bm.grad, br.grad) < tol * 2
    assert rel_l2(rm.grad, rr.grad) < tol * 3


def test_conv_residual_and_nchw_io():
    ops, plans = _ops()
    x = torch.randn(2, 3, 16, 16, device="cuda", requires_grad=True)          # NCHW fp32 image
    w = torch.randn(8, 3, 1, 1, device="cuda", requires_grad=True)
    res = _cl(torch.randn(2, 8, 16, 16, device="cuda")).requires_grad_()
    y = ops.conv_act(x, w, None, None, res, wscale=0.3, plan=plans.conv(1, 1, 16, 16), gain=0.7)
    ref = F.conv2d(x, w * 0.3) * 0.7 + res
    assert rel_l2(y, ref) < FP32_TOL
    g = torch.randn_like(ref)
    gx, gw, gr = torch.autograd.grad(y, (x, w, res), _cl(g))
    rx, rw, rr = torch.autograd.grad(ref, (x, w, res), g)
    assert gx.is_contiguous() and rel_l2(gx, rx) < FP32_TOL
    assert rel_l2(gw, rw) < FP32_TOL and rel_l2(gr, rr) < FP32_TOL
    y2 = ops.conv_act(_cl(x.detach()), w, None, None, None, wscale=0.3, plan=plans.conv(1, 1, 16, 16),
                      out_dtype=torch.float32, out_nchw=True)
    assert y2.is_contiguous() and rel_l2(y2, F.conv2d(x, w * 0.3)) < FP32_TOL


def test_linear_act():
    ops, _ = _ops()
    x = torch.randn(5, 300, device="cuda", requires_grad=True)
    w = torch.randn(70, 300, device="cuda", requires_grad=True)
    b = torch.randn(70, device="cuda", requires_grad=True)
    y = ops.linear_act(x, w, b, wscale=0.05, bias_scale=0.01, slope=0.2)
    ref = F.leaky_relu(F.linear(x, w * 0.05, b * 0.01), 0.2)
    assert rel_l2(y, ref) < FP32_TOL
    g = torch.randn_like(ref)
    for a, r in zip(torch.autograd.grad(y, (x, w, b), g), torch.autograd.grad(ref, (x, w, b), g)):
        assert rel_l2(a, r) < FP32_TOL


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("C", [16, 6])
def test_resample_ops(mode, C):
    ops, _ = _ops()
    dt = torch.float32 if mode == "fp32" else torch.bfloat16
    tol = FP32_TOL if mode == "fp32" else BF16_TOL
    x = _cl(torch.randn(2, C, 8, 12, device="cuda").to(dt))
    # NCHW-contiguous fp32 copy for the torch reference: torch's CUDA backward of
    # interpolate(nearest)+avg_pool2d on channels_last inputs disagrees with its own CPU result
    xf = x.float().contiguous()
    assert rel_l2(ops.Box3.apply(x).float(), F.avg_pool2d(xf, 3, 1, 1)) < tol
    assert rel_l2(ops.Pool2.apply(x, 0.25).float(), F.avg_pool2d(xf, 2, 2)) < tol
    assert rel_l2(ops.Up2.apply(x, 1.0).float(), F.interpolate(xf, scale_factor=2, mode="nearest")) < tol
    assert rel_l2(ops.Box3Act.apply(x, 0.2, 1.4).float(), F.leaky_relu(F.avg_pool2d(xf, 3, 1, 1), 0.2) * 1.4) < tol
    t = _cl(torch.randn(2, C, 16, 24, device="cuda").to(dt))
    ref = F.avg_pool2d(F.interpolate(xf, scale_factor=2, mode="nearest"), 3, 1, 1) + t.float()
    assert rel_l2(ops.Up2BoxAdd.apply(x, t).float(), ref) < tol
    # gradients
    xr = xf.clone().requires_grad_(); tr = t.float().contiguous().clone().requires_grad_()
    xm = x.clone().requires_grad_(); tm = t.clone().requires_grad_()
    g = torch.randn(2, C, 16, 24, device="cuda").to(dt)
    (F.avg_pool2d(F.interpolate(xr, scale_factor=2, mode="nearest"), 3, 1, 1) + tr).backward(g.float())
    ops.Up2BoxAdd.apply(xm, tm).backward(_cl(g))
    assert rel_l2(xm.grad.float(), xr.grad) < tol * 2 and rel_l2(tm.grad.float(), tr.grad) < tol
    xr.grad = None; xm.grad = None
    g = torch.randn(2, C, 8, 12, device="cuda").to(dt)
    (F.leaky_relu(F.avg_pool2d(xr, 3, 1, 1), 0.2) * 1.4).backward(g.float())
    ops.Box3Act.apply(xm, 0.2, 1.4).backward(_cl(g))
    assert rel_l2(xm.grad.float(), xr.grad) < tol * 2


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("shape", [(4, 64, 64, 48), (2, 32, 44, 72), (2, 24, 64, 48)])
def test_box_filter_strip_kernel(mode, shape):
    """Shapes of the shared-memory tiled kernel (C % 32 == 0 in bf16 / C % 16 in fp32, W >= 32, H >= 16;
    ragged tiles included) and of the row-reusing strip kernel (H % 8 == 0, many threads)."""
    ops, _ = _ops()
    dt = torch.float32 if mode == "fp32" else torch.bfloat16
    tol = FP32_TOL if mode == "fp32" else BF16_TOL
    x = _cl(torch.randn(*shape, device="cuda").to(dt))
    xf = x.float().contiguous()
    assert rel_l2(ops.Box3.apply(x).float(), F.avg_pool2d(xf, 3, 1, 1)) < tol
    xr = xf.clone().requires_grad_(); xm = x.clone().requires_grad_()
    g = torch.randn(*shape, device="cuda").to(dt)
    ref = F.leaky_relu(F.avg_pool2d(xr, 3, 1, 1), 0.2) * 1.4
    ref.backward(g.float())
    y = ops.Box3Act.apply(xm, 0.2, 1.4)
    y.backward(_cl(g))
    assert rel_l2(y.float(), ref.detach()) < tol
    assert rel_l2(xm.grad.float(), xr.grad) < tol * 2


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("C", [32, 6, 512])
def test_modulate_and_warp(mode, C):
    ops, _ = _ops()
    dt = torch.float32 if mode == "fp32" else torch.bfloat16
    tol = FP32_TOL * 5 if mode == "fp32" else BF16_TOL
    N, H, W = 2, 8, 16
    x = _cl(torch.randn(N, C, H, W, device="cuda").to(dt))
    s = torch.randn(N, C, device="cuda")
    xr = x.float().clone().requires_grad_(); sr = s.clone().requires_grad_()
    xm = x.clone().requires_grad_(); sm = s.clone().requires_grad_()
    g = torch.randn(N, C, H, W, device="cuda").to(dt)
    (xr * sr[:, :, None, None]).backward(g.float())
    ym = ops.Modulate.apply(xm, sm)
    ym.backward(_cl(g))
    assert rel_l2(ym.float(), (xr * sr[:, :, None, None]).detach()) < tol
    assert rel_l2(xm.grad.float(), xr.grad) < tol and rel_l2(sm.grad, sr.grad) < tol * 2

    flow = _cl(torch.randn(N, 2, H, W, device="cuda") * 1.5)
    fr = flow.clone().requires_grad_(); fm = flow.clone().requires_grad_()
    xr.grad = None; xm.grad = None
    ys = 2 * torch.arange(H, device="cuda", dtype=torch.float32) / (H - 1) - 1
    xs = 2 * torch.arange(W, device="cuda", dtype=torch.float32) / (W - 1) - 1
    gy, gx = torch.meshgrid(ys, xs, indexing="ij")
    grid = (torch.stack((gx, gy))[None] + torch.tanh(fr) * 0.1).permute(0, 2, 3, 1)
    ref = F.grid_sample(xr, grid, mode="bicubic", padding_mode="zeros", align_corners=False)
    ref.backward(g.float())
    out = ops.Warp.apply(xm, fm, 0.1)
    out.backward(_cl(g))
    assert rel_l2(out.float(), ref.detach()) < tol
    assert rel_l2(xm.grad.float(), xr.grad) < tol * 2
    assert rel_l2(fm.grad, fr.grad) < (1e-3 if mode == "fp32" else 3e-2)


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("case", [
    # (N, C, H, W, flow std, flow scale, flow mean): eligible for the shared-memory tiled kernels
    # (C % 32, W >= 32, H >= 16); displacement = tanh(flow) * scale * W/2 pixels
    (2, 32, 48, 80, 0.3, 0.1, 0.0),      # small flow, ragged tiles: one window per tile
    (1, 64, 16, 32, 1.5, 0.1, 0.0),      # two channel chunks, one tile
    (1, 32, 48, 96, 0.05, 0.5, 0.8),     # large SMOOTH flow (~16 px / 8 px): displaced single windows
    (1, 32, 48, 96, 0.05, 0.9, -1.5),    # ~ -39 px: most samples leave the image
    (2, 32, 40, 64, 1.5, 0.6, 0.0),      # rough +-19 px: several windows per source tile, per-pixel global fallback
    (1, 32, 32, 64, 1.5, 0.2, 0.0),      # rough +-6 px
    (1, 32, 64, 128, 1.5, 0.9, 0.0),     # rough +-57 px: the dx kernel gives up, scatter kernels take over
])
def test_warp_tiled(mode, case):
    ops, _ = _ops()
    N, C, H, W, fstd, scale, fmean = case
    dt = torch.float32 if mode == "fp32" else torch.bfloat16
    tol = FP32_TOL * 5 if mode == "fp32" else BF16_TOL
    torch.manual_seed(5)
    x = _cl(torch.randn(N, C, H, W, device="cuda").to(dt))
    g = torch.randn(N, C, H, W, device="cuda").to(dt)
    flow = _cl(torch.randn(N, 2, H, W, device="cuda") * fstd + fmean)
    xr = x.float().clone().requires_grad_(); fr = flow.clone().requires_grad_()
    xm = x.clone().requires_grad_(); fm = flow.clone().requires_grad_()
    ys = 2 * torch.arange(H, device="cuda", dtype=torch.float32) / (H - 1) - 1
    xs = 2 * torch.arange(W, device="cuda", dtype=torch.float32) / (W - 1) - 1
    gy, gx = torch.meshgrid(ys, xs, indexing="ij")
    grid = (torch.stack((gx, gy))[None] + torch.tanh(fr) * scale).permute(0, 2, 3, 1)
    ref = F.grid_sample(xr, grid, mode="bicubic", padding_mode="zeros", align_corners=False)
    ref.backward(g.float())
    out = ops.Warp.apply(xm, fm, scale)
    out.backward(_cl(g))
    assert rel_l2(out.float(), ref.detach()) < tol
    assert rel_l2(xm.grad.float(), xr.grad) < tol * 2
    assert rel_l2(fm.grad, fr.grad) < (1e-3 if mode == "fp32" else 3e-2)


def test_loss_kernels():
    ops, _ = _ops()
    from lcgan_b200 import loss
    a, p, n = (torch.randn(6, 256, device="cuda", requires_grad=True) for _ in range(3))
    an, pn, nn_ = (ops.L2Normalize.apply(t) for t in (a, p, n))
    l = loss.contrastive_loss(an, pn, nn_, 0.05)
    ar, pr, nr = (F.normalize(t.detach().clone().requires_grad_()) for t in (a, p, n))
    ep = torch.exp((ar * pr).sum(1) / 0.05); en = torch.exp((ar * nr).sum(1) / 0.05)
    ref = (-torch.log(ep / (ep + en))).mean()
    assert abs(float(l) - float(ref)) < 1e-5 * max(1, abs(float(ref)))
    a2, p2, n2 = (t.detach().clone().requires_grad_() for t in (a, p, n))
    an2, pn2, nn2 = (F.normalize(t) for t in (a2, p2, n2))
    F.softplus(((an2 * nn2).sum(1) - (an2 * pn2).sum(1)) / 0.05).mean().backward()
    l.backward()
    for m, r in ((a, a2), (p, p2), (n, n2)):
        assert rel_l2(m.grad, r.grad) < 1e-4
    x = torch.randn(3, 3 * 32 * 32 + 5, device="cuda", requires_grad=True)
    s = ops.SumSq.apply(x)
    assert rel_l2(s, x.detach().square().sum(1)) < 1e-5
    s.sum().backward()
    assert rel_l2(x.grad, 2 * x.detach()) < 1e-6


def test_ema_multi_tensor():
    ops, _ = _ops()
    src = [torch.randn(n, device="cuda") for n in (1, 7, 1000, 70000)]
    dst = [torch.randn_like(s) for s in src]
    ref = [s.lerp(d, 0.99) for s, d in zip(src, dst)]
    ops.ema_lerp_(dst, src, 0.99)
    for d, r in zip(dst, ref):
        assert rel_l2(d, r) < 1e-6


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_double_backward_through_conv_stack(mode):
    """R1-style: grad of (sum of squared input-gradient) w.r.t. weights, through conv+lrelu, box,
    stride-2 conv, pool/skip - against torch autograd on the same graph."""
    ops, plans = _ops()
    ops.set_precision(mode)
    dt = torch.float32 if mode == "fp32" else torch.bfloat16
    tol = 2e-4 if mode == "fp32" else 5e-2
    N, C, H = 2, 16, 8
    img = torch.randn(N, 3, H, H, device="cuda")
    w0 = torch.randn(C, 3, 1, 1, device="cuda"); w1 = torch.randn(C, C, 3, 3, device="cuda")
    w2 = torch.randn(2 * C, C, 3, 3, device="cuda"); ws = torch.randn(2 * C, C, 1, 1, device="cuda")
    b1 = torch.randn(C, device="cuda")

    def ref_fn(img, w0, w1, w2, ws, b1):
        h = F.leaky_relu(F.conv2d(img, w0 * 0.5), 0.2)
        t = F.leaky_relu(F.conv2d(h, w1 * 0.1, b1, padding=1), 0.2) * 1.4
        t = F.avg_pool2d(t, 3, 1, 1)
        t = F.leaky_relu(F.conv2d(t, w2 * 0.1, stride=2, padding=1), 0.2)
        return F.conv2d(F.avg_pool2d(h, 2), ws * 0.2) * 0.7 + t

    def my_fn(img, w0, w1, w2, ws, b1):
        h = ops.conv_act(img, w0, None, None, None, wscale=0.5, plan=plans.conv(1, 1, H, H), slope=0.2)
        t = ops.conv_act(h, w1, b1, None, None, wscale=0.1, plan=plans.conv(3, 1, H, H), slope=0.2, gain=1.4)
        t = ops.Box3.apply(t)
        t = ops.conv_act(t, w2, None, None, None, wscale=0.1, plan=plans.conv(3, 2, H, H), slope=0.2)
        return ops.conv_act(ops.Pool2.apply(h, 0.25), ws, None, None, t, wscale=0.2,
                            plan=plans.conv(1, 1, H // 2, H // 2), gain=0.7)

    outs = []
    for fn in (ref_fn, my_fn):
        args = [t.clone().requires_grad_() for t in (img, w0, w1, w2, ws, b1)]
        y = fn(*args)
        ctx = ops.no_weight_gradients() if fn is my_fn else torch.enable_grad()
        with ctx:
            (gimg,) = torch.autograd.grad(y.float().sum(), args[0], create_graph=True)
        pen = gimg.square().sum() if fn is ref_fn else ops.SumSq.apply(gimg.reshape(N, -1)).sum()
        grads = torch.autograd.grad(pen + y.float().square().mean(), args[1:])
        outs.append((y.detach().float(), gimg.detach(), grads))
    (yr, gr, wr), (ym, gm, wm) = outs
    assert rel_l2(ym, yr) < tol and rel_l2(gm, gr) < tol
    for a, r in zip(wm, wr):
        assert rel_l2(a, r) < tol * 2
