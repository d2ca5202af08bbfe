"""Host logic on CPU: the tap-conv plans reproduce F.conv2d / F.conv_transpose2d / F.linear and
their gradients (against torch autograd), for the strides/kernels the reference uses."""
import pytest
import torch
import torch.nn.functional as F

from lcgan_b200 import plans
from lcgan_b200.ops import unpack_wgrad
import tapconv_emu as E

torch.manual_seed(0)


def _rand(*s):
    return torch.randn(*s, dtype=torch.float64)


@pytest.mark.parametrize("k,stride,H,W", [(3, 1, 8, 8), (3, 2, 8, 8), (1, 1, 4, 6), (3, 1, 4, 4), (3, 2, 16, 8)])
def test_conv_plan_forward_dgrad_wgrad(k, stride, H, W):
    x, w = _rand(2, 5, H, W).requires_grad_(), _rand(7, 5, k, k).requires_grad_()
    ref = F.conv2d(x, w, stride=stride, padding=k // 2)
    p = plans.conv(k, stride, H, W)
    assert (p.OH, p.OW) == tuple(ref.shape[2:])
    torch.testing.assert_close(E.tapconv(x.detach(), E.pack(w.detach(), False), p), ref.detach())
    g = _rand(*ref.shape)
    dx, dw = torch.autograd.grad(ref, (x, w), g)
    torch.testing.assert_close(E.tapconv(g, E.pack(w.detach(), True), plans.adjoint(p)), dx)
    torch.testing.assert_close(unpack_wgrad(E.wgrad(x.detach(), g, p, 5, 7), tuple(w.shape), False), dw)


@pytest.mark.parametrize("H,W", [(4, 4), (8, 6)])
def test_up2_plan_forward_dgrad_wgrad(H, W):
    x, w = _rand(2, 5, H, W).requires_grad_(), _rand(7, 5, 3, 3).requires_grad_()
    ref = F.conv_transpose2d(x, w.transpose(0, 1), stride=2, padding=1, output_padding=1)
    p = plans.conv_transpose_up2(3, H, W)
    assert (p.OH, p.OW) == tuple(ref.shape[2:]) == (2 * H, 2 * W)
    assert sorted(len(l.taps) for l in p.launches) == [1, 2, 2, 4]
    torch.testing.assert_close(E.tapconv(x.detach(), E.pack(w.detach(), False), p), ref.detach())
    g = _rand(*ref.shape)
    dx, dw = torch.autograd.grad(ref, (x, w), g)
    adj = plans.adjoint(p)
    assert len(adj.launches) == 1 and adj.launches[0].is_ == 2 and len(adj.launches[0].taps) == 9
    torch.testing.assert_close(E.tapconv(g, E.pack(w.detach(), True), adj), dx)
    torch.testing.assert_close(unpack_wgrad(E.wgrad(x.detach(), g, p, 5, 7), tuple(w.shape), False), dw)


def test_adjoint_is_involution_on_maps():
    """adjoint(adjoint(p)) computes the same map as p (launch decomposition may differ)."""
    x, w = _rand(1, 3, 8, 8), _rand(4, 3, 3, 3)
    for p in (plans.conv(3, 1, 8, 8), plans.conv(3, 2, 8, 8), plans.conv_transpose_up2(3, 8, 8)):
        pp = plans.adjoint(plans.adjoint(p))
        torch.testing.assert_close(E.tapconv(x, E.pack(w, False), pp), E.tapconv(x, E.pack(w, False), p))


def test_transposed_wgrad_convention():
    """ConvWgrad(transposed=True): dw[o,c,t] = sum x[p+t, o] g[p, c] (backward of the dgrad node)."""
    w = _rand(7, 5, 3, 3).requires_grad_()
    gy = _rand(2, 7, 4, 4)                                   # plays x of the transposed conv
    p = plans.adjoint(plans.conv(3, 2, 8, 8))                # dgrad plan of a stride-2 conv
    out = E.tapconv(gy, E.pack(w, True), p)                  # [2,5,8,8]
    gg = _rand(*out.shape)
    (dw,) = torch.autograd.grad(out, w, gg)
    dw2 = E.wgrad(gy, gg, p, 7, 5)
    torch.testing.assert_close(unpack_wgrad(dw2, tuple(w.shape), True), dw)


def test_linear_plan():
    x, w = _rand(3, 10), _rand(6, 10)
    y = E.tapconv(x[:, :, None, None], E.pack(w, False), plans.linear())[:, :, 0, 0]
    torch.testing.assert_close(y, x @ w.t())


def test_macs_model_counts_transposed_conv_without_zero_insert():
    # 2.25 * Cin * Cout MAC per output pixel
    p = plans.conv_transpose_up2(3, 16, 16)
    assert plans.macs(p, 1, 1, 1) == int(2.25 * 32 * 32)


def test_fused_up2_weight_layout_reproduces_conv_transpose():
    """The single-launch form of the x2 transposed conv (ops.fused_up2_weights + blocked output channels):
    a 4-tap conv over the input lattice with 4*Cout phase-major channels equals F.conv_transpose2d."""
    import torch.nn.functional as F
    from lcgan_b200 import ops
    torch.manual_seed(0)
    N, Cin, Cout, H, W = 2, 5, 3, 4, 6
    x = torch.randn(N, Cin, H, W, dtype=torch.float64)
    w = torch.randn(Cout, Cin, 3, 3, dtype=torch.float64)
    ref = F.conv_transpose2d(x, w.transpose(0, 1), stride=2, padding=1, output_padding=1)
    w2 = w.permute(0, 2, 3, 1).reshape(Cout, 9 * Cin)                       # pack_weight layout [o][t*Cin + c]
    wf = ops.fused_up2_weights(w2, Cin).view(4, Cout, 4, Cin)               # [phase][o][tap][c]
    xp = F.pad(x, (0, 1, 0, 1))                                             # taps read x[m+dy, n+dx], zero beyond
    out = torch.zeros(N, Cout, 2 * H, 2 * W, dtype=torch.float64)
    for py in (0, 1):
        for px in (0, 1):
            acc = torch.zeros(N, Cout, H, W, dtype=torch.float64)
            for dy in (0, 1):
                for dx in (0, 1):
                    acc += torch.einsum("nchw,oc->nohw", xp[:, :, dy:dy + H, dx:dx + W], wf[py * 2 + px, :, dy * 2 + dx, :])
            out[:, :, py::2, px::2] = acc
    assert torch.allclose(out, ref, atol=1e-12)
