"""Pure-torch emulation of the tap-convolution contraction (include/lcgan_b200.h) used to check
the host-side plans on CPU: same lattice / tap / weight-pack semantics as the CUDA kernels."""
import torch


def pack(w, transposed):
    if w.dim() == 2:
        w = w[:, :, None, None]
    perm = (1, 2, 3, 0) if transposed else (0, 2, 3, 1)
    return w.permute(*perm).reshape(w.shape[perm[0]], -1).contiguous()


def _gather(x, l, dy, dx):
    """x [N,C,IH,IW] -> [N,C,MH,MW] sampled at (m*is+dy, n*is+dx), zero outside."""
    N, C, IH, IW = x.shape
    m = torch.arange(l.MH) * l.is_ + dy
    n = torch.arange(l.MW) * l.is_ + dx
    ok = ((m >= 0) & (m < IH))[:, None] & ((n >= 0) & (n < IW))[None, :]
    v = x[:, :, m.clamp(0, IH - 1)][:, :, :, n.clamp(0, IW - 1)]
    return v * ok.to(x.dtype)


def tapconv(x, w2, plan):
    N, cin = x.shape[:2]
    cout = w2.shape[0]
    y = torch.zeros(N, cout, plan.OH, plan.OW, dtype=x.dtype)
    for l in plan.launches:
        acc = torch.zeros(N, cout, l.MH, l.MW, dtype=x.dtype)
        for dy, dx, wt in l.taps:
            acc += torch.einsum("nchw,oc->nohw", _gather(x, l, dy, dx), w2[:, wt * cin:(wt + 1) * cin])
        y[:, :, l.py::l.os, l.px::l.os][:, :, :l.MH, :l.MW] = acc
    return y


def wgrad(x, g, plan, cin, cout):
    dw2 = torch.zeros(cout, plan.k * plan.k * cin, dtype=x.dtype)
    for l in plan.launches:
        gl = g[:, :, l.py::l.os, l.px::l.os][:, :, :l.MH, :l.MW]
        for dy, dx, wt in l.taps:
            dw2[:, wt * cin:(wt + 1) * cin] += torch.einsum("nohw,nchw->oc", gl, _gather(x, l, dy, dx))
    return dw2
