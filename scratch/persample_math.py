"""Design check for DESIGN.md section 9 item 4 (CPU, fp64): with per-sample weights w_b = w * s[b] the
modulated conv needs no modulate / modulate_bwd passes - the per-sample weight gradient
dW_b = wgrad(x_b, g_b * d_b) (UNmodulated x) yields both
    dL/dw = sum_b dW_b * s[b]          and          dL/ds[b,c] = sum_{o,k} dW_b[o,c,k] * w[o,c,k]
(conv term only; the demodulation term d(s) keeps its existing closed form)."""
import torch
import torch.nn.functional as F

torch.manual_seed(0)
B, Ci, Co, H = 3, 5, 4, 6
x = torch.randn(B, Ci, H, H, dtype=torch.float64, requires_grad=True)
w = torch.randn(Co, Ci, 3, 3, dtype=torch.float64, requires_grad=True)
s = torch.randn(B, Ci, dtype=torch.float64, requires_grad=True)
d = torch.rand(B, Co, dtype=torch.float64) + 0.5           # demod coefficients, held constant here
g = torch.randn(B, Co, H, H, dtype=torch.float64)

y = F.conv2d(x * s[:, :, None, None], w, padding=1) * d[:, :, None, None]      # shared-weight form (today)
gx, gw, gs = torch.autograd.grad(y, (x, w, s), g)

gh = g * d[:, :, None, None]
dx2 = torch.stack([F.conv_transpose2d(gh[b:b + 1], w * s[b][None, :, None, None], padding=1)[0] for b in range(B)])
dWb = torch.stack([torch.autograd.grad(F.conv2d(x[b:b + 1].detach(), wb, padding=1), wb, gh[b:b + 1])[0]
                   for b, wb in ((b, w.detach().clone().requires_grad_()) for b in range(B))])
gw2 = (dWb * s.detach()[:, None, :, None, None]).sum(0)
gs2 = (dWb * w.detach()[None]).sum(dim=(1, 3, 4))
for name, a, b_ in (("dx", gx, dx2), ("dw", gw, gw2), ("ds", gs, gs2)):
    print(f"{name}: max |diff| = {(a - b_).abs().max().item():.3e}")
    assert torch.allclose(a, b_, atol=1e-10)
print("per-sample-weight formulation reproduces autograd")
