import sys, torch
sys.path.insert(0, '.')
from lcgan_b200 import cnn, ops
from oracle.lcgan_oracle import Config
res = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
cfg = Config(img_resolution=res)
torch.manual_seed(0)
dev = 'cuda'
G = cnn.Generator(cfg.namespace()).to(dev)
orig = ops.Warp.apply
def spy(y, flow, scale):
    H, W = flow.shape[2], flow.shape[3]
    d = torch.tanh(flow.float()) * scale * W / 2          # displacement in pixels, [b,2,H,W] logical
    dx = (d[:, :, :, 1:] - d[:, :, :, :-1]).abs()
    dy = (d[:, :, 1:, :] - d[:, :, :-1, :]).abs()
    print(f"res {H:5d} C {y.shape[1]:4d}: |disp| mean {d.abs().mean():7.3f} max {d.abs().max():7.3f} px; "
          f"neighbour diff x mean {dx.mean():6.3f} max {dx.max():6.3f}; y mean {dy.mean():6.3f}", flush=True)
    return orig(y, flow, scale)
ops.Warp.apply = spy
with torch.no_grad():
    G(torch.randn(2, 64, device=dev), torch.randn(2, 64, device=dev))
