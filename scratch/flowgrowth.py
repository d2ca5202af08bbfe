"""How fast do the flow displacements grow during training? (decides the tiled warp kernels' range)"""
import sys, torch
sys.path.insert(0, '.')
from lcgan_b200 import cnn, ops, train_step as T
from oracle.lcgan_oracle import Config, Hyper
res, b, iters = 1024, 8, 57
cfg = Config(img_resolution=res); hp = Hyper(lr=1e-3)
torch.manual_seed(0); dev = torch.device('cuda')
ops.set_precision("bf16")
G, D = cnn.Generator(cfg.namespace()).to(dev), cnn.Discriminator(cfg.namespace()).to(dev)
tr = T.Trainer(G, D, hp)
orig = ops.Warp.apply
stats = {}
def spy(y, flow, scale):
    with torch.no_grad():
        H, W = flow.shape[2], flow.shape[3]
        d = (torch.tanh(flow.float()) * scale * W / 2).abs()
        s = stats.setdefault(W, [0.0, 0.0, 0.0, 0.0])
        s[0] = max(s[0], float(d.max())); s[1] = max(s[1], float(d.mean()))
        s[2] = max(s[2], float((d > 6.0).float().mean())); s[3] = max(s[3], float((d > 3.0).float().mean()))
    return orig(y, flow, scale)
ops.Warp.apply = spy
g = torch.Generator().manual_seed(1)
data = {k: (torch.rand(b, 3, res, res, generator=g) * 2 - 1).to(dev) for k in ("image", "geometry_change", "appearance_change")}
for it in range(iters):
    z = {k: torch.randn(b, 64, device=dev) for k in ("rand1", "rand2", "resample1", "resample2")}
    zd = {k: torch.randn(b, 64, device=dev) for k in ("rand1", "rand2")}
    stats.clear()
    tr.g_step(it, z); tr.ema.update(it); tr.d_step(it, zd, data)
    if it % 8 == 0 or it < 4:
        print(f"it {it:3d}: " + "  ".join(f"W{W}: max {s[0]:.2f} mean {s[1]:.2f} >6px {100*s[2]:.3f}% >3px {100*s[3]:.2f}%" for W, s in sorted(stats.items()) if W >= 128), flush=True)
