import sys, time, torch
sys.path.insert(0,'.')
from lcgan_b200 import cnn, ops, train_step as T, _lib
from oracle.lcgan_oracle import Config, Hyper
res=int(sys.argv[1]) if len(sys.argv)>1 else 64; b=int(sys.argv[2]) if len(sys.argv)>2 else 8
dev=torch.device('cuda'); cfg=Config(img_resolution=res); hp=Hyper()
torch.manual_seed(0)
G,D=cnn.Generator(cfg.namespace()).to(dev),cnn.Discriminator(cfg.namespace()).to(dev)
tr=T.GraphedTrainer(G,D,hp,b,dev)
for k in tr.z: tr.z[k].normal_()
for k in tr.zd: tr.zd[k].normal_()
for k in tr.data: tr.data[k].uniform_(-1,1)
t0=time.time(); tr.capture(); print("captured in",time.time()-t0, tr.launches)
# compare: eager trainer with same init and inputs for a few iterations
torch.manual_seed(0)
G2,D2=cnn.Generator(cfg.namespace()).to(dev),cnn.Discriminator(cfg.namespace()).to(dev)
G2.load_state_dict(G.state_dict()); D2.load_state_dict(D.state_dict())
tr2=T.Trainer(G2,D2,hp)
# note: optimizer moments differ (graphed trainer had warm-up steps) -> compare loss of the first replay only loosely
for it in range(8):
    tr.replay_g(it); tr.replay_d(it)
torch.cuda.synchronize()
print("graph losses", float(tr.g_loss), float(tr.d_loss))
e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
e0.record()
for it in range(8,24):
    tr.replay_g(it); tr.replay_d(it)
e1.record(); torch.cuda.synchronize()
print(f"graph replay: {e0.elapsed_time(e1)/16:.2f} ms/iter  -> {b*16/(e0.elapsed_time(e1)/1e3):.1f} img/s", "finite", bool(torch.isfinite(tr.g_loss)), bool(torch.isfinite(tr.d_loss)))
for it in range(3): tr2.iteration(it, tr.z, tr.zd, tr.data)
e0.record()
for it in range(8,24): tr2.iteration(it, tr.z, tr.zd, tr.data)
e1.record(); torch.cuda.synchronize()
print(f"eager: {e0.elapsed_time(e1)/16:.2f} ms/iter")
print("peak mem GB", torch.cuda.max_memory_allocated()/1e9)
