import sys, torch, json
sys.path.insert(0,'.')
from lcgan_b200 import cnn, ops, train_step as T
from oracle import lcgan_oracle as O
torch.backends.cudnn.allow_tf32=False; torch.backends.cuda.matmul.allow_tf32=False
res,b,steps=32,8,100
dev='cuda'
cfg=O.Config(img_resolution=res); hp=O.Hyper()
def run(mode):
    gen=torch.Generator().manual_seed(5)
    gsd,dsd=O.make_generator_state(cfg,0),O.make_discriminator_state(cfg,1)
    if mode=='oracle':
        tr=O.OracleTrainer(cfg,hp,{k:v.to(dev) for k,v in gsd.items()},{k:v.to(dev) for k,v in dsd.items()})
    else:
        ops.set_precision(mode)
        G,D=cnn.Generator(cfg.namespace()),cnn.Discriminator(cfg.namespace())
        G.load_state_dict(gsd); D.load_state_dict(dsd)
        tr=T.Trainer(G.to(dev),D.to(dev),hp)
    out=[]
    for it in range(steps):
        zg,zd=O.synthetic_latents(b,cfg,gen,dev),O.synthetic_latents(b,cfg,gen,dev)
        data=O.synthetic_data(b,cfg,gen,dev)
        out.append(tr.iteration(it,zg,zd,data))
    return out
r={m:run(m) for m in ('oracle','fp32','bf16')}
json.dump(r,open('gpurun_out/traj.json','w'))
import statistics
for m in ('fp32','bf16'):
    for j,name in ((0,'g'),(1,'d')):
        rel=[abs(a[j]-o[j])/max(abs(o[j]),1e-6) for a,o in zip(r[m],r['oracle'])]
        print(m,name,'max rel',max(rel),'median',statistics.median(rel),'first10 max',max(rel[:10]),'mean-of-100 rel', abs(sum(a[j] for a in r[m])-sum(o[j] for o in r['oracle']))/abs(sum(o[j] for o in r['oracle'])))
print([ (round(o[0],3),round(a[0],3)) for o,a in list(zip(r['oracle'],r['bf16']))[::10]])
print([ (round(o[1],3),round(a[1],3)) for o,a in list(zip(r['oracle'],r['bf16']))[::10]])
