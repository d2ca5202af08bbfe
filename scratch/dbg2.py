import torch, sys
sys.path.insert(0,'.')
from lcgan_b200 import ops, cnn
from oracle import lcgan_oracle as O
torch.backends.cudnn.allow_tf32=False; torch.backends.cuda.matmul.allow_tf32=False
def rel(a,b): return float((a.double().cpu()-b.double().cpu()).norm()/b.double().cpu().norm())
ops.set_precision("fp32")
cfg=O.Config(img_resolution=64); gsd=O.make_generator_state(cfg,5)
G=cnn.Generator(cfg.namespace()); G.load_state_dict(gsd); G=G.cuda()
torch.manual_seed(1); b=4
glat,alat=torch.randn(b,64),torch.randn(b,512)
for i,(cin,cout,res) in enumerate(cfg.g_channels()[:3]):
    x=torch.randn(b,cin,res//2,res//2); gy=torch.randn(b,cout,res,res)
    outs={}
    for name,dev,dt in (('cpu64','cpu',torch.float64),('gpu32','cuda',torch.float32)):
        sd={k:v.to(dev).to(dt).requires_grad_(True) for k,v in gsd.items()}
        xo=x.to(dev).to(dt).requires_grad_()
        yo=O.synthesis_block(sd,f"model.{i}",xo,glat.to(dev).to(dt),alat.to(dev).to(dt),cfg.max_flow_scale)
        yo.backward(gy.to(dev).to(dt))
        outs[name]=(yo.detach(),xo.grad,{k:v.grad for k,v in sd.items() if v.grad is not None})
    xm=x.cuda().contiguous(memory_format=torch.channels_last).requires_grad_()
    G.zero_grad()
    ym=G.model[i](xm,glat.cuda()[:,None],alat.cuda()[:,None].expand(-1,2,-1))
    ym.backward(gy.cuda().contiguous(memory_format=torch.channels_last))
    t=outs['cpu64']
    print(f"block {i}: fwd mine {rel(ym,t[0]):.2e} gpu-oracle {rel(outs['gpu32'][0],t[0]):.2e} | dx mine {rel(xm.grad,t[1]):.2e} gpu-oracle {rel(outs['gpu32'][1],t[1]):.2e}")
    for k,p in G.model[i].named_parameters():
        kk=f"model.{i}.{k}"
        print(f"    {k:50s} mine {rel(p.grad,t[2][kk]):.2e} gpu-oracle {rel(outs['gpu32'][2][kk],t[2][kk]):.2e}")
