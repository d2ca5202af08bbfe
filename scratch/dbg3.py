import torch, sys
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
from lcgan_b200 import ops, cnn
from oracle import lcgan_oracle as O
torch.backends.cudnn.allow_tf32=False; torch.backends.cuda.matmul.allow_tf32=False
def rel(a,b): return float((a.double().cpu()-b.double().cpu()).norm()/b.double().cpu().norm())
for mode in ("fp32","bf16"):
    ops.set_precision(mode); dt=torch.float32 if mode=="fp32" else torch.bfloat16
    cfg=O.Config(img_resolution=64); gsd=O.make_generator_state(cfg,5); dsd=O.make_discriminator_state(cfg,6)
    G=cnn.Generator(cfg.namespace()); G.load_state_dict(gsd); G=G.cuda()
    D=cnn.Discriminator(cfg.namespace()); D.load_state_dict(dsd); D=D.cuda()
    def q(k,v):
        return v.to(dt).float().cuda() if (k.endswith("weight.weight") and v.dim()==4) else v.cuda()
    gc={k:q(k,v) for k,v in gsd.items()}; dc={k:q(k,v) for k,v in dsd.items()}
    torch.manual_seed(1); b=4
    glat,alat=torch.randn(b,64,device='cuda'),torch.randn(b,512,device='cuda')
    for i,(cin,cout,res) in enumerate(cfg.g_channels()):
        x=torch.randn(b,cin,res//2,res//2,device='cuda').to(dt).float(); xo=x.clone().requires_grad_()
        for v in gc.values(): v.requires_grad_(True); v.grad=None
        yo=O.synthesis_block(gc,f"model.{i}",xo,glat,alat,cfg.max_flow_scale); gy=torch.randn_like(yo).to(dt).float(); yo.backward(gy)
        xm=x.to(dt).contiguous(memory_format=torch.channels_last).requires_grad_(); G.zero_grad()
        ym=G.model[i](xm,glat[:,None],alat[:,None].expand(-1,2,-1)); ym.backward(gy.to(dt).contiguous(memory_format=torch.channels_last))
        errs={k:rel(p.grad,gc[f"model.{i}.{k}"].grad) for k,p in G.model[i].named_parameters()}
        print(mode,f"G{i} fwd {rel(ym,yo):.1e} dx {rel(xm.grad,xo.grad):.1e}",' '.join(f"{k.replace('modulated_','m').replace('.weight.weight','.w')}={v:.1e}" for k,v in errs.items()))
    for i,(cin,cout) in enumerate(cfg.d_channels()):
        res=cfg.img_resolution>>i
        x=torch.randn(b,cin,res,res,device='cuda').to(dt).float(); xo=x.clone().requires_grad_()
        for v in dc.values(): v.requires_grad_(True); v.grad=None
        yo=O.discriminator_block(dc,f"shared_model.{i+2}",xo); gy=torch.randn_like(yo).to(dt).float(); yo.backward(gy)
        xm=x.to(dt).contiguous(memory_format=torch.channels_last).requires_grad_(); D.zero_grad()
        ym=D.shared_model[i+2](xm); ym.backward(gy.to(dt).contiguous(memory_format=torch.channels_last))
        errs={k:rel(p.grad,dc[f"shared_model.{i+2}.{k}"].grad) for k,p in D.shared_model[i+2].named_parameters()}
        print(mode,f"D{i} fwd {rel(ym,yo):.1e} dx {rel(xm.grad,xo.grad):.1e}",' '.join(f"{k.replace('.weight.weight','.w')}={v:.1e}" for k,v in errs.items()))
