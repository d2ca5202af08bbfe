import torch, torch.nn.functional as F, sys
sys.path.insert(0,'.')
from lcgan_b200 import ops
torch.manual_seed(0)
def rel(a,b): return float((a.double()-b.double()).norm()/b.double().norm())
cl=lambda x: x.contiguous(memory_format=torch.channels_last)
ops.set_precision("fp32")
C=16
x=cl(torch.randn(2,C,8,12,device='cuda'))
t=cl(torch.randn(2,C,16,24,device='cuda'))
g=cl(torch.randn(2,C,16,24,device='cuda'))
xr=x.clone().requires_grad_()
(F.avg_pool2d(F.interpolate(xr,scale_factor=2,mode='nearest'),3,1,1)).backward(g)
formula=F.avg_pool2d(F.avg_pool2d(g,3,1,1),2,2)*4
print('torch formula vs autograd', rel(formula,xr.grad))
b3=ops.Box3.apply(g); print('box3', rel(b3,F.avg_pool2d(g,3,1,1)))
p2=ops.Pool2.apply(b3,1.0); print('pool2(box3)', rel(p2,formula))
xm=x.clone().requires_grad_(); tm=t.clone().requires_grad_()
out=ops.Up2BoxAdd.apply(xm,tm); out.backward(g)
print('Up2BoxAdd ds', rel(xm.grad,xr.grad), 'dt', rel(tm.grad,g))
# conv cases
from lcgan_b200 import plans
import os
for mode in ('bf16',):
    ops.set_precision(mode)
    for case in [(1,1,1,2,3,32,16,16),(3,1,1,4,64,128,16,16),(3,2,1,4,64,128,16,16)]:
        k,stride,up,N,Cin,Cout,H,W=case
        dt=torch.bfloat16
        x=torch.randn(N,Cin,H,W,device='cuda'); w=torch.randn(Cout,Cin,k,k,device='cuda'); bias=torch.randn(Cout,device='cuda'); rs=torch.rand(N,Cout,device='cuda')+.5
        wscale=1/(Cin*k*k)**.5
        plan=plans.conv(k,stride,H,W)
        xq=x.to(dt).float()
        xr,wr,br,rr=(t.clone().requires_grad_() for t in (xq,w,bias,rs))
        ref=F.leaky_relu(F.conv2d(xr,wr*wscale,stride=stride,padding=k//2)*rr[:,:,None,None]+br[None,:,None,None]*.5,.2)*1.4
        g=torch.randn_like(ref); gq=g.to(dt).float(); ref.backward(gq)
        xm=cl(xq.to(dt)).requires_grad_(); wm,bm,rm=(t.clone().requires_grad_() for t in (w,bias,rs))
        y=ops.conv_act(xm,wm,bm,rm,None,wscale=wscale,plan=plan,slope=.2,gain=1.4,bias_scale=.5)
        y.backward(cl(gq.to(dt)))
        print(case,'y',rel(y.float(),ref),'dx',rel(xm.grad.float(),xr.grad),'dw',rel(wm.grad,wr.grad),'db',rel(bm.grad,br.grad),'drs',rel(rm.grad,rr.grad))
