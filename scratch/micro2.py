"""One launch each of the memory-bound kernels at 1024^2-model shapes (for ncu)."""
import sys, torch
sys.path.insert(0, '.')
from lcgan_b200 import ops, plans
ops.set_precision("bf16")
dev='cuda'; N=32
def cl(x): return x.contiguous(memory_format=torch.channels_last)
def t(fn, name, nbytes):
    fn(); torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1); print(f"{name:22s} {ms:8.3f} ms {nbytes/ms/1e6:8.0f} GB/s")
# flow forward C64 -> 2, up2 (4 phase launches), fp32 out
x64=cl(torch.randn(N,64,512,512,device=dev).bfloat16()); w=torch.randn(2,9*64,device=dev).bfloat16()
yf=ops.empty_cl(N,2,1024,1024,torch.float32,dev); rs=torch.rand(N,2,device=dev)+.5; b2=torch.randn(2,device=dev)
t(lambda: ops.tapconv(x64,w,yf,plans.conv_transpose_up2(3,512,512),rs,b2), "flow_fwd C64->2 up2", x64.numel()*2*4+yf.numel()*4)
# flow dgrad C2 -> 64 (is2, 9 taps), g fp32
g2=cl(torch.randn(N,2,1024,1024,device=dev)); wT=torch.randn(64,9*2,device=dev).bfloat16(); dx=ops.empty_cl(N,64,512,512,torch.bfloat16,dev)
t(lambda: ops.tapconv(g2,wT,dx,plans.adjoint(plans.conv_transpose_up2(3,512,512))), "flow_dgrad C2->64", g2.numel()*4+dx.numel()*2)
# flow wgrad
t(lambda: ops.tapconv_wgrad(x64,g2,plans.conv_transpose_up2(3,512,512),64,2), "flow_wgrad C64->2", (x64.numel()*2+g2.numel()*4/4)*4)
# from-RGB fwd 3->32 (NCHW fp32 in), toRGB 32->3 (NCHW fp32 out)
img=torch.randn(N,3,1024,1024,device=dev); w3=torch.randn(32,3,device=dev); y32=ops.empty_cl(N,32,1024,1024,torch.bfloat16,dev); b32=torch.randn(32,device=dev)
t(lambda: ops.tapconv(img,w3,y32,plans.conv(1,1,1024,1024),None,b32,None,slope=0.2), "fromRGB 3->32", img.numel()*4+y32.numel()*2)
x32=cl(torch.randn(N,32,1024,1024,device=dev).bfloat16()); w32=torch.randn(3,32,device=dev).bfloat16(); rgb=torch.empty(N,3,1024,1024,device=dev)
t(lambda: ops.tapconv(x32,w32,rgb,plans.conv(1,1,1024,1024),torch.rand(N,3,device=dev),torch.randn(3,device=dev)), "toRGB 32->3", x32.numel()*2+rgb.numel()*4)
g32=cl(torch.randn(N,32,1024,1024,device=dev).bfloat16())
t(lambda: ops.Box3.apply(x32), "box3", 2*x32.numel()*2)
t(lambda: ops._act_bwd_raw(g32,x32,None,0.2,1.4,True,False), "act_bwd", 3*x32.numel()*2)
t(lambda: ops._modulate_raw(x32, torch.rand(N,32,device=dev)), "modulate", 2*x32.numel()*2)
flow=cl(torch.randn(N,2,1024,1024,device=dev)*0.3)
t(lambda: ops.Warp.apply(x32,flow,0.1), "warp_fwd", 2*x32.numel()*2)
xr=x32.clone().requires_grad_(); fr=flow.clone().requires_grad_(); out=ops.Warp.apply(xr,fr,0.1)
t(lambda: torch.autograd.grad(out,(xr,fr),g32,retain_graph=True), "warp_bwd(+zero+cast)", 4*x32.numel()*2)
t(lambda: ops.Up2BoxAdd.apply(cl(torch.empty(N,32,512,512,device=dev).bfloat16()), x32), "up2box_add", 2.25*x32.numel()*2)
