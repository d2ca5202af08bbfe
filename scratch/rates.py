import sys, ctypes as C, torch
sys.path.insert(0,'.')
from lcgan_b200 import _lib
lib=_lib.lib()
fn=lib.lcgan_debug_tc_rate
fn.restype=C.c_int
fn.argtypes=[C.c_int,C.c_int,C.c_int,C.c_void_p,C.c_int,C.c_int,C.c_int,C.c_int,C.c_int,C.c_void_p,C.c_int,C.c_void_p]
dev='cuda'
out=torch.zeros(148,dtype=torch.int64,device=dev)
st=C.c_void_p(torch.cuda.current_stream().cuda_stream)
for Cc in (64,32):
    x=torch.randn(32,1024,1024,Cc,device=dev).bfloat16() if Cc==32 else torch.randn(32,512,512,Cc,device=dev).bfloat16()
    H=x.shape[1]
    for mode in (0,1):
        for n in (16,32,64,96,128,192,256):
            if mode==1 and Cc==32: continue
            it=4096
            rc=fn(mode,n,it,x.data_ptr(),32,H,H,Cc,8,out.data_ptr(),148,st); torch.cuda.synchronize()
            assert rc==0, lib.lcgan_last_error()
            print(f"C{Cc} mode{mode} MMA N={n:3d}: {out.float().mean().item()/it:7.1f} cycles/MMA")
    for ht in (8,10):
        for blocks in (1,148):
            it=2048
            rc=fn(2,0,it,x.data_ptr(),32,H,H,Cc,ht,out.data_ptr(),blocks,st); torch.cuda.synchronize()
            assert rc==0, lib.lcgan_last_error()
            cyc=out[:blocks].float().mean().item()/it
            nbytes=16*ht*Cc*2
            print(f"C{Cc} TMA box 16x{ht}x{Cc}ch ({nbytes}B) blocks={blocks}: {cyc:7.1f} cycles/box  {nbytes/cyc:6.1f} B/clk/SM  rows/box={16*ht} -> {cyc/(16*ht):.2f} cyc/row")
    del x
