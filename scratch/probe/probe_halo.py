"""Runs scratch/probe/probe_halo.cu (see its header): prints, per channel count and base_offset value,
which of the nine taps the tensor core read correctly from ONE haloed shared-memory tile."""
import ctypes as C
import os
import subprocess
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
so = os.path.join(HERE, "libprobe_halo.so")
if not os.path.exists(so) or "--rebuild" in sys.argv:
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-shared",
                           "-Xcompiler", "-fPIC", os.path.join(HERE, "probe_halo.cu"),
                           os.path.join(ROOT, "lcgan_b200", "csrc", "api.cu"), "-lcuda", "-o", so])
lib = C.CDLL(so)
lib.probe_halo.restype = C.c_int
lib.probe_halo.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_int] * 9 + [C.c_void_p]
lib.lcgan_last_error.restype = C.c_char_p
dev = "cuda"
torch.manual_seed(0)
N, H, W = 2, 32, 32
for Cc in (32, 64):
    x = torch.randn(N, H, W, Cc, device=dev).bfloat16().contiguous()          # NHWC storage
    eye = torch.eye(Cc, device=dev).bfloat16().contiguous()
    for (n0, m0, b) in ((8, 16, 1), (0, 0, 0)):                               # interior tile / corner tile (zero fill)
        xp = torch.zeros(H + 2, W + 2, Cc, device=dev)
        xp[1:-1, 1:-1] = x[b].float()
        for base_off in range(8):
            out = torch.full((9, 128, Cc), float("nan"), device=dev)
            for nrow0 in range(0, Cc, 32):
                part = torch.full((9, 128, 32), float("nan"), device=dev)
                rc = lib.probe_halo(x.data_ptr(), eye.data_ptr(), part.data_ptr(), N, H, W, Cc, n0, m0, b, base_off,
                                    nrow0, torch.cuda.current_stream().cuda_stream)
                assert rc == 0, lib.lcgan_last_error()
                torch.cuda.synchronize()
                out[:, :, nrow0:nrow0 + 32] = part
            ok = []
            for t in range(9):
                dy, dx = t // 3, t % 3
                exp = xp[m0 + dy:m0 + dy + 16, n0 + dx:n0 + dx + 8].reshape(128, Cc)   # padded coords: -1 + 1
                ok.append(bool(torch.equal(out[t], exp)))
            print(f"C={Cc} tile(n0={n0},m0={m0},b={b}) base_offset={base_off}: taps ok = "
                  + "".join("1" if o else "0" for o in ok) + ("   <-- all nine" if all(ok) else ""), flush=True)
