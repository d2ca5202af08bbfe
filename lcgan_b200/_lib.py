"""ctypes binding of liblcgan_b200.so (the C ABI declared in include/lcgan_b200.h).

The library is loaded lazily at module level and never stored on an nn.Module, so modules stay
picklable (worker.py:40 deep-copies a DDP-wrapped generator).  There is no CPU fallback: if the
shared library is missing, every op raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liblcgan_b200.so")
_SRCS = ["api.cu", "conv_simt.cu", "conv_tc.cu", "thin.cu", "resample.cu", "warp.cu", "loss.cu", "optim.cu", "demod.cu"]
_lock = threading.Lock()
_lib = None

F32, BF16 = 0, 1
MAX_TAPS = 9


class TapConvDesc(C.Structure):
    """Mirror of `struct lcgan_tapconv` (include/lcgan_b200.h)."""
    _fields_ = [
        ("N", C.c_int32), ("IH", C.c_int32), ("IW", C.c_int32), ("Cin", C.c_int32),
        ("OH", C.c_int32), ("OW", C.c_int32), ("Cout", C.c_int32),
        ("xs_n", C.c_int64), ("xs_h", C.c_int64), ("xs_w", C.c_int64), ("xs_c", C.c_int64),
        ("ys_n", C.c_int64), ("ys_h", C.c_int64), ("ys_w", C.c_int64), ("ys_c", C.c_int64),
        ("x_dtype", C.c_int32), ("y_dtype", C.c_int32), ("w_dtype", C.c_int32),
        ("MH", C.c_int32), ("MW", C.c_int32),
        ("os", C.c_int32), ("py", C.c_int32), ("px", C.c_int32),
        ("is_", C.c_int32),
        ("ntaps", C.c_int32),
        ("dy", C.c_int32 * MAX_TAPS), ("dx", C.c_int32 * MAX_TAPS), ("wtap", C.c_int32 * MAX_TAPS),
        ("w_ld", C.c_int64),
        ("acc_scale", C.c_float), ("bias_scale", C.c_float), ("slope", C.c_float), ("gain", C.c_float),
        ("noise", C.c_void_p), ("noise_scale", C.c_float),
        ("colscale", C.c_void_p),
    ]


MT_MAX = 48


class AdamChunk(C.Structure):
    """Mirror of `struct lcgan_adam_chunk`."""
    _fields_ = [("p", C.c_void_p * MT_MAX), ("g", C.c_void_p * MT_MAX), ("m", C.c_void_p * MT_MAX),
                ("v", C.c_void_p * MT_MAX), ("step", C.c_void_p * MT_MAX), ("numel", C.c_int64 * MT_MAX),
                ("count", C.c_int32)]


class PackChunk(C.Structure):
    """Mirror of `struct lcgan_pack_chunk`."""
    _fields_ = [("src", C.c_void_p * MT_MAX), ("dst", C.c_void_p * MT_MAX), ("O", C.c_int32 * MT_MAX),
                ("I", C.c_int32 * MT_MAX), ("K", C.c_int32 * MT_MAX), ("mode", C.c_int32 * MT_MAX),
                ("scale", C.c_float * MT_MAX), ("count", C.c_int32)]


def build(verbose: bool = False) -> str:
    """Compile csrc/*.cu for sm_100a into lcgan_b200/liblcgan_b200.so (nvcc cross-compiles
    without a GPU).  One object per source (compiled in parallel, rebuilt only when the source or a
    header is newer), then one link."""
    from concurrent.futures import ThreadPoolExecutor
    src_dir = os.path.join(_HERE, "csrc")
    obj_dir = os.path.join(src_dir, "_build")
    os.makedirs(obj_dir, exist_ok=True)
    hdrs = [os.path.join(src_dir, f) for f in os.listdir(src_dir) if f.endswith(".cuh")]
    hdrs.append(os.path.join(os.path.dirname(_HERE), "include", "lcgan_b200.h"))
    hdr_time = max(os.path.getmtime(h) for h in hdrs)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    flags = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC"]
    srcs = [os.path.join(src_dir, n) for n in _SRCS]
    if os.path.exists(_SO) and all(os.path.getmtime(_SO) >= os.path.getmtime(d) for d in srcs + hdrs):
        return _SO
    jobs, objs = [], []
    for name in _SRCS:
        src, obj = os.path.join(src_dir, name), os.path.join(obj_dir, name[:-3] + ".o")
        objs.append(obj)
        if not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), hdr_time):
            jobs.append([nvcc] + flags + ["-c", src, "-o", obj])

    def run(cmd):
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.run(cmd, check=True)

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        list(ex.map(run, jobs))
    run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", _SO] + objs)
    return _SO


_VOIDP = C.c_void_p
_FP = C.c_void_p   # float* passed as raw address
_SIGS = {
    "lcgan_version": ([], C.c_int),
    "lcgan_tapconv_tc_eligible": ([C.POINTER(TapConvDesc)], C.c_int),
    "lcgan_tapconv_up2_thin_eligible": ([C.POINTER(TapConvDesc)], C.c_int),
    "lcgan_tapconv_up2_thin": ([C.POINTER(TapConvDesc), _VOIDP, _VOIDP, _VOIDP, _FP, _FP, _VOIDP], C.c_int),
    "lcgan_tapconv_up2_thin_wgrad": ([C.POINTER(TapConvDesc), _VOIDP, _VOIDP, _FP, C.c_float, _VOIDP], C.c_int),
    "lcgan_tapconv_tc_blocked": ([C.POINTER(TapConvDesc), _VOIDP, _VOIDP, _VOIDP, _FP, _FP, C.c_int, C.c_int64, C.c_int,
                                  _VOIDP], C.c_int),
    "lcgan_tapconv_simt": ([C.POINTER(TapConvDesc), _VOIDP, _VOIDP, _VOIDP, _FP, _FP, _VOIDP, _VOIDP], C.c_int),
    "lcgan_tapconv_tc": ([C.POINTER(TapConvDesc), _VOIDP, _VOIDP, _VOIDP, _FP, _FP, _VOIDP, _VOIDP], C.c_int),
    "lcgan_flow_grad_im2col": ([_FP, _VOIDP, C.c_int, C.c_int, C.c_int, _VOIDP], C.c_int),
    "lcgan_pw_wgrad32": ([C.POINTER(TapConvDesc), _VOIDP, _VOIDP, _FP, _VOIDP], C.c_int),
    "lcgan_tapconv_wgrad_simt": ([C.POINTER(TapConvDesc), _VOIDP, _VOIDP, _FP, C.c_float, _VOIDP], C.c_int),
    "lcgan_tapconv_wgrad_tc": ([C.POINTER(TapConvDesc), _VOIDP, _VOIDP, _FP, C.c_float, _VOIDP], C.c_int),
    "lcgan_box3": ([_VOIDP, _VOIDP, _VOIDP, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                    C.c_float, C.c_float, C.c_float, C.c_float, _VOIDP], C.c_int),
    "lcgan_box3_postmask": ([_VOIDP, _VOIDP, _VOIDP, _FP, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                             C.c_float, C.c_float, _VOIDP], C.c_int),
    "lcgan_box3_cs": ([_VOIDP, _VOIDP, _VOIDP, _FP, _FP, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                       C.c_float, C.c_float, C.c_float, C.c_float, _VOIDP], C.c_int),
    "lcgan_pool2": ([_VOIDP, _VOIDP, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, _VOIDP], C.c_int),
    "lcgan_up2": ([_VOIDP, _VOIDP, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, _VOIDP], C.c_int),
    "lcgan_debug_tc_rate": ([C.c_int, C.c_int, C.c_int, _VOIDP, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _VOIDP, C.c_int,
                             _VOIDP], C.c_int),
    "lcgan_up2_add": ([_VOIDP, _VOIDP, _VOIDP, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, _VOIDP], C.c_int),
    "lcgan_box3_pool2": ([_VOIDP, _VOIDP, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _VOIDP], C.c_int),
    "lcgan_up2box_add": ([_VOIDP, _VOIDP, _VOIDP, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _VOIDP], C.c_int),
    "lcgan_act_bwd": ([_VOIDP, _VOIDP, _VOIDP, _FP, _FP, _FP, C.c_int, C.c_int, C.c_int, C.c_int,
                       C.c_float, C.c_float, _VOIDP], C.c_int),
    "lcgan_modulate": ([_VOIDP, _FP, _VOIDP, C.c_int, C.c_int, C.c_int, C.c_int, _VOIDP], C.c_int),
    "lcgan_modulate_bwd": ([_VOIDP, _VOIDP, _FP, _VOIDP, _FP, C.c_int, C.c_int, C.c_int, C.c_int, _VOIDP], C.c_int),
    "lcgan_warp_fwd": ([_VOIDP, _FP, _VOIDP, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, _VOIDP], C.c_int),
    "lcgan_warp_fwd_cs": ([_VOIDP, _FP, _VOIDP, _FP, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, _VOIDP], C.c_int),
    "lcgan_warp_bwd": ([_VOIDP, _FP, _VOIDP, _FP, _FP, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                        C.c_float, _VOIDP], C.c_int),
    "lcgan_warp_bwd_tiled": ([_VOIDP, _FP, _VOIDP, _VOIDP, _FP, _FP, _VOIDP, C.c_int, C.c_int, C.c_int, C.c_int,
                              C.c_int, C.c_float, _VOIDP], C.c_int),
    "lcgan_cast": ([_VOIDP, _VOIDP, C.c_int, C.c_int, C.c_int64, _VOIDP], C.c_int),
    "lcgan_l2norm_fwd": ([_FP, _FP, _FP, C.c_int, C.c_int, _VOIDP], C.c_int),
    "lcgan_l2norm_bwd": ([_FP, _FP, _FP, _FP, C.c_int, C.c_int, _VOIDP], C.c_int),
    "lcgan_contrastive_fwd": ([_FP, _FP, _FP, _FP, _FP, C.c_int, C.c_int, C.c_float, _VOIDP], C.c_int),
    "lcgan_contrastive_bwd": ([_FP, _FP, _FP, _FP, _FP, _FP, _FP, _FP, C.c_int, C.c_int, C.c_float, _VOIDP], C.c_int),
    "lcgan_sumsq": ([_FP, _FP, C.c_int, C.c_int64, _VOIDP], C.c_int),
    "lcgan_rowscale": ([_FP, _FP, _FP, C.c_int, C.c_int64, _VOIDP], C.c_int),
    "lcgan_ema_lerp": ([_VOIDP, _VOIDP, _VOIDP, C.c_int, C.c_float, _FP, _VOIDP], C.c_int),
    "lcgan_adam_step": ([C.POINTER(AdamChunk), C.c_float, C.c_float, C.c_float, C.c_float, _VOIDP], C.c_int),
    "lcgan_pack_weights": ([C.POINTER(PackChunk), C.c_int, _VOIDP], C.c_int),
    "lcgan_set_deterministic": ([C.c_int], C.c_int),
    "lcgan_demod_fwd": ([_FP, _FP, _FP, C.c_int, C.c_int, C.c_int, C.c_float, _VOIDP], C.c_int),
    "lcgan_demod_bwd": ([_FP, _FP, _FP, _FP, _FP, _FP, _FP, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int,
                         _VOIDP], C.c_int),
    "lcgan_epilogue_grads": ([_FP, _FP, _FP, _FP, C.c_float, _FP, _FP, C.c_int, C.c_int, _VOIDP], C.c_int),
}
EXPORTS = tuple(_SIGS) + ("lcgan_last_error",)


def lib():
    """The loaded library (raises if it has not been built)."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(_SO):
                    raise RuntimeError(
                        f"{_SO} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(lcgan_b200 has no CPU fallback)")
                l = C.CDLL(_SO)
                for name, (args, res) in _SIGS.items():
                    fn = getattr(l, name)
                    fn.argtypes, fn.restype = args, res
                l.lcgan_last_error.argtypes, l.lcgan_last_error.restype = [], C.c_char_p
                _lib = l
    return _lib


def check(rc: int, what: str):
    if rc != 0:
        raise RuntimeError(f"{what} failed (code {rc}): {lib().lcgan_last_error().decode()}")


# launch counters: bench.py reports how many of OUR kernels ran inside the timed region; `counts` keeps
# them per entry point (tests assert e.g. that no weight-gradient kernel runs inside cal_derivative)
launches = 0
counts = {}


_prof = None


def profile_begin():
    """Start recording a CUDA-event pair around every launch (bench.py roofline pass)."""
    global _prof
    _prof = []


def profile_end():
    """Stop recording; returns {tag: {n, ms, flops, bytes}} (synchronises the device)."""
    global _prof
    import torch
    torch.cuda.synchronize()
    out = {}
    for tag, e0, e1, flops, nbytes in _prof:
        s = out.setdefault(tag, {"n": 0, "ms": 0.0, "flops": 0.0, "bytes": 0.0})
        s["n"] += 1
        s["ms"] += e0.elapsed_time(e1)
        s["flops"] += flops
        s["bytes"] += nbytes
    _prof = None
    return out


def call(name: str, *args, flops=0, nbytes=0, tag=None):
    """Launch one C-ABI entry point on the caller's current stream.  flops / nbytes are the
    ALGORITHMIC work of the launch, used only by the profiling pass."""
    global launches
    launches += 1
    counts[name] = counts.get(name, 0) + 1
    if _prof is None:
        check(getattr(lib(), name)(*args), name)
        return
    import torch
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    check(getattr(lib(), name)(*args), name)
    e1.record()
    _prof.append((tag or name.replace("lcgan_", ""), e0, e1, flops, nbytes))
