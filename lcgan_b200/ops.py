"""torch.autograd.Functions over the lcgan_b200 C ABI.

Everything activation-sized runs in our CUDA kernels; torch is used for memory (torch.empty),
streams (the caller's current stream is passed to every launch, so DDP's bucket hooks order
correctly) and for a few weight-/[b,C]-sized scalar ops.

Second order (R1, reference loss.py:18-34): the discriminator-side Functions implement their
backward by *calling other Functions* (ConvFwd <-> ConvWgrad, ActBwd, Box3, Pool2 <-> Up2), so
autograd.grad(create_graph=True) builds a graph whose nodes are again our kernels - the
conv2d_gradfix pattern.  Generator-only Functions (Modulate, Warp, Up2BoxAdd, Box3Act) are
once-differentiable.
"""
from __future__ import annotations

import contextlib
import ctypes as C
import math
import os
import threading
import weakref

import torch
from torch.autograd.function import once_differentiable

from . import _lib, plans
from ._lib import BF16, F32, TapConvDesc

# LCGAN_PRECISION=fp32 selects the accurate mode at import (the reference's unmodified main.py has no flag for it)
_ACT_DTYPE = torch.float32 if os.environ.get("LCGAN_PRECISION", "bf16") == "fp32" else torch.bfloat16
_USE_TC = os.environ.get("LCGAN_DISABLE_TC", "0") != "1"


def set_precision(mode: str):
    """'bf16': bf16 activations, tcgen05 tensor-core convs (fp32 accumulate).
    'fp32': fp32 activations and CUDA-core kernels (the <=1e-4 parity mode)."""
    global _ACT_DTYPE
    assert mode in ("bf16", "fp32")
    _ACT_DTYPE = torch.bfloat16 if mode == "bf16" else torch.float32


def get_precision() -> str:
    return "bf16" if _ACT_DTYPE == torch.bfloat16 else "fp32"


def act_dtype():
    return _ACT_DTYPE


def set_tensor_cores(flag: bool):
    global _USE_TC
    _USE_TC = bool(flag)


# A plain module-level flag, NOT a threading.local: backward nodes of CUDA tensors run on the autograd
# engine's device thread, not on the thread that entered the context manager (conv2d_gradfix keeps a
# module global for the same reason).
_NO_WGRAD = False


@contextlib.contextmanager
def no_weight_gradients():
    """Skip weight/bias gradients inside (used for the R1 first-order pass, which only needs
    d logit / d image; same role as conv2d_gradfix.no_weight_gradients)."""
    global _NO_WGRAD
    old = _NO_WGRAD
    _NO_WGRAD = True
    try:
        yield
    finally:
        _NO_WGRAD = old


def _wgrad_enabled():
    return not _NO_WGRAD


def _dt(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise TypeError(f"lcgan_b200: unsupported dtype {t.dtype}")


def _stream(t: torch.Tensor):
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _need_cuda(*ts):
    """Every op is a launch on the CURRENT device with a stream taken from its tensors: refuse CPU tensors (there is no
    fallback) and tensors of another device (one process per GPU, as in the reference)."""
    cur = None
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("lcgan_b200 ops need CUDA tensors (there is no CPU fallback)")
        if cur is None:
            cur = torch.cuda.current_device()
        if t.device.index != cur:
            raise RuntimeError(f"lcgan_b200 ops launch on the current device (cuda:{cur}); got a tensor on {t.device} - "
                               "call torch.cuda.set_device() first (one process per GPU)")


def _cl(x: torch.Tensor, dtype=None) -> torch.Tensor:
    """Dense channels-last view/copy of a logical NCHW tensor in `dtype`."""
    dtype = dtype or x.dtype
    if x.dtype != dtype:
        x = x.to(dtype)
    return x.contiguous(memory_format=torch.channels_last)


def _is_cl(x):
    n, c, h, w = x.shape
    return x.stride() == (h * w * c, 1, w * c, c) or x.is_contiguous(memory_format=torch.channels_last)


def empty_cl(n, c, h, w, dtype, device):
    return torch.empty((n, c, h, w), dtype=dtype, device=device, memory_format=torch.channels_last)


def _strides_nhwc(t):
    """element strides (n, h, w, c) of a logical NCHW tensor"""
    s = t.stride()
    return s[0], s[2], s[3], s[1]


# ------------------------------------------------------------------------------------------
# weight packing and derived weight tables (cached per parameter version)
# ------------------------------------------------------------------------------------------
_pack_cache = {}      # (id(w), kind) -> (tag, tensor);  kind = ("pack", transposed, dtype) | ("wsq", dtype, c)
_demand = {}          # id(w) -> set of kinds ever requested for that parameter (what prepack refreshes)
_pack_lock = threading.Lock()
# Bumped by everything that rewrites parameters through raw pointers (ema_lerp_, the fused Adam kernel,
# CUDA-graph replays that contain an optimizer step): those writes do not move Tensor._version, so the
# generation is part of every cache tag.
_generation = 0


def bump_generation():
    """Invalidate every cached weight pack / derived weight tensor."""
    global _generation
    _generation += 1


def _tag(w):
    return (w._version, w.data_ptr(), _generation)


def _forget(wid):
    with _pack_lock:
        for k in [k for k in _pack_cache if k[0] == wid]:
            _pack_cache.pop(k, None)
        _demand.pop(wid, None)


def _cache_get(w, kind):
    with _pack_lock:
        hit = _pack_cache.get((id(w), kind))
        if hit is not None and hit[0] == _tag(w):
            return hit[1]
    return None


def _cache_put(w, kind, t):
    with _pack_lock:
        if id(w) not in _demand:
            weakref.finalize(w, _forget, id(w))
            _demand[id(w)] = set()
        _demand[id(w)].add(kind)
        _pack_cache[(id(w), kind)] = (_tag(w), t)


def _kernel_packable(w, dtype):
    return (w.is_cuda and w.dtype == torch.float32 and w.is_contiguous() and w.dim() in (2, 4)
            and dtype in (torch.float32, torch.bfloat16))


def _pack_shape(w, kind):
    o, i = w.shape[0], w.shape[1]
    k = w.shape[2] * w.shape[3] if w.dim() == 4 else 1
    if kind[0] == "wsq":
        return (o, i), torch.float32
    if kind[0] == "up2f":
        return (4 * o, 4 * i), kind[1]
    return ((i, k * o) if kind[1] else (o, k * i)), kind[2]


def _kind_dtype(kind):
    """compute dtype of a cached kind: ("pack", transposed, dtype) | ("wsq", dtype, c) | ("up2f", dtype)"""
    return kind[2] if kind[0] == "pack" else kind[1]


def _launch_packs(entries):
    """entries: [(w, kind, dst)] -> lcgan_pack_weights launches of up to MT_MAX tensors per compute dtype."""
    by_dt = {}
    for e in entries:
        by_dt.setdefault(_kind_dtype(e[1]), []).append(e)
    for dt, lst in by_dt.items():
        for i0 in range(0, len(lst), _lib.MT_MAX):
            part = lst[i0:i0 + _lib.MT_MAX]
            ch = _lib.PackChunk()
            for j, (w, kind, dst) in enumerate(part):
                ch.src[j], ch.dst[j] = w.data_ptr(), dst.data_ptr()
                ch.O[j], ch.I[j] = w.shape[0], w.shape[1]
                ch.K[j] = w.shape[2] * w.shape[3] if w.dim() == 4 else 1
                if kind[0] == "wsq":
                    ch.mode[j], ch.scale[j] = 2, kind[2]
                elif kind[0] == "up2f":
                    ch.mode[j], ch.scale[j] = 3, 1.0
                else:
                    ch.mode[j], ch.scale[j] = (1 if kind[1] else 0), 1.0
            ch.count = len(part)
            _lib.call("lcgan_pack_weights", C.byref(ch), BF16 if dt == torch.bfloat16 else F32, _stream(part[0][0]),
                      tag="pack_weights", nbytes=sum(w.numel() * 6 for w, _, _ in part))


def _derive(w, kind):
    """One cached weight-derived tensor, computed by our pack kernel (torch fallback for odd inputs)."""
    cacheable = isinstance(w, torch.nn.Parameter)
    if cacheable:
        hit = _cache_get(w, kind)
        if hit is not None:
            return hit
    wd = w.detach()
    dtype = _kind_dtype(kind)
    if _kernel_packable(wd, dtype) and (kind[0] != "up2f" or (wd.dim() == 4 and wd.shape[2] == wd.shape[3] == 3)):
        shape, odt = _pack_shape(wd, kind)
        dst = torch.empty(shape, dtype=odt, device=wd.device)
        _launch_packs([(wd, kind, dst)])
    elif kind[0] == "up2f":
        dst = fused_up2_weights(_derive(w, ("pack", False, dtype)), wd.shape[1])
    elif kind[0] == "pack":
        w4 = wd[:, :, None, None] if wd.dim() == 2 else wd
        perm = (1, 2, 3, 0) if kind[1] else (0, 2, 3, 1)
        dst = w4.permute(*perm).reshape(w4.shape[perm[0]], -1).to(dtype).contiguous()
    else:
        w4 = wd[:, :, None, None] if wd.dim() == 2 else wd
        dst = (w4.to(dtype).float() * kind[2]).square().sum(dim=(2, 3))
    if cacheable:
        _cache_put(w, kind, dst)
    return dst


def pack_weight(w: torch.Tensor, transposed: bool, dtype: torch.dtype) -> torch.Tensor:
    """W2[o][tap*Cin + c] = w[o, c, kh, kw]  (transposed: rows = c, contraction over o), cast to the
    compute dtype.  The equalized-lr constant is NOT folded in: the parameter values themselves are
    rounded to bf16 and the constant is applied to the fp32 accumulator (lcgan_tapconv.acc_scale).
    Packs of nn.Parameters are cached until the parameter changes (version counter, or the generation
    bumped by raw-pointer writers); the cache entry dies with the parameter."""
    if w.dim() == 2 and not transposed and w.dtype == dtype and w.is_contiguous():
        return w.detach()                        # a plain [O][I] matrix already is its forward pack
    return _derive(w, ("pack", bool(transposed), dtype))


def weight_sq(w: torch.Tensor, wscale: float, dtype: torch.dtype) -> torch.Tensor:
    """Wsq[o,c] = sum_k (round_dtype(w[o,c,k]) * wscale)^2 (f32) - the demodulation table of
    custom_layers.py:65-67, from the weights as the conv sees them."""
    return _derive(w, ("wsq", dtype, float(wscale)))


def prepack(*modules):
    """Refresh, in a few multi-tensor launches, every cached pack / table of the modules' parameters that
    is stale (optimizer step, EMA, load_state_dict).  Only kinds requested before are refreshed, so the
    first iteration packs lazily and later ones in bulk."""
    entries = []
    for m in modules:
        for w in m.parameters():
            kinds = _demand.get(id(w))
            if not kinds or not w.is_cuda:
                continue
            wd = w.detach()
            for kind in kinds:
                dtype = _kind_dtype(kind)
                if _cache_get(w, kind) is None and _kernel_packable(wd, dtype):
                    shape, odt = _pack_shape(wd, kind)
                    dst = torch.empty(shape, dtype=odt, device=wd.device)
                    entries.append((wd, kind, dst))
                    _cache_put(w, kind, dst)
    if entries:
        _launch_packs(entries)


def clear_pack_cache():
    """Drop every cached weight pack (called around CUDA-graph capture so that each graph records
    its own packing kernels and no eager path keeps a graph-pool tensor).  The demand sets survive."""
    with _pack_lock:
        _pack_cache.clear()


def unpack_wgrad(dw2: torch.Tensor, wshape, transposed: bool) -> torch.Tensor:
    if len(wshape) == 2:
        o, i = wshape
        return dw2.t() if transposed else dw2
    o, i, kh, kw = wshape
    if kh == 1 and kw == 1:      # keep plain contiguous strides (DDP's bucket views expect them)
        return (dw2.t().contiguous() if transposed else dw2).reshape(wshape)
    if transposed:
        return dw2.view(i, kh, kw, o).permute(3, 0, 1, 2)
    return dw2.view(o, kh, kw, i).permute(0, 3, 1, 2)


# ------------------------------------------------------------------------------------------
# raw launches
# ------------------------------------------------------------------------------------------
def _fill_desc(d: TapConvDesc, l: plans.Launch, x, y, cin, cout, w2, slope, gain, bias_scale, acc_scale=1.0,
               noise=None, colscale=None):
    n = x.shape[0]
    d.N, d.IH, d.IW, d.Cin = n, x.shape[2], x.shape[3], cin
    d.OH, d.OW, d.Cout = y.shape[2], y.shape[3], cout
    d.xs_n, d.xs_h, d.xs_w, d.xs_c = _strides_nhwc(x)
    d.ys_n, d.ys_h, d.ys_w, d.ys_c = _strides_nhwc(y)
    d.x_dtype, d.y_dtype = _dt(x), _dt(y)
    d.w_dtype = _dt(w2) if w2 is not None else F32
    d.MH, d.MW, d.os, d.py, d.px, d.is_ = l.MH, l.MW, l.os, l.py, l.px, l.is_
    d.ntaps = len(l.taps)
    for t, (dy, dx, wt) in enumerate(l.taps):
        d.dy[t], d.dx[t], d.wtap[t] = dy, dx, wt
    d.w_ld = w2.shape[1] if w2 is not None else len(l.taps) * cin
    d.acc_scale, d.bias_scale, d.slope, d.gain = acc_scale, bias_scale, slope, gain
    d.noise, d.noise_scale = (noise.data_ptr(), 1.0) if noise is not None else (None, 0.0)
    d.colscale = colscale.data_ptr() if colscale is not None else None


_PROFILE_SHAPES = os.environ.get("LCGAN_PROFILE_SHAPES", "0") == "1"


def set_profile_shapes(flag: bool):
    """Tag tap-conv launches with their shape in the profiling pass (bench.py reports the dominant shape)."""
    global _PROFILE_SHAPES
    _PROFILE_SHAPES = bool(flag)


def _shape_tag(fn, d):
    name = fn.replace("lcgan_", "")
    if not _PROFILE_SHAPES:
        return name
    return (f"{name}|N{d.N} {d.IH}x{d.IW}->{d.OH}x{d.OW} C{d.Cin}->{d.Cout} taps{d.ntaps} is{d.is_} os{d.os} "
            f"x{d.x_dtype}y{d.y_dtype}")


# Experimental (off by default, LCGAN_UP2_FUSED=1 or set_up2_fused): memory-bound x2 transposed convs
# (Cin <= 128) as ONE tensor-core launch over the input lattice instead of four phase launches.
_UP2_FUSED = os.environ.get("LCGAN_UP2_FUSED", "0") == "1"
# measured (scratch/up2fused.py, batch 32): 64->32 @512->1024 1.84 -> 1.37 ms; 128->64 @256->512 0.88 -> 0.87 ms
# (the zero weight blocks cost 16/9 of the MMA work, which only the HBM-bound layers can absorb)
_UP2_FUSED_MAX_CIN = 64


def set_up2_fused(flag: bool):
    global _UP2_FUSED
    _UP2_FUSED = bool(flag)


def _is_cl_dense(t):
    n, c, h, w = t.shape
    return t.stride() == (c * h * w, 1, w * c, c)


def fused_up2_weights(w2: torch.Tensor, cin: int) -> torch.Tensor:
    """W2 [Cout][9*Cin] of a 3x3 x2 transposed conv -> W' [4*Cout][4*Cin]: row (py*2+px)*Cout + o, column
    (dy*2+dx)*Cin + c holds w[o, c, py+1-2dy, px+1-2dx] (zero when that kernel index is outside 0..2)."""
    cout = w2.shape[0]
    w3 = w2.view(cout, 9, cin)
    wf = torch.zeros((4 * cout, 4 * cin), dtype=w2.dtype, device=w2.device)
    for py in (0, 1):
        for px in (0, 1):
            for dy in (0, 1):
                for dx in (0, 1):
                    ki, kj = py + 1 - 2 * dy, px + 1 - 2 * dx
                    if 0 <= ki <= 2 and 0 <= kj <= 2:
                        ph, tap = py * 2 + px, dy * 2 + dx
                        wf[ph * cout:(ph + 1) * cout, tap * cin:(tap + 1) * cin] = w3[:, ki * 3 + kj, :]
    return wf


_FLOW_TC = os.environ.get("LCGAN_NO_FLOW_TC", "0") != "1"
_UP2_HALO = os.environ.get("LCGAN_NO_UP2_HALO", "0") != "1"


def up2_halo_eligible(cin, cout, h, w) -> bool:
    """x2 transposed convs the haloed single-launch tensor-core form takes (conv_tc.cu, rowshare 4)."""
    return bool(_UP2_HALO and cin in (32, 64) and cout % 16 == 0 and cout <= 32 and h % 16 == 0 and w % 8 == 0)


def _tapconv_up2_fused(lib, d, x, w2, y, plan, rowscale, bias, slope, gain, bias_scale, acc_scale, st, wf=None):
    """conv_transpose2d(k3, s2, p1, op1) as a 4-tap conv over the input lattice with 4*Cout output channels
    (phase-major): out(2m+py, 2n+px) = sum_{dy,dx in {0,1}} x[m+dy, n+dx] w[py+1-2dy, px+1-2dx]; the
    (phase, tap) blocks with a kernel index outside 0..2 are zero.  The epilogue scatters channel block
    (py, px) to output pixel (2m+py, 2n+px) (lcgan_tapconv_tc_blocked)."""
    cout, cin = w2.shape[0], x.shape[1]
    n, _, h, w = x.shape
    if wf is None:
        wf = fused_up2_weights(w2, cin)
    d.N, d.IH, d.IW, d.Cin = n, h, w, cin
    d.OH, d.OW, d.Cout = h, w, 4 * cout
    d.xs_n, d.xs_h, d.xs_w, d.xs_c = _strides_nhwc(x)
    ysn, ysh, ysw, ysc = _strides_nhwc(y)
    d.ys_n, d.ys_h, d.ys_w, d.ys_c = ysn, 2 * ysh, 2 * ysw, 1
    d.x_dtype, d.y_dtype, d.w_dtype = _dt(x), _dt(y), _dt(wf)
    d.MH, d.MW, d.os, d.py, d.px, d.is_ = h, w, 1, 0, 0, 1
    d.ntaps = 4
    for t, (dy, dx) in enumerate(((0, 0), (0, 1), (1, 0), (1, 1))):
        d.dy[t], d.dx[t], d.wtap[t] = dy, dx, t
    d.w_ld = 4 * cin
    d.acc_scale, d.bias_scale, d.slope, d.gain = acc_scale, bias_scale, slope, gain
    if cout != 2 and not lib.lcgan_tapconv_tc_eligible(C.byref(d)):
        raise RuntimeError("fused x2 transposed conv: descriptor not eligible for the tensor-core path")
    _lib.call("lcgan_tapconv_tc_blocked", C.byref(d), _ptr(x), _ptr(wf), _ptr(y), _ptr(rowscale), _ptr(bias),
              2 * cout, C.c_int64(ysh), cout, st, tag=_shape_tag("lcgan_tapconv_tc_up2fused", d),
              flops=2.0 * n * h * w * 9 * cin * cout,
              nbytes=x.numel() * x.element_size() + y.numel() * y.element_size())


def tapconv(x, w2, y, plan: plans.Plan, rowscale=None, bias=None, residual=None, slope=1.0, gain=1.0,
            bias_scale=1.0, acc_scale=1.0, noise=None, up2f=None, colscale=None):
    """y = epilogue(tapconv(x, w2)) for every launch of the plan; x, y logical NCHW.  noise: optional
    contiguous f32 [OH, OW] plane added before the activation (custom_layers.py:108-110)."""
    _need_cuda(x, w2, y)
    if noise is not None:
        assert noise.dtype == torch.float32 and noise.is_contiguous() and tuple(noise.shape) == (plan.OH, plan.OW)
    cout, cin = w2.shape[0], x.shape[1]
    assert w2.shape[1] == plan.k * plan.k * cin, (w2.shape, plan.k, cin)
    assert x.shape[2:] == (plan.IH, plan.IW) and y.shape[2:] == (plan.OH, plan.OW) and y.shape[1] == cout
    if residual is not None:
        assert residual.shape == y.shape and residual.dtype == y.dtype
        assert all(a == b for a, b, n in zip(residual.stride(), y.stride(), y.shape) if n > 1)
    lib = _lib.lib()
    st = _stream(x)
    d = TapConvDesc()
    if (_USE_TC and len(plan.launches) == 4 and plan.launches[0].os == 2 and residual is None
            and noise is None and cout % 16 == 0 and cin % 32 == 0 and x.dtype == torch.bfloat16
            and w2.dtype == torch.bfloat16 and _is_cl_dense(x) and _is_cl_dense(y)
            and ((_UP2_FUSED and cin <= _UP2_FUSED_MAX_CIN) or up2_halo_eligible(cin, cout, plan.IH, plan.IW))):
        # one launch over the input lattice; with 32 / 64 input channels the haloed form (one box per tile, resident
        # weights, no zero blocks)
        _tapconv_up2_fused(lib, d, x, w2, y, plan, rowscale, bias, slope, gain, bias_scale, acc_scale, st, wf=up2f)
        return y
    if (up2f is not None and _USE_TC and _FLOW_TC and cout == 2 and len(plan.launches) == 4 and plan.launches[0].os == 2
            and residual is None and noise is None and cin % 32 == 0 and x.dtype == torch.bfloat16
            and y.dtype == torch.float32 and _is_cl_dense(x) and _is_cl_dense(y)):
        # flow layer (C -> 2, x2): one tensor-core launch over the input lattice, 8 = 4 phases x 2 channels columns
        _tapconv_up2_fused(lib, d, x, w2, y, plan, rowscale, bias, slope, gain, bias_scale, acc_scale, st, wf=up2f)
        return y
    if len(plan.launches) == 4 and cout <= 4 and residual is None and noise is None and plan.launches[0].os == 2:
        # x2 transposed conv of a flow layer: one fused launch for the four output phases
        _fill_desc(d, plan.launches[0], x, y, cin, cout, w2, slope, gain, bias_scale, acc_scale)
        if lib.lcgan_tapconv_up2_thin_eligible(C.byref(d)):
            _lib.call("lcgan_tapconv_up2_thin", C.byref(d), _ptr(x), _ptr(w2), _ptr(y), _ptr(rowscale), _ptr(bias), st,
                      tag=_shape_tag("lcgan_tapconv_up2_thin", d),
                      flops=2.0 * x.shape[0] * plan.IH * plan.IW * 9 * cin * cout,
                      nbytes=x.numel() * x.element_size() + y.numel() * y.element_size())
            return y
    for l in plan.launches:
        _fill_desc(d, l, x, y, cin, cout, w2, slope, gain, bias_scale, acc_scale, noise, colscale)
        fn = "lcgan_tapconv_tc" if (_USE_TC and colscale is None and lib.lcgan_tapconv_tc_eligible(C.byref(d))) \
            else "lcgan_tapconv_simt"
        rows = x.shape[0] * l.MH * l.MW
        _lib.call(fn, C.byref(d), _ptr(x), _ptr(w2), _ptr(y), _ptr(rowscale), _ptr(bias), _ptr(residual), st,
                  tag=_shape_tag(fn, d),
                  flops=2.0 * rows * len(l.taps) * cin * cout,
                  nbytes=(x.numel() * x.element_size() + y.numel() * y.element_size()) / len(plan.launches)
                  + w2.numel() * w2.element_size() * len(l.taps) / (plan.k * plan.k))
    return y


def tapconv_wgrad(x, g, plan: plans.Plan, cin, cout, scale=1.0):
    """dW2[o][tap*cin + c] (f32) for x at the plan's input positions and g at its output positions."""
    _need_cuda(x, g)
    assert x.shape[1] == cin and g.shape[1] == cout
    assert x.shape[2:] == (plan.IH, plan.IW) and g.shape[2:] == (plan.OH, plan.OW)
    dw2 = torch.zeros((cout, plan.k * plan.k * cin), dtype=torch.float32, device=x.device)
    lib = _lib.lib()
    st = _stream(x)
    d = TapConvDesc()
    cv4 = cin // 4
    if (_FLOW_TC and _USE_TC and _USE_WGRAD_TC and not _DET and len(plan.launches) == 4 and cout == 2
            and plan.launches[0].os == 2 and cin % 32 == 0 and x.dtype == torch.bfloat16 and g.dtype == torch.float32
            and _is_cl_dense(x) and _is_cl_dense(g) and plan.IH * plan.IW >= 128
            and plan.IH & (plan.IH - 1) == 0 and plan.IW & (plan.IW - 1) == 0):
        # flow layer: gather the 9 x 2 gradient taps into a 32-channel bf16 tensor, then a pointwise tensor-core wgrad
        n = x.shape[0]
        g18 = empty_cl(n, 32, plan.IH, plan.IW, torch.bfloat16, x.device)
        _lib.call("lcgan_flow_grad_im2col", _ptr(g), _ptr(g18), n, plan.IH, plan.IW, st, tag="flow_grad_im2col",
                  nbytes=g.numel() * 4 + g18.numel() * 2)
        p1 = plans.conv(1, 1, plan.IH, plan.IW)
        l = p1.launches[0]
        t = torch.zeros((32, cin), dtype=torch.float32, device=x.device)
        _fill_desc(d, l, x, g18, cin, 32, None, 1.0, 1.0, 1.0)
        d.w_ld = cin
        _lib.call("lcgan_tapconv_wgrad_tc", C.byref(d), _ptr(x), _ptr(g18), _ptr(t), C.c_float(scale), st,
                  tag=_shape_tag("lcgan_tapconv_wgrad_tc", d), flops=2.0 * n * plan.IH * plan.IW * 18 * cin,
                  nbytes=x.numel() * 2 + g18.numel() * 2)
        # rows (tap, o) -> dW2[o][tap*cin + c]
        return t[:18].view(9, 2, cin).permute(1, 0, 2).reshape(2, 9 * cin).contiguous()
    if (len(plan.launches) == 4 and cout <= 2 and plan.launches[0].os == 2 and cin % 4 == 0 and cv4 <= 256
            and 256 % cv4 == 0):
        # x2 transposed conv of a flow layer: one fused pass for all 9 taps
        _fill_desc(d, plan.launches[0], x, g, cin, cout, None, 1.0, 1.0, 1.0)
        d.w_ld = plan.k * plan.k * cin
        if lib.lcgan_tapconv_up2_thin_eligible(C.byref(d)):
            _lib.call("lcgan_tapconv_up2_thin_wgrad", C.byref(d), _ptr(x), _ptr(g), _ptr(dw2), C.c_float(scale), st,
                      tag=_shape_tag("lcgan_tapconv_up2_thin_wgrad", d),
                      flops=2.0 * x.shape[0] * plan.IH * plan.IW * 9 * cin * cout,
                      nbytes=x.numel() * x.element_size() + g.numel() * g.element_size())
            return dw2
    for l in plan.launches:
        _fill_desc(d, l, x, g, cin, cout, None, 1.0, 1.0, 1.0)
        d.w_ld = plan.k * plan.k * cin
        fn = "lcgan_tapconv_wgrad_tc" if (_USE_TC and lib.lcgan_tapconv_tc_eligible(C.byref(d))
                                          and _wgrad_tc_ok(d)) else "lcgan_tapconv_wgrad_simt"
        rows = x.shape[0] * l.MH * l.MW
        _lib.call(fn, C.byref(d), _ptr(x), _ptr(g), _ptr(dw2), C.c_float(scale), st,
                  tag=_shape_tag(fn, d),
                  flops=2.0 * rows * len(l.taps) * cin * cout,
                  nbytes=(x.numel() * x.element_size() + g.numel() * g.element_size()) / len(plan.launches))
    return dw2


_USE_WGRAD_TC = _USE_TC


def set_wgrad_tensor_cores(flag: bool):
    global _USE_WGRAD_TC
    _USE_WGRAD_TC = bool(flag)


def _wgrad_tc_ok(d):
    # the wgrad kernel additionally needs G dense channels-last bf16 with Cout % 64 == 0
    return (_USE_WGRAD_TC and d.y_dtype == BF16 and d.Cout % 32 == 0 and d.ys_c == 1
            and (d.OW == 1 or d.ys_w == d.Cout) and (d.OH == 1 or d.ys_h == d.OW * d.Cout)
            and (d.N == 1 or d.ys_n == d.OH * d.OW * d.Cout))


def _alloc_out(n, c, h, w, dtype, device, nchw):
    if nchw:
        return torch.empty((n, c, h, w), dtype=dtype, device=device)
    return empty_cl(n, c, h, w, dtype, device)


# ------------------------------------------------------------------------------------------
# convolution family (closed under differentiation)
# ------------------------------------------------------------------------------------------
class ConvFwd(torch.autograd.Function):
    """Pure linear tap conv.  transposed=False contracts w over dim 1 (input channels);
    transposed=True contracts over dim 0 (this is the data gradient of the former)."""

    @staticmethod
    def forward(ctx, x, w, wscale, plan, transposed, out_dtype, out_nchw):
        ctx.plan, ctx.transposed, ctx.wscale = plan, transposed, wscale
        ctx.x_dtype, ctx.x_nchw = x.dtype, (x.is_contiguous() and not _is_cl(x))
        ctx.save_for_backward(x, w)
        w2 = pack_weight(w, transposed, torch.float32 if x.dtype == torch.float32 else torch.bfloat16)
        y = _alloc_out(x.shape[0], w2.shape[0], plan.OH, plan.OW, out_dtype, x.device, out_nchw)
        return tapconv(x, w2, y, plan, acc_scale=wscale)

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dx = dw = None
        if ctx.needs_input_grad[0]:
            dx = ConvFwd.apply(dy, w, ctx.wscale, plans.adjoint(ctx.plan), not ctx.transposed,
                               ctx.x_dtype, ctx.x_nchw)
        if ctx.needs_input_grad[1] and _wgrad_enabled():
            dw = ConvWgrad.apply(x, dy, ctx.plan, ctx.transposed, tuple(w.shape), ctx.wscale)
        return dx, dw, None, None, None, None, None


class ConvWgrad(torch.autograd.Function):
    """dw[o,c,t] = scale * sum x[p+t, c] g[p, o]   (transposed: sum x[p+t, o] g[p, c]); scale carries the
    equalized-lr constant, applied by the kernel's epilogue."""

    @staticmethod
    def forward(ctx, x, g, plan, transposed, wshape, scale=1.0):
        ctx.plan, ctx.transposed, ctx.wshape, ctx.scale = plan, transposed, wshape, scale
        ctx.x_fmt = (x.dtype, x.is_contiguous() and not _is_cl(x))
        ctx.g_fmt = (g.dtype, g.is_contiguous() and not _is_cl(g))
        ctx.save_for_backward(x, g)
        o, i = wshape[0], wshape[1]
        cin, cout = (o, i) if transposed else (i, o)
        dw2 = tapconv_wgrad(x, g, plan, cin, cout, scale=scale)
        return unpack_wgrad(dw2, wshape, transposed).contiguous()

    @staticmethod
    def backward(ctx, ggw):
        x, g = ctx.saved_tensors
        dx = dg = None
        if ctx.needs_input_grad[0]:
            dx = ConvFwd.apply(g, ggw, ctx.scale, plans.adjoint(ctx.plan), not ctx.transposed, *ctx.x_fmt)
        if ctx.needs_input_grad[1]:
            dg = ConvFwd.apply(x, ggw, ctx.scale, ctx.plan, ctx.transposed, *ctx.g_fmt)
        return dx, dg, None, None, None, None


class ActBwd(torch.autograd.Function):
    """gout = dy * gain * (y>0 ? 1 : slope) * d[b,c];  r0 = sum_p dz, r1 = sum_p dz*z (f32 [N,C]).
    Linear (and self-adjoint) in dy, so its own backward is the same kernel."""

    @staticmethod
    def forward(ctx, dy, y, d, slope, gain, want_r0, want_r1):
        _need_cuda(dy, y)
        n, c, h, w = y.shape
        dy = _cl(dy, y.dtype)
        assert _is_cl(y)
        gout, r0, r1 = _act_bwd_raw(dy, y, d, slope, gain, want_r0, want_r1)
        ctx.save_for_backward(y, d)
        ctx.slope, ctx.gain = slope, gain
        outs = (gout, r0 if want_r0 else _placeholder(y.device), r1 if want_r1 else _placeholder(y.device))
        ctx.mark_non_differentiable(outs[1], outs[2])
        return outs

    @staticmethod
    def backward(ctx, gg, _g0, _g1):
        y, d = ctx.saved_tensors
        out = ActBwd.apply(gg, y, d, ctx.slope, ctx.gain, False, False)[0]
        return out, None, None, None, None, None, None


class ConvAct(torch.autograd.Function):
    """y = lrelu(tapconv(x, w*wscale) * rowscale[b,o] + bias*bias_scale, slope) * gain (+ residual).

    One fused kernel forward (custom_layers.py:41-43 / :83-85 + the F.leaky_relu that follows).
    Backward composes ActBwd, ConvFwd(adjoint) and ConvWgrad, so it is differentiable again.
    """

    @staticmethod
    def forward(ctx, x, w, bias, rowscale, residual, wscale, plan, slope, gain, bias_scale, out_dtype,
                out_nchw):
        assert residual is None or slope == 1.0, "residual fusion only without activation"
        compute = torch.float32 if x.dtype == torch.float32 else torch.bfloat16
        w2 = pack_weight(w, False, compute)
        y = _alloc_out(x.shape[0], w2.shape[0], plan.OH, plan.OW, out_dtype, x.device, out_nchw)
        assert rowscale is None or (rowscale.is_contiguous() and rowscale.dtype == torch.float32)
        assert bias is None or (bias.is_contiguous() and bias.dtype == torch.float32)
        tapconv(x, w2, y, plan, rowscale, bias, residual, slope, gain, bias_scale, wscale)
        ctx.save_for_backward(x, w, bias, rowscale, y)
        ctx.cfg = (wscale, plan, slope, gain, bias_scale, residual is not None)
        ctx.x_fmt = (x.dtype, x.is_contiguous() and not _is_cl(x))
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, bias, rowscale, y = ctx.saved_tensors
        wscale, plan, slope, gain, bias_scale, has_res = ctx.cfg
        need_x, need_w, need_b, need_rs, need_res = ctx.needs_input_grad[:5]
        wg = _wgrad_enabled()
        need_w, need_b = need_w and wg, need_b and wg
        dres = dy if (has_res and need_res) else None
        # no activation, no row scale, no bias gradient (the 1x1 skip convs): the epilogue was y = acc * gain, so the
        # gain moves into the scale of the data- and weight-gradient kernels and no pass over dy is needed at all
        trivial = slope == 1.0 and rowscale is None
        if trivial and not need_b:
            g, r0, r1 = (dy if (dy.dtype == y.dtype and _is_cl(dy)) else _cl(dy, y.dtype)), None, None
            wscale = wscale * gain
        else:
            ycl = y if _is_cl(y) else _cl(y)
            if has_res:
                # slope == 1 here, so the mask is all-ones and y is only used when rowscale needs z;
                # residual + rowscale are never combined by the layers.
                assert rowscale is None
            g, r0, r1 = ActBwd.apply(dy, ycl, rowscale, slope, gain, bool(need_b or need_rs), bool(need_rs))
        dx = dw = db = drs = None
        if need_x:
            dx = ConvFwd.apply(g, w, wscale, plans.adjoint(plan), True, *ctx.x_fmt)
        if need_w:
            dw = ConvWgrad.apply(x, g, plan, False, tuple(w.shape), wscale)
        if need_b or need_rs:
            db, drs = _epilogue_grads(r0, r1, bias, rowscale, bias_scale, need_b, need_rs)
        return dx, dw, db, drs, dres, None, None, None, None, None, None, None


def _epilogue_grads(r0, r1, bias, d, bias_scale, want_db, want_dd):
    """db[o] = bias_scale sum_b r0;  dd[b,o] = (r1 - bias*bias_scale*r0) / d  - one launch."""
    n, c = r0.shape
    db = torch.empty((c,), dtype=torch.float32, device=r0.device) if want_db else None
    dd = torch.empty((n, c), dtype=torch.float32, device=r0.device) if want_dd else None
    _lib.call("lcgan_epilogue_grads", _ptr(r0), _ptr(r1 if want_dd else None), _ptr(bias), _ptr(d if want_dd else None),
              C.c_float(bias_scale), _ptr(db), _ptr(dd), n, c, _stream(r0))
    return db, dd


class BoxMask(torch.autograd.Function):
    """g = box3(dt) * lrelu'(y) * gain, r0 = sum_p g: from the gradient of a box filter's OUTPUT to the gradient of
    the pre-activation of the conv in front of it (y = that conv's stored output), one pass instead of Box3 + ActBwd.
    Linear in dt; its adjoint is MaskBox, so R1's create_graph pass builds a graph of our kernels again."""

    @staticmethod
    def forward(ctx, dt, y, slope, gain, want_r0):
        _need_cuda(dt, y)
        assert _is_cl(y)
        dt = _cl(dt, y.dtype)
        n, c, h, w = y.shape
        r0 = torch.zeros((n, c), dtype=torch.float32, device=y.device) if want_r0 else None
        if box3_mod_eligible(y):
            g = torch.empty_like(y)
            _lib.call("lcgan_box3_postmask", _ptr(dt), _ptr(y), _ptr(g), _ptr(r0), _dt(y), n, h, w, c,
                      C.c_float(slope), C.c_float(gain), _stream(y), nbytes=3 * y.numel() * y.element_size(),
                      tag="box3_act_bwd")
        else:                                   # small maps: the two separate kernels
            tmp = torch.empty_like(y)
            _lib.call("lcgan_box3", _ptr(dt), None, _ptr(tmp), _dt(y), n, h, w, c, C.c_float(1.0), C.c_float(1.0),
                      C.c_float(1.0), C.c_float(1.0), _stream(y), nbytes=2 * y.numel() * y.element_size())
            g = torch.empty_like(y)
            _lib.call("lcgan_act_bwd", _ptr(tmp), _ptr(y), _ptr(g), None, _ptr(r0), None, _dt(y), n, h * w, c,
                      C.c_float(slope), C.c_float(gain), _stream(y), nbytes=3 * y.numel() * y.element_size())
        ctx.save_for_backward(y)
        ctx.cfg = (slope, gain)
        outs = (g, r0 if want_r0 else _placeholder(y.device))
        ctx.mark_non_differentiable(outs[1])
        return outs

    @staticmethod
    def backward(ctx, gg, _g0):
        (y,) = ctx.saved_tensors
        return MaskBox.apply(gg, y, *ctx.cfg), None, None, None, None


class MaskBox(torch.autograd.Function):
    """h = box3(gg * lrelu'(y) * gain): the adjoint of BoxMask (the box filter is self-adjoint)."""

    @staticmethod
    def forward(ctx, gg, y, slope, gain):
        _need_cuda(gg, y)
        gg = _cl(gg, y.dtype)
        n, c, h, w = y.shape
        out = torch.empty_like(y)
        _lib.call("lcgan_box3", _ptr(gg), _ptr(y), _ptr(out), _dt(y), n, h, w, c, C.c_float(slope), C.c_float(gain),
                  C.c_float(1.0), C.c_float(1.0), _stream(y), nbytes=3 * y.numel() * y.element_size(), tag="box3_act_bwd")
        ctx.save_for_backward(y)
        ctx.cfg = (slope, gain)
        return out

    @staticmethod
    def backward(ctx, hgrad):
        (y,) = ctx.saved_tensors
        return BoxMask.apply(hgrad, y, ctx.cfg[0], ctx.cfg[1], False)[0], None, None, None


class ConvActBox(torch.autograd.Function):
    """t = box3(lrelu(tapconv(x, w*wscale) + bias*bias_scale, slope) * gain): the first conv of a discriminator block
    with the box filter that follows it (custom_layers.py:204-206) as ONE autograd node, so that backward can run the
    box filter's and the activation's backward as one pass (BoxMask).  Twice differentiable like ConvAct."""

    @staticmethod
    def forward(ctx, x, w, bias, wscale, plan, slope, gain, bias_scale):
        compute = torch.float32 if x.dtype == torch.float32 else torch.bfloat16
        w2 = pack_weight(w, False, compute)
        y = _alloc_out(x.shape[0], w2.shape[0], plan.OH, plan.OW, _ACT_DTYPE, x.device, False)
        tapconv(x, w2, y, plan, None, bias, None, slope, gain, bias_scale, wscale)
        n, c, h, wd = y.shape
        t = torch.empty_like(y)
        _lib.call("lcgan_box3", _ptr(y), None, _ptr(t), _dt(y), n, h, wd, c, C.c_float(1.0), C.c_float(1.0),
                  C.c_float(1.0), C.c_float(1.0), _stream(y), nbytes=2 * y.numel() * y.element_size())
        ctx.save_for_backward(x, w, bias, y)
        ctx.cfg = (wscale, plan, slope, gain, bias_scale)
        ctx.x_fmt = (x.dtype, x.is_contiguous() and not _is_cl(x))
        return t

    @staticmethod
    def backward(ctx, dt):
        x, w, bias, y = ctx.saved_tensors
        wscale, plan, slope, gain, bias_scale = ctx.cfg
        need_x, need_w, need_b = ctx.needs_input_grad[:3]
        wg = _wgrad_enabled()
        need_w, need_b = need_w and wg, need_b and wg
        g, r0 = BoxMask.apply(dt, y, slope, gain, bool(need_b))
        dx = dw = db = None
        if need_x:
            dx = ConvFwd.apply(g, w, wscale, plans.adjoint(plan), True, *ctx.x_fmt)
        if need_w:
            dw = ConvWgrad.apply(x, g, plan, False, tuple(w.shape), wscale)
        if need_b:
            db, _ = _epilogue_grads(r0, None, bias, None, bias_scale, True, False)
        return dx, dw, db, None, None, None, None, None


def conv_act(x, w, bias=None, rowscale=None, residual=None, *, wscale=1.0, plan, slope=1.0, gain=1.0,
             bias_scale=1.0, out_dtype=None, out_nchw=False):
    return ConvAct.apply(x, w, bias, rowscale, residual, wscale, plan, slope, gain, bias_scale,
                         out_dtype or _ACT_DTYPE, out_nchw)


def linear_act(x2d, w, bias, *, wscale, bias_scale, slope=1.0, gain=1.0, out_dtype=torch.float32):
    """EqualizedLinear (+ optional leaky-relu) as a 1x1 tap conv; x2d [b, in] -> [b, out]."""
    y = ConvAct.apply(x2d[:, :, None, None], w, bias, None, None, wscale, plans.linear(), slope, gain,
                      bias_scale, out_dtype, True)
    return y[:, :, 0, 0]


# ------------------------------------------------------------------------------------------
# resampling (discriminator side: differentiable twice)
# ------------------------------------------------------------------------------------------
class Box3(torch.autograd.Function):
    """F.avg_pool2d(x, 3, 1, 1) (count_include_pad) - linear and self-adjoint."""

    @staticmethod
    def forward(ctx, x):
        _need_cuda(x)
        x = _cl(x)
        n, c, h, w = x.shape
        out = torch.empty_like(x)
        _lib.call("lcgan_box3", _ptr(x), None, _ptr(out), _dt(x), n, h, w, c,
                  C.c_float(1.0), C.c_float(1.0), C.c_float(1.0), C.c_float(1.0), _stream(x),
                  nbytes=2 * x.numel() * x.element_size())
        return out

    @staticmethod
    def backward(ctx, dy):
        return Box3.apply(dy)


class Pool2(torch.autograd.Function):
    """y = scale * sum over 2x2 (F.avg_pool2d(2,2) with scale .25); adjoint is Up2."""

    @staticmethod
    def forward(ctx, x, scale):
        _need_cuda(x)
        x = _cl(x)
        n, c, h, w = x.shape
        ctx.scale = scale
        out = empty_cl(n, c, h // 2, w // 2, x.dtype, x.device)
        _lib.call("lcgan_pool2", _ptr(x), _ptr(out), _dt(x), n, h, w, c, C.c_float(scale), _stream(x),
                  nbytes=1.25 * x.numel() * x.element_size())
        return out

    @staticmethod
    def backward(ctx, dy):
        return Up2.apply(dy, ctx.scale), None


class AddUp2(torch.autograd.Function):
    """out = a + scale * nearest_up2(s): the two gradients of a pooled-and-used tensor summed in one pass."""

    @staticmethod
    def forward(ctx, a, s, scale):
        _need_cuda(a, s)
        a, s = _cl(a), _cl(s)
        n, c, h, w = s.shape
        ctx.scale = scale
        out = torch.empty_like(a)
        _lib.call("lcgan_up2_add", _ptr(a), _ptr(s), _ptr(out), _dt(a), n, h, w, c, C.c_float(scale), _stream(a),
                  nbytes=2.25 * a.numel() * a.element_size())
        return out

    @staticmethod
    def backward(ctx, g):
        return g, (Pool2.apply(g, ctx.scale) if ctx.needs_input_grad[1] else None), None


class PoolFork(torch.autograd.Function):
    """(x, scale * sumpool2x2(x)): the input of a block that uses x AND its 2x2 average (DiscriminatorBlock,
    custom_layers.py:206-216).  As ONE node with two outputs its backward receives the two gradients separately and
    adds them in a single pass (AddUp2) - autograd's own sum of the two branches costs up2 (write N) + add (read 2N,
    write N); every piece is again a Function, so the R1 double backward goes through our kernels."""

    @staticmethod
    def forward(ctx, x, scale):
        _need_cuda(x)
        xc = _cl(x)
        n, c, h, w = xc.shape
        ctx.scale = scale
        pooled = empty_cl(n, c, h // 2, w // 2, xc.dtype, xc.device)
        _lib.call("lcgan_pool2", _ptr(xc), _ptr(pooled), _dt(xc), n, h, w, c, C.c_float(scale), _stream(xc),
                  nbytes=1.25 * xc.numel() * xc.element_size())
        return xc.view_as(xc), pooled

    @staticmethod
    def backward(ctx, g_main, g_pool):
        if g_pool is None:
            return g_main, None
        if g_main is None:
            return Up2.apply(g_pool, ctx.scale), None
        if g_main.dtype != g_pool.dtype:
            return g_main + Up2.apply(g_pool, ctx.scale).to(g_main.dtype), None
        return AddUp2.apply(g_main, g_pool, ctx.scale), None


class Up2(torch.autograd.Function):
    """y[2i+a, 2j+b] = scale * x[i, j]; adjoint is Pool2."""

    @staticmethod
    def forward(ctx, x, scale):
        _need_cuda(x)
        x = _cl(x)
        n, c, h, w = x.shape
        ctx.scale = scale
        out = empty_cl(n, c, h * 2, w * 2, x.dtype, x.device)
        _lib.call("lcgan_up2", _ptr(x), _ptr(out), _dt(x), n, h, w, c, C.c_float(scale), _stream(x),
                  nbytes=5 * x.numel() * x.element_size())
        return out

    @staticmethod
    def backward(ctx, dy):
        return Pool2.apply(dy, ctx.scale), None


# ------------------------------------------------------------------------------------------
# generator-side fused ops (first order only)
# ------------------------------------------------------------------------------------------
class Box3Act(torch.autograd.Function):
    """y = lrelu(box3(x), slope) * gain   (custom_layers.py:154-155)."""

    @staticmethod
    def forward(ctx, x, slope, gain):
        _need_cuda(x)
        x = _cl(x)
        n, c, h, w = x.shape
        y = torch.empty_like(x)
        _lib.call("lcgan_box3", _ptr(x), None, _ptr(y), _dt(x), n, h, w, c,
                  C.c_float(1.0), C.c_float(1.0), C.c_float(slope), C.c_float(gain), _stream(x),
                  nbytes=2 * x.numel() * x.element_size(), tag="box3_act")
        ctx.save_for_backward(y)
        ctx.cfg = (slope, gain)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        (y,) = ctx.saved_tensors
        slope, gain = ctx.cfg
        dy = _cl(dy, y.dtype)
        n, c, h, w = y.shape
        dx = torch.empty_like(y)
        _lib.call("lcgan_box3", _ptr(dy), _ptr(y), _ptr(dx), _dt(y), n, h, w, c,
                  C.c_float(slope), C.c_float(gain), C.c_float(1.0), C.c_float(1.0), _stream(y),
                  nbytes=3 * y.numel() * y.element_size(), tag="box3_act_bwd")
        return dx, None, None


def box3_mod_eligible(x) -> bool:
    """Shapes the tiled box kernel with a folded style scale takes (lcgan_box3_cs)."""
    n, c, h, w = x.shape
    return (x.is_cuda and w >= 32 and h >= 16 and c % (32 if x.dtype == torch.bfloat16 else 16) == 0 and _FOLD_STYLE
            and not _DET)


_FOLD_STYLE = os.environ.get("LCGAN_NO_FOLD_STYLE", "0") != "1"
# deterministic mode (set_deterministic): the fused passes whose reductions finish with atomics (style gradient in the
# box pass, bias gradient in BoxMask, the per-image pointwise weight gradient) fall back to the separate kernels, whose
# reductions take ordered turns
_DET = False


def fold_style_eligible(h, w, c, is_cuda=True) -> bool:
    """Shapes for which the tiled box / warp kernels can fold a style scale into their pass."""
    return bool(is_cuda and _FOLD_STYLE and not _DET and w >= 32 and h >= 16
                and c % (32 if _ACT_DTYPE == torch.bfloat16 else 16) == 0)


class Box3ActMod(torch.autograd.Function):
    """y = lrelu(box3(x), slope) * gain * s[b,c]: Box3Act with the style modulation of the FOLLOWING modulated
    conv (custom_layers.py:62-64, shared-weight form) folded into the same pass - the conv then reads y directly
    (ModConvAct with premodulated=True), so neither x*s nor its recomputation in backward ever makes a pass over
    HBM, and the style gradient ds = sum_p dy * a (a = y / s) is reduced inside the backward box pass."""

    @staticmethod
    def forward(ctx, x, s, slope, gain):
        _need_cuda(x, s)
        x = _cl(x)
        n, c, h, w = x.shape
        assert s.dtype == torch.float32 and s.is_contiguous() and tuple(s.shape) == (n, c)
        y = torch.empty_like(x)
        _lib.call("lcgan_box3_cs", _ptr(x), None, _ptr(y), _ptr(s), None, _dt(x), n, h, w, c,
                  C.c_float(1.0), C.c_float(1.0), C.c_float(slope), C.c_float(gain), _stream(x),
                  nbytes=2 * x.numel() * x.element_size(), tag="box3_act")
        ctx.save_for_backward(y, s)
        ctx.cfg = (slope, gain)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        y, s = ctx.saved_tensors
        slope, gain = ctx.cfg
        dy = _cl(dy, y.dtype)
        n, c, h, w = y.shape
        dx = torch.empty_like(y)
        red = torch.zeros_like(s) if ctx.needs_input_grad[1] else None
        _lib.call("lcgan_box3_cs", _ptr(dy), _ptr(y), _ptr(dx), _ptr(s), _ptr(red), _dt(y), n, h, w, c,
                  C.c_float(slope), C.c_float(gain), C.c_float(1.0), C.c_float(1.0), _stream(y),
                  nbytes=3 * y.numel() * y.element_size(), tag="box3_act_bwd")
        # ds = sum_p dy * a with a = y / s (a style that is exactly zero gives no gradient here; measure zero)
        ds = None if red is None else torch.where(s != 0, red / s, torch.zeros_like(red))
        return dx, ds, None, None


class Up2BoxAdd(torch.autograd.Function):
    """out = box3(nearest_up2(s)) + t   (custom_layers.py:146-147,159)."""

    @staticmethod
    def forward(ctx, s, t):
        _need_cuda(s, t)
        s, t = _cl(s), _cl(t)
        n, c, h, w = s.shape
        out = torch.empty_like(t)
        _lib.call("lcgan_up2box_add", _ptr(s), _ptr(t), _ptr(out), _dt(s), n, h, w, c, _stream(s),
                  nbytes=(s.numel() + 2 * t.numel()) * s.element_size())
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        dy = _cl(dy)
        n, c, h, w = dy.shape
        ds = None
        if ctx.needs_input_grad[0]:
            ds = empty_cl(n, c, h // 2, w // 2, dy.dtype, dy.device)
            _lib.call("lcgan_box3_pool2", _ptr(dy), _ptr(ds), _dt(dy), n, h, w, c, _stream(dy),
                      nbytes=1.25 * dy.numel() * dy.element_size())
        return ds, (dy if ctx.needs_input_grad[1] else None)


class Modulate(torch.autograd.Function):
    """xs = x * s[b,c]   (style modulation in shared-weight form, custom_layers.py:62-64)."""

    @staticmethod
    def forward(ctx, x, s):
        _need_cuda(x, s)
        x = _cl(x)
        s = s.contiguous().float()
        n, c, h, w = x.shape
        xs = torch.empty_like(x)
        _lib.call("lcgan_modulate", _ptr(x), _ptr(s), _ptr(xs), _dt(x), n, h * w, c, _stream(x),
                  nbytes=2 * x.numel() * x.element_size())
        ctx.save_for_backward(x, s)
        return xs

    @staticmethod
    @once_differentiable
    def backward(ctx, t):
        x, s = ctx.saved_tensors
        t = _cl(t, x.dtype)
        n, c, h, w = x.shape
        dx = torch.empty_like(x)
        ds = torch.zeros_like(s)
        _lib.call("lcgan_modulate_bwd", _ptr(x), _ptr(t), _ptr(s), _ptr(dx), _ptr(ds), _dt(x), n, h * w, c,
                  _stream(x), nbytes=3 * x.numel() * x.element_size())
        return dx, ds


_placeholders = {}


def _placeholder(device):
    """A shared 0-dim zero standing in for an output that was not requested (no fill kernel per call)."""
    t = _placeholders.get(device)
    if t is None:
        t = _placeholders[device] = torch.zeros((), device=device)
    return t


def _act_bwd_raw(dy, y, d, slope, gain, want_r0, want_r1):
    n, c, h, w = y.shape
    gout = torch.empty_like(y)
    r0 = r1 = None
    if want_r0 or want_r1:                      # one zero-fill for both reduction buffers
        r = torch.zeros((2, n, c), dtype=torch.float32, device=y.device)
        r0, r1 = (r[0] if want_r0 else None), (r[1] if want_r1 else None)
    _lib.call("lcgan_act_bwd", _ptr(dy), _ptr(y), _ptr(gout), _ptr(d), _ptr(r0), _ptr(r1), _dt(y),
              n, h * w, c, C.c_float(slope), C.c_float(gain), _stream(y), nbytes=3 * y.numel() * y.element_size())
    return gout, r0, r1


def _modulate_raw(x, s):
    n, c, h, w = x.shape
    xs = torch.empty_like(x)
    _lib.call("lcgan_modulate", _ptr(x), _ptr(s), _ptr(xs), _dt(x), n, h * w, c, _stream(x),
              nbytes=2 * x.numel() * x.element_size())
    return xs


class ModConvAct(torch.autograd.Function):
    """Modulated convolution of the generator in one autograd node (custom_layers.py:60-86 + the
    leaky-relu that follows):  y = lrelu(d[b,o] * conv(x * s[b,c], w*c) + bias, slope) * gain.
    The modulated activations x*s are a temporary: they are NOT saved for backward but recomputed
    there (one elementwise pass) - at 1024x1024 they would be 40 % of the generator's saved bytes.
    First order only (the generator is never on the R1 path)."""

    @staticmethod
    def forward(ctx, x, s, w, bias, d, noise, wscale, plan, slope, gain, bias_scale, out_dtype, out_nchw,
                premodulated=False):
        """premodulated: x already IS x*s (its producer folded the style in - Box3ActMod); s is then ignored here
        and gets its gradient from that producer."""
        _need_cuda(x, s, w)
        assert _is_cl(x) and s.is_contiguous() and d.is_contiguous() and s.dtype == d.dtype == torch.float32
        compute = torch.float32 if x.dtype == torch.float32 else torch.bfloat16
        ctx.premod = bool(premodulated)
        # pointwise 32 -> (<= 4) layer (to-RGB 1x1 at 1024^2): the thin kernel multiplies by the style as it reads x,
        # and one per-image weight-gradient pass yields dW and ds - no modulate / modulate_bwd passes at all
        ctx.pwmod = bool(not premodulated and _FOLD_STYLE and not _DET and plan.k == 1 and len(plan.launches) == 1
                         and plan.launches[0].is_ == 1 and x.shape[1] == 32 and w.shape[0] <= 4
                         and x.dtype == torch.bfloat16 and _is_cl_dense(x) and noise is None)
        xs = x if (premodulated or ctx.pwmod) else _modulate_raw(x, s)
        w2 = pack_weight(w, False, compute)
        y = _alloc_out(x.shape[0], w2.shape[0], plan.OH, plan.OW, out_dtype, x.device, out_nchw)
        up2f = None
        if len(plan.launches) == 4 and compute == torch.bfloat16 and _USE_TC and not out_nchw and (
                (w.shape[0] == 2 and _FLOW_TC and out_dtype == torch.float32)
                or up2_halo_eligible(x.shape[1], w.shape[0], plan.IH, plan.IW)):
            up2f = _derive(w, ("up2f", compute))
        tapconv(xs, w2, y, plan, d, bias, None, slope, gain, bias_scale, wscale, noise, up2f,
                colscale=s if ctx.pwmod else None)
        del xs
        ctx.save_for_backward(x, s, w, bias, d, y)
        ctx.cfg = (wscale, plan, slope, gain, bias_scale)
        ctx.has_noise, ctx.noise = noise is not None, noise
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        x, s, w, bias, d, y = ctx.saved_tensors
        wscale, plan, slope, gain, bias_scale = ctx.cfg
        need_x, need_s, need_w, need_b, need_d, need_nz = ctx.needs_input_grad[:6]
        compute = torch.float32 if x.dtype == torch.float32 else torch.bfloat16
        ycl = y if _is_cl(y) else _cl(y)
        g, r0, r1 = _act_bwd_raw(_cl(dy, ycl.dtype), ycl, d, slope, gain, need_b or need_d, need_d)
        # (g stays fp32 for the fp32-output layers - flow field, RGB - the thin kernels mix dtypes)
        dx = ds = dw = db = dd = dnz = None
        if ctx.has_noise and need_nz:
            # d noise[h,w] = sum_{b,o} dz, dz = g / d[b,o]: a torch reduction, only on noise-enabled layers
            # (cnn.py never enables them; custom_layers.py:98-101 keeps the option)
            dnz = (g.float() / d[:, :, None, None]).sum(dim=(0, 1))
        if ctx.pwmod:
            if need_x:
                dx = torch.empty_like(x)
                tapconv(g, pack_weight(w, True, compute), dx, plans.adjoint(plan), rowscale=s, acc_scale=wscale)
            if (need_w and _wgrad_enabled()) or need_s:
                n, cin, cout = x.shape[0], x.shape[1], w.shape[0]
                dwp = torch.zeros((n, cout, cin), dtype=torch.float32, device=x.device)
                dsc = TapConvDesc()
                _fill_desc(dsc, plan.launches[0], x, g, cin, cout, None, 1.0, 1.0, 1.0)
                _lib.call("lcgan_pw_wgrad32", C.byref(dsc), _ptr(x), _ptr(g), _ptr(dwp), _stream(x), tag="pw_wgrad32",
                          nbytes=x.numel() * x.element_size() + g.numel() * g.element_size())
                if need_w and _wgrad_enabled():
                    dw = ((dwp * s[:, None, :]).sum(0) * wscale).reshape(w.shape)
                if need_s:
                    wq = pack_weight(w, False, compute).float()
                    ds = (dwp * wq[None]).sum(1) * wscale
        elif ctx.premod:
            # x is the modulated activation itself: the data gradient w.r.t. it is the plain adjoint conv, the
            # weight gradient reads it as is, and the style gradient belongs to the producer of x
            if need_x:
                dx = empty_cl(x.shape[0], x.shape[1], x.shape[2], x.shape[3], x.dtype, x.device)
                tapconv(g, pack_weight(w, True, compute), dx, plans.adjoint(plan), acc_scale=wscale)
            if need_w and _wgrad_enabled():
                dw2 = tapconv_wgrad(x, g, plan, x.shape[1], w.shape[0], scale=wscale)
                dw = unpack_wgrad(dw2, tuple(w.shape), False).contiguous()
        else:
            if need_x or need_s:
                t = empty_cl(x.shape[0], x.shape[1], x.shape[2], x.shape[3], x.dtype, x.device)
                tapconv(g, pack_weight(w, True, compute), t, plans.adjoint(plan), acc_scale=wscale)
                dx = torch.empty_like(x)
                ds = torch.zeros_like(s)
                n, c, h, wd = x.shape
                _lib.call("lcgan_modulate_bwd", _ptr(x), _ptr(t), _ptr(s), _ptr(dx), _ptr(ds), _dt(x), n, h * wd, c,
                          _stream(x), nbytes=3 * x.numel() * x.element_size())
                del t
            if need_w and _wgrad_enabled():
                xs = _modulate_raw(x, s)
                dw2 = tapconv_wgrad(xs, g, plan, x.shape[1], w.shape[0], scale=wscale)
                dw = unpack_wgrad(dw2, tuple(w.shape), False).contiguous()
                del xs
        if need_b or need_d:
            # r1 = sum_p dz * z with z = d*acc + bias (+ noise): remove the additive terms to get sum_p dz * acc
            if need_d and ctx.has_noise:
                r1 = r1 - (g.float() * ctx.noise[None, None]).sum(dim=(2, 3)) / d
            db, dd = _epilogue_grads(r0, r1, bias, d, bias_scale, need_b, need_d)
        return dx, ds, dw, db, dd, dnz, None, None, None, None, None, None, None, None


class Demod(torch.autograd.Function):
    """Demodulation coefficients d[b,o] = rsqrt(sum_c s[b,c]^2 Wsq[o,c] + eps), Wsq = sum_k (w c)^2
    (custom_layers.py:65-67), from the weights as the conv sees them (rounded to the compute dtype, scaled
    in fp32).  Forward and backward are one launch each (plus the cached Wsq table); first order only."""

    @staticmethod
    def forward(ctx, s, w, wscale, eps, dtype):
        _need_cuda(s, w)
        assert s.dtype == torch.float32 and s.is_contiguous() and w.dtype == torch.float32 and w.is_contiguous()
        wsq = weight_sq(w, wscale, dtype)
        b, i = s.shape
        o = w.shape[0]
        d = torch.empty((b, o), dtype=torch.float32, device=s.device)
        _lib.call("lcgan_demod_fwd", _ptr(s), _ptr(wsq), _ptr(d), b, o, i, C.c_float(eps), _stream(s))
        ctx.save_for_backward(s, w, d)
        ctx.cfg = (wscale, dtype)
        return d

    @staticmethod
    @once_differentiable
    def backward(ctx, dd):
        s, w, d = ctx.saved_tensors
        wscale, dtype = ctx.cfg
        b, i = s.shape
        o = w.shape[0]
        k = w.shape[2] * w.shape[3] if w.dim() == 4 else 1
        dd = dd.contiguous().float()
        ds = torch.empty_like(s) if ctx.needs_input_grad[0] else None
        dw = torch.empty_like(w) if (ctx.needs_input_grad[1] and _wgrad_enabled()) else None
        if ds is not None or dw is not None:
            _lib.call("lcgan_demod_bwd", _ptr(dd), _ptr(d), _ptr(s), _ptr(weight_sq(w, wscale, dtype)), _ptr(w),
                      _ptr(ds), _ptr(dw), b, o, i, k, C.c_float(wscale), BF16 if dtype == torch.bfloat16 else F32,
                      _stream(s))
        return ds, dw, None, None, None


class Warp(torch.autograd.Function):
    """Flow warp: bicubic grid_sample at linspace coords + tanh(flow)*scale
    (custom_layers.py:127-134,151,161-165).  flow is the box-filtered, pre-tanh field [b,2,H,W] f32."""

    @staticmethod
    def forward(ctx, x, flow, scale):
        _need_cuda(x, flow)
        x = _cl(x)
        flow = _cl(flow, torch.float32)
        n, c, h, w = x.shape
        out = torch.empty_like(x)
        _lib.call("lcgan_warp_fwd", _ptr(x), _ptr(flow), _ptr(out), _dt(x), n, h, w, c, C.c_float(scale),
                  _stream(x), nbytes=2 * x.numel() * x.element_size() + flow.numel() * 4)
        ctx.save_for_backward(x, flow)
        ctx.scale = scale
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        x, flow = ctx.saved_tensors
        dout = _cl(dout, x.dtype)
        n, c, h, w = x.shape
        dflow = torch.empty_like(flow)
        dx = torch.empty_like(x)
        # scratch for the large-flow (scatter) path only; small flows never touch it
        ws_acc = None if x.dtype == torch.float32 else torch.empty(x.numel(), dtype=torch.float32, device=x.device)
        ws_bounds = torch.empty(4 * n * ((h + 15) // 16) * ((w + 31) // 32) + 4, dtype=torch.int32, device=x.device)
        _lib.call("lcgan_warp_bwd_tiled", _ptr(x), _ptr(flow), _ptr(dout), _ptr(dx), _ptr(dflow), _ptr(ws_acc),
                  _ptr(ws_bounds), _dt(x), n, h, w, c, C.c_float(ctx.scale), _stream(x), tag="warp_bwd",
                  nbytes=3 * x.numel() * x.element_size() + 2 * flow.numel() * 4)
        return dx, dflow, None


class WarpMod(torch.autograd.Function):
    """out = warp(x, flow) * s[b,c]: the flow warp with the style of the modulated conv that consumes its output
    folded in (the to-RGB block reads the last synthesis block's output and nothing else does).  Backward: one
    pass turns the incoming gradient t into dout = t * s and the style gradient ds = sum_p t * out / s, then the
    plain warp backward runs on dout."""

    @staticmethod
    def forward(ctx, x, flow, s, scale):
        _need_cuda(x, flow, s)
        x = _cl(x)
        flow = _cl(flow, torch.float32)
        n, c, h, w = x.shape
        assert s.dtype == torch.float32 and s.is_contiguous() and tuple(s.shape) == (n, c)
        out = torch.empty_like(x)
        _lib.call("lcgan_warp_fwd_cs", _ptr(x), _ptr(flow), _ptr(out), _ptr(s), _dt(x), n, h, w, c, C.c_float(scale),
                  _stream(x), nbytes=2 * x.numel() * x.element_size() + flow.numel() * 4, tag="warp_fwd")
        ctx.save_for_backward(x, flow, s, out)
        ctx.scale = scale
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, t):
        x, flow, s, out = ctx.saved_tensors
        t = _cl(t, x.dtype)
        n, c, h, w = x.shape
        dout = torch.empty_like(x)
        ds = torch.zeros_like(s)
        _lib.call("lcgan_modulate_bwd", _ptr(out), _ptr(t), _ptr(s), _ptr(dout), _ptr(ds), _dt(x), n, h * w, c,
                  _stream(x), nbytes=3 * x.numel() * x.element_size())
        ds = torch.where(s != 0, ds / s, torch.zeros_like(ds))
        dflow = torch.empty_like(flow)
        dx = torch.empty_like(x)
        ws_acc = None if x.dtype == torch.float32 else torch.empty(x.numel(), dtype=torch.float32, device=x.device)
        ws_bounds = torch.empty(4 * n * ((h + 15) // 16) * ((w + 31) // 32) + 4, dtype=torch.int32, device=x.device)
        _lib.call("lcgan_warp_bwd_tiled", _ptr(x), _ptr(flow), _ptr(dout), _ptr(dx), _ptr(dflow), _ptr(ws_acc),
                  _ptr(ws_bounds), _dt(x), n, h, w, c, C.c_float(ctx.scale), _stream(x), tag="warp_bwd",
                  nbytes=3 * x.numel() * x.element_size() + 2 * flow.numel() * 4)
        return dx, dflow, ds, None


# ------------------------------------------------------------------------------------------
# losses
# ------------------------------------------------------------------------------------------
class L2Normalize(torch.autograd.Function):
    """F.normalize(x, dim=1) (cnn.py:40-41)."""

    @staticmethod
    def forward(ctx, x):
        _need_cuda(x)
        x = x.contiguous().float()
        b, d = x.shape
        y = torch.empty_like(x)
        inv = torch.empty((b,), dtype=torch.float32, device=x.device)
        _lib.call("lcgan_l2norm_fwd", _ptr(x), _ptr(y), _ptr(inv), b, d, _stream(x))
        ctx.save_for_backward(y, inv)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        y, inv = ctx.saved_tensors
        dy = dy.contiguous().float()
        dx = torch.empty_like(y)
        _lib.call("lcgan_l2norm_bwd", _ptr(y), _ptr(inv), _ptr(dy), _ptr(dx), y.shape[0], y.shape[1], _stream(y))
        return dx


class Contrastive(torch.autograd.Function):
    """Per-sample softplus((a.n - a.p)/tau)  (loss.py:9-14)."""

    @staticmethod
    def forward(ctx, a, p, n, tau):
        _need_cuda(a, p, n)
        a, p, n = (t.contiguous().float() for t in (a, p, n))
        b, d = a.shape
        l = torch.empty((b,), dtype=torch.float32, device=a.device)
        sig = torch.empty_like(l)
        _lib.call("lcgan_contrastive_fwd", _ptr(a), _ptr(p), _ptr(n), _ptr(l), _ptr(sig), b, d, C.c_float(tau),
                  _stream(a))
        ctx.save_for_backward(a, p, n, sig)
        ctx.tau = tau
        return l

    @staticmethod
    @once_differentiable
    def backward(ctx, dl):
        a, p, n, sig = ctx.saved_tensors
        dl = dl.contiguous().float()
        da, dp, dn = torch.empty_like(a), torch.empty_like(a), torch.empty_like(a)
        _lib.call("lcgan_contrastive_bwd", _ptr(a), _ptr(p), _ptr(n), _ptr(sig), _ptr(dl), _ptr(da), _ptr(dp),
                  _ptr(dn), a.shape[0], a.shape[1], C.c_float(ctx.tau), _stream(a))
        return da, dp, dn, None


class SumSq(torch.autograd.Function):
    """out[b] = sum_i x[b,i]^2  (R1 penalty reduction, loss.py:21-23)."""

    @staticmethod
    def forward(ctx, x):
        _need_cuda(x)
        x = x.contiguous().float()
        b = x.shape[0]
        out = torch.zeros((b,), dtype=torch.float32, device=x.device)
        _lib.call("lcgan_sumsq", _ptr(x), _ptr(out), b, C.c_int64(x.numel() // b), _stream(x))
        ctx.save_for_backward(x)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        (x,) = ctx.saved_tensors
        s = (dout.float() * 2.0).contiguous()
        dx = torch.empty_like(x)
        _lib.call("lcgan_rowscale", _ptr(x), _ptr(s), _ptr(dx), x.shape[0], C.c_int64(x.numel() // x.shape[0]),
                  _stream(x))
        return dx


_ema_tables = {}


def ema_lerp_(dst_tensors, src_tensors, decay: float, decay_dev=None):
    """dst = src.lerp(dst, decay) for every tensor pair, in ONE launch (ema.py:26-32).  decay_dev: optional
    device f32 scalar that overrides `decay` (lets a captured CUDA graph follow the start_iter schedule)."""
    pairs = [(d, s) for d, s in zip(dst_tensors, src_tensors) if d.numel() > 0]
    if not pairs:
        return
    dev = pairs[0][0].device
    _need_cuda(pairs[0][0])
    for d, s in pairs:
        assert d.dtype == torch.float32 and s.dtype == torch.float32 and d.is_contiguous() and s.is_contiguous()
    key = tuple((d.data_ptr(), s.data_ptr(), d.numel()) for d, s in pairs)
    table = _ema_tables.get(key)
    if table is None:
        # pointer table built once per parameter set (H2D copy outside any graph capture)
        table = torch.tensor([[k[0] for k in key], [k[1] for k in key], [k[2] for k in key]],
                             dtype=torch.int64).to(dev)
        _ema_tables.clear()
        _ema_tables[key] = table
    _lib.call("lcgan_ema_lerp", _ptr(table[0]), _ptr(table[1]), _ptr(table[2]), len(pairs), C.c_float(decay),
              _ptr(decay_dev), _stream(pairs[0][0]), nbytes=3 * 4 * sum(k[2] for k in key))
    bump_generation()        # the destination parameters changed behind autograd's version counters


def set_deterministic(flag: bool) -> bool:
    """Ordered (turn-taking) reductions instead of fp32 atomics in every kernel that has them: two runs
    on the same inputs are bit-identical.  Returns the previous setting.  (The rough-flow scatter
    fallback of the warp backward stays atomic; smooth flows - everything at and near initialisation -
    take the gather path, which is deterministic by construction.)"""
    global _DET
    _DET = bool(flag)
    return bool(_lib.lib().lcgan_set_deterministic(1 if flag else 0))
