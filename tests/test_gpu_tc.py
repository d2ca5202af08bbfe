"""tcgen05 kernels against the CUDA-core kernels on identical bf16 operands (both accumulate in
fp32, so they must agree to accumulation-order noise), over the lattice shapes the networks use:
stride-1/2, the four transposed-conv phases, 1x1, linear, small maps with batch folded into the
tile, partial batch tiles, every epilogue feature."""
import pytest
import torch

from conftest import rel_l2

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _setup():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    torch.manual_seed(0)
    from lcgan_b200 import ops
    ops.set_precision("bf16")
    yield
    ops.set_tensor_cores(True)
    ops.set_wgrad_tensor_cores(True)
    ops.set_up2_fused(False)
    ops._UP2_FUSED_MAX_CIN = 64
    ops._UP2_HALO = True


def _cl(x):
    return x.contiguous(memory_format=torch.channels_last)


def _plan(kind, k, H, W):
    from lcgan_b200 import plans
    if kind == "up2":
        return plans.conv_transpose_up2(3, H, W)
    if kind == "s2":
        return plans.conv(k, 2, H, W)
    if kind == "s2_adj":
        return plans.adjoint(plans.conv(k, 2, 2 * H, 2 * W))
    if kind == "up2_adj":
        return plans.adjoint(plans.conv_transpose_up2(3, H // 2, W // 2))
    return plans.conv(k, 1, H, W)


CASES = [  # kind, k, N, Cin, Cout, H, W
    ("s1", 3, 2, 64, 64, 16, 16), ("s1", 3, 3, 128, 256, 8, 8), ("s1", 3, 5, 64, 32, 4, 4),
    ("s1", 1, 2, 64, 128, 32, 32), ("s2", 3, 2, 64, 128, 16, 16), ("up2", 3, 2, 128, 64, 8, 8),
    ("s2_adj", 3, 2, 128, 64, 8, 8), ("up2_adj", 3, 2, 64, 128, 16, 16), ("s1", 1, 32, 8192, 512, 1, 1),
    ("s1", 3, 1, 192, 48, 64, 32),
    # 32-channel layers of the 1024x1024 / 512x512 models: 64-byte swizzle rows
    ("s1", 3, 2, 32, 32, 32, 32), ("s2", 3, 2, 32, 64, 32, 32), ("up2", 3, 2, 64, 32, 16, 16),
    ("s1", 1, 2, 32, 64, 16, 16), ("s2_adj", 3, 2, 64, 32, 16, 16), ("s1", 3, 1, 96, 160, 16, 16),
    # many tiles per persistent CTA (ring wrap-around, TMEM stage reuse, the two-issuer resident mode)
    ("s1", 3, 8, 32, 32, 128, 128), ("s1", 3, 4, 64, 64, 128, 128), ("s1", 3, 4, 128, 128, 128, 64),
    ("s2", 3, 4, 64, 128, 128, 128),
    # haloed single-copy tiles (3x3 stride 1, Cin 32 / 64, Cout <= 128): 8 x 16 lattice tiles, image borders on every
    # side of a tile, one tile per image, 16 / 64 output channels, many images, non-square maps
    ("s1", 3, 3, 32, 16, 16, 8), ("s1", 3, 2, 32, 64, 32, 16), ("s1", 3, 33, 64, 64, 16, 32), ("s1", 3, 2, 32, 32, 256, 256),
    ("s1", 3, 2, 64, 48, 64, 128),
    # haloed stride-2 mode (3x3, 32 input channels, pixel pairs as 128-byte rows): one tile, many tiles, non-square,
    # 32 / 128 output channels, odd image counts
    ("s2", 3, 3, 32, 64, 64, 128), ("s2", 3, 2, 32, 32, 32, 16), ("s2", 3, 2, 32, 128, 64, 64), ("s2", 3, 5, 32, 64, 256, 256),
    # widths that are not multiples of the tile width, a single row of tiles, mixed 32 / 64 channel counts
    ("s1", 3, 3, 32, 32, 8, 24), ("s1", 3, 2, 64, 64, 24, 40), ("s1", 3, 1, 64, 32, 40, 72), ("s1", 3, 2, 32, 64, 8, 16),
    ("s1", 3, 5, 32, 32, 64, 1024),
]


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("out_f32", [False, True])
def test_tc_forward_matches_simt(case, out_f32):
    from lcgan_b200 import ops, _lib
    kind, k, N, Cin, Cout, H, W = case
    plan = _plan(kind, k, H, W)
    x = _cl((torch.randn(N, Cin, plan.IH, plan.IW, device="cuda")).bfloat16())
    w2 = (torch.randn(Cout, k * k * Cin, device="cuda") / (k * k * Cin) ** 0.5).bfloat16()
    rs = torch.rand(N, Cout, device="cuda") + 0.5
    bias = torch.randn(Cout, device="cuda")
    dt = torch.float32 if out_f32 else torch.bfloat16
    res = _cl(torch.randn(N, Cout, plan.OH, plan.OW, device="cuda").to(dt))
    outs = []
    for tc in (False, True):
        ops.set_tensor_cores(tc)
        before = _lib.launches
        y = ops.empty_cl(N, Cout, plan.OH, plan.OW, dt, "cuda").fill_(float("nan"))
        ops.tapconv(x, w2, y, plan, rs, bias, None, slope=0.2, gain=1.4, bias_scale=0.5)
        y2 = ops.empty_cl(N, Cout, plan.OH, plan.OW, dt, "cuda").fill_(float("nan"))
        ops.tapconv(x, w2, y2, plan, None, None, res, slope=1.0, gain=0.7)
        outs.append((y.float(), y2.float()))
    torch.cuda.synchronize()
    (a, a2), (b, b2) = outs
    assert torch.isfinite(b).all() and torch.isfinite(b2).all()
    tol = 1e-5 if out_f32 else 4e-3
    assert rel_l2(b, a) < tol, f"{case}: {rel_l2(b, a)}"
    assert rel_l2(b2, a2) < tol, f"{case} residual: {rel_l2(b2, a2)}"


@pytest.mark.parametrize("case", [("s1", 3, 2, 32, 32, 32, 32), ("s1", 3, 33, 64, 64, 16, 32), ("s1", 3, 2, 32, 32, 256, 256),
                                  ("s1", 3, 3, 32, 32, 8, 24), ("s1", 3, 2, 64, 64, 24, 40), ("s1", 3, 1, 64, 32, 40, 72),
                                  ("s1", 3, 2, 32, 64, 8, 16), ("s1", 3, 5, 32, 32, 64, 1024)])
def test_tc_forward_dx_on_n_mode_matches(case, monkeypatch):
    """The opt-in dx-on-N mode (LCGAN_DXN=1: one MMA per dy with N = 3 * Cout, the epilogue adds the dx column groups of
    neighbouring pixels; conv_tc.cu explains why it is not the default) against the CUDA-core path."""
    monkeypatch.setenv("LCGAN_DXN", "1")
    test_tc_forward_matches_simt(case, False)


@pytest.mark.parametrize("case", [("s1", 3, 2, 32, 32, 32, 32), ("s1", 3, 33, 64, 64, 16, 32), ("s1", 3, 2, 32, 64, 32, 16)])
def test_tc_wgrad_halo_nine_view_form_still_matches(case, monkeypatch):
    """The haloed weight gradient puts the dy taps on N by default (one MMA per K step); the older form with one MMA per
    dy stays selectable (LCGAN_WG_NO_DYN=1) and must stay right."""
    monkeypatch.setenv("LCGAN_WG_NO_DYN", "1")
    test_tc_wgrad_matches_simt(case)


@pytest.mark.parametrize("case", [c for c in CASES if c[4] % 32 == 0])
def test_tc_wgrad_matches_simt(case):
    from lcgan_b200 import ops
    kind, k, N, Cin, Cout, H, W = case
    plan = _plan(kind, k, H, W)
    x = _cl(torch.randn(N, Cin, plan.IH, plan.IW, device="cuda").bfloat16())
    g = _cl(torch.randn(N, Cout, plan.OH, plan.OW, device="cuda").bfloat16())
    ops.set_wgrad_tensor_cores(False)
    ref = ops.tapconv_wgrad(x, g, plan, Cin, Cout)
    ops.set_wgrad_tensor_cores(True)
    out = ops.tapconv_wgrad(x, g, plan, Cin, Cout)
    torch.cuda.synchronize()
    assert rel_l2(out, ref) < 1e-5, f"{case}: {rel_l2(out, ref)}"


@pytest.mark.parametrize("case", [  # N, Cin, Cout, H, W
    (2, 64, 32, 16, 16), (1, 128, 64, 8, 16), (3, 32, 16, 16, 8), (2, 64, 32, 64, 64), (5, 32, 32, 4, 4),
    # haloed form (Cin 32 / 64, Cout <= 32, H % 16 == 0, W % 8 == 0): many tiles, many images, non-square
    (3, 64, 32, 128, 64), (33, 32, 16, 16, 8), (2, 64, 16, 32, 32),
])
@pytest.mark.parametrize("out_f32", [False, True])
def test_fused_up2_matches_four_phase_path(case, out_f32):
    """Experimental single-launch x2 transposed conv (blocked output channels) against the four phase
    launches on identical bf16 operands, with every epilogue feature."""
    from lcgan_b200 import ops, plans, _lib
    N, Cin, Cout, H, W = case
    plan = plans.conv_transpose_up2(3, H, W)
    x = _cl(torch.randn(N, Cin, H, W, device="cuda").bfloat16())
    w2 = (torch.randn(Cout, 9 * Cin, device="cuda") / (9 * Cin) ** 0.5).bfloat16()
    rs = torch.rand(N, Cout, device="cuda") + 0.5
    bias = torch.randn(Cout, device="cuda")
    dt = torch.float32 if out_f32 else torch.bfloat16
    outs = []
    for fused in (False, True):
        ops.set_up2_fused(fused)
        ops._UP2_HALO = fused                 # the four-phase reference path must not take the haloed form either
        ops._UP2_FUSED_MAX_CIN = 128          # exercise the two-n-tile case as well
        before = _lib.launches
        y = ops.empty_cl(N, Cout, 2 * H, 2 * W, dt, "cuda").fill_(float("nan"))
        ops.tapconv(x, w2, y, plan, rs, bias, None, slope=0.2, gain=1.4, bias_scale=0.5)
        assert _lib.launches - before == (1 if fused else 4)
        outs.append(y.float())
    torch.cuda.synchronize()
    a, b = outs
    assert torch.isfinite(b).all()
    assert rel_l2(b, a) < (1e-5 if out_f32 else 4e-3), f"{case}: {rel_l2(b, a)}"


@pytest.mark.parametrize("case", [(2, 64, 16, 16), (3, 128, 8, 16), (1, 32, 64, 32), (5, 64, 4, 4)])   # N, Cin, H, W
def test_flow_layer_on_tensor_cores_matches_thin_kernel(case):
    """The flow layers (conv_transpose2d k3 s2, C -> 2, fp32 out: custom_layers.py:78,150) as ONE tensor-core launch
    (narrow blocked epilogue: 4 phases x 2 channels) against the CUDA-core all-phase kernel, with every epilogue term;
    the fused weight comes from the multi-tensor pack kernel (mode 3) and must equal the torch construction."""
    from lcgan_b200 import ops, plans, _lib
    N, Cin, H, W = case
    plan = plans.conv_transpose_up2(3, H, W)
    x = _cl(torch.randn(N, Cin, H, W, device="cuda").bfloat16())
    wparam = torch.nn.Parameter(torch.randn(2, Cin, 3, 3, device="cuda") / (9 * Cin) ** 0.5)
    w2 = ops.pack_weight(wparam, False, torch.bfloat16)
    wf = ops._derive(wparam, ("up2f", torch.bfloat16))
    assert torch.equal(wf, ops.fused_up2_weights(w2, Cin))
    rs = torch.rand(N, 2, device="cuda") + 0.5
    bias = torch.randn(2, device="cuda")
    outs = []
    for fused in (None, wf):
        before = dict(_lib.counts)
        y = ops.empty_cl(N, 2, 2 * H, 2 * W, torch.float32, "cuda").fill_(float("nan"))
        ops.tapconv(x, w2, y, plan, rs, bias, None, slope=0.2, gain=1.4, bias_scale=0.5, acc_scale=0.7, up2f=fused)
        key = "lcgan_tapconv_tc_blocked" if fused is not None else "lcgan_tapconv_up2_thin"
        assert _lib.counts.get(key, 0) == before.get(key, 0) + 1
        outs.append(y)
    torch.cuda.synchronize()
    assert torch.isfinite(outs[1]).all()
    assert rel_l2(outs[1], outs[0]) < 1e-5, f"{case}: {rel_l2(outs[1], outs[0])}"


@pytest.mark.parametrize("case", [(2, 64, 16, 16), (3, 128, 8, 16), (1, 32, 64, 32), (4, 64, 32, 32)])   # N, Cin, H, W
def test_flow_layer_wgrad_on_tensor_cores_matches_thin_kernel(case):
    """Weight gradient of the flow layers (C -> 2, x2 transposed, fp32 gradient): gather of the 9 x 2 gradient taps +
    a pointwise tensor-core wgrad against the CUDA-core all-tap kernel (the gather rounds the gradient to bf16)."""
    from lcgan_b200 import ops, plans, _lib
    N, Cin, H, W = case
    plan = plans.conv_transpose_up2(3, H, W)
    x = _cl(torch.randn(N, Cin, H, W, device="cuda").bfloat16())
    g = _cl(torch.randn(N, 2, 2 * H, 2 * W, device="cuda"))
    outs = []
    for tc in (False, True):
        ops._FLOW_TC = tc
        before = _lib.counts.get("lcgan_flow_grad_im2col", 0)
        outs.append(ops.tapconv_wgrad(x, g if tc else g.bfloat16().float(), plan, Cin, 2, scale=0.37))
        assert _lib.counts.get("lcgan_flow_grad_im2col", 0) == before + (1 if tc else 0)
    ops._FLOW_TC = True
    torch.cuda.synchronize()
    assert outs[1].shape == outs[0].shape == (2, 9 * Cin)
    assert rel_l2(outs[1], outs[0]) < 1e-5, f"{case}: {rel_l2(outs[1], outs[0])}"
