"""B200-native layer library with the constructor / call / state_dict surface of the reference's
custom_layers.py (class by class; reference line numbers in each docstring).

Logical tensors are NCHW like the reference's; physically activations are dense channels-last in
the activation dtype (bf16, or fp32 in the accurate mode - see ops.set_precision) and every
activation-sized operation is one of our CUDA kernels (ops.py).  Fusions relative to the reference
graph: bias + leaky-relu + gain (+ residual add) live in the conv epilogue; the style modulation is
applied to the activations and the demodulation to the accumulator rows, so per-sample weights are
never materialised; box filter + activation, nearest-up + box + add, and tanh + coordinate grid +
bicubic warp are single kernels.
"""
import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops, plans

SQRT2 = math.sqrt(2.0)
SQRT_HALF = math.sqrt(0.5)


def _as_act(x):
    """Logical NCHW -> dense channels-last in the activation dtype (no-op when already so)."""
    return ops._cl(x, ops.act_dtype())


class EqualizedWeight(nn.Module):
    """Reference custom_layers.py:7-14.  Parameter stored as randn/lr_mul under `.weight`; the
    runtime scale c = lr_mul/sqrt(fan_in) is applied to the fp32 accumulator in the conv epilogue."""

    def __init__(self, shape, lr_mul=1.0):
        super().__init__()
        self.c = 1 / np.sqrt(np.prod(shape[1:])) * lr_mul
        self.weight = nn.Parameter(torch.randn(shape).div_(lr_mul))

    def forward(self):
        return self.weight * self.c


class EqualizedLinear(nn.Module):
    """Reference custom_layers.py:17-25: x @ (W c)^T + bias*lr_mul, as a 1x1 tap conv."""

    def __init__(self, in_features, out_features, bias=0.0, lr_mul=1.0):
        super().__init__()
        self.weight = EqualizedWeight([out_features, in_features], lr_mul)
        self.bias = nn.Parameter(torch.ones(out_features) * bias)
        self.lr_mul = lr_mul

    def forward(self, x, slope=1.0, gain=1.0):
        return ops.linear_act(x, self.weight.weight, self.bias, wscale=float(self.weight.c),
                              bias_scale=float(self.lr_mul), slope=slope, gain=gain)


class EqualizedConv2d(nn.Module):
    """Reference custom_layers.py:28-44: conv k in {1,3}, stride in {1,2}, padding k//2."""

    def __init__(self, in_features, out_features, kernel_size, stride=1, no_bias=False, lr_mul=1.0):
        super().__init__()
        self.padding = kernel_size // 2
        self.kernel_size = kernel_size
        self.weight = EqualizedWeight([out_features, in_features, kernel_size, kernel_size], lr_mul)
        self.no_bias = no_bias
        if not self.no_bias:
            self.bias = nn.Parameter(torch.zeros([out_features]))
        self.stride = stride
        self.lr_mul = lr_mul

    def forward(self, x, slope=1.0, gain=1.0, residual=None, box=False):
        """slope/gain/residual: fused epilogue  lrelu(conv + bias, slope)*gain + residual.  box: follow with the 3x3
        box filter inside the same autograd node (DiscriminatorBlock: conv0 -> lrelu -> box_filter)."""
        plan = plans.conv(self.kernel_size, self.stride, x.shape[2], x.shape[3])
        if box:
            assert residual is None and not self.no_bias
            return ops.ConvActBox.apply(x, self.weight.weight, self.bias, float(self.weight.c), plan, slope, gain,
                                        float(self.lr_mul))
        return ops.conv_act(x, self.weight.weight, None if self.no_bias else self.bias, None, residual,
                            wscale=float(self.weight.c), plan=plan, slope=slope, gain=gain,
                            bias_scale=float(self.lr_mul))


class ModulatedConv2d(nn.Module):
    """Reference custom_layers.py:47-86 in shared-weight form:
         y = d[b,o] * conv(x * s[b,c], w) + bias,   d = rsqrt(sum_c s^2 Wsq[o,c] + eps),
       up=2 -> conv_transpose2d(stride 2, padding 1, output_padding 1) as four phase launches."""

    def __init__(self, in_features, out_features, kernel_size, up=1, eps=1e-8, lr_mul=1.0):
        super().__init__()
        self.in_features = in_features
        self.out_features = out_features
        self.kernel_size = kernel_size
        self.padding = (kernel_size - 1) // 2
        self.up = up
        self.weight = EqualizedWeight([out_features, in_features, kernel_size, kernel_size], lr_mul)
        self.bias = nn.Parameter(torch.zeros([out_features]))
        self.lr_mul = lr_mul
        self.eps = eps

    def forward(self, x, s, slope=1.0, gain=1.0, out_dtype=None, out_nchw=False, noise=None, premodulated=False):
        """noise: optional f32 [H_out, W_out] plane added between the conv (+bias) and the activation
        (SynthesisLayer's noise injection, fused into the conv epilogue).  premodulated: x is already x * s (its
        producer folded the style in, ops.Box3ActMod)."""
        w = self.weight.weight
        c = float(self.weight.c)
        # demodulation coefficients, fp32: d[b,o] = rsqrt(s^2 @ Wsq^T + eps), Wsq cached per weight version
        # (from the weights as the conv sees them: rounded to the compute dtype, scaled in fp32)
        s = s.float().contiguous()
        d = ops.Demod.apply(s, w, c, float(self.eps), ops.act_dtype())
        if self.up > 1:
            plan = plans.conv_transpose_up2(self.kernel_size, x.shape[2], x.shape[3])
        else:
            plan = plans.conv(self.kernel_size, 1, x.shape[2], x.shape[3])
        if noise is not None:
            noise = noise.float().contiguous()
        return ops.ModConvAct.apply(_as_act(x), s, w, self.bias, d, noise, c, plan, slope,
                                    gain, float(self.lr_mul), out_dtype or ops.act_dtype(), out_nchw, premodulated)


class SynthesisLayer(nn.Module):
    """Reference custom_layers.py:89-111: style affine (bias init 1) -> modulated conv -> noise."""

    def __init__(self, in_features, out_features, latent_dim, resolution, kernel_size=3, up=1, lr_mul=1.0,
                 use_noise=False):
        super().__init__()
        self.latent_dim = latent_dim
        self.up = up
        self.resolution = resolution
        self.use_noise = use_noise
        self.linear = EqualizedLinear(self.latent_dim, in_features, bias=1.0, lr_mul=1.0)
        self.modulated_conv = ModulatedConv2d(in_features, out_features, kernel_size, up=self.up, lr_mul=1.0)
        if self.use_noise:
            self.noise_gain = 0.01
            self.noise_strength = nn.Parameter(torch.zeros([]))
            self.register_buffer("noise_const", torch.randn([self.resolution, self.resolution]))

    def style(self, latent):
        """The per-sample channel scales of this layer (custom_layers.py:105): [b, in_features] f32."""
        return self.linear(latent.float()).float().contiguous()

    def forward(self, x, latent, slope=1.0, gain=1.0, out_dtype=None, out_nchw=False, style=None):
        """style: the result of self.style(latent) when the caller already folded it into x (x = activation * style)."""
        s = self.linear(latent.float()) if style is None else style
        if not self.use_noise:
            return self.modulated_conv(x, s, slope, gain, out_dtype, out_nchw, premodulated=style is not None)
        # custom_layers.py:108-110: x + noise_const * noise_strength * noise_gain, between the conv and any
        # activation -> a [res, res] plane (a weight-sized torch op, differentiable w.r.t. noise_strength)
        # that the conv epilogue adds before the leaky-relu.  cnn.py never enables it (use_noise=False).
        plane = self.noise_const * self.noise_strength * self.noise_gain
        return self.modulated_conv(x, s, slope, gain, out_dtype, out_nchw, noise=plane, premodulated=style is not None)


class SynthesisBlock(nn.Module):
    """Reference custom_layers.py:114-166: skip / flow / main branches, then the flow warp."""

    def __init__(self, in_features, out_features, g_latent_dim, a_latent_dim, resolution, max_flow_scale,
                 use_noise=False):
        super().__init__()
        self.resolution = resolution
        self.use_noise = use_noise
        self.max_flow_scale = max_flow_scale
        self.modulated_conv0 = SynthesisLayer(in_features, out_features, a_latent_dim, resolution, up=2,
                                              use_noise=self.use_noise)
        self.modulated_conv1 = SynthesisLayer(out_features, out_features, a_latent_dim, resolution, up=1,
                                              use_noise=self.use_noise)
        self.skip_layer = EqualizedConv2d(in_features, out_features, kernel_size=1, no_bias=True, lr_mul=1.0)
        self.flow_layer = SynthesisLayer(in_features, 2, g_latent_dim, resolution, up=2, use_noise=False)
        self.gain = np.sqrt(2)
        self.skip_gain = np.sqrt(0.5)

    def get_coordinates(self, b, h, w, device):
        """Reference :127-134 (kept for the surface; the warp kernel generates these on the fly)."""
        gy, gx = torch.meshgrid(torch.arange(h, dtype=torch.float32, device=device),
                                torch.arange(w, dtype=torch.float32, device=device), indexing='ij')
        return torch.stack(((2 * gx / (w - 1)) - 1, (2 * gy / (h - 1)) - 1)).unsqueeze(0).repeat([b, 1, 1, 1])

    def box_filter(self, x):
        return ops.Box3.apply(x)

    def forward(self, x, g_latent, a_latent, out_style=None):
        """out_style [b, out_features] (optional): the style of the ONE modulated conv that consumes this block's
        output (the to-RGB block after the last synthesis block); it is then folded into the warp pass and the block
        returns the modulated features."""
        g_lat = g_latent[:, 0]
        a_lat0, a_lat1 = a_latent[:, 0], a_latent[:, 1]
        x = _as_act(x)
        skip_lo = self.skip_layer(x, gain=float(self.skip_gain))               # 1x1 at the low resolution
        flow = self.flow_layer(x, g_lat, out_dtype=torch.float32)                # fp32: sub-pixel offsets
        flow = ops.Box3.apply(flow)
        t = self.modulated_conv0(x, a_lat0)                                       # x2 up-conv + bias
        if ops.box3_mod_eligible(t):
            # box -> lrelu * sqrt2 -> * style of the next conv, one pass; the conv reads the modulated tensor
            s1 = self.modulated_conv1.style(a_lat1)
            t = ops.Box3ActMod.apply(t, s1, 0.2, float(self.gain))
            t = self.modulated_conv1(t, a_lat1, slope=0.2, style=s1)              # conv -> lrelu
        else:
            t = ops.Box3Act.apply(t, 0.2, float(self.gain))                       # box -> lrelu * sqrt2
            t = self.modulated_conv1(t, a_lat1, slope=0.2)                        # conv -> lrelu
        y = ops.Up2BoxAdd.apply(skip_lo, t)                                       # box(up2(skip)) + t
        if out_style is not None:
            return ops.WarpMod.apply(y, flow, out_style, float(self.max_flow_scale))
        return ops.Warp.apply(y, flow, float(self.max_flow_scale))                # tanh + grid + bicubic


class ToRGBBlock(nn.Module):
    """Reference custom_layers.py:169-182: modconv3x3 -> lrelu -> modconv1x1 (demodulated)."""

    def __init__(self, in_features, out_features, a_latent_dim, resolution, use_noise=False):
        super().__init__()
        self.resolution = resolution
        self.use_noise = use_noise
        self.modulated_conv0 = SynthesisLayer(in_features, in_features, a_latent_dim, resolution,
                                              use_noise=self.use_noise)
        self.modulated_conv1 = SynthesisLayer(in_features, out_features, a_latent_dim, resolution, kernel_size=1,
                                              use_noise=False)

    def forward(self, x, a_latent, style0=None):
        """style0: self.modulated_conv0.style(...) when x already carries it (SynthesisBlock(out_style=...))."""
        x = self.modulated_conv0(_as_act(x), a_latent[:, 0], slope=0.2, style=style0)
        # the image leaves the generator as NCHW fp32, like the reference's
        return self.modulated_conv1(x, a_latent[:, 1], out_dtype=torch.float32, out_nchw=True)


class DiscriminatorBlock(nn.Module):
    """Reference custom_layers.py:185-217."""

    def __init__(self, in_features, out_features, skip=False):
        super().__init__()
        self.conv0 = EqualizedConv2d(in_features, in_features, kernel_size=3, lr_mul=1.0)
        self.conv1 = EqualizedConv2d(in_features, out_features, kernel_size=3, stride=2, lr_mul=1.0)
        self.skip = skip
        if self.skip:
            self.skip_layer = EqualizedConv2d(in_features, out_features, kernel_size=1, no_bias=True, lr_mul=1.0)
            self.gain = np.sqrt(2)
            self.skip_gain = np.sqrt(0.5)

    def box_filter(self, x):
        return ops.Box3.apply(x)

    def forward(self, x):
        x = _as_act(x)
        if self.skip:
            x, pooled = ops.PoolFork.apply(x, 0.25)                            # (x, avg_pool2d(x, 2)): one backward node
            t = self.conv0(x, slope=0.2, gain=float(self.gain), box=True)      # conv -> lrelu * sqrt2 -> box filter
            t = self.conv1(t, slope=0.2)
            # skip*sqrt(.5) + t, with the add fused into the (activation-free) skip conv epilogue
            return self.skip_layer(pooled, gain=float(self.skip_gain), residual=t)
        t = self.conv0(x, slope=0.2)
        t = ops.Box3.apply(t)
        return self.conv1(t, slope=0.2)


class MinibatchStdLayer(nn.Module):
    """Reference custom_layers.py:237-256.  [b,C,4,4] only: a handful of tiny fp32 torch ops
    (autograd supplies the genuine second derivative R1 needs); group members are strided."""

    def __init__(self, group_size, num_channels=1):
        super().__init__()
        self.group_size = group_size
        self.num_channels = num_channels

    def forward(self, x):
        N, C, H, W = x.shape
        G = min(self.group_size, N) if self.group_size is not None else N
        Fc = self.num_channels
        c = C // Fc
        y = x.float().reshape(G, -1, Fc, c, H, W)
        y = y - y.mean(dim=0)
        y = y.square().mean(dim=0)
        y = (y + 1e-8).sqrt()
        y = y.mean(dim=[2, 3, 4])
        y = y.reshape(-1, Fc, 1, 1)
        y = y.repeat(G, 1, H, W)
        return torch.cat([x, y.to(x.dtype)], dim=1)


class DiscriminatorEpilogue(nn.Module):
    """Reference custom_layers.py:220-234."""

    def __init__(self, in_features, resolution, mbstd_group_size=4):
        super().__init__()
        self.resolution = resolution
        self.mb_std = MinibatchStdLayer(group_size=mbstd_group_size)
        self.conv = EqualizedConv2d(in_features + 1, in_features, kernel_size=3, lr_mul=1.0)
        self.linear = EqualizedLinear(in_features * (resolution ** 2), in_features, lr_mul=0.01)

    def forward(self, x):
        x = self.mb_std(_as_act(x))
        pad = (-x.shape[1]) % 64
        if pad and ops.act_dtype() == torch.bfloat16:
            # 513 -> 576 zero channels (a [b,.,4,4] tensor and a 2.6 M-element weight: tiny) so the
            # layer runs on the tensor-core path, which needs Cin % 64 == 0
            x = _as_act(F.pad(x, (0, 0, 0, 0, 0, pad)))
            w = F.pad(self.conv.weight.weight, (0, 0, 0, 0, 0, pad))
            x = ops.conv_act(x, w, self.conv.bias, None, None, wscale=float(self.conv.weight.c),
                             plan=plans.conv(3, 1, x.shape[2], x.shape[3]), slope=0.2,
                             bias_scale=float(self.conv.lr_mul))
        else:
            x = self.conv(x, slope=0.2)
        return self.linear(x.flatten(1), slope=0.2)


class MappingNetwork(nn.Module):
    """Reference custom_layers.py:259-287: L = Q(tanh(basis)) diag(|d|+1e-6); x = L z; 12 linear
    layers (lr_mul .01, no activation).  All fp32; the 64x64 QR stays on torch.linalg."""

    def __init__(self, channels_list, lr_mul=0.01):
        super().__init__()
        self.eps = 1e-6
        self.matrix_size = channels_list[0]
        self.diagonal_params = nn.Parameter(torch.randn([self.matrix_size]))
        self.basis_params = nn.Parameter(torch.randn([self.matrix_size, self.matrix_size]))
        self.num_layers = len(channels_list) - 1
        self.mlp = nn.Sequential(*[EqualizedLinear(channels_list[i], channels_list[i + 1], lr_mul=lr_mul)
                                   for i in range(self.num_layers)])

    def orthogonalize(self, matrix):
        Q, _ = torch.linalg.qr(matrix)
        return Q

    def forward(self, z):
        B = self.orthogonalize(torch.tanh(self.basis_params))
        L = B * (torch.abs(self.diagonal_params) + self.eps)[None, :]
        x = ops.linear_act(z.float(), L, None, wscale=1.0, bias_scale=1.0)
        return self.mlp(x)


class ProjectionHead(nn.Module):
    """Reference custom_layers.py:290-306.  The nn.LeakyReLU entries keep the state_dict indices
    (mlp.0, mlp.2, mlp.4); at run time each is fused into the preceding linear's epilogue."""

    def __init__(self, channels_list, lr_mul=0.01):
        super().__init__()
        self.num_layers = len(channels_list) - 1
        if self.num_layers > 0:
            mlp = []
            for idx in range(self.num_layers):
                mlp += [EqualizedLinear(channels_list[idx], channels_list[idx + 1], lr_mul=lr_mul)]
                if idx < self.num_layers - 1:
                    mlp += [nn.LeakyReLU(0.2)]
            self.mlp = nn.Sequential(*mlp)

    def forward(self, z):
        mods = list(self.mlp)
        x, i = z, 0
        while i < len(mods):
            m = mods[i]
            if isinstance(m, EqualizedLinear) and i + 1 < len(mods) and isinstance(mods[i + 1], nn.LeakyReLU):
                x = m(x, slope=float(mods[i + 1].negative_slope))
                i += 2
            else:
                x = m(x)
                i += 1
        return x
