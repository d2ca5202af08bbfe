"""Generator / Discriminator with the surface of the reference's cnn.py (cnn.py:7-115): same
constructor argument (`args` namespace), same call signatures and return values, same
parameter/buffer names, shapes and dtypes (so reference checkpoints and worker.py's freezeD /
diagonal_params attribute paths keep working) - built from the lcgan_b200 CUDA layers.
"""
import math

import torch
import torch.nn as nn

from . import ops
from .custom_layers import *  # noqa: F401,F403  (reference cnn.py:4 re-exports the layer classes too)
from .custom_layers import (DiscriminatorBlock, DiscriminatorEpilogue, EqualizedConv2d, MappingNetwork,
                            ProjectionHead, SynthesisBlock, ToRGBBlock)

MAX_NF = 512


def _base_nf(resolution):
    # cnn.py:17,54
    return {1024: 32, 512: 64}.get(resolution, 128)


def _num_blocks(resolution, edge=4):
    return int(math.log2(resolution)) - int(math.log2(edge))


class Discriminator(torch.nn.Module):
    """cnn.py:7-43.  shared_model children stay ['0' from-RGB conv, '1' LeakyReLU, '2'.. blocks]
    (worker.py:128-131 counts on that order for freezeD); at run time the LeakyReLU is fused into
    the from-RGB conv's epilogue."""

    def __init__(self, args):
        super().__init__()
        self.img_resolution = args.img_resolution
        self.last_block_resolution = 4
        self.log_last_block_resolution = 2
        self.num_blocks = _num_blocks(self.img_resolution)
        self.geo_projection_dim = args.geo_projection_dim
        self.app_projection_dim = args.app_projection_dim
        self.max_nf = MAX_NF
        self.base_nf = _base_nf(self.img_resolution)

        widths = [min(self.base_nf << i, MAX_NF) for i in range(self.num_blocks + 1)]
        stem = [EqualizedConv2d(3, self.base_nf, kernel_size=1), nn.LeakyReLU(0.2)]
        body = [DiscriminatorBlock(cin, cout, skip=True) for cin, cout in zip(widths[:-1], widths[1:])]
        self.shared_model = nn.Sequential(*stem, *body)
        top = widths[-1]
        self.discriminator_epilogue = DiscriminatorEpilogue(top, resolution=4, mbstd_group_size=8)
        self.logit_mapper = ProjectionHead([top, 1])
        self.projection_header1 = ProjectionHead([top * 16, top * 4, top, self.geo_projection_dim])
        self.projection_header2 = ProjectionHead([top * 16, top * 4, top, self.app_projection_dim])

    def _features(self, image):
        mods = list(self.shared_model)
        h = mods[0](image, slope=float(mods[1].negative_slope))     # NCHW fp32 image read directly
        for blk in mods[2:]:
            h = blk(h)
        return h

    def forward(self, image, get_embedding_features=False):
        h = self._features(image)
        logit = self.logit_mapper(self.discriminator_epilogue(h))
        geometry_embedding = appearance_embedding = None
        if get_embedding_features:
            x = h.flatten(1)
            geometry_embedding = ops.L2Normalize.apply(self.projection_header1(x))
            appearance_embedding = ops.L2Normalize.apply(self.projection_header2(x))
        return logit, geometry_embedding, appearance_embedding


class Generator(torch.nn.Module):
    """cnn.py:46-115."""

    def __init__(self, args):
        super().__init__()
        self.img_resolution = args.img_resolution
        self.first_block_resolution = 4
        self.log_first_block_resolution = 2
        self.num_blocks = _num_blocks(self.img_resolution)
        self.max_nf = MAX_NF
        self.base_nf = _base_nf(self.img_resolution)
        self.geo_latent_dim = args.geo_latent_dim
        self.app_latent_dim = args.app_latent_dim
        self.geo_noise_dim = args.geo_noise_dim
        self.app_noise_dim = args.app_noise_dim
        self.max_flow_scale = args.max_flow_scale

        self.w_avg_beta = 0.998
        self.register_buffer("avg_latent1", torch.zeros([self.geo_latent_dim]))
        self.register_buffer("avg_latent2", torch.zeros([self.app_latent_dim]))

        g, a = self.geo_latent_dim, self.app_latent_dim
        self.geometry_mapping = MappingNetwork([self.geo_noise_dim] + [g] * 12)
        self.appearance_mapping = MappingNetwork([self.app_noise_dim, a // 4, a // 2] + [a] * 10)
        self.const = torch.nn.Parameter(torch.randn([MAX_NF, 4, 4]))

        widths = [MAX_NF] + [min(self.base_nf << (self.num_blocks - 1 - i), MAX_NF) for i in range(self.num_blocks)]
        self.model = nn.Sequential(*[
            SynthesisBlock(widths[i], widths[i + 1], g, a, 8 << i, self.max_flow_scale, use_noise=False)
            for i in range(self.num_blocks)])
        self.rgb_layer = ToRGBBlock(widths[-1], 3, a, self.img_resolution, use_noise=False)

    def forward(self, rand_noise1, rand_noise2, w_psi=-1.0):
        batch_size = rand_noise1.size(0)
        geometry_code = self.geometry_mapping(rand_noise1)
        appearance_code = self.appearance_mapping(rand_noise2)

        if w_psi <= 0:   # running latent average, every training-mode forward (cnn.py:95-97)
            self.avg_latent1.copy_(geometry_code.detach().mean(0).lerp(self.avg_latent1, self.w_avg_beta))
            self.avg_latent2.copy_(appearance_code.detach().mean(0).lerp(self.avg_latent2, self.w_avg_beta))
        if w_psi > 0.0:  # truncation trick (cnn.py:99-101)
            geometry_code = self.avg_latent1.lerp(geometry_code, w_psi)
            appearance_code = self.avg_latent2.lerp(appearance_code, w_psi)

        # one geometry code per block, two appearance codes per block + two for to-RGB (cnn.py:103-104);
        # expand() instead of repeat(): the codes are only read
        g_codes = geometry_code.unsqueeze(1)
        a_codes = appearance_code.unsqueeze(1).expand(-1, 2, -1)
        x = self.const.unsqueeze(0).expand(batch_size, -1, -1, -1)
        blocks = list(self.model)
        for block in blocks[:-1]:
            x = block(x, g_codes, a_codes)
        # the to-RGB block is the only reader of the last block's output: its first style is folded into that
        # block's warp pass (x * s never makes a pass of its own over HBM)
        last = blocks[-1]
        res, c = last.resolution, self.rgb_layer.modulated_conv0.modulated_conv.in_features
        style0 = None
        if ops.fold_style_eligible(res, res, c, x.is_cuda):
            style0 = self.rgb_layer.modulated_conv0.style(a_codes[:, 0])
        x = last(x, g_codes, a_codes, out_style=style0)
        return self.rgb_layer(x, a_codes, style0=style0)
