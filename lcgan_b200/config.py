"""Configuration containers of the training hot path (no compute).

`Config` carries the fields the reference constructors read from `args` (main.py:25-31,
cnn.py:10-15, 49-60); `Hyper` the loss weights / optimizer settings of the README recipes
(README.md:29/45/49, worker.py:98-110).  bench.py, the inference runner and the tests build the
models from these; the reference's own main.py passes its argparse namespace instead.
"""
import math
import types


class Config:
    def __init__(self, img_resolution=256, geo_noise_dim=64, app_noise_dim=64, geo_latent_dim=64,
                 app_latent_dim=512, geo_projection_dim=256, app_projection_dim=256, max_flow_scale=0.1):
        self.img_resolution = img_resolution
        self.geo_noise_dim = geo_noise_dim
        self.app_noise_dim = app_noise_dim
        self.geo_latent_dim = geo_latent_dim
        self.app_latent_dim = app_latent_dim
        self.geo_projection_dim = geo_projection_dim
        self.app_projection_dim = app_projection_dim
        self.max_flow_scale = max_flow_scale

    @property
    def num_blocks(self):                    # cnn.py:13, 52
        return int(math.log2(self.img_resolution)) - 2

    @property
    def base_nf(self):                       # cnn.py:17, 54
        return {1024: 32, 512: 64}.get(self.img_resolution, 128)

    def g_channels(self):
        """[(in, out, out_resolution)] per synthesis block (cnn.py:79-84)."""
        out, cin = [], 512
        for i in range(self.num_blocks):
            cout = min(self.base_nf << (self.num_blocks - 1 - i), 512)
            out.append((cin, cout, 8 << i))
            cin = cout
        return out

    def d_channels(self):
        """[(in, out)] per discriminator block (cnn.py:22-25)."""
        return [(min(self.base_nf << i, 512), min(self.base_nf << (i + 1), 512)) for i in range(self.num_blocks)]

    def namespace(self):
        return types.SimpleNamespace(**self.__dict__)


class Hyper:
    def __init__(self, tau=0.05, l_aux=0.5, l_r1=10.0, l_s=1e-7, lr=2e-3, beta1=0.0, beta2=0.99):
        self.tau, self.l_aux, self.l_r1, self.l_s = tau, l_aux, l_r1, l_s
        self.lr, self.beta1, self.beta2 = lr, beta1, beta2


def recipe(resolution):
    """(Hyper, freezeD_layer) of the README recipe for a resolution (README.md:27-57)."""
    return Hyper(lr=1e-3 if resolution == 1024 else 2e-3), {256: 3, 512: 4, 1024: 5}.get(resolution, 3)
