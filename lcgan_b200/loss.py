"""Losses with the surface of the reference's loss.py (loss.py:9-34), on lcgan_b200 kernels."""
import torch
from torch import autograd

from . import ops


def contrastive_loss(anchor, p_sample, n_sample, tau):
    """loss.py:9-15: mean_b -log(e^{a.p/tau} / (e^{a.p/tau} + e^{a.n/tau})), one warp per sample."""
    return ops.Contrastive.apply(anchor, p_sample, n_sample, float(tau)).mean()


def cal_r1_reg(adv_output, images, device):
    """loss.py:18-24: 0.5 * mean_b sum_chw (d sum(D(x)) / dx)^2 + images[:,0,0,0].mean()*0."""
    batch_size = images.size(0)
    grad_dout = cal_derivative(inputs=images, outputs=adv_output.sum(), device=device)
    assert grad_dout.size() == images.size()
    r1_reg = 0.5 * ops.SumSq.apply(grad_dout.reshape(batch_size, -1)).mean(0) + images[:, 0, 0, 0].mean() * 0
    return r1_reg


def cal_derivative(inputs, outputs, device):
    """loss.py:27-34.  The first-order pass only needs d outputs / d inputs, so weight/bias
    gradients are skipped inside it (they are produced by the later d_loss.backward())."""
    with ops.no_weight_gradients():
        grads = autograd.grad(outputs=outputs, inputs=inputs,
                              grad_outputs=torch.ones(outputs.size(), device=outputs.device),
                              create_graph=True, retain_graph=True, only_inputs=True)[0]
    return grads
