"""mlx.nn.layers.upsample subset."""
import torch

from ... import core as mx


def upsample_nearest(x, scale_factor):
    sh, sw = scale_factor
    x = torch.repeat_interleave(x, int(sh), dim=1)
    x = torch.repeat_interleave(x, int(sw), dim=2)
    return mx.array(x)
