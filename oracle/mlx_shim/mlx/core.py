"""mlx.core subset (see package docstring)."""
from __future__ import annotations

import math
from typing import Sequence

import numpy as np
import torch

float32 = torch.float32
float16 = torch.float16
bfloat16 = torch.bfloat16
int16 = torch.int16
int32 = torch.int32
int64 = torch.int64
uint8 = torch.uint8
bool_ = torch.bool
cpu = "cpu"


class array(torch.Tensor):
    """torch.Tensor with the handful of mx.array method signatures that differ."""

    @staticmethod
    def __new__(cls, data=None, dtype=None):
        if isinstance(data, torch.Tensor):
            t = data.detach()
            if dtype is not None:
                t = t.to(dtype)
        else:
            t = torch.as_tensor(np.asarray(data))
            if dtype is not None:
                t = t.to(dtype)
            elif t.dtype == torch.float64:
                t = t.to(torch.float32)
            elif t.dtype == torch.int64:
                t = t.to(torch.int32)
        return torch.Tensor._make_subclass(cls, t, False)

    def __init__(self, *a, **k):
        pass

    def transpose(self, *axes):
        if len(axes) == 1 and isinstance(axes[0], (tuple, list)):
            axes = tuple(axes[0])
        if not axes:
            axes = tuple(reversed(range(self.ndim)))
        return self.permute(*axes)

    def astype(self, dtype):
        return self.to(dtype)

    def __rtruediv__(self, other):
        # torch lowers scalar/tensor to reciprocal()*scalar (two roundings); MLX divides once
        if not isinstance(other, torch.Tensor):
            dt = self.dtype if self.is_floating_point() else torch.float32
            other = torch.full((), other, dtype=dt)
        return array(torch.div(other, self))

    def square(self):
        return self * self

    def squeeze(self, axis=None):
        if axis is None:
            return torch.Tensor.squeeze(self)
        return torch.Tensor.squeeze(self, axis)

    def flatten(self, start_axis=0, end_axis=-1):
        return torch.Tensor.flatten(self, start_axis, end_axis)

    @property
    def T(self):
        return self.transpose()

    def tolist(self):
        return torch.Tensor.tolist(self.float() if self.dtype == torch.bfloat16 else self)

    def item(self):
        return torch.Tensor.item(self)


def _a(x):
    return x if isinstance(x, array) else array(x)


def eval(*args):  # noqa: A001  (lazy-graph flush: nothing to do)
    return None


def compile(fun=None, **kwargs):  # noqa: A001
    if fun is None:
        return lambda f: f
    return fun


def stop_gradient(x):
    return x


def arange(start, stop=None, step=1, dtype=None):
    if stop is None:
        start, stop = 0, start
    is_f = any(isinstance(v, float) for v in (start, stop, step))
    dtype = dtype or (torch.float32 if is_f else torch.int32)
    return _a(torch.arange(start, stop, step, dtype=dtype))


def linspace(start, stop, num=50, dtype=float32):
    # MLX: (1 - i/(n-1)) * start + (i/(n-1)) * stop in f32
    i = torch.arange(num, dtype=torch.float32) / float(num - 1)
    return _a(((1.0 - i) * float(start) + i * float(stop)).to(dtype))


def zeros(shape, dtype=float32):
    return _a(torch.zeros(tuple(shape), dtype=dtype))


def full(shape, vals, dtype=None):
    return _a(torch.full(tuple(shape), vals, dtype=dtype or torch.float32))


def concatenate(arrs, axis=0):
    return _a(torch.cat(list(arrs), dim=axis))


def stack(arrs, axis=0):
    return _a(torch.stack(list(arrs), dim=axis))


def split(x, indices_or_sections, axis=0):
    if isinstance(indices_or_sections, int):
        return [_a(t) for t in torch.chunk(x, indices_or_sections, dim=axis)]
    idx = list(indices_or_sections)
    return [_a(t) for t in torch.tensor_split(x, idx, dim=axis)]


def meshgrid(*arrs, indexing="xy"):
    return [_a(t) for t in torch.meshgrid(*arrs, indexing=indexing)]


def repeat(x, repeats, axis=None):
    return _a(torch.repeat_interleave(x, repeats, dim=axis))


def broadcast_to(x, shape):
    return _a(torch.broadcast_to(x, tuple(shape)))


def where(c, a, b):
    return _a(torch.where(c, a, b))


def minimum(a, b):
    if not isinstance(b, torch.Tensor):
        b = torch.as_tensor(b, dtype=a.dtype)
    if not isinstance(a, torch.Tensor):
        a = torch.as_tensor(a, dtype=b.dtype)
    return _a(torch.minimum(a, b))


def clip(x, lo, hi):
    return _a(torch.clamp(x, lo, hi))


def pad(x, pad_width, constant_values=0):
    flat = []
    for lo, hi in reversed(list(pad_width)):
        flat += [lo, hi]
    return _a(torch.nn.functional.pad(x, flat, value=constant_values))


def cos(x):
    return _a(torch.cos(x))


def sin(x):
    return _a(torch.sin(x))


def exp(x):
    return _a(torch.exp(x))


def log(x):
    return _a(torch.log(x.to(torch.float32) if not x.is_floating_point() else x))


def sigmoid(x):
    return _a(torch.sigmoid(x))


def softmax(x, axis=-1):
    return _a(torch.softmax(x, dim=axis))


def addmm(c, a, b):
    return _a(torch.addmm(c, a.reshape(-1, a.shape[-1]), b).reshape(*a.shape[:-1], b.shape[-1]))


def load(path, return_metadata=False):
    from safetensors.torch import load_file
    w = {k: _a(v) for k, v in load_file(path).items()}
    return (w, {}) if return_metadata else w


class _Random:
    def __init__(self):
        self._g = torch.Generator().manual_seed(0)

    def seed(self, s):
        self._g.manual_seed(int(s))

    def normal(self, shape=(), dtype=float32, loc=0.0, scale=1.0, key=None):
        # NOT MLX's threefry stream: parity tests always pass x_T explicitly.
        return _a((torch.randn(tuple(shape), generator=self._g) * scale + loc).to(dtype))

    def uniform(self, low=0.0, high=1.0, shape=(), dtype=float32, key=None):
        return _a((torch.rand(tuple(shape), generator=self._g) * (high - low) + low).to(dtype))

    def randint(self, low, high, shape=(), dtype=int32, key=None):
        return _a(torch.randint(low, high, tuple(shape), generator=self._g).to(dtype))

    def permutation(self, n, key=None):
        return _a(torch.randperm(n, generator=self._g).to(torch.int32))


random = _Random()


class _Fast:
    @staticmethod
    def scaled_dot_product_attention(q, k, v, *, scale, mask=None):
        """softmax((q*scale) k^T + mask) v with fp32 softmax; output in q.dtype."""
        s = torch.matmul((q * scale).to(torch.float32), torch.transpose(k.to(torch.float32), -1, -2))
        if mask is not None:
            s = s + mask.to(torch.float32)
        p = torch.softmax(s, dim=-1)
        return _a(torch.matmul(p, v.to(torch.float32)).to(q.dtype))

    @staticmethod
    def rms_norm(x, weight, eps):
        xf = x.to(torch.float32)
        y = xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)
        if weight is not None:
            y = y * weight.to(torch.float32)
        return _a(y.to(torch.result_type(x, weight) if weight is not None else x.dtype))

    @staticmethod
    def layer_norm(x, weight, bias, eps):
        xf = x.to(torch.float32)
        y = torch.nn.functional.layer_norm(xf, (x.shape[-1],), None, None, eps)
        if weight is not None:
            y = y * weight.to(torch.float32)
        if bias is not None:
            y = y + bias.to(torch.float32)
        dt = x.dtype if weight is None else torch.result_type(x, weight)
        return _a(y.to(dt))


fast = _Fast()


class _Metal:
    @staticmethod
    def get_peak_memory():
        return 0

    @staticmethod
    def reset_peak_memory():
        return None


metal = _Metal()
