"""mlx.nn subset (see mlx/__init__.py)."""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from .. import core as mx
from . import layers  # noqa: F401


def silu(x):
    return mx.array(F.silu(x))


def relu(x):
    return mx.array(F.relu(x))


def gelu(x):
    return mx.array(F.gelu(x))


def gelu_approx(x):
    return mx.array(F.gelu(x, approximate="tanh"))


def gelu_fast_approx(x):
    return mx.array(x * torch.sigmoid(1.702 * x))


class Module:
    """Attribute-tree module with MLX's dotted-path weight loading."""
    training = False

    def __call__(self, *a, **k):
        raise NotImplementedError

    def _children(self):
        return {k: v for k, v in vars(self).items() if not k.startswith("_")}

    def parameters(self):
        def rec(v):
            if isinstance(v, Module):
                return {k: rec(c) for k, c in v._children().items() if rec(c) is not None}
            if isinstance(v, dict):
                return {k: rec(c) for k, c in v.items()}
            if isinstance(v, (list, tuple)):
                return [rec(c) for c in v]
            if isinstance(v, torch.Tensor):
                return v
            return None
        return rec(self)

    def load_weights(self, items, strict=True):
        if isinstance(items, dict):
            items = list(items.items())
        for key, val in items:
            parts = key.split(".")
            cur = self
            for p in parts[:-1]:
                if isinstance(cur, (list, tuple)):
                    cur = cur[int(p)]
                elif isinstance(cur, dict):
                    cur = cur[p]
                else:
                    cur = getattr(cur, p)
            leaf = parts[-1]
            old = cur[leaf] if isinstance(cur, dict) else getattr(cur, leaf, None)
            if old is None:
                if strict:
                    raise ValueError(f"unexpected weight {key}")
                continue
            val = mx.array(val)
            if strict and tuple(old.shape) != tuple(val.shape):
                raise ValueError(f"shape mismatch for {key}: {tuple(old.shape)} vs {tuple(val.shape)}")
            if isinstance(cur, dict):
                cur[leaf] = val
            else:
                setattr(cur, leaf, val)
        return self

    def named_modules(self):
        out = []

        def rec(prefix, v):
            if isinstance(v, Module):
                out.append((prefix, v))
                for k, c in v._children().items():
                    rec(f"{prefix}.{k}" if prefix else k, c)
            elif isinstance(v, dict):
                for k, c in v.items():
                    rec(f"{prefix}.{k}" if prefix else k, c)
            elif isinstance(v, (list, tuple)):
                for i, c in enumerate(v):
                    rec(f"{prefix}.{i}" if prefix else str(i), c)
        rec("", self)
        return out

    def eval(self):
        return self

    def __contains__(self, key):  # `"bias" in linear` (flux/lora.py:30)
        return key in vars(self)

    def update_modules(self, tree):
        """Replace sub-modules along dotted paths (mlx.nn.Module.update_modules; flux/flux.py:236,246)."""
        def rec(cur, sub):
            for k, v in sub.items():
                if isinstance(v, dict):
                    nxt = cur[int(k)] if isinstance(cur, (list, tuple)) else (cur[k] if isinstance(cur, dict) else getattr(cur, k))
                    rec(nxt, v)
                elif isinstance(cur, list):
                    cur[int(k)] = v
                elif isinstance(cur, dict):
                    cur[k] = v
                else:
                    setattr(cur, k, v)
        rec(self, tree)
        return self


class Identity(Module):
    def __init__(self, *a, **k):
        pass

    def __call__(self, x):
        return x


class Dropout(Identity):
    pass


class Linear(Module):
    def __init__(self, input_dims, output_dims, bias=True):
        s = math.sqrt(1.0 / input_dims)
        self.weight = mx.array(torch.empty(output_dims, input_dims).uniform_(-s, s))
        if bias:
            self.bias = mx.array(torch.empty(output_dims).uniform_(-s, s))

    def __call__(self, x):
        dt = torch.result_type(x, self.weight)  # MLX type promotion (bf16 (+) f32 -> f32)
        y = torch.matmul(x.to(dt), self.weight.to(dt).transpose())
        if hasattr(self, "bias"):
            y = y + self.bias.to(dt)
        return mx.array(y)


class Embedding(Module):
    def __init__(self, num_embeddings, dims):
        self.weight = mx.array(torch.randn(num_embeddings, dims) * math.sqrt(1.0 / dims))

    def __call__(self, x):
        return mx.array(self.weight[x.long()])


class LayerNorm(Module):
    def __init__(self, dims, eps=1e-5, affine=True, bias=True):
        self.dims, self.eps = dims, eps
        if affine:
            self.weight = mx.array(torch.ones(dims))
            if bias:
                self.bias = mx.array(torch.zeros(dims))

    def __call__(self, x):
        return mx.fast.layer_norm(x, getattr(self, "weight", None), getattr(self, "bias", None), self.eps)


class RMSNorm(Module):
    def __init__(self, dims, eps=1e-5):
        self.weight = mx.array(torch.ones(dims))
        self.eps = eps

    def __call__(self, x):
        return mx.fast.rms_norm(x, self.weight, self.eps)


class GroupNorm(Module):
    def __init__(self, num_groups, dims, eps=1e-5, affine=True, pytorch_compatible=False):
        assert pytorch_compatible, "shim implements the pytorch_compatible grouping only"
        self.num_groups, self.dims, self.eps = num_groups, dims, eps
        if affine:
            self.weight = mx.array(torch.ones(dims))
            self.bias = mx.array(torch.zeros(dims))

    def __call__(self, x):
        # channels-last input of any rank >= 3; statistics over all non-batch positions x C/G
        b = x.shape[0]
        c = x.shape[-1]
        dt = x.dtype
        xf = x.to(torch.float32).reshape(b, -1, self.num_groups, c // self.num_groups)
        mean = xf.mean(dim=(1, 3), keepdim=True)
        var = xf.var(dim=(1, 3), keepdim=True, unbiased=False)
        y = ((xf - mean) * torch.rsqrt(var + self.eps)).reshape(x.shape)
        if hasattr(self, "weight"):
            y = y * self.weight.to(torch.float32) + self.bias.to(torch.float32)
            dt = torch.result_type(x, self.weight)
        return mx.array(y.to(dt))


class Conv2d(Module):
    """NHWC input, weight [O, kH, kW, I]."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, bias=True):
        k = (kernel_size, kernel_size) if isinstance(kernel_size, int) else tuple(kernel_size)
        s = math.sqrt(1.0 / (in_channels * k[0] * k[1]))
        self.weight = mx.array(torch.empty(out_channels, k[0], k[1], in_channels).uniform_(-s, s))
        if bias:
            self.bias = mx.array(torch.zeros(out_channels))
        self.stride, self.padding = stride, padding

    def __call__(self, x):
        w = self.weight.permute(0, 3, 1, 2)
        b = getattr(self, "bias", None)
        dt = torch.result_type(x, self.weight)
        y = F.conv2d(x.to(dt).permute(0, 3, 1, 2), w.to(dt), None if b is None else b.to(dt),
                     stride=self.stride, padding=self.padding)
        return mx.array(y.permute(0, 2, 3, 1))


class MultiHeadAttention(Module):
    def __init__(self, dims, num_heads, query_input_dims=None, key_input_dims=None,
                 value_input_dims=None, value_dims=None, value_output_dims=None, bias=False):
        self.num_heads = num_heads
        self.query_proj = Linear(dims, dims, bias=bias)
        self.key_proj = Linear(dims, dims, bias=bias)
        self.value_proj = Linear(dims, dims, bias=bias)
        self.out_proj = Linear(dims, dims, bias=bias)

    def __call__(self, queries, keys, values, mask=None):
        q, k, v = self.query_proj(queries), self.key_proj(keys), self.value_proj(values)
        H = self.num_heads
        B, L, _ = q.shape
        S = k.shape[1]
        q = q.reshape(B, L, H, -1).transpose(0, 2, 1, 3)
        k = k.reshape(B, S, H, -1).transpose(0, 2, 1, 3)
        v = v.reshape(B, S, H, -1).transpose(0, 2, 1, 3)
        scale = math.sqrt(1 / q.shape[-1])
        o = mx.fast.scaled_dot_product_attention(q, k, v, scale=scale, mask=mask)
        return self.out_proj(o.transpose(0, 2, 1, 3).flatten(-2, -1))


class GELU(Module):
    def __init__(self, approx="none"):
        self._approx = approx

    def __call__(self, x):
        if self._approx in ("tanh", "precise"):
            return gelu_approx(x)
        if self._approx == "fast":
            return gelu_fast_approx(x)
        return gelu(x)


class SiLU(Module):
    def __init__(self):
        pass

    def __call__(self, x):
        return silu(x)


class Sequential(Module):
    def __init__(self, *modules):
        self.layers = list(modules)

    def __call__(self, x):
        for m in self.layers:
            x = m(x)
        return x


def quantize(*a, **k):
    raise NotImplementedError("nn.quantize is not part of the shim")
