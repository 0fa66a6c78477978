"""Flow-matching sampler: schedule, prior, Euler step (reference: flux/sampler.py:9-57).

Host-side only.  The schedule is bit-exact float32 arithmetic (numpy), the Euler update itself is
fused into the final-layer kernel on the device (csrc/elementwise.cu: fx_euler_step is the
stand-alone form used by this class).
"""
from __future__ import annotations

import math
from functools import lru_cache
from typing import List

import numpy as np


class FluxSampler:
    def __init__(self, name: str, base_shift: float = 0.5, max_shift: float = 1.15):
        self._base_shift = base_shift
        self._max_shift = max_shift
        self._schnell = "schnell" in name

    def _time_shift(self, x, t):
        # flux/sampler.py:15-20 -- mu in python double, tensor math in float32
        x1, x2 = 256, 4096
        t1, t2 = self._base_shift, self._max_shift
        exp_mu = np.float32(math.exp((x - x1) * (t2 - t1) / (x2 - x1) + t1))
        with np.errstate(divide="ignore"):
            t = exp_mu / (exp_mu + (np.float32(1.0) / t - np.float32(1.0)))
        return t.astype(np.float32)

    @lru_cache
    def timesteps(self, num_steps, image_sequence_length, start: float = 1, stop: float = 0) -> List[float]:
        # flux/sampler.py:22-31; mx.linspace = (1 - i/n) * start + (i/n) * stop in float32
        i = np.arange(num_steps + 1, dtype=np.float32) / np.float32(num_steps)
        t = ((np.float32(1.0) - i) * np.float32(start) + i * np.float32(stop)).astype(np.float32)
        if not self._schnell:
            t = self._time_shift(image_sequence_length, t)
        return [float(v) for v in t]

    def random_timesteps(self, B, L, dtype=None, key=None):
        """flux/sampler.py:33-42: schnell draws t from {1/4, 2/4, 3/4, 1}, dev a uniform t pushed through the time shift.
        Returns a torch tensor [B] (float32, or `dtype`); `key` seeds a torch generator."""
        import torch
        g = torch.Generator().manual_seed(0 if key is None else int(key)) if key is not None else None
        if self._schnell:
            t = torch.randint(1, 5, (B,), generator=g).to(torch.float32) / 4
        else:
            u = torch.rand((B,), generator=g, dtype=torch.float32).numpy()
            t = torch.from_numpy(self._time_shift(L, u))
        return t if dtype is None else t.to(dtype)

    def add_noise(self, x, t, noise=None, key=None):
        """flux/sampler.py:47-54: x * (1 - t) + t * noise, t broadcast over the trailing dimensions."""
        import torch
        if noise is None:
            g = torch.Generator(device=x.device).manual_seed(0 if key is None else int(key))
            noise = torch.randn(x.shape, generator=g, device=x.device, dtype=torch.float32).to(x.dtype)
        t = t.to(device=x.device, dtype=x.dtype).reshape([-1] + [1] * (x.ndim - 1))
        return x * (1 - t) + t * noise

    def sample_prior(self, shape, dtype=None, key=None, first_index: int = 0):
        """flux/sampler.py:44-45.  MLX's threefry stream cannot be reproduced offline;
        noise is keyed by (seed, global image index) instead -- see flux.synthetic.synthetic_prior."""
        from .synthetic import synthetic_prior
        seed = 0 if key is None else int(key)
        x = synthetic_prior(shape[0], tuple(shape[1:3]), seed=seed, first_index=first_index)
        return x if dtype is None else x.to(dtype)

    def step(self, pred, x_t, t, t_prev):
        # flux/sampler.py:56-57 (torch tensors; the device path fuses this into the last kernel)
        return (x_t + (t_prev - t) * pred).to(x_t.dtype)
