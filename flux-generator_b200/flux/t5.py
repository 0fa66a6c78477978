"""T5-v1.1 encoder on B200 (reference: flux/t5.py:70-244).  Runs once per prompt; results cached by
the pipeline.  Linears are tcgen05 GEMMs (q|k|v stacked into one), attention is the small generic
kernel with T5's additive relative-position bias and scale 1.0 -- pad tokens are attended, exactly
like the reference (no padding mask, flux/t5.py:219-223).  The gated FFN uses the exact-erf GELU the
reference selects for "gated-gelu" (flux/t5.py:172-176)."""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch

from . import ops
from .model import WeightArena
from .specs import T5Config, t5_manifest

bf16 = torch.bfloat16


def relative_position_bucket(rpos: torch.Tensor, num_buckets: int, max_distance: int) -> torch.Tensor:
    """flux/t5.py:78-97 (bidirectional): integer path, bit-exact on the host."""
    num_buckets //= 2
    max_exact = num_buckets // 2
    abspos = rpos.abs()
    is_small = abspos < max_exact
    scale = (num_buckets - max_exact) / math.log(max_distance / max_exact)
    large = torch.log(abspos.clamp(min=1).to(torch.float32) / max_exact) * scale
    large = large.to(torch.int16).to(torch.int64)
    large = torch.minimum(max_exact + large, torch.tensor(num_buckets - 1))
    buckets = torch.where(is_small, abspos, large)
    return buckets + (rpos > 0).to(torch.int64) * num_buckets


class T5Encoder:
    def __init__(self, config: T5Config, device: Optional[str] = None):
        if config.d_kv != 64:
            raise ValueError("the B200 text-encoder attention kernel is specialised for d_kv 64")
        if not config.feed_forward_proj.startswith("gated"):
            raise ValueError("only gated feed-forward T5 variants are supported")
        act = config.feed_forward_proj.removeprefix("gated-")
        if act not in ("gelu",):
            raise ValueError(f"Unknown activation: {act}")
        self.config = config
        self.device = torch.device(device or "cuda")
        self._manifest = t5_manifest(config)
        self._shapes = {k: s for k, s, _ in self._manifest}
        inner = config.d_kv * config.num_heads
        entries = [(k, s) for k, s, _ in self._manifest]
        entries += [(f"encoder.block.{i}.layer.0.SelfAttention.qkv.weight", (3 * inner, config.d_model))
                    for i in range(config.num_layers)]
        self.arena = WeightArena(entries, self.device)
        self._bias_cache: Dict[int, torch.Tensor] = {}

    def sanitize(self, weights):
        """The reference renames HF keys to its module tree (flux/t5.py:232-241); this class keeps HF names."""
        return dict(weights)

    def load_weights(self, weights, strict: bool = True) -> "T5Encoder":
        items = list(weights.items()) if isinstance(weights, dict) else list(weights)
        seen = self.arena.loaded
        for key, w in items:
            if key not in self._shapes:
                if strict and key.startswith(("encoder.", "shared.")):
                    raise ValueError(f"Received parameters not in model: {key}")
                continue
            self.arena[key].copy_(w.to(device=self.device, dtype=bf16))
            seen.add(key)
        if strict:
            missing = [k for k in self._shapes if k not in seen]
            if missing:
                raise ValueError(f"Missing {len(missing)} parameters, e.g. {missing[:3]}")
        if len(seen) == len(self._shapes):  # derived tensors: stacked q|k|v projections
            for i in range(self.config.num_layers):
                pre = f"encoder.block.{i}.layer.0.SelfAttention."
                self.arena[pre + "qkv.weight"].copy_(torch.cat([self.arena[pre + n + ".weight"] for n in "qkv"], 0))
        self._bias_cache = {}
        return self

    def parameters(self):
        return {"arena": self.arena.buffer}

    def position_bias(self, S: int) -> torch.Tensor:
        """RelativePositionBias (flux/t5.py:99-120) -> fp32 [heads, S, S]; shared by all layers."""
        if S not in self._bias_cache:
            c = self.config
            ctx = torch.arange(S)[:, None]
            mem = torch.arange(S)[None, :]
            bucket = relative_position_bucket(mem - ctx, c.relative_attention_num_buckets,
                                              c.relative_attention_max_distance).to(self.device)
            emb = self.arena["encoder.block.0.layer.0.SelfAttention.relative_attention_bias.weight"]
            self._bias_cache[S] = emb[bucket].permute(2, 0, 1).to(torch.float32).contiguous()
        return self._bias_cache[S]

    def __call__(self, tokens: torch.Tensor) -> torch.Tensor:
        """tokens int [B, S] -> [B, S, d_model] bf16."""
        c = self.config
        tokens = tokens.to(device=self.device, dtype=torch.int32)
        B, S = tokens.shape
        H, inner = c.num_heads, c.d_kv * c.num_heads
        A = self.arena
        x = ops.embedding(tokens, A["shared.weight"])
        bias = self.position_bias(S)
        for i in range(c.num_layers):
            pre = f"encoder.block.{i}.layer."
            y = ops.rownorm(x, 2, A[pre + "0.layer_norm.weight"], None, c.layer_norm_epsilon)
            qkv = ops.gemm(y, A[pre + "0.SelfAttention.qkv.weight"])
            a = ops.attention_small(qkv[..., :inner], qkv[..., inner:2 * inner], qkv[..., 2 * inner:], H, 1.0, bias=bias)
            ops.gemm(a, A[pre + "0.SelfAttention.o.weight"], resid=x, out=x)
            y = ops.rownorm(x, 2, A[pre + "1.layer_norm.weight"], None, c.layer_norm_epsilon)
            g = ops.gemm(y, A[pre + "1.DenseReluDense.wi_0.weight"])
            u = ops.gemm(y, A[pre + "1.DenseReluDense.wi_1.weight"])
            h = ops.act_mul(g, u, "gelu")
            ops.gemm(h, A[pre + "1.DenseReluDense.wo.weight"], resid=x, out=x)
        return ops.rownorm(x, 2, A["encoder.final_layer_norm.weight"], None, c.layer_norm_epsilon)
