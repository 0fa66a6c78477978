"""CLIP text encoder on B200 (reference: flux/clip.py:46-154).  Runs once per prompt.
Pre-LN transformer with causal attention (additive -1e9 mask == hard causal mask after softmax),
quick-GELU MLP, final LayerNorm, pooled output = hidden state at the first EOS (argmax of the
token ids, flux/clip.py:130,148)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional

import torch

from . import ops
from .model import WeightArena
from .specs import CLIPTextModelConfig, clip_manifest

bf16 = torch.bfloat16


@dataclass
class CLIPOutput:
    # flux/clip.py:33-43
    pooled_output: Optional[torch.Tensor] = None
    last_hidden_state: Optional[torch.Tensor] = None
    hidden_states: Optional[List[torch.Tensor]] = None


class CLIPTextModel:
    def __init__(self, config: CLIPTextModelConfig, device: Optional[str] = None):
        if config.model_dims // config.num_heads != 64:
            raise ValueError("the B200 text-encoder attention kernel is specialised for head_dim 64")
        if config.hidden_act not in ("quick_gelu", "gelu"):
            raise ValueError(f"unknown activation {config.hidden_act}")
        self.config = config
        self.device = torch.device(device or "cuda")
        self._manifest = clip_manifest(config)
        self._shapes = {k: s for k, s, _ in self._manifest}
        D = config.model_dims
        entries = [(k, s) for k, s, _ in self._manifest]
        for i in range(config.num_layers):
            pre = f"text_model.encoder.layers.{i}.self_attn.qkv"
            entries += [(pre + ".weight", (3 * D, D)), (pre + ".bias", (3 * D,))]
        self.arena = WeightArena(entries, self.device)

    def sanitize(self, weights):
        """The reference renames HF keys (flux/clip.py:96-125); this class keeps HF names."""
        return dict(weights)

    def load_weights(self, weights, strict: bool = True) -> "CLIPTextModel":
        items = list(weights.items()) if isinstance(weights, dict) else list(weights)
        seen = self.arena.loaded
        for key, w in items:
            if key not in self._shapes:
                if strict and key.startswith("text_model."):
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
                pre = f"text_model.encoder.layers.{i}.self_attn."
                for part in ("weight", "bias"):
                    self.arena[pre + "qkv." + part].copy_(
                        torch.cat([self.arena[pre + n + "_proj." + part] for n in "qkv"], 0))
        return self

    def parameters(self):
        return {"arena": self.arena.buffer}

    def __call__(self, x: torch.Tensor) -> CLIPOutput:
        c = self.config
        tokens = x.to(torch.int32)
        B, N = tokens.shape
        eos = tokens.to("cpu").argmax(-1)
        tokens = tokens.to(self.device)
        A = self.arena
        D, H = c.model_dims, c.num_heads
        pre = "text_model."
        h = ops.embedding(tokens, A[pre + "embeddings.token_embedding.weight"], A[pre + "embeddings.position_embedding.weight"])
        hidden = []
        for i in range(c.num_layers):
            lp = f"{pre}encoder.layers.{i}."
            y = ops.rownorm(h, 1, A[lp + "layer_norm1.weight"], A[lp + "layer_norm1.bias"], 1e-5)
            qkv = ops.gemm(y, A[lp + "self_attn.qkv.weight"], A[lp + "self_attn.qkv.bias"])
            a = ops.attention_small(qkv[..., :D], qkv[..., D:2 * D], qkv[..., 2 * D:], H, 64 ** -0.5, causal=True)
            h = ops.gemm(a, A[lp + "self_attn.out_proj.weight"], A[lp + "self_attn.out_proj.bias"], resid=h)
            y = ops.rownorm(h, 1, A[lp + "layer_norm2.weight"], A[lp + "layer_norm2.bias"], 1e-5)
            y = ops.gemm(y, A[lp + "mlp.fc1.weight"], A[lp + "mlp.fc1.bias"], act=c.hidden_act)
            h = ops.gemm(y, A[lp + "mlp.fc2.weight"], A[lp + "mlp.fc2.bias"], resid=h)
            hidden.append(h)
        last = ops.rownorm(h, 1, A[pre + "final_layer_norm.weight"], A[pre + "final_layer_norm.bias"], 1e-5)
        pooled = last[torch.arange(B, device=self.device), eos.to(self.device)]
        return CLIPOutput(pooled_output=pooled, last_hidden_state=last, hidden_states=hidden)
