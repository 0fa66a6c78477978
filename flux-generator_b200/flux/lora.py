"""LoRA adapters at inference (reference: txt2image.py:32-39, flux/flux.py:228-246, flux/lora.py).

The reference wraps every nn.Linear of the last `lora_blocks` transformer blocks in a LoRALinear, loads the adapter
file (safetensors with `lora_rank` / `lora_blocks` metadata, dreambooth.py:46-59) and either keeps the low-rank
branch `y + scale * (x @ lora_a) @ lora_b` (flux/lora.py:73-76) or fuses it, `W + (scale * lora_b.T) @ lora_a.T`
(flux/lora.py:28-43).  On this build an adapter is ALWAYS fused into the weight arena -- the two forms are the same
function of x up to the rounding of W + dW to bf16, and the fused form keeps the hot path a plain tcgen05 GEMM.
Training (`LoRALinear` as a trainable module, dreambooth.py) is out of scope.
"""
from __future__ import annotations

import json
import struct
from typing import Dict, List, Tuple

import torch


def lora_blocks(depth: int, depth_single: int, num_blocks: int) -> List[str]:
    """Block prefixes wrapped by linear_to_lora_layers (flux/flux.py:230-233): double + single blocks, REVERSED,
    the first `num_blocks` of them (all when num_blocks <= 0)."""
    allb = [f"double_blocks.{i}" for i in range(depth)] + [f"single_blocks.{i}" for i in range(depth_single)]
    allb.reverse()
    return allb[:num_blocks if num_blocks > 0 else len(allb)]


def read_adapter(path: str) -> Tuple[Dict[str, torch.Tensor], int, int]:
    """(tensors, lora_rank, lora_blocks) of an adapter file as `mx.load(file, return_metadata=True)` returns them
    (txt2image.py:33-35).  Raises ValueError when the metadata the reference requires is missing."""
    from safetensors import safe_open
    tensors = {}
    with safe_open(path, framework="pt") as f:
        meta = f.metadata() or {}
        for k in f.keys():
            tensors[k] = f.get_tensor(k)
    if "lora_rank" not in meta or "lora_blocks" not in meta:
        raise ValueError(f"{path}: adapter metadata must carry lora_rank and lora_blocks (dreambooth.py:53-58)")
    return tensors, int(meta["lora_rank"]), int(meta["lora_blocks"])


def adapter_deltas(adapter: Dict[str, torch.Tensor], rank: int, num_blocks: int, depth: int, depth_single: int,
                   scale: float = 1.0) -> Dict[str, torch.Tensor]:
    """checkpoint-side weight key -> fp32 delta [out, in] = (scale * lora_b.T) @ lora_a.T for every adapted Linear.
    Keys are MLX module paths (`...img_mlp.layers.0.lora_a`); `.layers.` is Flux.sanitize's nn.Sequential renaming
    (flux/model.py:92-95) and is dropped.  Entries outside the wrapped blocks are ignored, as the reference's
    load_weights(strict=False) ignores them; a wrapped entry of the wrong rank is an error (strict shape check)."""
    prefixes = tuple(p + "." for p in lora_blocks(depth, depth_single, num_blocks))
    out = {}
    for k, a in adapter.items():
        if not k.endswith(".lora_a") or not k.startswith(prefixes):
            continue
        mod = k[:-len(".lora_a")]
        if mod + ".lora_b" not in adapter:
            raise ValueError(f"adapter has {k} but no {mod}.lora_b")
        b = adapter[mod + ".lora_b"]
        if a.shape[1] != rank or b.shape[0] != rank:
            raise ValueError(f"Expected rank {rank} but received shapes {tuple(a.shape)}, {tuple(b.shape)} for {mod}")
        out[mod.replace(".layers.", ".") + ".weight"] = (scale * b.to(torch.float32).T) @ a.to(torch.float32).T
    return out


def load_adapter(flux, adapter_file: str, fuse: bool = False) -> None:
    """txt2image.py:32-39.  `fuse` is accepted for signature parity; the adapter is fused either way (module docstring)."""
    tensors, rank, blocks = read_adapter(adapter_file)
    flux.linear_to_lora_layers(rank, blocks)
    flux.flow.load_weights(list(tensors.items()), strict=False)
    flux.fuse_lora_layers()
