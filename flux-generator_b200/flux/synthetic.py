"""Deterministic synthetic checkpoints and prompts.

No Flux / T5 / CLIP / VAE weights or tokenizer files exist offline, so every measurement and parity
test runs on tensors drawn here.  Each tensor is keyed by its checkpoint name: the generator for
key k is seeded with ``seed + crc32(k)`` so a tensor's values do not depend on which other tensors
are generated, nor on model depth.  All values are bf16-representable (the CPU oracle reads the
same numbers in fp32).
"""
from __future__ import annotations

import zlib
from typing import Dict, Iterable, Optional, Tuple

import torch

from .specs import Manifest

BASE_SEED = 0xF1A5


def _fan_in(shape: Tuple[int, ...]) -> int:
    n = 1
    for s in shape[1:]:
        n *= s
    return max(n, 1)


def synthetic_tensor(key: str, shape: Tuple[int, ...], kind: str, seed: int = BASE_SEED,
                     device: str = "cpu", dtype: torch.dtype = torch.bfloat16) -> torch.Tensor:
    g = torch.Generator(device=device)
    g.manual_seed((seed + zlib.crc32(key.encode())) & 0x7FFFFFFF)
    x = torch.randn(shape, generator=g, device=device, dtype=torch.float32)
    if kind == "w":
        x *= _fan_in(shape) ** -0.5
    elif kind == "wmod":
        x *= 0.1 * _fan_in(shape) ** -0.5
    elif kind == "wsmall":
        x *= 0.125 * _fan_in(shape) ** -0.5
    elif kind == "b":
        x *= 0.02
    elif kind == "scale":
        x = 1.0 + 0.02 * x
    elif kind == "nb":
        x *= 0.02
    elif kind == "emb":
        pass
    else:
        raise ValueError(f"unknown tensor kind {kind!r}")
    return x.to(dtype)


def synthetic_state_dict(manifest: Manifest, seed: int = BASE_SEED, device: str = "cpu",
                         dtype: torch.dtype = torch.bfloat16) -> Dict[str, torch.Tensor]:
    out = {}
    for key, shape, kind in manifest:
        if kind == "w" and key.endswith("SelfAttention.q.weight"):
            kind = "wsmall"  # T5 uses attention scale 1.0 (flux/t5.py:153-155): keep logits O(1)
        out[key] = synthetic_tensor(key, shape, kind, seed, device, dtype)
    return out


def synthetic_prompt_tokens(t5_len: int, clip_len: int = 77, seed: int = 1234, n_tok: int = 32,
                            t5_vocab: int = 32100, clip_vocab: int = 49408,
                            pad: bool = True) -> Tuple[torch.Tensor, torch.Tensor]:
    """Token ids shaped like the tokenizers' output (flux/tokenizers.py:103-119,160-185):
    T5  = n_tok random ids + EOS(1), padded with 0 to t5_len;
    CLIP = BOS(clip_vocab-2) + ids + EOS(clip_vocab-1), never padded for a single prompt."""
    g = torch.Generator().manual_seed(seed)
    n_tok = min(n_tok, t5_len - 1)
    t5 = torch.randint(3, t5_vocab, (n_tok,), generator=g, dtype=torch.int64)
    t5 = torch.cat([t5, torch.tensor([1])])
    if pad and len(t5) < t5_len:
        t5 = torch.cat([t5, torch.zeros(t5_len - len(t5), dtype=torch.int64)])
    n_clip = min(n_tok, clip_len - 2)
    clip = torch.randint(0, clip_vocab - 3, (n_clip,), generator=g, dtype=torch.int64)
    clip = torch.cat([torch.tensor([clip_vocab - 2]), clip, torch.tensor([clip_vocab - 1])])
    return t5[None].to(torch.int32), clip[None].to(torch.int32)


def synthetic_prior(n_images: int, latent_size: Tuple[int, int], seed: int = 42,
                    first_index: int = 0) -> torch.Tensor:
    """x_T [B, h, w, 16] bf16 keyed by *global* image index, so a batch sharded over G ranks
    draws the same noise per image for every G."""
    h, w = latent_size
    out = []
    for i in range(first_index, first_index + n_images):
        g = torch.Generator().manual_seed((seed * 1000003 + i) & 0x7FFFFFFF)
        out.append(torch.randn((h, w, 16), generator=g))
    return torch.stack(out).to(torch.bfloat16)


def state_dict_checksum(sd: Dict[str, torch.Tensor]) -> int:
    """Order-independent CRC over raw bf16/f32 bytes; stored in golden fixtures to detect
    generator drift between the box that wrote a fixture and the box that replays it."""
    acc = 0
    for k in sorted(sd):
        t = sd[k].detach().cpu().contiguous()
        raw = t.view(torch.uint8).numpy().tobytes() if t.dtype != torch.bfloat16 else \
            t.view(torch.int16).numpy().tobytes()
        acc = zlib.crc32(raw, zlib.crc32(k.encode(), acc))
    return acc
