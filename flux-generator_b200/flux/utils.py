"""Model registry and loaders (reference: flux/utils.py).

Same names and environment overrides as the reference (``FLUX_DEV``, ``FLUX_SCHNELL``, ``AE`` point
at local safetensors files, flux/utils.py:35,50,67,82).  There is no network here, so instead of
``hf_hub_download`` the text encoders / tokenizers are looked up under ``FLUX_HF_DIR`` (a local
snapshot with the hub layout: text_encoder/, text_encoder_2/, tokenizer/, tokenizer_2/).  When a
file is absent the loader raises, unless synthetic weights were requested (``FLUX_B200_SYNTHETIC=1``
or ``synthetic=True``): then tensors come from flux.synthetic under the checkpoint key names.
"""
from __future__ import annotations

import json
import os
from dataclasses import dataclass
from pathlib import Path
from typing import Optional, Union

import torch

from .autoencoder import AutoEncoder
from .clip import CLIPTextModel
from .model import Flux
from .specs import (AutoEncoderParams, CLIPTextModelConfig, FluxParams, T5Config, ae_decoder_manifest, ae_encoder_manifest,
                    clip_manifest, flow_manifest, t5_manifest)
from .synthetic import synthetic_state_dict
from .t5 import T5Encoder
from .tokenizers import CLIPTokenizer, SyntheticTokenizer, T5Tokenizer


@dataclass
class ModelSpec:
    params: FluxParams
    ae_params: AutoEncoderParams
    ckpt_path: Optional[str]
    ae_path: Optional[str]
    repo_id: Optional[str]
    repo_flow: Optional[str]
    repo_ae: Optional[str]


def _spec(repo: str, flow_file: str, env: str, guidance_embed: bool) -> ModelSpec:
    return ModelSpec(
        repo_id=repo, repo_flow=flow_file, repo_ae="ae.safetensors", ckpt_path=os.getenv(env),
        params=FluxParams(in_channels=64, vec_in_dim=768, context_in_dim=4096, hidden_size=3072, mlp_ratio=4.0,
                          num_heads=24, depth=19, depth_single_blocks=38, axes_dim=[16, 56, 56], theta=10_000,
                          qkv_bias=True, guidance_embed=guidance_embed),
        ae_path=os.getenv("AE"),
        ae_params=AutoEncoderParams(resolution=256, in_channels=3, ch=128, out_ch=3, ch_mult=[1, 2, 4, 4],
                                    num_res_blocks=2, z_channels=16, scale_factor=0.3611, shift_factor=0.1159))


# flux/utils.py:30-95
configs = {
    "flux-dev": _spec("black-forest-labs/FLUX.1-dev", "flux1-dev.safetensors", "FLUX_DEV", True),
    "flux-schnell": _spec("black-forest-labs/FLUX.1-schnell", "flux1-schnell.safetensors", "FLUX_SCHNELL", False),
}


def want_synthetic(flag: Optional[bool] = None) -> bool:
    return bool(flag) if flag is not None else os.getenv("FLUX_B200_SYNTHETIC", "0") not in ("0", "")


def _hf_file(name: str, rel: str) -> Optional[str]:
    root = os.getenv("FLUX_HF_DIR")
    if root and os.path.exists(os.path.join(root, rel)):
        return os.path.join(root, rel)
    return None


def _need(path: Optional[str], what: str) -> str:
    if path is None or not os.path.exists(path):
        raise FileNotFoundError(
            f"{what} not found. This build has no network access: point FLUX_SCHNELL / FLUX_DEV / AE at local "
            "safetensors files and FLUX_HF_DIR at a local snapshot of the hub repo, or set FLUX_B200_SYNTHETIC=1.")
    return path


def _load_safetensors(path: str):
    from safetensors.torch import load_file
    return load_file(path)


def _rank_world():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def _finish(model, device_sd, synthetic_manifest, synthetic: bool, gen_device: str):
    """Rank 0 materialises the tensors; with more than one rank the whole arena then travels once over
    NCCL / NVLink (north star: "NCCL over NVLink used only to broadcast weights at load").
    device_sd: a state dict, or a zero-argument callable returning one -- the checkpoint READER, which only rank 0
    ever calls (the other ranks receive the arena over NVLink and never touch the 24 GB file)."""
    rank, world = _rank_world()
    if rank == 0 or world == 1:
        if device_sd is not None:
            if callable(device_sd):
                device_sd = device_sd()
            model.load_weights(list(model.sanitize(device_sd).items()))
        elif synthetic:
            for entry in synthetic_manifest:  # one tensor at a time: no second copy of the model
                one = model.sanitize(synthetic_state_dict([entry], device=gen_device))
                model.load_weights(list(one.items()), strict=False)
            model.load_weights([], strict=True)
    model.arena.broadcast(0)
    return model


def load_flow_model(name: str, hf_download: bool = True, synthetic: Optional[bool] = None,
                    params: Optional[FluxParams] = None, device: Optional[str] = None, gen_device: str = "cuda"):
    """flux/utils.py:98-120."""
    params = params or configs[name].params
    model = Flux(params, device=device)
    path = configs[name].ckpt_path if name in configs else None
    if want_synthetic(synthetic):
        return _finish(model, None, flow_manifest(params), True, gen_device)
    if path is None and not hf_download:
        return model  # reference: model left at its random init
    path = _need(path, f"{name} flow checkpoint")
    return _finish(model, lambda: _load_safetensors(path), None, False, gen_device)


def load_ae(name: str, hf_download: bool = True, synthetic: Optional[bool] = None,
            params: Optional[AutoEncoderParams] = None, device: Optional[str] = None, gen_device: str = "cuda"):
    """flux/utils.py:123-145."""
    params = params or configs[name].ae_params
    ae = AutoEncoder(params, device=device)
    path = configs[name].ae_path if name in configs else None
    if want_synthetic(synthetic):
        return _finish(ae, None, ae_decoder_manifest(params) + ae_encoder_manifest(params), True, gen_device)
    if path is None and not hf_download:
        return ae
    path = _need(path, "autoencoder checkpoint (AE)")
    return _finish(ae, lambda: _load_safetensors(path), None, False, gen_device)


def load_clip(name: str, synthetic: Optional[bool] = None, config: Optional[CLIPTextModelConfig] = None,
              device: Optional[str] = None, gen_device: str = "cuda"):
    """flux/utils.py:148-163."""
    if want_synthetic(synthetic):
        config = config or CLIPTextModelConfig()
        return _finish(CLIPTextModel(config, device=device), None, clip_manifest(config), True, gen_device)
    with open(_need(_hf_file(name, "text_encoder/config.json"), "text_encoder/config.json")) as f:
        config = CLIPTextModelConfig.from_dict(json.load(f))
    path = _need(_hf_file(name, "text_encoder/model.safetensors"), "text_encoder/model.safetensors")
    return _finish(CLIPTextModel(config, device=device), lambda: _load_safetensors(path), None, False, gen_device)


def load_t5(name: str, synthetic: Optional[bool] = None, config: Optional[T5Config] = None,
            device: Optional[str] = None, gen_device: str = "cuda"):
    """flux/utils.py:166-191."""
    if want_synthetic(synthetic):
        config = config or T5Config()
        return _finish(T5Encoder(config, device=device), None, t5_manifest(config), True, gen_device)
    with open(_need(_hf_file(name, "text_encoder_2/config.json"), "text_encoder_2/config.json")) as f:
        config = T5Config.from_dict(json.load(f))
    index = _need(_hf_file(name, "text_encoder_2/model.safetensors.index.json"), "text_encoder_2 index")
    with open(index) as f:
        files = sorted(set(json.load(f)["weight_map"].values()))
    paths = [_need(_hf_file(name, f"text_encoder_2/{w}"), w) for w in files]

    def read():
        sd = {}
        for w in paths:
            sd.update(_load_safetensors(w))
        return sd

    return _finish(T5Encoder(config, device=device), read, None, False, gen_device)


def load_clip_tokenizer(name: str, synthetic: Optional[bool] = None, vocab_size: int = 49408):
    """flux/utils.py:194-205 (merges rows [1 : 49152-256-2+1])."""
    vocab_file, merges_file = _hf_file(name, "tokenizer/vocab.json"), _hf_file(name, "tokenizer/merges.txt")
    if vocab_file is None or merges_file is None:
        if want_synthetic(synthetic):
            return SyntheticTokenizer("clip", 77, vocab_size)
        _need(vocab_file, "tokenizer/vocab.json")
        _need(merges_file, "tokenizer/merges.txt")
    with open(vocab_file, encoding="utf-8") as f:
        vocab = json.load(f)
    with open(merges_file, encoding="utf-8") as f:
        merges = f.read().strip().split("\n")[1: 49152 - 256 - 2 + 1]
    merges = [tuple(m.split()) for m in merges]
    return CLIPTokenizer({pair: i for i, pair in enumerate(merges)}, vocab, max_length=77)


def load_t5_tokenizer(name: str, pad: bool = True, synthetic: Optional[bool] = None, vocab_size: int = 32100):
    """flux/utils.py:208-210: max length 256 for schnell, 512 for dev."""
    max_len = 256 if "schnell" in name else 512
    model_file = _hf_file(name, "tokenizer_2/spiece.model")
    if model_file is None:
        if want_synthetic(synthetic):
            return SyntheticTokenizer("t5", max_len, vocab_size)
        _need(model_file, "tokenizer_2/spiece.model")
    return T5Tokenizer(model_file, max_len)


def save_config(config: dict, config_path: Union[str, Path]) -> None:
    """flux/utils.py:213-230."""
    with open(config_path, "w") as fid:
        json.dump(dict(sorted(config.items())), fid, indent=4)
