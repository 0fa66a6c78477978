"""Model hyper-parameter records and the checkpoint tensor manifest.

Mirrors the reference's dataclasses: FluxParams (flux/model.py:20-32), AutoEncoderParams
(flux/autoencoder.py:11-21), T5Config (flux/t5.py:34-67), CLIPTextModelConfig (flux/clip.py:12-30).
The manifests enumerate every tensor under the *checkpoint-side* key names the reference's
sanitizers accept (flux/model.py:85-97, flux/autoencoder.py:336-345, flux/t5.py:10-31,232-241,
flux/clip.py:96-125); Linear weights are [out, in], conv weights OIHW as stored in the files.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Tuple


@dataclass
class FluxParams:
    in_channels: int = 64
    vec_in_dim: int = 768
    context_in_dim: int = 4096
    hidden_size: int = 3072
    mlp_ratio: float = 4.0
    num_heads: int = 24
    depth: int = 19
    depth_single_blocks: int = 38
    axes_dim: List[int] = field(default_factory=lambda: [16, 56, 56])
    theta: int = 10_000
    qkv_bias: bool = True
    guidance_embed: bool = False

    def validate(self) -> None:
        # same checks and messages as Flux.__init__ (flux/model.py:42-50)
        if self.hidden_size % self.num_heads != 0:
            raise ValueError(
                f"Hidden size {self.hidden_size} must be divisible by num_heads {self.num_heads}")
        pe_dim = self.hidden_size // self.num_heads
        if sum(self.axes_dim) != pe_dim:
            raise ValueError(f"Got {self.axes_dim} but expected positional dim {pe_dim}")

    @property
    def mlp_hidden(self) -> int:
        return int(self.hidden_size * self.mlp_ratio)


@dataclass
class AutoEncoderParams:
    resolution: int = 256
    in_channels: int = 3
    ch: int = 128
    out_ch: int = 3
    ch_mult: List[int] = field(default_factory=lambda: [1, 2, 4, 4])
    num_res_blocks: int = 2
    z_channels: int = 16
    scale_factor: float = 0.3611
    shift_factor: float = 0.1159


@dataclass
class T5Config:
    vocab_size: int = 32128
    num_layers: int = 24
    num_heads: int = 64
    relative_attention_num_buckets: int = 32
    d_kv: int = 64
    d_model: int = 4096
    d_ff: int = 10240
    feed_forward_proj: str = "gated-gelu"
    tie_word_embeddings: bool = False
    relative_attention_max_distance: int = 128
    layer_norm_epsilon: float = 1e-6

    @classmethod
    def from_dict(cls, config: dict) -> "T5Config":
        # flux/t5.py:50-67
        return cls(
            vocab_size=config["vocab_size"], num_layers=config["num_layers"],
            num_heads=config["num_heads"],
            relative_attention_num_buckets=config["relative_attention_num_buckets"],
            d_kv=config["d_kv"], d_model=config["d_model"],
            feed_forward_proj=config["feed_forward_proj"],
            tie_word_embeddings=config["tie_word_embeddings"],
            d_ff=config.get("d_ff", 4 * config["d_model"]),
            relative_attention_max_distance=config.get("relative_attention_max_distance", 128),
            layer_norm_epsilon=config.get("layer_norm_epsilon", 1e-6))


@dataclass
class CLIPTextModelConfig:
    num_layers: int = 12
    model_dims: int = 768
    num_heads: int = 12
    max_length: int = 77
    vocab_size: int = 49408
    hidden_act: str = "quick_gelu"

    @classmethod
    def from_dict(cls, config: dict) -> "CLIPTextModelConfig":
        # flux/clip.py:21-30
        return cls(num_layers=config["num_hidden_layers"], model_dims=config["hidden_size"],
                   num_heads=config["num_attention_heads"],
                   max_length=config["max_position_embeddings"], vocab_size=config["vocab_size"],
                   hidden_act=config["hidden_act"])


Manifest = List[Tuple[str, Tuple[int, ...], str]]  # (key, shape, kind)
# kind in: "w" (linear/conv weight, N(0, 1/fan_in)), "wmod" (modulation weight, x0.1), "b" (bias),
#          "scale" (norm scale ~1), "nb" (norm bias ~0), "emb" (embedding table)


def _lin(m: Manifest, key: str, out: int, inp: int, bias: bool = True, kind: str = "w") -> None:
    m.append((key + ".weight", (out, inp), kind))
    if bias:
        m.append((key + ".bias", (out,), "b"))


def flow_manifest(p: FluxParams) -> Manifest:
    """BFL checkpoint keys (flux/model.py:56-83, flux/layers.py)."""
    D, M = p.hidden_size, p.mlp_hidden
    hd = D // p.num_heads
    m: Manifest = []
    _lin(m, "img_in", D, p.in_channels)
    for name, inp in (("time_in", 256), ("vector_in", p.vec_in_dim)) + (
            (("guidance_in", 256),) if p.guidance_embed else ()):
        _lin(m, f"{name}.in_layer", D, inp)
        _lin(m, f"{name}.out_layer", D, D)
    _lin(m, "txt_in", D, p.context_in_dim)
    for i in range(p.depth):
        for s in ("img", "txt"):
            pre = f"double_blocks.{i}.{s}"
            _lin(m, f"{pre}_mod.lin", 6 * D, D, kind="wmod")
            _lin(m, f"{pre}_attn.qkv", 3 * D, D, bias=p.qkv_bias)
            m.append((f"{pre}_attn.norm.query_norm.scale", (hd,), "scale"))
            m.append((f"{pre}_attn.norm.key_norm.scale", (hd,), "scale"))
            _lin(m, f"{pre}_attn.proj", D, D)
            _lin(m, f"{pre}_mlp.0", M, D)
            _lin(m, f"{pre}_mlp.2", D, M)
    for i in range(p.depth_single_blocks):
        pre = f"single_blocks.{i}"
        _lin(m, f"{pre}.linear1", 3 * D + M, D)
        _lin(m, f"{pre}.linear2", D, D + M)
        m.append((f"{pre}.norm.query_norm.scale", (hd,), "scale"))
        m.append((f"{pre}.norm.key_norm.scale", (hd,), "scale"))
        _lin(m, f"{pre}.modulation.lin", 3 * D, D, kind="wmod")
    _lin(m, "final_layer.linear", p.in_channels, D)
    _lin(m, "final_layer.adaLN_modulation.1", 2 * D, D, kind="wmod")
    return m


def _conv(m: Manifest, key: str, out: int, inp: int, k: int) -> None:
    m.append((key + ".weight", (out, inp, k, k), "w"))
    m.append((key + ".bias", (out,), "b"))


def _gn(m: Manifest, key: str, c: int) -> None:
    m.append((key + ".weight", (c,), "scale"))
    m.append((key + ".bias", (c,), "nb"))


def _res(m: Manifest, key: str, cin: int, cout: int) -> None:
    _gn(m, key + ".norm1", cin)
    _conv(m, key + ".conv1", cout, cin, 3)
    _gn(m, key + ".norm2", cout)
    _conv(m, key + ".conv2", cout, cout, 3)
    if cin != cout:
        _conv(m, key + ".nin_shortcut", cout, cin, 1)


def ae_decoder_manifest(a: AutoEncoderParams) -> Manifest:
    """`decoder.*` keys of ae.safetensors (flux/autoencoder.py:212-269)."""
    m: Manifest = []
    n = len(a.ch_mult)
    block_in = a.ch * a.ch_mult[n - 1]
    _conv(m, "decoder.conv_in", block_in, a.z_channels, 3)
    _res(m, "decoder.mid.block_1", block_in, block_in)
    _gn(m, "decoder.mid.attn_1.norm", block_in)
    for nm in ("q", "k", "v", "proj_out"):
        _conv(m, f"decoder.mid.attn_1.{nm}", block_in, block_in, 1)
    _res(m, "decoder.mid.block_2", block_in, block_in)
    for lvl in reversed(range(n)):
        block_out = a.ch * a.ch_mult[lvl]
        for b in range(a.num_res_blocks + 1):
            _res(m, f"decoder.up.{lvl}.block.{b}", block_in, block_out)
            block_in = block_out
        if lvl != 0:
            _conv(m, f"decoder.up.{lvl}.upsample.conv", block_in, block_in, 3)
    _gn(m, "decoder.norm_out", block_in)
    _conv(m, "decoder.conv_out", a.out_ch, block_in, 3)
    return m


def ae_encoder_manifest(a: AutoEncoderParams) -> Manifest:
    """`encoder.*` keys of ae.safetensors (flux/autoencoder.py:127-178)."""
    m: Manifest = []
    n = len(a.ch_mult)
    _conv(m, "encoder.conv_in", a.ch, a.in_channels, 3)
    in_mult = (1,) + tuple(a.ch_mult)
    block_in = a.ch
    for lvl in range(n):
        block_in, block_out = a.ch * in_mult[lvl], a.ch * a.ch_mult[lvl]
        for b in range(a.num_res_blocks):
            _res(m, f"encoder.down.{lvl}.block.{b}", block_in, block_out)
            block_in = block_out
        if lvl != n - 1:
            _conv(m, f"encoder.down.{lvl}.downsample.conv", block_in, block_in, 3)
    _res(m, "encoder.mid.block_1", block_in, block_in)
    _gn(m, "encoder.mid.attn_1.norm", block_in)
    for nm in ("q", "k", "v", "proj_out"):
        _conv(m, f"encoder.mid.attn_1.{nm}", block_in, block_in, 1)
    _res(m, "encoder.mid.block_2", block_in, block_in)
    _gn(m, "encoder.norm_out", block_in)
    _conv(m, "encoder.conv_out", 2 * a.z_channels, block_in, 3)
    return m


def t5_manifest(c: T5Config) -> Manifest:
    """HF T5 encoder keys (flux/t5.py:10-31 replacement patterns)."""
    inner = c.d_kv * c.num_heads
    m: Manifest = [("shared.weight", (c.vocab_size, c.d_model), "emb")]
    for i in range(c.num_layers):
        pre = f"encoder.block.{i}.layer."
        for nm in ("q", "k", "v"):
            _lin(m, f"{pre}0.SelfAttention.{nm}", inner, c.d_model, bias=False)
        _lin(m, f"{pre}0.SelfAttention.o", c.d_model, inner, bias=False)
        m.append((f"{pre}0.layer_norm.weight", (c.d_model,), "scale"))
        _lin(m, f"{pre}1.DenseReluDense.wi_0", c.d_ff, c.d_model, bias=False)
        _lin(m, f"{pre}1.DenseReluDense.wi_1", c.d_ff, c.d_model, bias=False)
        _lin(m, f"{pre}1.DenseReluDense.wo", c.d_model, c.d_ff, bias=False)
        m.append((f"{pre}1.layer_norm.weight", (c.d_model,), "scale"))
    m.append(("encoder.block.0.layer.0.SelfAttention.relative_attention_bias.weight",
              (c.relative_attention_num_buckets, c.num_heads), "b"))
    m.append(("encoder.final_layer_norm.weight", (c.d_model,), "scale"))
    return m


def clip_manifest(c: CLIPTextModelConfig) -> Manifest:
    """HF CLIP text-model keys (flux/clip.py:96-125)."""
    D = c.model_dims
    m: Manifest = [("text_model.embeddings.token_embedding.weight", (c.vocab_size, D), "emb"),
                   ("text_model.embeddings.position_embedding.weight", (c.max_length, D), "emb")]
    for i in range(c.num_layers):
        pre = f"text_model.encoder.layers.{i}."
        for nm in ("q_proj", "k_proj", "v_proj", "out_proj"):
            _lin(m, f"{pre}self_attn.{nm}", D, D)
        _gn(m, f"{pre}layer_norm1", D)
        _gn(m, f"{pre}layer_norm2", D)
        _lin(m, f"{pre}mlp.fc1", 4 * D, D)
        _lin(m, f"{pre}mlp.fc2", D, 4 * D)
    _gn(m, "text_model.final_layer_norm", D)
    return m
