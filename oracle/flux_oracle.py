"""CPU oracle for the Flux denoising hot path -- TEST INFRASTRUCTURE ONLY.

This file is a plain PyTorch-CPU (fp32, optional fp64 / bf16-emulating) restatement of the
reference's algorithm for the path named by BASELINE.json:north_star.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import it; the product package (``flux-generator_b200/flux``) never does and fails loudly when its
CUDA library is missing.

Parity status: the reference executes on Apple MLX, which is not installable here (no wheel, no
network), and the reference's own tests pin no numeric result on this path (every test mocks
``FluxPipeline``: test/test_api.py:51-63, test/test_generation.py:183-214).  This oracle is
therefore pinned against the *reference's own Python code* executed over a small MLX-API shim
(``oracle/mlx_shim``; fixtures in ``tests/golden`` written by ``oracle/gen_golden.py``), i.e. the
module structure, split/concat orders, reshapes and constants are the reference's; the arithmetic
inside each MLX primitive is restated from MLX's public documentation.  "parity unpinned" against
a real MLX run -- stated here, in DESIGN.md and in the tests.

Every function cites the reference file:line it follows (paths relative to /root/reference).
Weights are passed as a flat ``dict[str, Tensor]`` under the *checkpoint-side* key names the
reference's sanitizers accept (flux/model.py:85-97, flux/autoencoder.py:336-345,
flux/t5.py:10-31,232-241, flux/clip.py:96-125); Linear weights are ``[out, in]``, conv weights
``OIHW`` as in the checkpoint files.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import re

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# --------------------------------------------------------------------------------------------
# precision modes
# --------------------------------------------------------------------------------------------
class Mode:
    """fp32: every op in float32.  fp64: float64.  bf16: float32 math, result of every reference
    op rounded to bfloat16 (how an unfused bf16 MLX graph behaves: flux/flux.py:24)."""

    def __init__(self, name: str = "fp32", quantize: bool = False, quantize_attention: Optional[bool] = None, bits: int = 8,
                 fp4_scope: str = "all", fp4_fused: bool = True):
        assert name in ("fp32", "fp64", "bf16")
        self.name = name
        self.dtype = torch.float64 if name == "fp64" else torch.float32
        # quantize_attention (default: follows `quantize`): restates Flux.quantize(attention=True) -- q, k (after
        # QK-norm and RoPE) and v are cast to e4m3 without a scale, the softmax numerators are cast to e4m3 after a
        # 2^4 scale (fx_attention fp8 mode), the row sum and everything else stay fp32
        self.quantize_attention = quantize if quantize_attention is None else quantize_attention
        # bits = 4: restates Flux.quantize(bits=4) -- NVFP4 activations and weights (nvfp4_quant_rows) for every block Linear
        # (fp4_scope "all": FP8_LINEARS) or only for the K-long ones that read the attention | GELU(mlp) buffer (fp4_scope
        # "cat": FP4_LINEARS, the other block Linears then stay FP8)
        self.bits = bits
        assert fp4_scope in ("all", "cat")
        self.fp4_scope = fp4_scope
        # fp4_fused (with fp4_scope "all"): the operands of proj / mlp.2 / linear2 are emitted by their producers (attention
        # epilogue, GELU epilogue) in the chunked form of nvfp4_quant_rows_chunked instead of the row quantiser's
        self.fp4_fused = fp4_fused
        # quantize: restates THIS repo's --quantize path (not the reference's MLX 4-bit nn.quantize, which cannot be
        # restated without MLX's packed group format): the block Linears matched by FP8_LINEARS see row-quantised
        # e4m3 activations and weights (fx_quantize_rows in include/flux_b200.h), everything else is unchanged.
        self.quantize = quantize

    def r(self, x: Tensor) -> Tensor:
        if self.name == "bf16":
            return x.to(torch.bfloat16).to(torch.float32)
        return x

    def w(self, x: Tensor) -> Tensor:
        return x.to(self.dtype)


FP32 = Mode("fp32")


FP8_LINEARS = re.compile(r"^(double_blocks\.\d+\.(img|txt)_(attn\.qkv|attn\.proj|mlp\.0|mlp\.2)|single_blocks\.\d+\.linear[12])$")


FP4_FUSED_LINEARS = re.compile(r"^(double_blocks\.\d+\.(img|txt)_(attn\.proj|mlp\.2)|single_blocks\.\d+\.linear2)$")


FP4_LINEARS = re.compile(r"^(double_blocks\.\d+\.(img|txt)_(attn\.proj|mlp\.2)|single_blocks\.\d+\.linear2)$")


def fp8_quant_rows(x: Tensor) -> Tuple[Tensor, Tensor]:
    """Row-wise e4m3 quantisation as fx_quantize_rows does it (all in fp32): inv = 448 / absmax, q = e4m3_rn_sat(x * inv),
    dequantisation scale = absmax * (1/448); a zero row gets inv = scale = 1.  Returns (q as float32 values, scale [..., 1])."""
    x32 = x.to(torch.float32)
    amax = x32.abs().amax(dim=-1, keepdim=True)
    one = torch.ones_like(amax)
    scale = torch.where(amax > 0, amax * torch.tensor(1.0 / 448.0, dtype=torch.float32), one)
    inv = torch.where(amax > 0, torch.tensor(448.0, dtype=torch.float32) / amax, one)
    q = (x32 * inv).clamp(-448.0, 448.0).to(torch.float8_e4m3fn).to(torch.float32)
    return q, scale


E2M1_VALUES = (0.0, 0.5, 1.0, 1.5, 2.0, 3.0, 4.0, 6.0)


def e2m1_round(v: Tensor) -> Tensor:
    """Round to nearest e2m1 value (ties to the even mantissa), saturating at +-6: cvt.rn.satfinite.e2m1x2.f32."""
    a = v.abs()
    q = torch.zeros_like(a)
    q = torch.where(a > 0.25, torch.full_like(a, 0.5), q)
    q = torch.where(a >= 0.75, torch.full_like(a, 1.0), q)
    q = torch.where(a > 1.25, torch.full_like(a, 1.5), q)
    q = torch.where(a >= 1.75, torch.full_like(a, 2.0), q)
    q = torch.where(a > 2.5, torch.full_like(a, 3.0), q)
    q = torch.where(a >= 3.5, torch.full_like(a, 4.0), q)
    q = torch.where(a > 5.0, torch.full_like(a, 6.0), q)
    return torch.copysign(q, v)


def nvfp4_quant_rows(x: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
    """THIS repo's NVFP4 quantiser (fx_quantize_rows_fp4), every step one IEEE fp32 operation:
    g = absmax(row) * (1/2688) (1 for a zero row); per block of 16: sf = e4m3_rn(absmax(block) * (1/6) * rcp(g)), d = sf * g,
    q = e2m1_rn_sat(x * rcp(d)) (0 where d == 0), rcp = correctly rounded reciprocal.  Returns (q values, sf values
    [.., K/16], g [.., 1]); x ~= q * sf * g."""
    x32 = x.to(torch.float32)
    K = x32.shape[-1]
    amax = x32.abs().amax(dim=-1, keepdim=True)
    g = torch.where(amax > 0, amax * torch.tensor(1.0 / 2688.0, dtype=torch.float32), torch.ones_like(amax))
    xb = x32.reshape(*x32.shape[:-1], K // 16, 16)
    bmax = xb.abs().amax(dim=-1)
    one = torch.ones((), dtype=torch.float32)
    u = (bmax * torch.tensor(1.0 / 6.0, dtype=torch.float32)) * (one / g)
    sf = u.clamp(max=448.0).to(torch.float8_e4m3fn).to(torch.float32)
    d = (sf * g).unsqueeze(-1)
    rd = torch.where(d > 0, one / torch.where(d > 0, d, torch.ones_like(d)), torch.zeros_like(d))
    q = e2m1_round(xb * rd)
    return q.reshape(x32.shape), sf, g


def nvfp4_quant_rows_chunked(x: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
    """The NVFP4 operand as the PRODUCING kernels emit it (fx_gemm_fp4 with q_out / fx_quantize_chunks_fp4, then fx_fp4_finalize):
    the first-level scale cannot be the row's absmax (a GEMM epilogue sees 32 columns of a row at a time), so every chunk of 32
    columns takes a power-of-two scale g_c = 2^e_c, e_c = ceil(log2(absmax(chunk) * (1/2688))) (>= -100), the two blocks of 16
    quantise against it exactly like nvfp4_quant_rows does against g (sf = e4m3_rn(absmax(block) * (1/6) * 2^-e_c), d = sf * g_c,
    q = e2m1_rn_sat(x * rcp(d))), and a finalise pass lifts everything to the row's scale g = 2^max(e_c): sf' = e4m3_rn(sf *
    2^(e_c - e_row)) -- an exact exponent shift unless sf' leaves the normal range (blocks 2^-14 below the row maximum).
    Returns (q values, sf' values [.., K/16], g [.., 1]); x ~= q * sf' * g."""
    x32 = x.to(torch.float32)
    K = x32.shape[-1]
    cmax = x32.reshape(*x32.shape[:-1], K // 32, 32).abs().amax(dim=-1)
    t = cmax * torch.tensor(1.0 / 2688.0, dtype=torch.float32)
    mant, ex = torch.frexp(t)
    e_c = torch.where(mant == 0.5, ex - 1, ex).clamp(min=-100)
    e_c = torch.where(t > 0, e_c, torch.full_like(e_c, -100))
    e_row = e_c.amax(dim=-1, keepdim=True)
    one = torch.ones((), dtype=torch.float32)
    ones = torch.ones_like(t)
    g_c = torch.ldexp(ones, e_c).repeat_interleave(2, dim=-1)                # per block of 16
    inv_g = torch.ldexp(ones, -e_c).repeat_interleave(2, dim=-1)
    xb = x32.reshape(*x32.shape[:-1], K // 16, 16)
    bmax = xb.abs().amax(dim=-1)
    u = (bmax * torch.tensor(1.0 / 6.0, dtype=torch.float32)) * inv_g
    sf = u.clamp(max=448.0).to(torch.float8_e4m3fn).to(torch.float32)
    d = (sf * g_c).unsqueeze(-1)
    rd = torch.where(d > 0, one / torch.where(d > 0, d, torch.ones_like(d)), torch.zeros_like(d))
    q = e2m1_round(xb * rd)
    shift = torch.ldexp(ones, (e_c - e_row).clamp(min=-60)).repeat_interleave(2, dim=-1)
    sf_final = (sf * shift).to(torch.float8_e4m3fn).to(torch.float32)
    return q.reshape(x32.shape), sf_final, torch.ldexp(torch.ones_like(e_row, dtype=torch.float32), e_row)


def _linear(m: Mode, x: Tensor, sd: Dict[str, Tensor], key: str, bias: bool = True) -> Tensor:
    w = m.w(sd[key + ".weight"])
    b = m.w(sd[key + ".bias"]) if bias and (key + ".bias") in sd else None
    if m.quantize and m.bits == 4 and (FP8_LINEARS if m.fp4_scope == "all" else FP4_LINEARS).match(key):
        chunked = m.fp4_fused and m.fp4_scope == "all" and FP4_FUSED_LINEARS.match(key)
        qx, sx, gx = nvfp4_quant_rows_chunked(x) if chunked else nvfp4_quant_rows(x)
        qw, sw, gw = nvfp4_quant_rows(w)
        xd = qx * sx.repeat_interleave(16, dim=-1) * gx
        wd = qw * sw.repeat_interleave(16, dim=-1) * gw
        y = F.linear(xd, wd).to(m.dtype)
        return m.r(y + b if b is not None else y)
    if m.quantize and FP8_LINEARS.match(key):
        xq, xs = fp8_quant_rows(x)
        wq, ws = fp8_quant_rows(w)
        y = F.linear(xq, wq) * xs * ws.squeeze(-1)
        return m.r(y.to(m.dtype) + b if b is not None else y.to(m.dtype))
    return m.r(F.linear(x, w, b))


# --------------------------------------------------------------------------------------------
# N4: LoRA adapters at inference   (txt2image.py:32-39, flux/flux.py:228-246, flux/lora.py:28-43)
# --------------------------------------------------------------------------------------------
def lora_blocks(depth: int, depth_single: int, num_blocks: int) -> List[str]:
    """Prefixes of the blocks FluxPipeline.linear_to_lora_layers wraps (flux/flux.py:230-233): the list
    double_blocks + single_blocks REVERSED, first `num_blocks` entries (all of them when num_blocks <= 0)."""
    allb = [f"double_blocks.{i}" for i in range(depth)] + [f"single_blocks.{i}" for i in range(depth_single)]
    allb.reverse()
    return allb[:num_blocks if num_blocks > 0 else len(allb)]


def lora_checkpoint_key(module_path: str) -> str:
    """MLX module path -> checkpoint-side Linear name: the inverse of Flux.sanitize's nn.Sequential renaming
    (flux/model.py:92-95: `.img_mlp.` -> `.img_mlp.layers.`)."""
    return module_path.replace(".layers.", ".")


def lora_fuse(sd: Dict[str, Tensor], adapter: Dict[str, Tensor], num_blocks: int, depth: int, depth_single: int,
              scale: float = 1.0) -> Dict[str, Tensor]:
    """LoRALinear.fuse (flux/lora.py:28-43) for every adapted Linear: W' = W + ((scale * lora_b.T) @ lora_a.T).astype(W.dtype).
    `adapter` holds `<module path>.lora_a` [in, r] / `.lora_b` [r, out] as dreambooth.py:46-59 saves them; entries
    outside the wrapped blocks are ignored exactly like load_weights(strict=False) ignores them (txt2image.py:37)."""
    out = dict(sd)
    prefixes = tuple(p + "." for p in lora_blocks(depth, depth_single, num_blocks))
    for k, a in adapter.items():
        if not k.endswith(".lora_a") or not k.startswith(prefixes):
            continue
        mod = k[:-len(".lora_a")]
        b = adapter[mod + ".lora_b"]
        wk = lora_checkpoint_key(mod) + ".weight"
        w = sd[wk]
        delta = (scale * b.to(torch.float32).T) @ a.to(torch.float32).T
        out[wk] = w + delta.to(w.dtype)
    return out


# --------------------------------------------------------------------------------------------
# a1: schedule   (flux/sampler.py:15-31)
# --------------------------------------------------------------------------------------------
def timesteps(num_steps: int, image_seq_len: int, schnell: bool, start: float = 1.0,
              stop: float = 0.0, base_shift: float = 0.5, max_shift: float = 1.15) -> List[float]:
    """flux/sampler.py:22-31.  mx.linspace in f32 as (1-i/n)*start + (i/n)*stop; dev applies
    _time_shift (flux/sampler.py:15-20) with a python-double mu and f32 tensor math."""
    n = num_steps
    i = np.arange(n + 1, dtype=np.float32)
    step = i / np.float32(n)
    t = (np.float32(1.0) - step) * np.float32(start) + step * np.float32(stop)
    t = t.astype(np.float32)
    if not schnell:
        x1, x2 = 256, 4096
        mu = (image_seq_len - x1) * (max_shift - base_shift) / (x2 - x1) + base_shift
        exp_mu = np.float32(math.exp(mu))
        with np.errstate(divide="ignore"):
            t = exp_mu / (exp_mu + (np.float32(1.0) / t - np.float32(1.0)))
        t = t.astype(np.float32)
    return [float(v) for v in t]


def euler_step(m: Mode, pred: Tensor, x_t: Tensor, t: float, t_prev: float) -> Tensor:
    """flux/sampler.py:56-57."""
    return m.r(x_t + m.r((t_prev - t) * pred))


# --------------------------------------------------------------------------------------------
# a4: patchify / ids   (flux/flux.py:53-71, inverse flux/flux.py:157-160)
# --------------------------------------------------------------------------------------------
def prepare_latent_images(x: Tensor) -> Tuple[Tensor, Tensor]:
    b, h, w, c = x.shape
    x = x.reshape(b, h // 2, 2, w // 2, 2, c).permute(0, 1, 3, 5, 2, 4).reshape(b, h * w // 4, c * 4)
    i = torch.zeros((h // 2, w // 2), dtype=torch.int32)
    j, k = torch.meshgrid(torch.arange(h // 2, dtype=torch.int32),
                          torch.arange(w // 2, dtype=torch.int32), indexing="ij")
    ids = torch.stack([i, j, k], dim=-1).reshape(1, h * w // 4, 3).repeat(b, 1, 1)
    return x, ids


def unpatchify(x: Tensor, latent_size: Tuple[int, int]) -> Tensor:
    h, w = latent_size
    b = x.shape[0]
    x = x.reshape(b, h // 2, w // 2, -1, 2, 2).permute(0, 1, 4, 2, 5, 3).reshape(b, h, w, -1)
    return x


# --------------------------------------------------------------------------------------------
# a7, a9, a10, a11: embeddings, RoPE, attention   (flux/layers.py:12-57)
# --------------------------------------------------------------------------------------------
def timestep_embedding(m: Mode, t: Tensor, dim: int, max_period: int = 10000,
                       time_factor: float = 1000.0) -> Tensor:
    """flux/layers.py:46-57.  dtype-driven exactly like the reference: `time_factor * t` is
    evaluated in t's dtype (bf16 for every pipeline call, flux/flux.py:101-102 -> the model sees
    bf16(1000*bf16(t))), the sinusoid in f32, and the result is cast back to t's dtype."""
    half = dim // 2
    freqs = torch.arange(0, half, dtype=torch.float32) / half
    freqs = torch.exp(freqs * (-math.log(max_period)))
    tt = (time_factor * t).to(torch.float32)
    x = tt[:, None] * freqs[None]
    x = torch.cat([torch.cos(x), torch.sin(x)], dim=-1)
    return x.to(t.dtype).to(m.dtype)


def rope(pos: Tensor, dim: int, theta: float) -> Tensor:
    """flux/layers.py:12-21 -> [..., dim/2, 2, 2] rotation matrices (f32)."""
    scale = torch.arange(0, dim, 2, dtype=torch.float32) / dim
    omega = 1.0 / (theta ** scale)
    x = pos[..., None].to(torch.float32) * omega
    cosx, sinx = torch.cos(x), torch.sin(x)
    pe = torch.stack([cosx, -sinx, sinx, cosx], dim=-1)
    return pe.reshape(*pe.shape[:-1], 2, 2)


def embed_nd(ids: Tensor, axes_dim: Sequence[int], theta: float) -> Tensor:
    """flux/layers.py:60-75 -> [B, 1, N, sum(axes)/2, 2, 2]."""
    pe = torch.cat([rope(ids[..., i], axes_dim[i], theta) for i in range(ids.shape[-1])], dim=-3)
    return pe[:, None]


def apply_rope(m: Mode, x: Tensor, pe: Tensor) -> Tensor:
    """flux/layers.py:24-33: adjacent pairs; out = x0*pe[...,0] + x1*pe[...,1]."""
    s = x.shape
    x = x.reshape(*s[:-1], -1, 1, 2)
    out = m.r(x[..., 0] * pe[..., 0] + x[..., 1] * pe[..., 1])
    return out.reshape(s)


def sdpa(m: Mode, q: Tensor, k: Tensor, v: Tensor, scale: float, mask: Optional[Tensor] = None) -> Tensor:
    """mx.fast.scaled_dot_product_attention: softmax((q*scale) k^T + mask) v, fp32 softmax."""
    s = torch.matmul(q * scale, k.transpose(-1, -2))
    if mask is not None:
        s = s + mask
    p = torch.softmax(s, dim=-1)
    return m.r(torch.matmul(p, v))


def _e4m3(x: Tensor) -> Tensor:
    return x.to(torch.float32).clamp(-448.0, 448.0).to(torch.float8_e4m3fn).to(x.dtype)


def sdpa_fp8(m: Mode, q: Tensor, k: Tensor, v: Tensor, scale: float) -> Tensor:
    """THIS repo's FP8 attention (fx_attention, fp8 = 1), not a reference function: e4m3 q / k / v, fp32 scores and softmax
    statistics, numerators exp(s - max) * 2^4 cast to e4m3 for the P V product, the row sum taken from the unrounded
    numerators.  (The kernel's running maximum may lag the true one by up to 2^4, which moves some roundings by a
    binade; the tolerance of the tests covers it.)"""
    q, k, v = _e4m3(q), _e4m3(k), _e4m3(v)
    s = torch.matmul(q, k.transpose(-1, -2)) * scale
    e = torch.exp(s - s.amax(dim=-1, keepdim=True))
    p8 = _e4m3(e * 16.0) / 16.0
    return m.r(torch.matmul(p8, v) / e.sum(dim=-1, keepdim=True))


def attention(m: Mode, q: Tensor, k: Tensor, v: Tensor, pe: Tensor) -> Tensor:
    """flux/layers.py:36-43."""
    B, H, L, D = q.shape
    q = apply_rope(m, q, pe)
    k = apply_rope(m, k, pe)
    x = sdpa_fp8(m, q, k, v, D ** (-0.5)) if m.quantize_attention else sdpa(m, q, k, v, D ** (-0.5))
    return x.transpose(1, 2).reshape(B, L, -1)


def layer_norm(m: Mode, x: Tensor, eps: float = 1e-6) -> Tensor:
    """nn.LayerNorm(affine=False, eps=1e-6)  (flux/layers.py:156)."""
    return m.r(F.layer_norm(x, (x.shape[-1],), eps=eps))


def rms_norm(m: Mode, x: Tensor, w: Tensor, eps: float) -> Tensor:
    """nn.RMSNorm (mx.fast.rms_norm): x * rsqrt(mean(x^2) + eps) * w, fp32 accumulation."""
    v = x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps)
    return m.r(v * m.w(w))


def gelu_tanh(x: Tensor) -> Tensor:
    return F.gelu(x, approximate="tanh")


def silu(x: Tensor) -> Tensor:
    return F.silu(x)


# --------------------------------------------------------------------------------------------
# a6, a8, a12-a16: the MMDiT   (flux/model.py:99-136, flux/layers.py:78-302)
# --------------------------------------------------------------------------------------------
@dataclass
class FluxParams:
    """flux/model.py:20-32 (values flux/utils.py:36-49)."""
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


QK_RMS_EPS = 1e-5  # MLX nn.RMSNorm default eps (flux/layers.py:91-92 pass none)


def mlp_embedder(m: Mode, sd, key: str, x: Tensor) -> Tensor:
    """flux/layers.py:78-85."""
    return _linear(m, m.r(silu(_linear(m, x, sd, key + ".in_layer"))), sd, key + ".out_layer")


def modulation(m: Mode, sd, key: str, vec: Tensor, multiplier: int) -> List[Tensor]:
    """flux/layers.py:129-143: lin(silu(vec)) split into (shift, scale, gate)[x2]."""
    x = _linear(m, m.r(silu(vec)), sd, key + ".lin")
    return list(torch.chunk(x[:, None, :], multiplier, dim=-1))


def _heads(x: Tensor, H: int) -> Tensor:
    B, L, _ = x.shape
    return x.reshape(B, L, H, -1).transpose(1, 2)


def _modulate(m: Mode, x: Tensor, shift: Tensor, scale: Tensor) -> Tensor:
    return m.r(m.r(m.r(1 + scale) * layer_norm(m, x)) + shift)


def double_block(m: Mode, sd, p: FluxParams, i: int, img: Tensor, txt: Tensor, vec: Tensor,
                 pe: Tensor) -> Tuple[Tensor, Tensor]:
    """flux/layers.py:181-231."""
    pre = f"double_blocks.{i}."
    H = p.num_heads
    S = txt.shape[1]
    i_sh1, i_sc1, i_g1, i_sh2, i_sc2, i_g2 = modulation(m, sd, pre + "img_mod", vec, 6)
    t_sh1, t_sc1, t_g1, t_sh2, t_sc2, t_g2 = modulation(m, sd, pre + "txt_mod", vec, 6)

    def qkv(x, sh, sc, name):
        xm = _modulate(m, x, sh, sc)
        q, k, v = torch.chunk(_linear(m, xm, sd, pre + name + ".qkv", p.qkv_bias), 3, dim=-1)
        q, k, v = _heads(q, H), _heads(k, H), _heads(v, H)
        q = rms_norm(m, q, sd[pre + name + ".norm.query_norm.scale"], QK_RMS_EPS)
        k = rms_norm(m, k, sd[pre + name + ".norm.key_norm.scale"], QK_RMS_EPS)
        return q, k, v

    iq, ik, iv = qkv(img, i_sh1, i_sc1, "img_attn")
    tq, tk, tv = qkv(txt, t_sh1, t_sc1, "txt_attn")
    q = torch.cat([tq, iq], dim=2)
    k = torch.cat([tk, ik], dim=2)
    v = torch.cat([tv, iv], dim=2)
    attn = attention(m, q, k, v, pe)
    txt_attn, img_attn = attn[:, :S], attn[:, S:]

    def tail(x, a, g1, sh2, sc2, g2, attn_name, mlp_name):
        x = m.r(x + m.r(g1 * _linear(m, a, sd, pre + attn_name + ".proj")))
        h = _linear(m, _modulate(m, x, sh2, sc2), sd, pre + mlp_name + ".0")
        h = _linear(m, m.r(gelu_tanh(h)), sd, pre + mlp_name + ".2")
        return m.r(x + m.r(g2 * h))

    img = tail(img, img_attn, i_g1, i_sh2, i_sc2, i_g2, "img_attn", "img_mlp")
    txt = tail(txt, txt_attn, t_g1, t_sh2, t_sc2, t_g2, "txt_attn", "txt_mlp")
    return img, txt


def single_block(m: Mode, sd, p: FluxParams, i: int, x: Tensor, vec: Tensor, pe: Tensor) -> Tensor:
    """flux/layers.py:262-284."""
    pre = f"single_blocks.{i}."
    H, D = p.num_heads, p.hidden_size
    sh, sc, g = modulation(m, sd, pre + "modulation", vec, 3)
    xm = _modulate(m, x, sh, sc)
    y = _linear(m, xm, sd, pre + "linear1")
    q, k, v, mlp = y[..., :D], y[..., D:2 * D], y[..., 2 * D:3 * D], y[..., 3 * D:]
    q, k, v = _heads(q, H), _heads(k, H), _heads(v, H)
    q = rms_norm(m, q, sd[pre + "norm.query_norm.scale"], QK_RMS_EPS)
    k = rms_norm(m, k, sd[pre + "norm.key_norm.scale"], QK_RMS_EPS)
    a = attention(m, q, k, v, pe)
    y = _linear(m, torch.cat([a, m.r(gelu_tanh(mlp))], dim=2), sd, pre + "linear2")
    return m.r(x + m.r(g * y))


def last_layer(m: Mode, sd, x: Tensor, vec: Tensor) -> Tensor:
    """flux/layers.py:287-302 (shift first, then scale)."""
    mod = _linear(m, m.r(silu(vec)), sd, "final_layer.adaLN_modulation.1")
    shift, scale = torch.chunk(mod, 2, dim=1)
    x = _modulate(m, x, shift[:, None, :], scale[:, None, :])
    return _linear(m, x, sd, "final_layer.linear")


def flux_vec(m: Mode, sd, p: FluxParams, timesteps_: Tensor, y: Tensor,
             guidance: Optional[Tensor]) -> Tensor:
    """flux/model.py:113-120."""
    vec = mlp_embedder(m, sd, "time_in", timestep_embedding(m, timesteps_, 256))
    if p.guidance_embed:
        if guidance is None:
            raise ValueError("Didn't get guidance strength for guidance distilled model.")
        vec = m.r(vec + mlp_embedder(m, sd, "guidance_in", timestep_embedding(m, guidance, 256)))
    return m.r(vec + mlp_embedder(m, sd, "vector_in", y))


def flux_forward(sd: Dict[str, Tensor], p: FluxParams, img: Tensor, img_ids: Tensor, txt: Tensor,
                 txt_ids: Tensor, timesteps_: Tensor, y: Tensor, guidance: Optional[Tensor] = None,
                 mode: Mode = FP32, taps: Optional[dict] = None) -> Tensor:
    """flux/model.py:99-136.  `taps`, if given, receives intermediate activations by name."""
    m = mode
    if img.ndim != 3 or txt.ndim != 3:
        raise ValueError("Input img and txt tensors must have 3 dimensions.")
    img, txt, y = m.w(img), m.w(txt), m.w(y)

    img = _linear(m, img, sd, "img_in")
    vec = flux_vec(m, sd, p, timesteps_, y, guidance)
    txt = _linear(m, txt, sd, "txt_in")
    ids = torch.cat([txt_ids, img_ids], dim=1)
    pe = m.r(embed_nd(ids, p.axes_dim, p.theta)).to(m.dtype)
    if taps is not None:
        taps.update(vec=vec, img_in=img, txt_in=txt)

    for i in range(p.depth):
        img, txt = double_block(m, sd, p, i, img, txt, vec, pe)
        if taps is not None:
            taps[f"double.{i}.img"] = img
            taps[f"double.{i}.txt"] = txt
    x = torch.cat([txt, img], dim=1)
    for i in range(p.depth_single_blocks):
        x = single_block(m, sd, p, i, x, vec, pe)
        if taps is not None:
            taps[f"single.{i}"] = x
    x = x[:, txt.shape[1]:, ...]
    return last_layer(m, sd, x, vec)


def denoise(sd, p: FluxParams, x_T: Tensor, x_ids: Tensor, txt: Tensor, txt_ids: Tensor, vec: Tensor,
            num_steps: int, guidance: float, schnell: bool, mode: Mode = FP32) -> List[Tensor]:
    """flux/flux.py:87-126: returns the list of num_steps latents.  t and guidance pass through
    `mx.full((B,), x, bf16)` (flux/flux.py:101-102) -> rounded to bf16 in every mode."""
    m = mode
    B = x_T.shape[0]

    def scalar(v: float) -> Tensor:
        return torch.full((B,), v, dtype=torch.bfloat16)

    ts = timesteps(num_steps, x_T.shape[1], schnell)
    g = scalar(guidance)
    x_t = m.w(x_T)
    out = []
    for i in range(num_steps):
        t, t_prev = ts[i], ts[i + 1]
        pred = flux_forward(sd, p, x_t, x_ids, txt, txt_ids, scalar(t), vec, g, mode=m)
        x_t = euler_step(m, pred, x_t, t, t_prev)
        out.append(x_t)
    return out


# --------------------------------------------------------------------------------------------
# a18: VAE decoder   (flux/autoencoder.py:24-124, 212-297, 352-354; flux/flux.py:157-162)
# --------------------------------------------------------------------------------------------
@dataclass
class AutoEncoderParams:
    """flux/autoencoder.py:11-21 (values flux/utils.py:51-61)."""
    resolution: int = 256
    in_channels: int = 3
    ch: int = 128
    out_ch: int = 3
    ch_mult: List[int] = field(default_factory=lambda: [1, 2, 4, 4])
    num_res_blocks: int = 2
    z_channels: int = 16
    scale_factor: float = 0.3611
    shift_factor: float = 0.1159


def _conv(m: Mode, x: Tensor, sd, key: str) -> Tensor:
    """NHWC activations, OIHW checkpoint weights (sanitize -> OHWI, flux/autoencoder.py:336-345)."""
    w = m.w(sd[key + ".weight"])
    b = m.w(sd[key + ".bias"])
    y = F.conv2d(x.permute(0, 3, 1, 2), w, b, padding=w.shape[-1] // 2)
    return m.r(y.permute(0, 2, 3, 1))


def _group_norm(m: Mode, x: Tensor, sd, key: str) -> Tensor:
    """nn.GroupNorm(32, eps=1e-6, pytorch_compatible=True) on NHWC (flux/autoencoder.py:29-35)."""
    y = F.group_norm(x.permute(0, 3, 1, 2), 32, m.w(sd[key + ".weight"]), m.w(sd[key + ".bias"]), eps=1e-6)
    return m.r(y.permute(0, 2, 3, 1))


def _resnet(m: Mode, x: Tensor, sd, key: str) -> Tensor:
    """flux/autoencoder.py:85-98."""
    h = _conv(m, m.r(silu(_group_norm(m, x, sd, key + ".norm1"))), sd, key + ".conv1")
    h = _conv(m, m.r(silu(_group_norm(m, h, sd, key + ".norm2"))), sd, key + ".conv2")
    if (key + ".nin_shortcut.weight") in sd:
        x = _conv(m, x, sd, key + ".nin_shortcut")
    return m.r(x + h)


def _attn_block(m: Mode, x: Tensor, sd, key: str) -> Tensor:
    """flux/autoencoder.py:41-52: single head, scale C^-0.5, 1x1 convs as Linears."""
    B, H, W, C = x.shape
    y = _group_norm(m, x, sd, key + ".norm").reshape(B, 1, H * W, C)

    def lin(t, name):
        w = m.w(sd[f"{key}.{name}.weight"]).reshape(C, C)
        return m.r(F.linear(t, w, m.w(sd[f"{key}.{name}.bias"])))

    q, k, v = lin(y, "q"), lin(y, "k"), lin(y, "v")
    y = lin(sdpa(m, q, k, v, C ** (-0.5)), "proj_out")
    return m.r(x + y.reshape(B, H, W, C))


def vae_decode(sd: Dict[str, Tensor], ap: AutoEncoderParams, z: Tensor, mode: Mode = FP32) -> Tensor:
    """AutoEncoder.decode + Decoder.__call__ (flux/autoencoder.py:352-354, 271-297).  z NHWC."""
    m = mode
    z = m.r(m.r(m.w(z) / ap.scale_factor) + ap.shift_factor)
    h = _conv(m, z, sd, "decoder.conv_in")
    h = _resnet(m, h, sd, "decoder.mid.block_1")
    h = _attn_block(m, h, sd, "decoder.mid.attn_1")
    h = _resnet(m, h, sd, "decoder.mid.block_2")
    n_res = len(ap.ch_mult)
    for lvl in reversed(range(n_res)):
        for blk in range(ap.num_res_blocks + 1):
            h = _resnet(m, h, sd, f"decoder.up.{lvl}.block.{blk}")
        if lvl != 0:
            h = h.repeat_interleave(2, dim=1).repeat_interleave(2, dim=2)  # upsample_nearest (2,2)
            h = _conv(m, h, sd, f"decoder.up.{lvl}.upsample.conv")
    h = m.r(silu(_group_norm(m, h, sd, "decoder.norm_out")))
    return _conv(m, h, sd, "decoder.conv_out")


def vae_encode(sd: Dict[str, Tensor], ap: AutoEncoderParams, x: Tensor, mode: Mode = FP32,
               noise: Optional[Tensor] = None) -> Tensor:
    """AutoEncoder.encode + Encoder.__call__ + DiagonalGaussian (flux/autoencoder.py:347-350, 180-209, 300-309).
    x NHWC [B, H, W, 3] -> z [B, H/8, W/8, z_channels].  Downsample (flux/autoencoder.py:101-113) = pad (0,1,0,1) on H
    and W, then a 3x3 stride-2 convolution without padding.  noise=None is eval mode (the mean)."""
    m = mode
    h = _conv(m, m.w(x), sd, "encoder.conv_in")
    n_res = len(ap.ch_mult)
    for lvl in range(n_res):
        for blk in range(ap.num_res_blocks):
            h = _resnet(m, h, sd, f"encoder.down.{lvl}.block.{blk}")
        if lvl != n_res - 1:
            key = f"encoder.down.{lvl}.downsample.conv"
            hp = F.pad(h.permute(0, 3, 1, 2), (0, 1, 0, 1))
            h = m.r(F.conv2d(hp, m.w(sd[key + ".weight"]), m.w(sd[key + ".bias"]), stride=2).permute(0, 2, 3, 1))
    h = _resnet(m, h, sd, "encoder.mid.block_1")
    h = _attn_block(m, h, sd, "encoder.mid.attn_1")
    h = _resnet(m, h, sd, "encoder.mid.block_2")
    h = m.r(silu(_group_norm(m, h, sd, "encoder.norm_out")))
    h = _conv(m, h, sd, "encoder.conv_out")
    mean, logvar = torch.chunk(h, 2, dim=-1)
    z = mean if noise is None else m.r(mean + m.r(m.r(torch.exp(m.r(0.5 * logvar))) * m.w(noise)))
    return m.r(ap.scale_factor * m.r(z - ap.shift_factor))


def add_noise(m: Mode, x: Tensor, t: Tensor, noise: Tensor) -> Tensor:
    """flux/sampler.py:47-54."""
    t = t.reshape([-1] + [1] * (x.ndim - 1))
    return m.r(m.r(x * m.r(1 - t)) + m.r(t * noise))


def training_loss(sd, p: FluxParams, x_0_nhwc: Tensor, t5_features: Tensor, clip_features: Tensor, guidance: Tensor,
                  t: Tensor, eps: Tensor, mode: Mode = FP32) -> Tensor:
    """FluxPipeline.training_loss (flux/flux.py:195-226) with the two random draws (t, eps) supplied by the caller:
    mean((flow(x_t) + x_0 - eps)^2), x_t = add_noise(x_0, t, eps).  t is a bf16 tensor like the pipeline's dtype."""
    m = mode
    x_0, x_ids = prepare_latent_images(m.w(x_0_nhwc))
    txt_ids = torch.zeros(t5_features.shape[:-1] + (3,), dtype=torch.int32)
    x_t = add_noise(m, x_0, m.w(t), m.w(eps))
    pred = flux_forward(sd, p, x_t, x_ids, m.w(t5_features), txt_ids, t, m.w(clip_features), guidance, mode=m)
    return m.r(m.r(m.r(pred + x_0) - m.w(eps)).square()).mean()


def decode(sd, ap: AutoEncoderParams, x: Tensor, latent_size: Tuple[int, int], mode: Mode = FP32) -> Tensor:
    """FluxPipeline.decode (flux/flux.py:157-162) -> [k, 8h, 8w, 3] in [0, 1]."""
    m = mode
    img = vae_decode(sd, ap, unpatchify(m.w(x), latent_size), m)
    return m.r(torch.clip(m.r(img + 1), 0, 2) * 0.5)


def to_uint8(img: Tensor) -> Tensor:
    """txt2image.py:133,144: (x*255).astype(uint8) -- truncation."""
    return (img * 255).to(torch.uint8)


# --------------------------------------------------------------------------------------------
# a19: text encoders   (flux/t5.py:70-244, flux/clip.py:46-154)
# --------------------------------------------------------------------------------------------
@dataclass
class T5Config:
    """flux/t5.py:34-48; T5-v1.1-XXL public values."""
    vocab_size: int = 32128
    num_layers: int = 24
    num_heads: int = 64
    relative_attention_num_buckets: int = 32
    d_kv: int = 64
    d_model: int = 4096
    d_ff: int = 10240
    relative_attention_max_distance: int = 128
    layer_norm_epsilon: float = 1e-6


def t5_relative_position_bucket(rpos: Tensor, num_buckets: int, max_distance: int) -> Tensor:
    """flux/t5.py:78-97, bidirectional=True."""
    num_buckets = num_buckets // 2
    max_exact = num_buckets // 2
    abspos = rpos.abs()
    is_small = abspos < max_exact
    scale = (num_buckets - max_exact) / math.log(max_distance / max_exact)
    with np.errstate(divide="ignore"):
        large = (torch.log(abspos.to(torch.float32) / max_exact) * scale)
    large = torch.where(abspos > 0, large, torch.zeros_like(large)).to(torch.int16).to(torch.int64)
    large = torch.minimum(max_exact + large, torch.tensor(num_buckets - 1))
    buckets = torch.where(is_small, abspos, large)
    return buckets + (rpos > 0).to(torch.int64) * num_buckets


def t5_position_bias(sd, cfg: T5Config, S: int) -> Tensor:
    """flux/t5.py:99-120 -> [heads, S, S]."""
    ctx = torch.arange(S)[:, None]
    mem = torch.arange(S)[None, :]
    bucket = t5_relative_position_bucket(mem - ctx, cfg.relative_attention_num_buckets,
                                         cfg.relative_attention_max_distance)
    emb = sd["encoder.block.0.layer.0.SelfAttention.relative_attention_bias.weight"].to(torch.float32)
    return emb[bucket].permute(2, 0, 1)


def t5_encode(sd: Dict[str, Tensor], cfg: T5Config, tokens: Tensor, mode: Mode = FP32) -> Tensor:
    """flux/t5.py:203-244: embed -> N x [RMSNorm, MHA(scale 1.0, +bias, no pad mask), RMSNorm,
    gated exact-erf GELU FFN] -> RMSNorm.  tokens [B, S] int."""
    m = mode
    x = m.w(sd["shared.weight"])[tokens.long()]
    B, S, _ = x.shape
    H = cfg.num_heads
    bias = m.r(t5_position_bias(sd, cfg, S)).to(m.dtype)
    eps = cfg.layer_norm_epsilon
    for i in range(cfg.num_layers):
        pre = f"encoder.block.{i}.layer."
        y = rms_norm(m, x, sd[pre + "0.layer_norm.weight"], eps)
        q = _heads(_linear(m, y, sd, pre + "0.SelfAttention.q", False), H)
        k = _heads(_linear(m, y, sd, pre + "0.SelfAttention.k", False), H)
        v = _heads(_linear(m, y, sd, pre + "0.SelfAttention.v", False), H)
        a = sdpa(m, q, k, v, 1.0, bias).transpose(1, 2).reshape(B, S, -1)
        x = m.r(x + _linear(m, a, sd, pre + "0.SelfAttention.o", False))
        y = rms_norm(m, x, sd[pre + "1.layer_norm.weight"], eps)
        g = m.r(F.gelu(_linear(m, y, sd, pre + "1.DenseReluDense.wi_0", False)))
        h = m.r(g * _linear(m, y, sd, pre + "1.DenseReluDense.wi_1", False))
        x = m.r(x + _linear(m, h, sd, pre + "1.DenseReluDense.wo", False))
    return rms_norm(m, x, sd["encoder.final_layer_norm.weight"], eps)


@dataclass
class CLIPConfig:
    """flux/clip.py:12-30; CLIP-L text tower public values."""
    num_layers: int = 12
    model_dims: int = 768
    num_heads: int = 12
    max_length: int = 77
    vocab_size: int = 49408
    hidden_act: str = "quick_gelu"


def clip_encode(sd: Dict[str, Tensor], cfg: CLIPConfig, tokens: Tensor, mode: Mode = FP32) -> Tuple[Tensor, Tensor]:
    """flux/clip.py:127-154 -> (pooled_output [B, D], last_hidden_state [B, N, D]).
    nn.MultiHeadAttention: q*scale, scores + mask, softmax in fp32 (precise), out_proj."""
    m = mode
    B, N = tokens.shape
    eos = tokens.argmax(-1)
    pre = "text_model."
    x = m.w(sd[pre + "embeddings.token_embedding.weight"])[tokens.long()]
    x = m.r(x + m.w(sd[pre + "embeddings.position_embedding.weight"])[:N])
    idx = torch.arange(N)
    mask = (idx[:, None] < idx[None]).to(m.dtype) * -1e9
    H = cfg.num_heads
    D = cfg.model_dims

    def ln(t, key):
        return m.r(F.layer_norm(t, (D,), m.w(sd[key + ".weight"]), m.w(sd[key + ".bias"]), eps=1e-5))

    for i in range(cfg.num_layers):
        lp = f"{pre}encoder.layers.{i}."
        y = ln(x, lp + "layer_norm1")
        q = _heads(_linear(m, y, sd, lp + "self_attn.q_proj"), H)
        k = _heads(_linear(m, y, sd, lp + "self_attn.k_proj"), H)
        v = _heads(_linear(m, y, sd, lp + "self_attn.v_proj"), H)
        a = sdpa(m, q, k, v, (D // H) ** -0.5, mask).transpose(1, 2).reshape(B, N, D)
        x = m.r(_linear(m, a, sd, lp + "self_attn.out_proj") + x)
        y = _linear(m, ln(x, lp + "layer_norm2"), sd, lp + "mlp.fc1")
        if cfg.hidden_act == "quick_gelu":
            y = m.r(y * torch.sigmoid(1.702 * y))
        else:
            y = m.r(F.gelu(y))
        x = m.r(_linear(m, y, sd, lp + "mlp.fc2") + x)
    x = ln(x, pre + "final_layer_norm")
    return x[torch.arange(B), eos], x


# --------------------------------------------------------------------------------------------
# end-to-end restatement used by tests and the CPU baseline
# --------------------------------------------------------------------------------------------
def generate_images(flow_sd, ae_sd, p: FluxParams, ap: AutoEncoderParams, x_T_nhwc: Tensor, txt: Tensor,
                    vec: Tensor, num_steps: int, guidance: float, schnell: bool,
                    mode: Mode = FP32) -> Tuple[List[Tensor], Tensor]:
    """flux/flux.py:128-193 with caller-supplied prior and conditioning.
    Returns (latents per step, images [B, H, W, 3] in [0,1])."""
    B, h, w, _ = x_T_nhwc.shape
    x_T, x_ids = prepare_latent_images(x_T_nhwc)
    txt_ids = torch.zeros((B, txt.shape[1], 3), dtype=torch.int32)
    lat = denoise(flow_sd, p, x_T, x_ids, txt, txt_ids, vec, num_steps, guidance, schnell, mode)
    img = decode(ae_sd, ap, lat[-1], (h, w), mode)
    return lat, img
