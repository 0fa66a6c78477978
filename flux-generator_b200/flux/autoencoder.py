"""Flux VAE decoder on B200 (reference: flux/autoencoder.py:24-124, 212-297, 311-357).

``AutoEncoder.decode(z)`` with the reference's NHWC convention.  3x3 convolutions run as implicit
GEMMs on tcgen05 (4-D TMA boxes over the NHWC activation, OHWI weights = the reference's sanitized
layout), GroupNorm(32)+SiLU as a two-pass HBM-bound kernel pair, the single-head 512-wide mid
attention as two tcgen05 GEMMs around a row softmax.  Activations are bf16, accumulation fp32 (the
reference promotes to the AE file's dtype; tolerance stated in tests/test_gpu_parity.py::test_vae_decode_vs_golden).
Only the decoder is on the hot path; the encoder is training-only and not built.
"""
from __future__ import annotations

import os
from typing import Dict, Optional

import torch

from . import ops
from .model import WeightArena
from .specs import AutoEncoderParams, ae_decoder_manifest

bf16 = torch.bfloat16


class AutoEncoder:
    def __init__(self, params: AutoEncoderParams, device: Optional[str] = None):
        self.params = params
        self.scale_factor = params.scale_factor
        self.shift_factor = params.shift_factor
        self.device = torch.device(device or "cuda")
        self._manifest = ae_decoder_manifest(params)
        self._shapes = {k: s for k, s, _ in self._manifest}
        entries = []
        for k, s, _ in self._manifest:
            entries.append((k, self._stored_shape(k, s)))
        # fused q|k|v projection of the mid attention block
        c = params.ch * params.ch_mult[-1]
        entries += [("decoder.mid.attn_1.qkv.weight", (3 * c, c)), ("decoder.mid.attn_1.qkv.bias", (3 * c,))]
        self.arena = WeightArena(entries, self.device)

    def _stored_shape(self, key, shape):
        if len(shape) == 4:
            o, i, kh, kw = shape
            if kh == 1:
                return (o, i)
            return (o, kh * kw * max(i, 64))  # conv_in's 16 input channels are zero-padded to one 64-block
        return shape

    def sanitize(self, weights: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        """OIHW -> OHWI, 1x1 squeezed to Linear (flux/autoencoder.py:336-345); then flattened [O, kh*kw*I]."""
        out = {}
        for k, w in weights.items():
            if w.ndim == 4:
                w = w.permute(0, 2, 3, 1)
                if w.shape[1:3] == (1, 1):
                    w = w.reshape(w.shape[0], w.shape[3])
            out[k] = w
        return out

    def load_weights(self, weights, strict: bool = True) -> "AutoEncoder":
        items = list(weights.items()) if isinstance(weights, dict) else list(weights)
        seen = self.arena.loaded
        for key, w in items:
            if key not in self._shapes:
                if strict and key.startswith("decoder."):
                    raise ValueError(f"Received parameters not in model: {key}")
                continue  # encoder.* tensors of ae.safetensors are not on the decode path
            w = w.to(device=self.device, dtype=bf16)
            if w.ndim == 4:  # OHWI from sanitize
                o, kh, kw, i = w.shape
                if i < 64:
                    w = torch.nn.functional.pad(w, (0, 64 - i))
                w = w.reshape(o, -1)
            self.arena[key].copy_(w.reshape(self.arena[key].shape))
            seen.add(key)
        if strict:
            missing = [k for k in self._shapes if k not in seen]
            if missing:
                raise ValueError(f"Missing {len(missing)} parameters, e.g. {missing[:3]}")
        if len(seen) == len(self._shapes):  # derived tensors: fused q|k|v projection
            pre = "decoder.mid.attn_1."
            self.arena[pre + "qkv.weight"].copy_(torch.cat([self.arena[pre + n + ".weight"] for n in "qkv"], 0))
            self.arena[pre + "qkv.bias"].copy_(torch.cat([self.arena[pre + n + ".bias"] for n in "qkv"], 0))
        return self

    def parameters(self):
        return {"arena": self.arena.buffer}

    # ------------------------------------------------------------------ blocks
    # Every GroupNorm input except the one after the mid attention is the output of a 3x3 convolution: that
    # convolution's epilogue accumulates the GroupNorm partial sums (ops.conv3x3(gn_stats=True)), so the statistics
    # pass over the activation (a third of the GroupNorm time) disappears.  `st` carries (partials, blocks per image).
    FUSE_GN_STATS = os.environ.get("FLUX_B200_GN_FUSED", "1") not in ("0", "")

    def _gn(self, x, key, silu, st=None):
        return ops.groupnorm(x, self.arena[key + ".weight"], self.arena[key + ".bias"], 1e-6, silu, partials=st)

    def _conv(self, x, key, resid=None, out_dtype=bf16, stats=False):
        w = self.arena[key + ".weight"]
        stats = stats and self.FUSE_GN_STATS and w.shape[0] % 128 == 0
        r = ops.conv3x3(x, w, self.arena[key + ".bias"], resid=resid, out_dtype=out_dtype, gn_stats=stats)
        return r if stats else (r, None)

    def _resnet(self, x, key, st=None):
        # flux/autoencoder.py:85-98
        h, hst = self._conv(self._gn(x, key + ".norm1", True, st), key + ".conv1", stats=True)
        h = self._gn(h, key + ".norm2", True, hst)
        if (key + ".nin_shortcut.weight") in self.arena:
            B, H, W, C = x.shape
            sk = key + ".nin_shortcut"
            x = ops.gemm(x.view(B, H * W, C), self.arena[sk + ".weight"], self.arena[sk + ".bias"]).view(B, H, W, -1)
        return self._conv(h, key + ".conv2", resid=x, stats=True)

    def _attn(self, x, key, st=None):
        # flux/autoencoder.py:41-52: single head, scale C^-0.5
        B, H, W, C = x.shape
        n = H * W
        y = self._gn(x, key + ".norm", False, st).view(B, n, C)
        qkv = ops.gemm(y, self.arena[key + ".qkv.weight"], self.arena[key + ".qkv.bias"])  # [B, n, 3C]
        o = torch.empty((B, n, C), device=x.device, dtype=bf16)
        s = torch.empty((n, n), device=x.device, dtype=torch.float32)
        p = torch.empty((n, n), device=x.device, dtype=bf16)
        for b in range(B):
            q, k, v = qkv[b, :, :C], qkv[b, :, C:2 * C], qkv[b, :, 2 * C:]
            ops.gemm(q, k, out=s)                       # S = q k^T (fp32)
            ops.softmax_rows(s, C ** -0.5, out=p)       # P = softmax(scale * S)
            ops.gemm(p, ops.transpose(v), out=o[b])     # O = P v
        pk = key + ".proj_out"
        xr = x.view(B, n, C)
        return ops.gemm(o, self.arena[pk + ".weight"], self.arena[pk + ".bias"], resid=xr).view(B, H, W, C)

    # ------------------------------------------------------------------ decode
    def decode_packed(self, packed: torch.Tensor, latent_size, want_u8: bool = True):
        """packed latents [k, L, 64] -> (image float32 [k, 8h, 8w, 3] in [0,1], uint8 copy).
        Fuses FluxPipeline.decode's unpatchify, AutoEncoder.decode's affine (flux/autoencoder.py:353),
        the decoder, and clip(x+1,0,2)*0.5 (flux/flux.py:162)."""
        a = self.params
        z = ops.unpatchify_scale(packed.to(bf16), latent_size, 64, a.scale_factor, a.shift_factor)
        return ops.finish_image(self._decoder(z), want_u8)

    def decode(self, z: torch.Tensor) -> torch.Tensor:
        """AutoEncoder.decode (flux/autoencoder.py:352-354): z [B, h, w, 16] NHWC -> [B, 8h, 8w, 3] float32."""
        a = self.params
        B, h, w, c = z.shape
        zz = (z.to(bf16) / a.scale_factor + a.shift_factor).to(bf16)
        zp = torch.zeros((B, h, w, 64), device=self.device, dtype=bf16)
        zp[..., :c] = zz
        return self._decoder(zp)

    def _decoder(self, z: torch.Tensor) -> torch.Tensor:
        # flux/autoencoder.py:271-297
        a = self.params
        h, st = self._conv(z, "decoder.conv_in", stats=True)
        h, st = self._resnet(h, "decoder.mid.block_1", st)
        h = self._attn(h, "decoder.mid.attn_1", st)
        h, st = self._resnet(h, "decoder.mid.block_2")  # (its input comes from a GEMM epilogue: standalone statistics)
        for lvl in reversed(range(len(a.ch_mult))):
            for blk in range(a.num_res_blocks + 1):
                h, st = self._resnet(h, f"decoder.up.{lvl}.block.{blk}", st)
            if lvl != 0:
                h, st = self._conv(ops.upsample2x(h), f"decoder.up.{lvl}.upsample.conv", stats=True)
        h = self._gn(h, "decoder.norm_out", True, st)
        return self._conv(h, "decoder.conv_out", out_dtype=torch.float32)[0]
