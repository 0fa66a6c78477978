"""Flux VAE decoder on B200 (reference: flux/autoencoder.py:24-124, 212-297, 311-357).

``AutoEncoder.decode(z)`` with the reference's NHWC convention.  3x3 convolutions run as implicit
GEMMs on tcgen05 (4-D TMA boxes over the NHWC activation, OHWI weights = the reference's sanitized
layout), GroupNorm(32)+SiLU as a two-pass HBM-bound kernel pair, the single-head 512-wide mid
attention as two tcgen05 GEMMs around a row softmax.  Activations are bf16, accumulation fp32 (the
reference promotes to the AE file's dtype; tolerance stated in tests/test_gpu_parity.py::test_vae_decode_vs_golden).
The decoder is the hot path.  The encoder (flux/autoencoder.py:127-209,347-350; SURVEY 8-f N4: the step before the path
in DreamBooth training) runs on the same kernels: its stride-2 "Downsample" convolutions (pad (0,1,0,1), stride 2,
flux/autoencoder.py:101-113) are the stride-1 convolution sampled at the odd pixels -- 4x the arithmetic of a native
stride-2 kernel, accepted because encoding happens once per training image, never in the denoising loop.
"""
from __future__ import annotations

import os
from typing import Dict, Optional

import torch

from . import ops
from .model import WeightArena
from .specs import AutoEncoderParams, ae_decoder_manifest, ae_encoder_manifest

bf16 = torch.bfloat16


class AutoEncoder:
    def __init__(self, params: AutoEncoderParams, device: Optional[str] = None):
        self.params = params
        self.scale_factor = params.scale_factor
        self.shift_factor = params.shift_factor
        self.device = torch.device(device or "cuda")
        self._manifest = ae_decoder_manifest(params)
        self._enc_manifest = ae_encoder_manifest(params)
        self._shapes = {k: s for k, s, _ in self._manifest + self._enc_manifest}
        self._dec_keys = [k for k, _, _ in self._manifest]
        self._enc_keys = [k for k, _, _ in self._enc_manifest]
        entries = []
        for k, s, _ in self._manifest + self._enc_manifest:
            entries.append((k, self._stored_shape(k, s)))
        # fused q|k|v projection of the mid attention blocks
        c = params.ch * params.ch_mult[-1]
        for side in ("decoder", "encoder"):
            entries += [(f"{side}.mid.attn_1.qkv.weight", (3 * c, c)), (f"{side}.mid.attn_1.qkv.bias", (3 * c,))]
        # parity-decomposed kernels of the Upsample convolutions (ops.upconv_weights): [4 * C, 4 * C]
        self._up_keys = [k[:-len(".weight")] for k in self._dec_keys if k.endswith(".upsample.conv.weight")]
        for k in self._up_keys:
            C_ = self._shapes[k + ".weight"][0]
            if C_ % 128 == 0:
                entries.append((k + ".weight4", (4 * C_, 4 * C_)))
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
                if strict and key.startswith(("decoder.", "encoder.")):
                    raise ValueError(f"Received parameters not in model: {key}")
                continue
            w = w.to(device=self.device, dtype=bf16)
            if w.ndim == 4:  # OHWI from sanitize
                o, kh, kw, i = w.shape
                if i < 64:
                    w = torch.nn.functional.pad(w, (0, 64 - i))
                w = w.reshape(o, -1)
            self.arena[key].copy_(w.reshape(self.arena[key].shape))
            seen.add(key)
        if strict:  # the decoder is mandatory; the encoder half of ae.safetensors is optional (decode-only deployments)
            missing = [k for k in self._dec_keys if k not in seen]
            if missing:
                raise ValueError(f"Missing {len(missing)} parameters, e.g. {missing[:3]}")
        if all(k in seen for k in self._dec_keys):
            for k in self._up_keys:
                if (k + ".weight4") in self.arena:
                    self.arena[k + ".weight4"].copy_(ops.upconv_weights(self.arena[k + ".weight"]))
        for side, keys in (("decoder", self._dec_keys), ("encoder", self._enc_keys)):
            if all(k in seen for k in keys):  # derived tensors: fused q|k|v projection
                pre = side + ".mid.attn_1."
                self.arena[pre + "qkv.weight"].copy_(torch.cat([self.arena[pre + n + ".weight"] for n in "qkv"], 0))
                self.arena[pre + "qkv.bias"].copy_(torch.cat([self.arena[pre + n + ".bias"] for n in "qkv"], 0))
        return self

    @property
    def has_encoder(self) -> bool:
        return all(k in self.arena.loaded for k in self._enc_keys)

    def parameters(self):
        return {"arena": self.arena.buffer}

    # ------------------------------------------------------------------ blocks
    # Every GroupNorm input except the one after the mid attention is the output of a 3x3 convolution: that
    # convolution's epilogue accumulates the GroupNorm partial sums (ops.conv3x3(gn_stats=True)), so the statistics
    # pass over the activation (a third of the GroupNorm time) disappears.  `st` carries (partials, blocks per image).
    FUSE_GN_STATS = os.environ.get("FLUX_B200_GN_FUSED", "1") not in ("0", "")
    # Upsample (nearest 2x + 3x3 conv, flux/autoencoder.py:121-124) as four parity-wise 2x2 convolutions of the
    # low-resolution tensor (fx_conv3x3 upsample2x): 16/36 of the MACs, no upsampled tensor in HBM
    FUSE_UPSAMPLE = os.environ.get("FLUX_B200_UPCONV", "1") not in ("0", "")

    def _upconv(self, x, key):
        if self.FUSE_UPSAMPLE and (key + ".weight4") in self.arena:
            stats = self.FUSE_GN_STATS
            r = ops.conv3x3(x, self.arena[key + ".weight4"], self.arena[key + ".bias"], gn_stats=stats, upsample=True)
            return r if stats else (r, None)
        return self._conv(ops.upsample2x(x), key, stats=True)

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
                h, st = self._upconv(h, f"decoder.up.{lvl}.upsample.conv")
        h = self._gn(h, "decoder.norm_out", True, st)
        return self._conv(h, "decoder.conv_out", out_dtype=torch.float32)[0]

    # ------------------------------------------------------------------ encode (training side)
    def encode(self, x: torch.Tensor, noise: torch.Tensor = None) -> torch.Tensor:
        """AutoEncoder.encode (flux/autoencoder.py:347-350): x [B, H, W, 3] NHWC -> z [B, H/8, W/8, 16] bf16,
        z = scale_factor * (reg(encoder(x)) - shift_factor).  DiagonalGaussian (flux/autoencoder.py:300-309) returns the
        mean in eval mode; pass `noise` (standard normal, shape of z) for the training-mode sample mean + exp(logvar/2) * eps."""
        if not self.has_encoder:
            raise RuntimeError("encoder weights were not loaded (ae.safetensors `encoder.*`)")
        a = self.params
        B, H, W, c = x.shape
        if H % 8 or W % 8:
            raise ValueError(f"image size {(H, W)} must be a multiple of 8")
        xp = torch.zeros((B, H, W, 64), device=self.device, dtype=bf16)  # conv_in reads one 64-channel block
        xp[..., :c] = x.to(device=self.device, dtype=bf16)
        h, st = self._conv(xp, "encoder.conv_in", stats=True)
        n = len(a.ch_mult)
        for lvl in range(n):
            for blk in range(a.num_res_blocks):
                h, st = self._resnet(h, f"encoder.down.{lvl}.block.{blk}", st)
            if lvl != n - 1:
                # Downsample: pad (0,1,0,1) + stride-2 conv == the pad-1 stride-1 conv at the odd pixels
                full, _ = self._conv(h, f"encoder.down.{lvl}.downsample.conv")
                h, st = full[:, 1::2, 1::2].contiguous(), None
        h, st = self._resnet(h, "encoder.mid.block_1", st)
        h = self._attn(h, "encoder.mid.attn_1", st)
        h, st = self._resnet(h, "encoder.mid.block_2")
        h = self._gn(h, "encoder.norm_out", True, st)
        moments = self._conv(h, "encoder.conv_out", out_dtype=torch.float32)[0]
        mean, logvar = moments[..., :a.z_channels], moments[..., a.z_channels:]
        z = mean if noise is None else mean + torch.exp(0.5 * logvar) * noise.to(mean)
        return (a.scale_factor * (z - a.shift_factor)).to(bf16)

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        """flux/autoencoder.py:356-357"""
        return self.decode(self.encode(x))
