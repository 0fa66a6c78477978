"""FluxPipeline on B200 (reference: flux/flux.py:22-193).

Same constructor, attributes and generator protocol as the reference: ``generate_latents`` first
yields the conditioning 5-tuple ``(x_T, x_ids, txt, txt_ids, vec)`` and then exactly ``num_steps``
latents ``[B, L, 64]``; ``decode`` maps packed latents to ``[k, 8h, 8w, 3]`` floats in [0, 1].
Tensors are torch CUDA tensors (bf16 latents, fp32 images); the callers' ``mx.eval(x)`` has no
equivalent -- kernels are enqueued on the current CUDA stream.

Differences that follow from the platform, not from choice:
  * prior noise: MLX's threefry stream is not reproducible offline; noise is keyed by
    (seed, global image index) so sharding a batch over G GPUs never changes an image.  Pass
    ``x_T=`` to supply the prior explicitly (parity tests do).
  * T5 / CLIP outputs are cached per prompt (north star: "run once per prompt and cached").
  * LoRA adapters (linear_to_lora_layers / fuse_lora_layers, flux/lora.py) are always run FUSED into the weights;
    training_loss evaluates the objective's forward value (no backward pass: these are inference kernels).
"""
from __future__ import annotations

import os
from collections import OrderedDict
from typing import Optional, Tuple

import torch

from . import ops
from .sampler import FluxSampler
from .utils import (load_ae, load_clip, load_clip_tokenizer, load_flow_model, load_t5, load_t5_tokenizer)

bf16 = torch.bfloat16


def _lru_put(cache: OrderedDict, key, value, limit: int):
    cache[key] = value
    cache.move_to_end(key)
    while len(cache) > limit:
        cache.popitem(last=False)
    return value


def _draw_seed(device) -> int:
    """A fresh 63-bit seed for a call without `seed=` (the reference's global MLX PRNG is simply not reseeded then,
    flux/flux.py:138-139: every call draws new noise).  With several ranks, rank 0's draw is broadcast so the shards of
    one batch stay distinct (the prior is keyed by (seed, global image index)) but consistent."""
    import torch.distributed as dist
    seed = int.from_bytes(os.urandom(8), "little") >> 1
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        t = torch.tensor([seed], dtype=torch.int64, device=device if dist.get_backend() == "nccl" else "cpu")
        dist.broadcast(t, src=0)
        seed = int(t.item())
    return seed


class FluxPipeline:
    COND_CACHE_PROMPTS = 16   # T5 / CLIP outputs kept (LRU)
    _b200_native = True  # callers that also accept reference-style pipelines (flux_app.py) key the GPU fast path on this

    def __init__(self, name: str, t5_padding: bool = True, synthetic: Optional[bool] = None,
                 device: Optional[str] = None, flow_params=None, ae_params=None, t5_config=None, clip_config=None,
                 first_image_index: int = 0):
        self.dtype = bf16
        self.name = name
        self.t5_padding = t5_padding
        self.device = torch.device(device or "cuda")
        self._synthetic = synthetic
        self._t5_config, self._clip_config = t5_config, clip_config
        self.first_image_index = first_image_index  # global index of this rank's first image (batch sharding)

        kw = dict(synthetic=synthetic, device=device)
        self.ae = load_ae(name, params=ae_params, **kw)
        self.flow = load_flow_model(name, params=flow_params, **kw)
        self.clip = load_clip(name, config=clip_config, **kw)
        self.clip_tokenizer = load_clip_tokenizer(name, synthetic=synthetic,
                                                  vocab_size=self.clip.config.vocab_size)
        self.t5 = load_t5(name, config=t5_config, **kw)
        self.t5_tokenizer = load_t5_tokenizer(name, synthetic=synthetic,
                                              vocab_size=min(self.t5.config.vocab_size, 32100))
        self.sampler = FluxSampler(name)
        # prompt cache (north star: T5 / CLIP "run once per prompt and cached"): a bounded LRU -- a long-running server
        # sees an unbounded stream of prompts (4 MB of T5 output each)
        self._cond_cache: "OrderedDict[tuple, Tuple[torch.Tensor, torch.Tensor]]" = OrderedDict()
        self._bcast_cache: "OrderedDict[tuple, tuple]" = OrderedDict()
        self._ids_cache: "OrderedDict[tuple, torch.Tensor]" = OrderedDict()
        self._txt_ids_cache: "OrderedDict[tuple, torch.Tensor]" = OrderedDict()
        self.use_graph = os.environ.get("FLUX_B200_GRAPH", "1") not in ("0", "")

    def ensure_models_are_loaded(self):
        torch.cuda.synchronize(self.device)

    def reload_text_encoders(self):
        kw = dict(synthetic=self._synthetic, device=str(self.device))
        self.t5 = load_t5(self.name, config=self._t5_config, **kw)
        self.clip = load_clip(self.name, config=self._clip_config, **kw)
        self._cond_cache.clear()  # outputs of the previous encoders
        self._bcast_cache.clear()

    def tokenize(self, text):
        t5_tokens = self.t5_tokenizer.encode(text, pad=self.t5_padding)
        clip_tokens = self.clip_tokenizer.encode(text)
        return t5_tokens, clip_tokens

    # ------------------------------------------------------------------ flux/flux.py:53-85
    def _prepare_latent_images(self, x: torch.Tensor):
        b, h, w, c = x.shape
        packed = ops.patchify(x.to(device=self.device, dtype=bf16))
        return packed, self._latent_ids(b, h, w)

    def _latent_ids(self, b: int, h: int, w: int) -> torch.Tensor:
        x_ids = self._ids_cache.get((b, h, w))
        if x_ids is None:
            i = torch.zeros((h // 2, w // 2), dtype=torch.int32)
            j, k = torch.meshgrid(torch.arange(h // 2, dtype=torch.int32), torch.arange(w // 2, dtype=torch.int32),
                                  indexing="ij")
            x_ids = torch.stack([i, j, k], dim=-1).reshape(1, h * w // 4, 3).repeat(b, 1, 1).to(self.device)
            _lru_put(self._ids_cache, (b, h, w), x_ids, 4)
        return x_ids

    def _prepare_conditioning(self, n_images, t5_tokens, clip_tokens):
        key = (tuple(t5_tokens.flatten().tolist()), tuple(clip_tokens.flatten().tolist()), tuple(t5_tokens.shape))
        hit = self._cond_cache.get(key)
        if hit is None:
            if not hasattr(self, "t5") or not hasattr(self, "clip"):
                raise RuntimeError("text encoders were deleted; call reload_text_encoders()")
            txt1 = self.t5(t5_tokens)
            vec1 = self.clip(clip_tokens).pooled_output
            hit = _lru_put(self._cond_cache, key, (txt1, vec1), self.COND_CACHE_PROMPTS)
        else:
            self._cond_cache.move_to_end(key)
        # the broadcast copies are cached too (the flow model keys its txt_in cache on the tensor)
        bkey = (key, n_images)
        bhit = self._bcast_cache.get(bkey)
        if bhit is None:
            txt, vec = hit
            if len(txt) == 1 and n_images > 1:
                txt = txt.expand(n_images, *txt.shape[1:]).contiguous()
            single = len(vec) == 1
            if single and n_images > 1:
                vec = vec.expand(n_images, *vec.shape[1:]).contiguous()
            if single:  # rows known identical without asking the device (see _denoising_loop)
                self._uniform_vecs = {vec.data_ptr()} | {p for p in getattr(self, "_uniform_vecs", set())
                                                         if any(p == b[2].data_ptr() for b in self._bcast_cache.values())}
            bhit = _lru_put(self._bcast_cache, bkey, (txt, self._txt_ids(n_images, txt.shape[1]), vec), 2)
        return bhit

    def _txt_ids(self, n_images: int, S: int) -> torch.Tensor:
        """All-zero text position ids (flux/flux.py:84): ONE tensor per (batch, S), so the flow model's RoPE table and
        CUDA graph (keyed on the id tensors) survive a change of prompt."""
        ids = self._txt_ids_cache.get((n_images, S))
        if ids is None:
            ids = _lru_put(self._txt_ids_cache, (n_images, S),
                           torch.zeros((n_images, S, 3), dtype=torch.int32, device=self.device), 4)
        return ids

    # ------------------------------------------------------------------ flux/flux.py:87-126
    def _denoising_loop(self, x_t, x_ids, txt, txt_ids, vec, num_steps: int = 35, guidance: float = 4.0,
                        start: float = 1, stop: float = 0):
        B = len(x_t)

        def scalar(x):
            return torch.full((B,), x, dtype=self.dtype, device=self.device)

        guidance_value = guidance
        guidance = scalar(guidance)
        timesteps = self.sampler.timesteps(num_steps, x_t.shape[1], start=start, stop=stop)
        # One prompt for the whole batch (what txt2image.py / flux_app.py do): (t, y, guidance) are the same for every
        # row and none depends on x_t, so the modulation of EVERY step is computed up front in one pass over the
        # modulation weights.  A list of prompts (one per image, which tokenize() accepts like the reference's does)
        # gives per-row CLIP vectors: then every row needs its own modulation and the per-row path runs.
        uniform = len(vec) == 1 or vec.data_ptr() in getattr(self, "_uniform_vecs", ()) or bool((vec == vec[:1]).all())
        table = None
        if uniform:
            table = self.flow.conditioning_table(timesteps[:num_steps], vec[:1],
                                                 guidance_value if self.flow.params.guidance_embed else None)
        x_t = x_t.clone()
        for i in range(num_steps):
            t = timesteps[i]
            t_prev = timesteps[i + 1]
            fwd = self.flow.forward_graphed if self.use_graph else self.flow.forward
            pred = fwd(img=x_t, img_ids=x_ids, txt=txt, txt_ids=txt_ids, y=vec, timesteps=scalar(t), guidance=guidance,
                       uniform=uniform, mod_row=None if table is None else table[i])
            x_t = ops.euler_step(x_t.clone(), pred, t_prev - t)  # sampler.step (flux/sampler.py:56-57)
            yield x_t

    def generate_latents(self, text: str, n_images: int = 1, num_steps: int = 35, guidance: float = 4.0,
                         latent_size: Tuple[int, int] = (64, 64), seed=None, x_T: Optional[torch.Tensor] = None,
                         x_T_packed: Optional[torch.Tensor] = None):
        """flux/flux.py:128-155.  `text` may be a list with one prompt per image.  x_T: an NHWC prior [B, h, w, 16] (the
        reference's layout); x_T_packed: the same prior already packed [B, L, 64] (flux_serve.py builds coalesced batches
        from per-request priors with ops.prior_packed)."""
        if x_T_packed is not None:
            h, w = latent_size
            x_T = x_T_packed.to(device=self.device, dtype=bf16)
            if x_T.shape != (n_images, h * w // 4, 64):
                raise ValueError(f"x_T_packed has shape {tuple(x_T.shape)}, expected {(n_images, h * w // 4, 64)}")
            x_ids = self._latent_ids(n_images, h, w)
        elif x_T is None:
            # prior drawn on the device straight into the packed layout (sample_prior + patchify fused)
            h, w = latent_size
            if h % 2 or w % 2:
                raise ValueError(f"latent size {latent_size} must be even")
            x_T = ops.prior_packed(n_images, latent_size, 16, _draw_seed(self.device) if seed is None else seed,
                                   self.first_image_index, device=self.device)
            x_ids = self._latent_ids(n_images, h, w)
        else:
            x_T, x_ids = self._prepare_latent_images(x_T)
        t5_tokens, clip_tokens = self.tokenize(text)
        txt, txt_ids, vec = self._prepare_conditioning(n_images, t5_tokens, clip_tokens)
        yield (x_T, x_ids, txt, txt_ids, vec)
        yield from self._denoising_loop(x_T, x_ids, txt, txt_ids, vec, num_steps=num_steps, guidance=guidance)

    def decode(self, x: torch.Tensor, latent_size: Tuple[int, int] = (64, 64)) -> torch.Tensor:
        """flux/flux.py:157-162 -> float32 [k, 8h, 8w, 3] in [0, 1]."""
        img, _ = self.ae.decode_packed(x, latent_size, want_u8=False)
        return img

    def decode_uint8(self, x: torch.Tensor, latent_size: Tuple[int, int] = (64, 64)) -> torch.Tensor:
        """decode + the CLI's truncating `(x * 255).astype(uint8)` (txt2image.py:133) in one pass."""
        _, u8 = self.ae.decode_packed(x, latent_size, want_u8=True)
        return u8

    def generate_images(self, text: str, n_images: int = 1, num_steps: int = 35, guidance: float = 4.0,
                        latent_size: Tuple[int, int] = (64, 64), seed=None, reload_text_encoders: bool = True,
                        progress: bool = True, x_T: Optional[torch.Tensor] = None, decoding_batch_size: int = 1):
        """flux/flux.py:164-193.  `reload_text_encoders` is accepted for signature parity; the encoders
        are never deleted here, so there is nothing to reload."""
        from tqdm import tqdm
        latents = self.generate_latents(text, n_images, num_steps, guidance, latent_size, seed, x_T=x_T)
        next(latents)
        x_t = None
        for x_t in tqdm(latents, total=num_steps, disable=not progress, leave=True):
            pass
        images = []
        for i in tqdm(range(0, len(x_t), decoding_batch_size), disable=not progress, desc="generate images"):
            images.append(self.decode(x_t[i:i + decoding_batch_size], latent_size))
        return torch.cat(images, dim=0)

    # ------------------------------------------------------------------ flux/flux.py:195-226
    def training_loss(self, x_0: torch.Tensor, t5_features: torch.Tensor, clip_features: torch.Tensor,
                      guidance: torch.Tensor, t: Optional[torch.Tensor] = None, eps: Optional[torch.Tensor] = None):
        """The rectified-flow training objective, FORWARD value only: mean((pred + x_0 - eps)^2) with
        x_t = (1 - t) x_0 + t eps (flux/flux.py:195-226).  x_0 are VAE latents [B, h, w, 16] (AutoEncoder.encode),
        t5_features / clip_features the cached text-encoder outputs, guidance [B].  `t` [B] and `eps` (packed, like the
        patchified x_0) may be supplied for reproducibility; otherwise they are drawn as the reference draws them.
        The kernels are inference kernels: there is no backward pass, so this evaluates / validates a loss (e.g. of a
        loaded adapter), it does not train one -- optimisation stays with the reference's MLX trainer (out of scope)."""
        txt = t5_features.to(device=self.device, dtype=bf16)
        vec = clip_features.to(device=self.device, dtype=bf16)
        txt_ids = self._txt_ids(txt.shape[0], txt.shape[1])
        x_0, x_ids = self._prepare_latent_images(x_0)
        B, L, _ = x_0.shape
        if t is None:
            t = self.sampler.random_timesteps(B, L)
        t = t.to(device=self.device, dtype=self.dtype)
        if eps is None:
            eps = torch.randn(x_0.shape, device=self.device, dtype=torch.float32)
        eps = eps.to(device=self.device, dtype=self.dtype)
        x_t = self.sampler.add_noise(x_0, t, noise=eps)
        pred = self.flow(img=x_t, img_ids=x_ids, txt=txt, txt_ids=txt_ids, y=vec, timesteps=t,
                         guidance=guidance.to(device=self.device, dtype=self.dtype))
        return (pred + x_0 - eps).square().mean()

    # ------------------------------------------------------------------ LoRA adapters at inference
    def linear_to_lora_layers(self, rank: int = 8, num_blocks: int = -1):
        """flux/flux.py:228-236: prepare the Linears of the last `num_blocks` blocks for an adapter of rank `rank`
        (`flow.load_weights(adapter, strict=False)` then stages its lora_a / lora_b entries)."""
        self.flow.enable_lora(rank, num_blocks)

    def fuse_lora_layers(self):
        """flux/flux.py:238-246: fold the staged adapter into the flow weights."""
        self.flow.fuse_lora()
