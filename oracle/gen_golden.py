#!/usr/bin/env python
"""Writes tests/golden/*.npz by executing the UNMODIFIED reference package (/root/reference/flux,
txt2image.py) over the MLX-API shim in oracle/mlx_shim -- TEST INFRASTRUCTURE ONLY.

Run in the build container only (``python oracle/gen_golden.py``): /root/reference does not exist
on the GPU box, so the fixtures are committed.  What a fixture pins: the reference's own module
structure, op order, split / concat / reshape conventions, constants and sanitizers, evaluated in
fp32 on bf16-representable synthetic weights.  What it cannot pin: the arithmetic inside MLX's
kernels (shim restates it from MLX's docs) and MLX's RNG stream (x_T is stored in the fixture).

Weights are not stored: they are re-drawn from flux.synthetic (seeded per key); the fixture keeps
a CRC of the state dict so a replay detects generator drift.
"""
from __future__ import annotations

import importlib.util
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")


def load_product_pkg():
    """Product package under an alias (its import name `flux` collides with the reference's)."""
    path = os.path.join(ROOT, "flux-generator_b200", "flux")
    pkg = types.ModuleType("fluxb200")
    pkg.__path__ = [path]
    sys.modules["fluxb200"] = pkg
    mods = {}
    for name in ("specs", "synthetic"):
        spec = importlib.util.spec_from_file_location(f"fluxb200.{name}", os.path.join(path, name + ".py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[f"fluxb200.{name}"] = mod
        spec.loader.exec_module(mod)
        mods[name] = mod
    return mods["specs"], mods["synthetic"]


specs, synthetic = load_product_pkg()
sys.path.insert(0, os.path.join(HERE, "mlx_shim"))
sys.path.insert(0, REF)
import mlx.core as mx  # noqa: E402  (the shim)
import flux as ref  # noqa: E402  (the reference package)
import flux.flux as ref_flux  # noqa: E402
from flux.autoencoder import AutoEncoder, AutoEncoderParams as RefAEParams  # noqa: E402
from flux.clip import CLIPTextModel, CLIPTextModelConfig as RefCLIPConfig  # noqa: E402
from flux.model import Flux, FluxParams as RefFluxParams  # noqa: E402
from flux.sampler import FluxSampler  # noqa: E402
from flux.t5 import T5Config as RefT5Config, T5Encoder  # noqa: E402
import txt2image as ref_cli  # noqa: E402

# ------------------------------------------------------------------ small configurations
SMALL_FLOW = dict(in_channels=64, vec_in_dim=128, context_in_dim=256, hidden_size=256, mlp_ratio=4.0,
                  num_heads=2, depth=2, depth_single_blocks=2, axes_dim=[16, 56, 56], theta=10_000,
                  qkv_bias=True)
SMALL_AE = dict(resolution=64, in_channels=3, ch=64, out_ch=3, ch_mult=[1, 2, 4, 4], num_res_blocks=2,
                z_channels=16, scale_factor=0.3611, shift_factor=0.1159)
SMALL_T5 = dict(vocab_size=1024, num_layers=2, num_heads=4, relative_attention_num_buckets=32, d_kv=64,
                d_model=256, d_ff=512, feed_forward_proj="gated-gelu", tie_word_embeddings=False,
                relative_attention_max_distance=128, layer_norm_epsilon=1e-6)
SMALL_CLIP = dict(num_layers=2, model_dims=128, num_heads=2, max_length=77, vocab_size=1024,
                  hidden_act="quick_gelu")


def f32(sd):
    return {k: mx.array(v.to(torch.float32)) for k, v in sd.items()}


def npy(x):
    if isinstance(x, torch.Tensor):
        x = x.detach()
        return x.to(torch.float32).numpy() if x.is_floating_point() else x.numpy()
    return np.asarray(x)


def build_flow(guidance_embed: bool):
    sp = specs.FluxParams(**SMALL_FLOW, guidance_embed=guidance_embed)
    sd = synthetic.synthetic_state_dict(specs.flow_manifest(sp))
    model = Flux(RefFluxParams(**SMALL_FLOW, guidance_embed=guidance_embed))
    model.load_weights(list(model.sanitize(f32(sd)).items()))
    return model, sd


def build_ae():
    ap = specs.AutoEncoderParams(**SMALL_AE)
    sd = synthetic.synthetic_state_dict(specs.ae_decoder_manifest(ap))
    ae = AutoEncoder(RefAEParams(**SMALL_AE))
    ae.load_weights(list(ae.sanitize(f32(sd)).items()), strict=False)
    return ae, sd


def build_t5():
    cfg = specs.T5Config(**SMALL_T5)
    sd = synthetic.synthetic_state_dict(specs.t5_manifest(cfg))
    t5 = T5Encoder(RefT5Config.from_dict(SMALL_T5))
    t5.load_weights(list(t5.sanitize(f32(sd)).items()))
    return t5, sd


def build_clip():
    cfg = specs.CLIPTextModelConfig(**SMALL_CLIP)
    sd = synthetic.synthetic_state_dict(specs.clip_manifest(cfg))
    hf = dict(num_hidden_layers=cfg.num_layers, hidden_size=cfg.model_dims,
              num_attention_heads=cfg.num_heads, max_position_embeddings=cfg.max_length,
              vocab_size=cfg.vocab_size, hidden_act=cfg.hidden_act)
    clip = CLIPTextModel(RefCLIPConfig.from_dict(hf))
    clip.load_weights(list(clip.sanitize(f32(sd)).items()))
    return clip, sd


class FakeTok:
    def __init__(self, ids):
        self.ids = ids

    def encode(self, text, pad=True):
        return mx.array(self.ids)


# ------------------------------------------------------------------ fixtures
def golden_schedule():
    """FluxSampler.timesteps (flux/sampler.py:22-31) for the BASELINE.json configs."""
    out = {}
    for name, steps, L in (("schnell", 2, 1024), ("schnell", 4, 4096), ("schnell", 1, 64),
                           ("dev", 50, 4096), ("dev", 50, 9216), ("dev", 50, 1024), ("dev", 28, 256)):
        s = FluxSampler("flux-" + name)
        out[f"{name}_{steps}_{L}"] = np.asarray(s.timesteps(steps, L), dtype=np.float64)
    np.savez(os.path.join(OUT, "schedule.npz"), **out)


def golden_patchify():
    """_prepare_latent_images / decode's unpatchify (flux/flux.py:53-71,157-160) and
    to_latent_size (txt2image.py:14-25)."""
    pipe = ref.FluxPipeline.__new__(ref.FluxPipeline)
    x = mx.array(torch.arange(2 * 4 * 6 * 16, dtype=torch.float32).reshape(2, 4, 6, 16))
    p, ids = pipe._prepare_latent_images(x)
    h, w = 4, 6
    back = p.reshape(len(p), h // 2, w // 2, -1, 2, 2).transpose(0, 1, 4, 2, 5, 3).reshape(len(p), h, w, -1)
    sizes = [(512, 512), (768, 512), (513, 513), (769, 769), (1024, 1024), (100, 260)]
    lat = [ref_cli.to_latent_size(s) for s in sizes]
    np.savez(os.path.join(OUT, "patchify.npz"), x=npy(x), packed=npy(p), ids=npy(ids), back=npy(back),
             sizes=np.asarray(sizes), latent_sizes=np.asarray(lat))


def golden_flow():
    """Flux.__call__ (flux/model.py:99-136) on the small config, schnell and dev variants."""
    for variant, ge in (("schnell", False), ("dev", True)):
        model, sd = build_flow(ge)
        g = torch.Generator().manual_seed(7)
        B, h, w, S = 2, 8, 12, 16
        x = mx.array(torch.randn((B, h, w, 16), generator=g).to(torch.bfloat16))
        pipe = ref.FluxPipeline.__new__(ref.FluxPipeline)
        img, img_ids = pipe._prepare_latent_images(x)
        txt = mx.array(torch.randn((B, S, SMALL_FLOW["context_in_dim"]), generator=g).to(torch.bfloat16))
        txt_ids = mx.zeros((B, S, 3), dtype=mx.int32)
        y = mx.array(torch.randn((B, SMALL_FLOW["vec_in_dim"]), generator=g).to(torch.bfloat16))
        t = mx.full((B,), 0.75, dtype=mx.bfloat16)
        gd = mx.full((B,), 3.5, dtype=mx.bfloat16)
        # float32 copies: fixture semantics = "reference code, fp32 tensors, bf16-valued inputs";
        # timesteps/guidance stay bf16 like FluxPipeline._denoising_loop's scalar()
        img32, txt32, y32 = img.astype(mx.float32), txt.astype(mx.float32), y.astype(mx.float32)
        out = model(img=img32, img_ids=img_ids, txt=txt32, txt_ids=txt_ids, timesteps=t, y=y32, guidance=gd)

        # stepwise taps through the reference's own sub-modules (same lines as model.py:112-134)
        from flux.layers import timestep_embedding
        taps = {}
        i_ = model.img_in(img32)
        vec = model.time_in(timestep_embedding(t, 256))
        if ge:
            vec = vec + model.guidance_in(timestep_embedding(gd, 256))
        vec = vec + model.vector_in(y32)
        t_ = model.txt_in(txt32)
        pe = model.pe_embedder(mx.concatenate([txt_ids, img_ids], axis=1)).astype(i_.dtype)
        taps["vec"], taps["img_in"], taps["txt_in"], taps["pe"] = vec, i_, t_, pe
        for n, blk in enumerate(model.double_blocks):
            i_, t_ = blk(img=i_, txt=t_, vec=vec, pe=pe)
            taps[f"double.{n}.img"], taps[f"double.{n}.txt"] = i_, t_
        xx = mx.concatenate([t_, i_], axis=1)
        for n, blk in enumerate(model.single_blocks):
            xx = blk(xx, vec=vec, pe=pe)
            taps[f"single.{n}"] = xx
        out2 = model.final_layer(xx[:, S:, ...], vec)
        assert torch.equal(out, out2)
        np.savez(os.path.join(OUT, f"flow_{variant}.npz"), config=json.dumps(SMALL_FLOW), guidance_embed=ge,
                 weights_crc=synthetic.state_dict_checksum(sd), x_nhwc=npy(x), img=npy(img),
                 img_ids=npy(img_ids), txt=npy(txt), txt_ids=npy(txt_ids), y=npy(y), t=0.75, guidance=3.5,
                 out=npy(out), **{"tap." + k: npy(v) for k, v in taps.items()})


def golden_ae():
    """FluxPipeline.decode -> AutoEncoder.decode (flux/flux.py:157-162, flux/autoencoder.py:352-354)."""
    ae, sd = build_ae()
    pipe = ref.FluxPipeline.__new__(ref.FluxPipeline)
    pipe.ae = ae
    g = torch.Generator().manual_seed(11)
    B, h, w = 2, 8, 12
    lat = mx.array(torch.randn((B, h * w // 4, 64), generator=g).to(torch.bfloat16).to(torch.float32))
    img = pipe.decode(lat, (h, w))
    u8 = (img * 255).astype(mx.uint8)
    np.savez(os.path.join(OUT, "ae_decode.npz"), config=json.dumps(SMALL_AE),
             weights_crc=synthetic.state_dict_checksum(sd), latents=npy(lat), latent_size=np.asarray([h, w]),
             image=npy(img), image_u8=npy(u8))


def golden_encoder_and_loss():
    """AutoEncoder.encode (flux/autoencoder.py:347-350, eval mode: DiagonalGaussian returns the mean) and
    FluxPipeline.training_loss (flux/flux.py:195-226) with its two random draws pinned (t, eps are stored)."""
    ap = specs.AutoEncoderParams(**SMALL_AE)
    sd = synthetic.synthetic_state_dict(specs.ae_decoder_manifest(ap) + specs.ae_encoder_manifest(ap))
    ae = AutoEncoder(RefAEParams(**SMALL_AE))
    ae.load_weights(list(ae.sanitize(f32(sd)).items()))
    ae.eval()
    g = torch.Generator().manual_seed(21)
    x = mx.array((torch.rand((2, 32, 48, 3), generator=g) * 2 - 1).to(torch.bfloat16).to(torch.float32))
    z = ae.encode(x)
    np.savez(os.path.join(OUT, "ae_encode.npz"), config=json.dumps(SMALL_AE), weights_crc=synthetic.state_dict_checksum(sd),
             image=npy(x), z=npy(z))
    # training loss on the small dev model (guidance embedded), conditioning and draws from a seeded generator
    model, fsd = build_flow(True)
    pipe = ref.FluxPipeline.__new__(ref.FluxPipeline)
    pipe.flow, pipe.dtype, pipe.sampler = model, mx.float32, FluxSampler("flux-dev")
    B, h, w, S = 2, 8, 12, 16
    x0 = mx.array(torch.randn((B, h, w, 16), generator=g).to(torch.bfloat16).to(torch.float32))
    t5f = mx.array(torch.randn((B, S, SMALL_FLOW["context_in_dim"]), generator=g).to(torch.bfloat16).to(torch.float32))
    clf = mx.array(torch.randn((B, SMALL_FLOW["vec_in_dim"]), generator=g).to(torch.bfloat16).to(torch.float32))
    gd = mx.array(torch.full((B,), 3.5))
    t = mx.array(torch.tensor([0.75, 0.3125]))
    eps = mx.array(torch.randn((B, h * w // 4, 64), generator=g).to(torch.bfloat16).to(torch.float32))
    pipe.sampler.random_timesteps = lambda *a, **k: t          # pin the two draws of training_loss
    real_normal = mx.random.normal
    mx.random.normal = lambda *a, **k: eps
    try:
        loss = pipe.training_loss(x0, t5f, clf, gd)
    finally:
        mx.random.normal = real_normal
    np.savez(os.path.join(OUT, "training_loss.npz"), config=json.dumps(SMALL_FLOW), weights_crc=synthetic.state_dict_checksum(fsd),
             x0=npy(x0), t5=npy(t5f), clip=npy(clf), guidance=3.5, t=npy(t), eps=npy(eps), loss=float(npy(loss)))


def golden_text():
    """T5Encoder (flux/t5.py:227-244) and CLIPTextModel (flux/clip.py:127-154)."""
    t5, t5_sd = build_t5()
    clip, clip_sd = build_clip()
    t5_tok, clip_tok = synthetic.synthetic_prompt_tokens(16, 77, seed=5, n_tok=9, t5_vocab=1000, clip_vocab=1024)
    txt = t5(mx.array(t5_tok))
    co = clip(mx.array(clip_tok))
    rpb = t5.encoder.relative_attention_bias(16, 16)
    np.savez(os.path.join(OUT, "text_encoders.npz"), t5_config=json.dumps(SMALL_T5),
             clip_config=json.dumps(SMALL_CLIP), t5_crc=synthetic.state_dict_checksum(t5_sd),
             clip_crc=synthetic.state_dict_checksum(clip_sd), t5_tokens=npy(t5_tok), clip_tokens=npy(clip_tok),
             t5_out=npy(txt), t5_bias=npy(rpb), clip_pooled=npy(co.pooled_output),
             clip_last=npy(co.last_hidden_state))


def golden_pipeline():
    """FluxPipeline.generate_latents + decode (flux/flux.py:128-162) end to end on the small models:
    tokens -> T5/CLIP -> 2 Euler steps -> VAE decode -> uint8 (txt2image.py:133)."""
    for variant, ge, steps in (("schnell", False, 2), ("dev", True, 3)):
        flow, flow_sd = build_flow(ge)
        ae, ae_sd = build_ae()
        t5, t5_sd = build_t5()
        clip, clip_sd = build_clip()
        t5_tok, clip_tok = synthetic.synthetic_prompt_tokens(16, 77, seed=5, n_tok=9, t5_vocab=1000, clip_vocab=1024)
        ref_flux.load_ae = lambda name: ae
        ref_flux.load_flow_model = lambda name: flow
        ref_flux.load_clip = lambda name: clip
        ref_flux.load_t5 = lambda name: t5
        ref_flux.load_clip_tokenizer = lambda name: FakeTok(clip_tok)
        ref_flux.load_t5_tokenizer = lambda name: FakeTok(t5_tok)
        pipe = ref.FluxPipeline("flux-" + variant)
        h, w, B = 8, 12, 2
        x_T = synthetic.synthetic_prior(B, (h, w), seed=42)
        pipe.sampler.sample_prior = lambda shape, dtype=None, key=None: mx.array(x_T)
        gen = pipe.generate_latents("a prompt", n_images=B, num_steps=steps, guidance=3.5, latent_size=(h, w), seed=3)
        x0, x_ids, txt, txt_ids, vec = next(gen)
        lats = [npy(x) for x in gen]
        assert len(lats) == steps
        img = pipe.decode(mx.array(torch.from_numpy(lats[-1])), (h, w))
        u8 = (img * 255).astype(mx.uint8)
        np.savez(os.path.join(OUT, f"pipeline_{variant}.npz"), steps=steps, guidance=3.5,
                 latent_size=np.asarray([h, w]), x_T_nhwc=npy(x_T), x_T=npy(x0), x_ids=npy(x_ids),
                 txt=npy(txt), txt_ids=npy(txt_ids), vec=npy(vec), t5_tokens=npy(t5_tok),
                 clip_tokens=npy(clip_tok), latents=np.stack(lats), image=npy(img), image_u8=npy(u8),
                 timesteps=np.asarray(pipe.sampler.timesteps(steps, x0.shape[1]), dtype=np.float64),
                 crc=np.asarray([synthetic.state_dict_checksum(s) for s in (flow_sd, ae_sd, t5_sd, clip_sd)]))


def golden_lora():
    """LoRA adapter at inference: txt2image.py:32-39 (load_adapter) -> FluxPipeline.linear_to_lora_layers /
    fuse_lora_layers (flux/flux.py:228-246) -> LoRALinear.__call__ / fuse (flux/lora.py:28-43,73-76), executed by the
    reference's own code on the small dev model.  The adapter file is written the way dreambooth.py:46-59 writes it
    (tree_flatten of the LoRA parameters + lora_rank / lora_blocks metadata)."""
    import zlib

    from flux.lora import LoRALinear
    from safetensors.torch import save_file
    model, sd = build_flow(True)
    pipe = ref.FluxPipeline.__new__(ref.FluxPipeline)
    pipe.flow = model
    rank, blocks = 4, 3  # the LAST three blocks: single.1, single.0, double.1 (flux/flux.py:230-232)
    pipe.linear_to_lora_layers(rank, blocks)
    adapter, names = {}, []
    for name, m in model.named_modules():
        if isinstance(m, LoRALinear):
            names.append(name)
            gen = torch.Generator().manual_seed(zlib.crc32(name.encode()))
            din, dout = m.lora_a.shape[0], m.lora_b.shape[1]
            adapter[name + ".lora_a"] = (torch.randn(din, rank, generator=gen) * din ** -0.5).to(torch.bfloat16).float()
            adapter[name + ".lora_b"] = (torch.randn(rank, dout, generator=gen) * 0.05).to(torch.bfloat16).float()
    model.load_weights([(k, mx.array(v)) for k, v in adapter.items()], strict=False)
    g = torch.Generator().manual_seed(11)
    B, h, w, S = 2, 8, 12, 16
    x = mx.array(torch.randn((B, h, w, 16), generator=g).to(torch.bfloat16))
    img, img_ids = pipe._prepare_latent_images(x)
    txt = mx.array(torch.randn((B, S, SMALL_FLOW["context_in_dim"]), generator=g).to(torch.bfloat16))
    txt_ids = mx.zeros((B, S, 3), dtype=mx.int32)
    y = mx.array(torch.randn((B, SMALL_FLOW["vec_in_dim"]), generator=g).to(torch.bfloat16))
    t = mx.full((B,), 0.5, dtype=mx.bfloat16)
    gd = mx.full((B,), 3.5, dtype=mx.bfloat16)
    call = lambda: model(img=img.astype(mx.float32), img_ids=img_ids, txt=txt.astype(mx.float32), txt_ids=txt_ids,  # noqa: E731
                         timesteps=t, y=y.astype(mx.float32), guidance=gd)
    out_unfused = call()
    pipe.fuse_lora_layers()
    assert not any(isinstance(m, LoRALinear) for _, m in model.named_modules())
    out_fused = call()
    fused = {"single_blocks.1.linear1.weight": model.single_blocks[1].linear1.weight,
             "double_blocks.1.img_mlp.layers.2.weight": model.double_blocks[1].img_mlp.layers[2].weight,
             "double_blocks.1.txt_mod.lin.weight": model.double_blocks[1].txt_mod.lin.weight,
             "double_blocks.0.img_attn.qkv.weight": model.double_blocks[0].img_attn.qkv.weight}  # untouched block
    save_file({k: v.to(torch.bfloat16).contiguous() for k, v in adapter.items()}, os.path.join(OUT, "lora_adapter.safetensors"),
              metadata={"lora_rank": str(rank), "lora_blocks": str(blocks)})
    np.savez(os.path.join(OUT, "lora.npz"), config=json.dumps(SMALL_FLOW), rank=rank, blocks=blocks,
             weights_crc=synthetic.state_dict_checksum(sd), lora_modules=np.asarray(names), img=npy(img), img_ids=npy(img_ids),
             txt=npy(txt), txt_ids=npy(txt_ids), y=npy(y), t=0.5, guidance=3.5, out_unfused=npy(out_unfused),
             out_fused=npy(out_fused), **{"fused." + k: npy(v[:48]) for k, v in fused.items()})  # first 48 rows of each


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    if len(sys.argv) > 1 and sys.argv[1] == "lora":  # add one fixture without touching the others
        torch.manual_seed(0)
        golden_lora()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "encoder":
        torch.manual_seed(0)
        golden_encoder_and_loss()
        sys.exit(0)
    torch.manual_seed(0)
    golden_schedule()
    golden_patchify()
    golden_flow()
    golden_ae()
    golden_text()
    golden_pipeline()
    golden_lora()
    golden_encoder_and_loss()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
