"""CPU: the on-disk formats either side of the path -- BFL / HF safetensors checkpoints through the loaders
(flux/utils.py:98-210 in the reference) and the SentencePiece T5 tokenizer -- with small synthetic files."""
import json
import os

import pytest
import torch
from safetensors.torch import save_file

from flux import specs, synthetic, utils
from flux.tokenizers import T5Tokenizer

SMALL_FLOW = dict(in_channels=64, vec_in_dim=128, context_in_dim=256, hidden_size=256, mlp_ratio=4.0, num_heads=2,
                  depth=1, depth_single_blocks=1, axes_dim=[16, 56, 56], theta=10_000, qkv_bias=True)
SMALL_AE = dict(resolution=64, in_channels=3, ch=64, out_ch=3, ch_mult=[1, 2, 4, 4], num_res_blocks=2, z_channels=16,
                scale_factor=0.3611, shift_factor=0.1159)


def test_flow_checkpoint_roundtrip(tmp_path, monkeypatch):
    p = specs.FluxParams(**SMALL_FLOW, guidance_embed=True)
    sd = synthetic.synthetic_state_dict(specs.flow_manifest(p))
    path = str(tmp_path / "flux1-dev.safetensors")
    save_file({"model.diffusion_model." + k: v for k, v in sd.items()}, path)   # optional prefix (flux/model.py:88-89)
    monkeypatch.setattr(utils.configs["flux-dev"], "ckpt_path", path)
    monkeypatch.delenv("FLUX_B200_SYNTHETIC", raising=False)
    model = utils.load_flow_model("flux-dev", params=p, device="cpu")
    for k, v in sd.items():
        assert torch.equal(model._dest(k), v), k
    # the modulation Linears sit back to back in the arena (one GEMV per step)
    off = model._mod_off["single_blocks.0.modulation.lin"]
    assert torch.equal(model.arena["__mod_w"][off:off + 3 * 256], sd["single_blocks.0.modulation.lin.weight"])
    # strictness: a missing tensor and a wrong shape are errors, like mlx's load_weights
    bad = dict(sd)
    bad.pop("img_in.bias")
    with pytest.raises(ValueError, match="Missing"):
        utils.Flux(p, device="cpu").load_weights(list(bad.items()))
    bad = dict(sd)
    bad["img_in.bias"] = torch.zeros(3)
    with pytest.raises(ValueError, match="shape"):
        utils.Flux(p, device="cpu").load_weights(list(bad.items()))
    monkeypatch.setattr(utils.configs["flux-dev"], "ckpt_path", None)
    with pytest.raises(FileNotFoundError, match="no network"):
        utils.load_flow_model("flux-dev", params=p, device="cpu")


def test_ae_checkpoint_layout(tmp_path, monkeypatch):
    ap = specs.AutoEncoderParams(**SMALL_AE)
    sd = synthetic.synthetic_state_dict(specs.ae_decoder_manifest(ap))
    full = dict(sd)
    full["encoder.conv_in.weight"] = torch.zeros(64, 3, 3, 3, dtype=torch.bfloat16)  # encoder tensors are ignored
    path = str(tmp_path / "ae.safetensors")
    save_file(full, path)
    monkeypatch.setattr(utils.configs["flux-schnell"], "ae_path", path)
    monkeypatch.delenv("FLUX_B200_SYNTHETIC", raising=False)
    ae = utils.load_ae("flux-schnell", params=ap, device="cpu")
    w = sd["decoder.up.1.block.0.conv1.weight"]                      # OIHW in the file
    assert torch.equal(ae.arena["decoder.up.1.block.0.conv1.weight"], w.permute(0, 2, 3, 1).reshape(w.shape[0], -1))
    w = sd["decoder.conv_in.weight"]                                # 16 input channels padded to a 64-channel block
    stored = ae.arena["decoder.conv_in.weight"].reshape(w.shape[0], 3, 3, 64)
    assert torch.equal(stored[..., :16], w.permute(0, 2, 3, 1)) and stored[..., 16:].abs().max() == 0
    w = sd["decoder.up.1.block.0.nin_shortcut.weight"]              # 1x1 conv squeezed to a Linear
    assert torch.equal(ae.arena["decoder.up.1.block.0.nin_shortcut.weight"], w[:, :, 0, 0])
    q = ae.arena["decoder.mid.attn_1.qkv.weight"]
    assert torch.equal(q[:256], sd["decoder.mid.attn_1.q.weight"][:, :, 0, 0])


def test_text_encoder_snapshot_layout(tmp_path, monkeypatch):
    t5c = dict(vocab_size=512, num_layers=2, num_heads=4, relative_attention_num_buckets=32, d_kv=64, d_model=256,
               d_ff=512, feed_forward_proj="gated-gelu", tie_word_embeddings=False)
    t5_sd = synthetic.synthetic_state_dict(specs.t5_manifest(specs.T5Config(**t5c)))
    clip_cfg = specs.CLIPTextModelConfig(num_layers=2, model_dims=128, num_heads=2, vocab_size=1000)
    clip_sd = synthetic.synthetic_state_dict(specs.clip_manifest(clip_cfg))
    root = tmp_path / "snapshot"
    (root / "text_encoder_2").mkdir(parents=True)
    (root / "text_encoder").mkdir()
    keys = sorted(t5_sd)
    shards = {"model-00001-of-00002.safetensors": keys[: len(keys) // 2], "model-00002-of-00002.safetensors": keys[len(keys) // 2:]}
    for name, ks in shards.items():
        save_file({k: t5_sd[k] for k in ks}, str(root / "text_encoder_2" / name))
    (root / "text_encoder_2" / "model.safetensors.index.json").write_text(
        json.dumps({"weight_map": {k: n for n, ks in shards.items() for k in ks}}))
    (root / "text_encoder_2" / "config.json").write_text(json.dumps(t5c))
    save_file(clip_sd, str(root / "text_encoder" / "model.safetensors"))
    (root / "text_encoder" / "config.json").write_text(json.dumps(dict(
        num_hidden_layers=2, hidden_size=128, num_attention_heads=2, max_position_embeddings=77, vocab_size=1000,
        hidden_act="quick_gelu")))
    monkeypatch.setenv("FLUX_HF_DIR", str(root))
    monkeypatch.delenv("FLUX_B200_SYNTHETIC", raising=False)
    t5 = utils.load_t5("flux-schnell", device="cpu")
    assert t5.config.d_ff == 512 and torch.equal(t5.arena["shared.weight"], t5_sd["shared.weight"])
    pre = "encoder.block.1.layer.0.SelfAttention."
    assert torch.equal(t5.arena[pre + "qkv.weight"][256:512], t5_sd[pre + "k.weight"])       # stacked q|k|v
    clip = utils.load_clip("flux-schnell", device="cpu")
    assert clip.config.num_layers == 2
    pre = "text_model.encoder.layers.0.self_attn."
    assert torch.equal(clip.arena[pre + "qkv.bias"][256:], clip_sd[pre + "v_proj.bias"])


def test_t5_tokenizer_semantics(tmp_path):
    spm = pytest.importorskip("sentencepiece")
    corpus = tmp_path / "corpus.txt"
    corpus.write_text("\n".join(["a photo of a cat", "a painting of a dog on the moon", "the quick brown fox",
                                 "jumps over the lazy dog", "flux makes images from text"] * 20))
    prefix = str(tmp_path / "spiece")
    spm.SentencePieceTrainer.train(input=str(corpus), model_prefix=prefix, vocab_size=36, model_type="unigram", minloglevel=2,
                                   pad_id=0, eos_id=1, unk_id=2, bos_id=-1)   # T5's special-token ids
    tok = T5Tokenizer(prefix + ".model", max_length=32)
    assert (tok.pad_token, tok.eos_token, tok.bos_token) == (0, 1, -1)
    ids = tok.encode("a photo of a cat")
    assert ids.shape == (1, 32) and ids.dtype == torch.int32
    n = int((ids[0] != 0).sum())
    assert ids[0, n - 1] == 1 and (ids[0, n:] == 0).all()                    # ... EOS, then pad id 0 (no BOS)
    assert tok.encode("a photo of a cat", pad=False).shape == (1, n)         # --no-t5-padding
    long = tok.encode("the quick brown fox jumps over the lazy dog " * 8)
    assert long.shape[1] > 32 and long[0, -1] == 1                           # never truncated (flux/tokenizers.py:160-173)
    both = tok.encode(["a cat", "a painting of a dog on the moon"], pad=False)
    assert both.shape[0] == 2 and both[0, -1] == 0                           # batch padded with the pad id
