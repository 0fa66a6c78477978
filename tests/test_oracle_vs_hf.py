"""CPU: independent pins of the oracle's text encoders against Hugging Face `transformers` (the code base
the reference's T5 / CLIP ports were written from), on small random-weight models with the checkpoint key
names the reference's sanitizers accept.  (The reference's T5 uses exact-erf GELU for "gated-gelu",
flux/t5.py:172-176, whereas HF maps it to gelu_new; the HF config is set to "gelu" to match the reference.)"""
import pytest
import torch

from flux import specs, synthetic
from oracle import flux_oracle as O

transformers = pytest.importorskip("transformers")


def test_clip_oracle_matches_hf():
    cfg = specs.CLIPTextModelConfig(num_layers=2, model_dims=128, num_heads=2, max_length=77, vocab_size=1000)
    sd = {k: v.float() for k, v in synthetic.synthetic_state_dict(specs.clip_manifest(cfg)).items()}
    hf_cfg = transformers.CLIPTextConfig(vocab_size=1000, hidden_size=128, intermediate_size=512, num_hidden_layers=2,
                                         num_attention_heads=2, max_position_embeddings=77, hidden_act="quick_gelu",
                                         eos_token_id=999, bos_token_id=998, pad_token_id=1)
    model = transformers.CLIPTextModel(hf_cfg).eval()
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected and all("position_ids" in m for m in missing)
    tokens = torch.tensor([[998, 5, 17, 250, 3, 999], [998, 9, 999, 999, 999, 999]])  # second row: EOS-padded batch
    with torch.no_grad():
        out = model(input_ids=tokens)
    pooled, last = O.clip_encode(sd, O.CLIPConfig(num_layers=2, model_dims=128, num_heads=2, vocab_size=1000), tokens)
    torch.testing.assert_close(last, out.last_hidden_state, rtol=1e-4, atol=1e-4)
    # reference pooling: hidden state at argmax(token id) = first EOS (flux/clip.py:130,148)
    torch.testing.assert_close(pooled, out.last_hidden_state[torch.arange(2), tokens.argmax(-1)], rtol=1e-4, atol=1e-4)


def test_t5_oracle_matches_hf():
    cfg = specs.T5Config(vocab_size=512, num_layers=2, num_heads=4, d_kv=64, d_model=256, d_ff=512)
    sd = {k: v.float() for k, v in synthetic.synthetic_state_dict(specs.t5_manifest(cfg)).items()}
    hf_cfg = transformers.T5Config(vocab_size=512, d_model=256, d_kv=64, d_ff=512, num_layers=2, num_heads=4,
                                   relative_attention_num_buckets=32, relative_attention_max_distance=128,
                                   feed_forward_proj="gated-gelu", layer_norm_epsilon=1e-6, dropout_rate=0.0,
                                   tie_word_embeddings=False)
    hf_cfg.dense_act_fn = "gelu"  # the reference's choice for gated-gelu (exact erf)
    model = transformers.T5EncoderModel(hf_cfg).eval()
    hf_sd = dict(sd)
    hf_sd["encoder.embed_tokens.weight"] = sd["shared.weight"]
    missing, unexpected = model.load_state_dict(hf_sd, strict=False)
    assert not unexpected and not missing
    tokens = torch.tensor([[5, 17, 250, 3, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0]])  # padded, pads attended
    with torch.no_grad():
        out = model(input_ids=tokens).last_hidden_state  # no attention mask: like the reference (flux/t5.py:219-223)
    ocfg = O.T5Config(vocab_size=512, num_layers=2, num_heads=4, d_kv=64, d_model=256, d_ff=512)
    torch.testing.assert_close(O.t5_encode(sd, ocfg, tokens), out, rtol=2e-4, atol=2e-4)
