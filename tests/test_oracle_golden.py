"""CPU: the oracle (oracle/flux_oracle.py) against the golden vectors written by the reference's
own code over the MLX shim (oracle/gen_golden.py).  fp32 on both sides -> tight tolerances."""
import json
import os

import numpy as np
import pytest
import torch

from flux import specs, synthetic
from oracle import flux_oracle as O

G = os.path.join(os.path.dirname(__file__), "golden")


def load(name):
    return np.load(os.path.join(G, name), allow_pickle=False)


def close(a, b, rtol=2e-5, atol=2e-5):
    a = a.detach().to(torch.float32).numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    np.testing.assert_allclose(a, b, rtol=rtol, atol=atol)


def test_schedule_bit_exact():
    g = load("schedule.npz")
    for key in g.files:
        name, steps, L = key.split("_")
        got = O.timesteps(int(steps), int(L), schnell=(name == "schnell"))
        assert np.array_equal(np.asarray(got, dtype=np.float64), g[key]), key
    # known answers from SURVEY 8-a1
    assert O.timesteps(2, 1024, True) == [1.0, 0.5, 0.0]
    assert O.timesteps(4, 4096, True) == [1.0, 0.75, 0.5, 0.25, 0.0]
    dev = O.timesteps(50, 4096, False)
    assert abs(dev[1] - 0.99357950687) < 1e-7 and dev[-1] == 0.0 and dev[0] == 1.0


def test_patchify_bit_exact():
    g = load("patchify.npz")
    x = torch.from_numpy(g["x"])
    p, ids = O.prepare_latent_images(x)
    assert np.array_equal(p.numpy(), g["packed"])
    assert np.array_equal(ids.numpy(), g["ids"])
    assert np.array_equal(O.unpatchify(p, (4, 6)).numpy(), g["back"])
    assert np.array_equal(g["back"], g["x"])
    # feature order c*4 + dy*2 + dx
    assert p[0, 0, :8].tolist() == [x[0, 0, 0, 0], x[0, 0, 1, 0], x[0, 1, 0, 0], x[0, 1, 1, 0],
                                    x[0, 0, 0, 1], x[0, 0, 1, 1], x[0, 1, 0, 1], x[0, 1, 1, 1]]


@pytest.mark.parametrize("variant", ["schnell", "dev"])
def test_flow_forward(variant):
    g = load(f"flow_{variant}.npz")
    cfg = json.loads(str(g["config"]))
    ge = bool(g["guidance_embed"])
    p = O.FluxParams(**cfg, guidance_embed=ge)
    sd = synthetic.synthetic_state_dict(specs.flow_manifest(specs.FluxParams(**cfg, guidance_embed=ge)))
    assert synthetic.state_dict_checksum(sd) == int(g["weights_crc"]), "synthetic weight generator drifted"
    B = g["img"].shape[0]
    t = torch.full((B,), float(g["t"]), dtype=torch.bfloat16)
    gd = torch.full((B,), float(g["guidance"]), dtype=torch.bfloat16)
    taps = {}
    out = O.flux_forward(sd, p, torch.from_numpy(g["img"]), torch.from_numpy(g["img_ids"]),
                         torch.from_numpy(g["txt"]), torch.from_numpy(g["txt_ids"]), t,
                         torch.from_numpy(g["y"]), gd, taps=taps)
    for k, v in taps.items():
        close(v, g["tap." + k], rtol=1e-4, atol=1e-4)
    close(out, g["out"], rtol=1e-4, atol=1e-4)


def test_ae_decode():
    g = load("ae_decode.npz")
    cfg = json.loads(str(g["config"]))
    sd = synthetic.synthetic_state_dict(specs.ae_decoder_manifest(specs.AutoEncoderParams(**cfg)))
    assert synthetic.state_dict_checksum(sd) == int(g["weights_crc"])
    h, w = (int(v) for v in g["latent_size"])
    img = O.decode(sd, O.AutoEncoderParams(**cfg), torch.from_numpy(g["latents"]), (h, w))
    close(img, g["image"], rtol=1e-4, atol=1e-4)
    u8 = O.to_uint8(img).numpy()
    assert (np.abs(u8.astype(int) - g["image_u8"].astype(int)) <= 1).all()


def test_ae_encode_and_training_loss():
    """N4: AutoEncoder.encode (flux/autoencoder.py:347-350) and FluxPipeline.training_loss (flux/flux.py:195-226) as
    the reference's own code computed them over the shim (t and eps of the loss are stored in the fixture)."""
    g = load("ae_encode.npz")
    ap = specs.AutoEncoderParams(**json.loads(str(g["config"])))
    sd = synthetic.synthetic_state_dict(specs.ae_decoder_manifest(ap) + specs.ae_encoder_manifest(ap))
    assert synthetic.state_dict_checksum(sd) == int(g["weights_crc"])
    oap = O.AutoEncoderParams(**json.loads(str(g["config"])))
    z = O.vae_encode({k: v.float() for k, v in sd.items()}, oap, torch.from_numpy(g["image"]))
    close(z, g["z"], rtol=1e-4, atol=1e-4)
    g = load("training_loss.npz")
    cfg = json.loads(str(g["config"]))
    fsd = synthetic.synthetic_state_dict(specs.flow_manifest(specs.FluxParams(**cfg, guidance_embed=True)))
    assert synthetic.state_dict_checksum(fsd) == int(g["weights_crc"])
    B = g["x0"].shape[0]
    loss = O.training_loss({k: v.float() for k, v in fsd.items()}, O.FluxParams(**cfg, guidance_embed=True),
                           torch.from_numpy(g["x0"]), torch.from_numpy(g["t5"]), torch.from_numpy(g["clip"]),
                           torch.full((B,), float(g["guidance"])), torch.from_numpy(g["t"]), torch.from_numpy(g["eps"]))
    assert abs(loss.item() - float(g["loss"])) <= 1e-4 * abs(float(g["loss"])), (loss.item(), float(g["loss"]))


def test_text_encoders():
    g = load("text_encoders.npz")
    t5c = json.loads(str(g["t5_config"]))
    clc = json.loads(str(g["clip_config"]))
    t5_sd = synthetic.synthetic_state_dict(specs.t5_manifest(specs.T5Config(**t5c)))
    clip_sd = synthetic.synthetic_state_dict(specs.clip_manifest(specs.CLIPTextModelConfig(**clc)))
    assert synthetic.state_dict_checksum(t5_sd) == int(g["t5_crc"])
    assert synthetic.state_dict_checksum(clip_sd) == int(g["clip_crc"])
    ocfg = O.T5Config(**{k: v for k, v in t5c.items() if k in O.T5Config.__dataclass_fields__})
    close(O.t5_position_bias(t5_sd, ocfg, 16), g["t5_bias"], rtol=0, atol=0)
    close(O.t5_encode(t5_sd, ocfg, torch.from_numpy(g["t5_tokens"])), g["t5_out"], rtol=1e-4, atol=1e-4)
    pooled, last = O.clip_encode(clip_sd, O.CLIPConfig(**clc), torch.from_numpy(g["clip_tokens"]))
    close(last, g["clip_last"], rtol=1e-4, atol=1e-4)
    close(pooled, g["clip_pooled"], rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("variant", ["schnell", "dev"])
def test_pipeline_end_to_end(variant):
    """tokens -> T5/CLIP -> Euler loop -> decode, all through the oracle, vs FluxPipeline over the shim."""
    g = load(f"pipeline_{variant}.npz")
    fcfg = json.loads(str(load(f"flow_{variant}.npz")["config"]))
    ge = variant == "dev"
    acfg = json.loads(str(load("ae_decode.npz")["config"]))
    te = load("text_encoders.npz")
    t5c, clc = json.loads(str(te["t5_config"])), json.loads(str(te["clip_config"]))
    flow_sd = synthetic.synthetic_state_dict(specs.flow_manifest(specs.FluxParams(**fcfg, guidance_embed=ge)))
    ae_sd = synthetic.synthetic_state_dict(specs.ae_decoder_manifest(specs.AutoEncoderParams(**acfg)))
    t5_sd = synthetic.synthetic_state_dict(specs.t5_manifest(specs.T5Config(**t5c)))
    clip_sd = synthetic.synthetic_state_dict(specs.clip_manifest(specs.CLIPTextModelConfig(**clc)))
    ocfg = O.T5Config(**{k: v for k, v in t5c.items() if k in O.T5Config.__dataclass_fields__})
    B = g["x_T"].shape[0]
    txt = O.t5_encode(t5_sd, ocfg, torch.from_numpy(g["t5_tokens"])).expand(B, -1, -1)
    vec = O.clip_encode(clip_sd, O.CLIPConfig(**clc), torch.from_numpy(g["clip_tokens"]))[0].expand(B, -1)
    close(txt, g["txt"], rtol=1e-4, atol=1e-4)
    close(vec, g["vec"], rtol=1e-4, atol=1e-4)
    steps = int(g["steps"])
    h, w = (int(v) for v in g["latent_size"])
    x_T = torch.from_numpy(g["x_T_nhwc"])
    packed, ids = O.prepare_latent_images(x_T)
    assert np.array_equal(packed.numpy(), g["x_T"]) and np.array_equal(ids.numpy(), g["x_ids"])
    ts = O.timesteps(steps, packed.shape[1], schnell=not ge)
    assert np.array_equal(np.asarray(ts, dtype=np.float64), g["timesteps"])
    lats, img = O.generate_images(flow_sd, ae_sd, O.FluxParams(**fcfg, guidance_embed=ge),
                                  O.AutoEncoderParams(**acfg), x_T, txt, vec, steps, float(g["guidance"]),
                                  schnell=not ge)
    for i in range(steps):
        close(lats[i], g["latents"][i], rtol=2e-4, atol=2e-4)
    close(img, g["image"], rtol=5e-4, atol=5e-4)


def test_fp8_quantiser_known_answers():
    """oracle.fp8_quant_rows (the restatement of fx_quantize_rows, --quantize): e4m3 known answers.
    Row absmax maps to 448 (0x7E), zero rows to scale 1 / byte 0, ties round to even, tiny values to subnormals."""
    import torch

    from oracle import flux_oracle as O
    x = torch.tensor([[448.0, -448.0, 224.0, 1.0, 0.0, 17.0, 19.0, 2.0 ** -9],
                      [0.0] * 8,
                      [-3.5, 3.5, 1.75, 0.21875, 0.109375, 7.0 / 512, 0.0, 0.0]])
    q, s = O.fp8_quant_rows(x)
    by = q.to(torch.float8_e4m3fn).view(torch.uint8)
    assert s.squeeze(-1).tolist() == [1.0, 1.0, float(torch.tensor(3.5) * torch.tensor(1.0 / 448.0))]
    # row 0 (scale 1): 448 = 0x7E, -448 = 0xFE, 224 = 0x76, 1 = 0x38, 0 = 0x00, 17 -> 16 (tie to even) = 0x58,
    # 19 -> 20 (nearest) = 0x5A, 2^-9 = smallest subnormal 0x01
    assert by[0].tolist() == [0x7E, 0xFE, 0x76, 0x38, 0x00, 0x58, 0x5A, 0x01]
    assert by[1].tolist() == [0] * 8
    # row 2: absmax 3.5 -> inv = 128: values * 128 = -448, 448, 224, 28, 14, 1.75, 0, 0
    assert by[2].tolist() == [0xFE, 0x7E, 0x76, 0x5E, 0x56, 0x3E, 0x00, 0x00]
    # dequantised values stay within half an e4m3 step
    assert ((q * s - x).abs() <= x.abs() * 2.0 ** -4 + s * 2.0 ** -10).all()


def test_quantised_oracle_stays_close_to_fp32():
    """Mode(quantize=True) only touches the block Linears named by FP8_LINEARS and stays within FP8 noise of fp32."""
    import json

    import torch

    from flux import specs, synthetic
    from oracle import flux_oracle as O
    assert O.FP8_LINEARS.match("double_blocks.3.img_attn.qkv") and O.FP8_LINEARS.match("single_blocks.37.linear2")
    assert O.FP8_LINEARS.match("double_blocks.3.img_attn.proj") and not O.FP8_LINEARS.match("txt_in")
    assert not O.FP8_LINEARS.match("double_blocks.3.img_mod.lin") and not O.FP8_LINEARS.match("final_layer.linear")
    g = load("flow_schnell.npz")
    cfg = json.loads(str(g["config"]))
    p = specs.FluxParams(**cfg, guidance_embed=False)
    sd = synthetic.synthetic_state_dict(specs.flow_manifest(p))
    op = O.FluxParams(**cfg, guidance_embed=False)
    a = [torch.from_numpy(np.asarray(g[k])) for k in ("img", "img_ids", "txt", "txt_ids")]
    B = a[0].shape[0]
    tt = torch.full((B,), float(g["t"]), dtype=torch.bfloat16)
    y = torch.from_numpy(g["y"])
    ref = O.flux_forward(sd, op, a[0].float(), a[1], a[2].float(), a[3], tt, y.float())
    qnt = O.flux_forward(sd, op, a[0].float(), a[1], a[2].float(), a[3], tt, y.float(), mode=O.Mode("fp32", quantize=True))
    rel = ((qnt - ref).norm() / ref.norm()).item()
    assert 1e-4 < rel < 5e-2, rel


def test_nvfp4_quantisers_known_answers():
    """oracle.nvfp4_quant_rows (fx_quantize_rows_fp4) and nvfp4_quant_rows_chunked (the producer-emitted form: fx_gemm_fp4 /
    fx_attention with q_out + fx_fp4_finalize): known answers and the properties the GPU kernels are held to bit-exactly."""
    import torch

    from oracle import flux_oracle as O
    # e2m1 rounding: ties to the even mantissa, saturation at 6
    v = torch.tensor([0.25, 0.2500001, 0.75, 1.25, 1.75, 2.5, 3.5, 5.0, 5.0001, 7.0, -0.75, -100.0])
    assert O.e2m1_round(v).tolist() == [0.0, 0.5, 1.0, 1.0, 2.0, 2.0, 4.0, 4.0, 6.0, 6.0, -1.0, -6.0]
    # row quantiser: the row maximum maps to block scale 448 and value 6 exactly; a zero row to scale 1, zeros
    x = torch.zeros(2, 64)
    x[0, :16] = torch.linspace(-2688.0, 2688.0, 16)
    x[0, 16:32] = 1.0
    q, sf, g = O.nvfp4_quant_rows(x)
    assert g.flatten().tolist() == [1.0, 1.0] and sf[0].tolist() == [448.0, float(torch.tensor(1.0 / 6.0).to(torch.float8_e4m3fn)), 0.0, 0.0]
    assert q[0, 0].item() == -6.0 and q[0, 15].item() == 6.0 and (q[1] == 0).all() and (sf[1] == 0).all()
    # chunked form: power-of-two scales; row scale = the largest chunk's; block scales of smaller chunks shifted exactly
    x = torch.zeros(3, 128)
    x[0, :32] = 2688.0 * 4          # chunk 0: e = 2
    x[0, 32:64] = 2688.0 / 8        # chunk 1: e = -3  -> its block scales are shifted by 2^-5
    x[0, 64:96] = 1.0               # chunk 2: 1/2688 -> e = ceil(log2(3.72e-4)) = -11
    x[1] = torch.randn(128, generator=torch.Generator().manual_seed(1))
    q, sf, g = O.nvfp4_quant_rows_chunked(x)
    assert g.flatten()[0].item() == 4.0 and g.flatten()[2].item() == 2.0 ** -100
    assert sf[0, :2].tolist() == [448.0, 448.0] and sf[0, 2:4].tolist() == [14.0, 14.0]       # 448 * 2^-5
    assert (q[0, :64].abs() == 6.0).all()
    deq = q * sf.repeat_interleave(16, dim=-1) * g
    assert torch.equal(deq[0, :64], x[0, :64])                                                  # exactly representable inputs
    assert abs(deq[0, 64].item() - 1.0) <= 0.07                                                 # 2^-13 of the row maximum: still there
    assert (deq[2] == 0).all()
    # against the row quantiser on Gaussian data: the same error to within a few percent (the chunk scale costs at most one bit
    # of the block scale's RANGE, none of its precision)
    x = torch.randn(64, 3072, generator=torch.Generator().manual_seed(2)) * torch.logspace(-2, 1, 3072)
    qa, sa, ga = O.nvfp4_quant_rows(x)
    qb, sb, gb = O.nvfp4_quant_rows_chunked(x)
    ea = ((qa * sa.repeat_interleave(16, dim=-1) * ga - x).norm() / x.norm()).item()
    eb = ((qb * sb.repeat_interleave(16, dim=-1) * gb - x).norm() / x.norm()).item()
    assert 0.05 < ea < 0.12 and abs(eb - ea) <= 0.05 * ea, (ea, eb)
    # powers of two only, and never below the row quantiser's scale
    assert torch.equal(torch.frexp(gb)[0], torch.full_like(gb, 0.5)) and (gb >= ga).all() and (gb < 2 * ga + 1e-30).all()
    assert O.FP4_FUSED_LINEARS.match("double_blocks.0.img_attn.proj") and O.FP4_FUSED_LINEARS.match("single_blocks.9.linear2")
    assert not O.FP4_FUSED_LINEARS.match("single_blocks.9.linear1") and not O.FP4_FUSED_LINEARS.match("double_blocks.0.txt_mlp.0")
