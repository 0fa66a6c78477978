"""GPU parity of the --quantize (FP8 e4m3) path through the C ABI.

The reference's --quantize is MLX 4-bit group quantisation of nn.Linear weights (txt2image.py:79-82); it cannot run
here (MLX absent), so this mode is pinned against the oracle's restatement of THIS repo's quantiser
(oracle.flux_oracle.fp8_quant_rows / Mode(quantize=True)) and, one level down, against fp32 torch matmuls on the
dequantised operands.  Tolerances:
  * quantiser: bytes and scales bit-exact against the oracle (integer / byte work);
  * FP8 GEMM family vs an fp32 matmul of the SAME dequantised operands: rel-L2 <= 2e-5 (fp32 out), 5e-3 (bf16 out)
    -- the tensor-core product of e4m3 values is exact in fp32, only the summation order differs;
  * quantised MMDiT forward vs the quantised oracle: rel-L2 <= 2e-2, cosine >= 0.9995 (the bf16 bar);
  * quantised vs UNquantised fp32 oracle (what FP8 costs): rel-L2 <= 8e-2, cosine >= 0.997, images mean |diff| <= 4/255.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu

from flux import ops, specs, synthetic  # noqa: E402
from flux.model import Flux  # noqa: E402
from helpers import cosine, rel_l2  # noqa: E402
from oracle import flux_oracle as O  # noqa: E402
from test_gpu_kernels import _qkv_ref  # noqa: E402

dev = "cuda"
bf = torch.bfloat16
F = torch.nn.functional


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(bf).to(dev)


def deq(q, s):
    return q.float() * s.float().unsqueeze(-1)


@pytest.mark.parametrize("shape", [(1, 8), (5, 264), (3, 200, 384), (2, 70, 15360), (1, 33, 3072)])
def test_quantize_rows_bit_exact_vs_oracle(shape):
    x = rnd(*shape, seed=1, scale=3.0)
    x.view(-1, shape[-1])[0].zero_()  # an all-zero row: scale 1, bytes 0
    q, s = ops.quantize_rows(x)
    oq, os_ = O.fp8_quant_rows(x.cpu().float())
    assert torch.equal(s.cpu(), os_.squeeze(-1))
    assert torch.equal(q.cpu().view(torch.uint8), oq.to(torch.float8_e4m3fn).view(torch.uint8))
    assert s.view(-1)[0].item() == 1.0 and q.view(torch.uint8).view(-1, shape[-1])[0].max().item() == 0
    # strided views: column slice in, column slice out
    if len(shape) == 3:
        B, R, K = shape
        wide = rnd(B, R, K + 64, seed=2)
        out = torch.zeros(B, R, K + 128, device=dev, dtype=ops.fp8)
        sc = torch.zeros(B, R, device=dev)
        ops.quantize_rows(wide[:, :, 64:], out=out[:, :, 128:], out_scale=sc)
        q2, s2 = ops.quantize_rows(wide[:, :, 64:].contiguous())
        assert torch.equal(out[:, :, 128:].view(torch.uint8), q2.view(torch.uint8)) and torch.equal(sc, s2)
        assert out[:, :, :128].view(torch.uint8).max().item() == 0


@pytest.mark.parametrize("M,N,K", [(128, 256, 128), (300, 520, 208), (1000, 136, 336), (129, 257, 80), (4096, 3072, 3072)])
def test_gemm_fp8_vs_dequantised_matmul(M, N, K):
    a, w = rnd(M, K, seed=3), rnd(N, K, seed=4, scale=K ** -0.5)
    qa, sa = ops.quantize_rows(a)
    qw, sw = ops.quantize_rows(w)
    ref = deq(qa, sa) @ deq(qw, sw).T
    assert rel_l2(ops.gemm(qa, qw, a_scale=sa, w_scale=sw, out_dtype=torch.float32), ref) <= 2e-5
    assert rel_l2(ops.gemm(qa, qw, a_scale=sa, w_scale=sw), ref) <= 5e-3
    # and FP8 itself stays within its quantisation noise of the unquantised product
    assert rel_l2(ref, a.float() @ w.float().T) <= 6e-2


def test_gemm_fp8_epilogues_views_and_errors():
    B, R, K, N = 3, 200, 256, 384
    a = rnd(B, R, K, seed=5)
    w, bias = rnd(N, K, seed=6, scale=K ** -0.5), rnd(N, seed=7)
    gate, resid = rnd(B, N, seed=8), rnd(B, R, N, seed=9)
    big = torch.zeros(B, R + 8, K + 128, device=dev, dtype=ops.fp8)
    scb = torch.zeros(B, R + 8, device=dev)
    qa, sa = big[:, 8:, 128:], scb[:, 8:]
    ops.quantize_rows(a, out=qa, out_scale=sa)
    qw, sw = ops.quantize_rows(w)
    lin = deq(qa, sa) @ deq(qw, sw).T + bias.float()
    assert rel_l2(ops.gemm(qa, qw, bias, act="gelu_tanh", a_scale=sa, w_scale=sw), F.gelu(lin, approximate="tanh")) <= 5e-3
    ref = resid.float() + gate.float()[:, None] * lin
    x = resid.clone()
    ops.gemm(qa, qw, bias, gate=gate, resid=x, out=x, a_scale=sa, w_scale=sw)  # in place on the residual stream
    assert rel_l2(x, ref) <= 5e-3
    with pytest.raises(ValueError):  # scales are mandatory
        ops.gemm(qa, qw)
    with pytest.raises(ValueError):  # mixed operand types
        ops.gemm(qa, w, a_scale=sa, w_scale=sw)
    with pytest.raises(ValueError):  # narrow outputs have no FP8 tile
        ops.gemm(qa, qw[:64], a_scale=sa, w_scale=sw[:64].contiguous())


@pytest.mark.parametrize("B,R,H,K,mlp", [(2, 200, 2, 256, 1024), (1, 130, 24, 3072, 12288)])
def test_gemm_qkv_fp8_epilogue(B, R, H, K, mlp):
    D, off = H * 128, 40
    a, w = rnd(B, R, K, seed=11), rnd(3 * D + mlp, K, seed=12, scale=K ** -0.5)
    bias, qs, ks = rnd(3 * D + mlp, seed=13, scale=0.1), (1 + rnd(128, seed=14, scale=0.1).float()).to(bf), \
        (1 + rnd(128, seed=15, scale=0.1).float()).to(bf)
    ang = torch.rand(R + off, 64, generator=torch.Generator().manual_seed(16)) * 6.28
    pe = torch.stack([torch.cos(ang), torch.sin(ang)], -1).to(bf).to(dev)
    q = torch.zeros(B, H, R + off, 128, device=dev, dtype=bf)
    k, v = torch.zeros_like(q), torch.zeros_like(q)
    mo = torch.zeros(B, R + off, D + mlp, device=dev, dtype=bf)
    qa, sa = ops.quantize_rows(a)
    qw, sw = ops.quantize_rows(w)
    ops.gemm_qkv(qa, qw, bias, qs, ks, pe, q, k, v, off, mlp_out=mo[:, :, D:], a_scale=sa, w_scale=sw)
    rq, rk, rv, rm = _qkv_ref(deq(qa, sa), deq(qw, sw), bias, qs, ks, pe[off:], H)
    assert rel_l2(q[:, :, off:], rq) <= 5e-3 and rel_l2(k[:, :, off:], rk) <= 5e-3 and rel_l2(v[:, :, off:], rv) <= 5e-3
    assert rel_l2(mo[:, off:, D:], rm) <= 5e-3 and mo[:, :, :D].abs().max().item() == 0


def test_rownorm_fp8_output():
    """The fused norm + quantise kernel is bit-identical to quantize_rows(rownorm(x)) and within half an e4m3 step
    of the fp32 result."""
    B, R, D = 2, 100, 3072
    x, sh, sc = rnd(B, R, D, seed=31), rnd(B, D, seed=32, scale=0.1), rnd(B, D, seed=33, scale=0.1)
    q = torch.zeros(B, R + 4, D, device=dev, dtype=ops.fp8)
    s = torch.zeros(B, R + 4, device=dev)
    ops.rownorm(x, 0, sh, sc, 1e-6, out=q[:, 4:], out_scale=s[:, 4:])
    q2, s2 = ops.quantize_rows(ops.rownorm(x, 0, sh, sc, 1e-6))
    assert torch.equal(q[:, 4:].view(torch.uint8), q2.view(torch.uint8)) and torch.equal(s[:, 4:], s2)
    assert q[:, :4].view(torch.uint8).max().item() == 0
    ref = (1 + sc.float()[:, None]) * F.layer_norm(x.float(), (D,), eps=1e-6) + sh.float()[:, None]
    err = (deq(q[:, 4:], s[:, 4:]) - ref).abs()
    bound = ref.abs() * (2.0 ** -4 + 2.0 ** -8) + s[:, 4:].unsqueeze(-1) * 2.0 ** -10  # e4m3 half step + bf16 rounding
    assert (err <= bound * 1.01).all()
    for mode, p0, p1 in ((1, (1 + rnd(D, seed=34, scale=0.1).float()).to(bf), rnd(D, seed=35, scale=0.1)),
                         (2, (1 + rnd(D, seed=34, scale=0.1).float()).to(bf), None)):
        qa, sa = torch.empty(B, R, D, device=dev, dtype=ops.fp8), torch.empty(B, R, device=dev)
        ops.rownorm(x, mode, p0, p1, 1e-6, out=qa, out_scale=sa)
        qb, sb = ops.quantize_rows(ops.rownorm(x, mode, p0, p1, 1e-6))
        assert torch.equal(qa.view(torch.uint8), qb.view(torch.uint8)) and torch.equal(sa, sb)


def _full_width(quantize):
    p = specs.FluxParams(depth=1, depth_single_blocks=1, guidance_embed=True)
    sd = synthetic.synthetic_state_dict(specs.flow_manifest(p))
    model = Flux(p, device=dev).load_weights(list(sd.items()))
    if quantize:
        model.quantize()
    g = torch.Generator().manual_seed(3)
    B, h, w, S = 2, 16, 40, 64
    x = torch.randn(B, h, w, 16, generator=g).to(bf)
    img, ids = O.prepare_latent_images(x)
    txt = torch.randn(B, S, 4096, generator=g).to(bf)
    y = torch.randn(B, 768, generator=g).to(bf)
    tids = torch.zeros(B, S, 3, dtype=torch.int32)
    ts = torch.full((B,), 0.5, dtype=bf)
    gd = torch.full((B,), 4.0, dtype=bf)
    args = (img, ids, txt, tids, ts, y, gd)
    return model, sd, args


@pytest.mark.parametrize("B,H,S", [(1, 2, 512), (2, 3, 1100), (1, 1, 77), (1, 24, 4352)])
def test_attention_fp8_vs_fp32_on_the_same_e4m3_operands(B, H, S):
    """fx_attention fp8 mode (e4m3 q / k / v, e4m3 P, fp32 softmax statistics) against softmax(q k^T / sqrt(128)) v in fp32
    on the SAME e4m3 operands: what is left is the rounding of P to e4m3 (3 mantissa bits, averaged over the keys) -- stated
    tolerance rel-L2 <= 4e-2 -- and against the oracle's restatement of the P rounding (sdpa_fp8)."""
    q, k, v = (rnd(B, H, S, 128, seed=s_).to(ops.fp8) for s_ in (81, 82, 83))
    out = torch.zeros(B, S, H * 128 + 64, device=dev, dtype=bf)
    ops.attention(q, k, v, out[:, :, :H * 128], 128 ** -0.5)
    qf, kf, vf = q.float(), k.float(), v.float()
    ref = F.scaled_dot_product_attention(qf, kf, vf).transpose(1, 2).reshape(B, S, -1)
    e = rel_l2(out[:, :, :H * 128], ref)
    orc = O.sdpa_fp8(O.FP32, qf.cpu(), kf.cpu(), vf.cpu(), 128 ** -0.5).transpose(1, 2).reshape(B, S, -1) if S <= 1100 else None
    print(f"fp8 attention B={B} H={H} S={S}: rel-L2 vs fp32 {e:.3e}" + ("" if orc is None else f", vs oracle sdpa_fp8 {rel_l2(out[:, :, :H * 128], orc):.3e}"))
    assert e <= 4e-2 and out[:, :, H * 128:].abs().max().item() == 0
    if orc is not None:
        assert rel_l2(out[:, :, :H * 128], orc) <= 4e-2
    with pytest.raises(ValueError):  # fp8 operands run on the persistent kernel only
        ops.attention(q, k, v, out[:, :, :H * 128], 128 ** -0.5, variant=1)


def test_gemm_qkv_writes_e4m3_q_k_v():
    """fx_gemm_qkv with e4m3 q / k / v outputs (the operands of the FP8 attention): the same values as the bf16 outputs,
    rounded to e4m3 (within one e4m3 step: rel-L2 <= 4e-2), for bf16 and for FP8 projections."""
    B, R, H, K, off = 2, 200, 3, 256, 32
    D = H * 128
    a, w = rnd(B, R, K, seed=11), rnd(3 * D, K, seed=12, scale=K ** -0.5)
    bias, qs, ks = rnd(3 * D, seed=13, scale=0.1), (1 + rnd(128, seed=14, scale=0.1).float()).to(bf), (1 + rnd(128, seed=15, scale=0.1).float()).to(bf)
    ang = torch.rand(R + off, 64, generator=torch.Generator().manual_seed(16)) * 6.28
    pe = torch.stack([torch.cos(ang), torch.sin(ang)], -1).to(bf).to(dev)
    ref = [torch.zeros(B, H, R + off, 128, device=dev, dtype=bf) for _ in range(3)]
    ops.gemm_qkv(a, w, bias, qs, ks, pe, *ref, off)
    got = [torch.zeros(B, H, R + off, 128, device=dev, dtype=ops.fp8) for _ in range(3)]
    ops.gemm_qkv(a, w, bias, qs, ks, pe, *got, off)
    for g8, r16 in zip(got, ref):
        assert rel_l2(g8.float()[:, :, off:], r16[:, :, off:]) <= 4e-2
        assert g8.view(torch.uint8)[:, :, :off].max().item() == 0          # rows before seq_off untouched
        step = (g8.float() - r16.float().to(ops.fp8).float()).abs()[:, :, off:]
        assert (step <= 0.13 * r16.float().abs()[:, :, off:] + 2e-3).all()  # at most one e4m3 step apart
    qa, sa = ops.quantize_rows(a)
    qw, sw = ops.quantize_rows(w)
    got8 = [torch.zeros(B, H, R + off, 128, device=dev, dtype=ops.fp8) for _ in range(3)]
    ops.gemm_qkv(qa, qw, bias, qs, ks, pe, *got8, off, a_scale=sa, w_scale=sw)
    for g8, r16 in zip(got8, ref):
        assert rel_l2(g8.float()[:, :, off:], r16[:, :, off:]) <= 8e-2
    with pytest.raises(ValueError):
        ops.gemm_qkv(a, w, bias, qs, ks, pe, got[0], ref[1], ref[2], off)    # q / k / v must agree in dtype


def test_flow_fp8_full_width_vs_quantised_oracle():
    """hidden 3072 / 24 heads, depth 1+1, batch 2: quantised CUDA forward vs the quantised and the plain oracle, with the
    FP8 attention (the default of Flux.quantize) and without it."""
    model, sd, (img, ids, txt, tids, ts, y, gd) = _full_width(True)
    assert model.quantized and len(model.quantized_keys()) == 2 * 4 + 2
    op = O.FluxParams(depth=1, depth_single_blocks=1, guidance_embed=True)
    ref = O.flux_forward(sd, op, img.float(), ids, txt.float(), tids, ts, y.float(), gd)
    for attn8, tol in ((True, 3e-2), (False, 2e-2)):
        model.quantize(attention=attn8)
        ref_q = O.flux_forward(sd, op, img.float(), ids, txt.float(), tids, ts, y.float(), gd,
                               mode=O.Mode("fp32", quantize=True, quantize_attention=attn8))
        out = model(img.to(dev), ids.to(dev), txt.to(dev), tids.to(dev), ts.to(dev), y.to(dev), gd.to(dev))
        print(f"fp8 (attention fp8={attn8}) vs quantised oracle", rel_l2(out, ref_q), cosine(out, ref_q), "| vs fp32 oracle",
              rel_l2(out, ref), cosine(out, ref), "| oracle fp8 vs fp32", rel_l2(ref_q, ref))
        assert rel_l2(out, ref_q) <= tol and cosine(out, ref_q) >= 0.9995
        assert rel_l2(out, ref) <= 8e-2 and cosine(out, ref) >= 0.997
    model.quantize()
    # graph replay of the quantised forward is bit-identical to eager
    a = [t_.to(dev) for t_ in (img, ids, txt, tids, ts, y, gd)]
    e = model.forward(*a).clone()
    g1 = model.forward_graphed(*a).clone()
    assert torch.equal(e, g1)


def test_quantize_then_reload_requantises():
    model, sd, _ = _full_width(True)
    k = "single_blocks.0.linear2"
    before = model._q8[k][0].view(torch.uint8).clone()
    sd2 = {kk: (v * 2 if kk == k + ".weight" else v) for kk, v in sd.items()}
    model.load_weights(list(sd2.items()))
    assert torch.equal(model._q8[k][0].view(torch.uint8), before)  # doubled weights: same bytes ...
    assert torch.allclose(model._q8[k][1], ops.quantize_rows(model._w(k))[1])  # ... doubled scales


@pytest.mark.parametrize("variant", ["schnell", "dev"])
def test_pipeline_quantised_vs_golden(variant):
    """The whole pipeline with Flux.quantize() against the reference-generated golden run (unquantised): what
    --quantize costs end to end.  Stated FP8 tolerance: latents rel-L2 <= 8e-2 / cosine >= 0.997, image mean |diff| <= 4/255."""
    import numpy as np

    from flux import FluxPipeline
    from helpers import FixedTokenizer, golden, small_configs
    g = golden(f"pipeline_{variant}.npz")
    fcfg, acfg, t5c, clc = small_configs()
    pipe = FluxPipeline("flux-" + variant, synthetic=True, device=dev,
                        flow_params=specs.FluxParams(**fcfg, guidance_embed=variant == "dev"),
                        ae_params=specs.AutoEncoderParams(**acfg), t5_config=specs.T5Config(**t5c),
                        clip_config=specs.CLIPTextModelConfig(**clc))
    for mod, man in ((pipe.flow, specs.flow_manifest(pipe.flow.params)), (pipe.ae, specs.ae_decoder_manifest(pipe.ae.params)),
                     (pipe.t5, specs.t5_manifest(pipe.t5.config)), (pipe.clip, specs.clip_manifest(pipe.clip.config))):
        sd = synthetic.synthetic_state_dict(man)
        mod.load_weights(list(mod.sanitize(sd).items()) if mod is pipe.ae else list(sd.items()))
    pipe.flow.quantize()
    pipe.t5_tokenizer = FixedTokenizer(g["t5_tokens"])
    pipe.clip_tokenizer = FixedTokenizer(g["clip_tokens"])
    steps = int(g["steps"])
    h, w = (int(v) for v in g["latent_size"])
    gen = pipe.generate_latents("a prompt", n_images=g["x_T"].shape[0], num_steps=steps, guidance=float(g["guidance"]),
                                latent_size=(h, w), seed=3, x_T=torch.from_numpy(g["x_T_nhwc"]).to(bf))
    next(gen)
    lats = list(gen)
    for i in range(steps):
        assert rel_l2(lats[i], g["latents"][i]) <= 8e-2 and cosine(lats[i], g["latents"][i]) >= 0.997
    d = np.abs(pipe.decode(lats[-1], (h, w)).cpu().numpy() - g["image"])
    print("quantised pipeline", variant, "latent rel-L2", rel_l2(lats[-1], g["latents"][-1]), "image mean |diff| * 255", d.mean() * 255)
    assert d.mean() <= 4 / 255


def test_cli_quantize_flag(tmp_path, monkeypatch):
    """txt2image.py --quantize (txt2image.py:56,79-82) switches the flow model to the quantised path (4-bit by default; this
    tiny shape -- 24 image tokens -- is not tileable by the NVFP4 kernels and runs on the FP8 Linears) and still writes images."""
    import flux
    import txt2image
    from helpers import small_configs
    from PIL import Image
    fcfg, acfg, t5c, clc = small_configs()
    real = flux.FluxPipeline
    made = []

    def small(name, **kw):
        p = real(name, flow_params=specs.FluxParams(**fcfg, guidance_embed="dev" in name),
                 ae_params=specs.AutoEncoderParams(**acfg), t5_config=specs.T5Config(**t5c),
                 clip_config=specs.CLIPTextModelConfig(**clc), **kw)
        made.append(p)
        return p

    monkeypatch.setattr(flux, "FluxPipeline", small)
    args = ["a cat", "--synthetic", "--n-images", "2", "--image-size", "64x96", "--steps", "2", "--seed", "3", "--save-raw"]
    txt2image.main(args + ["--output", str(tmp_path / "q.png"), "--quantize"])
    assert made[-1].flow.quantized
    txt2image.main(args + ["--output", str(tmp_path / "b.png")])
    assert not made[-1].flow.quantized
    import numpy as np
    a = np.asarray(Image.open(tmp_path / "q.0.png")).astype(np.int32)
    b = np.asarray(Image.open(tmp_path / "b.0.png")).astype(np.int32)
    assert a.shape == b.shape == (64, 96, 3) and np.abs(a - b).mean() <= 4.0
