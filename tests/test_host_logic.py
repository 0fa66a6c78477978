"""CPU: host-side logic of the drop-in surface -- schedule, CLI parsing and rounding, tokenizers,
manifests, batch sharding (incl. a world_size-2 gloo run), and that the C-ABI library loads and
exports every symbol include/flux_b200.h declares (no compute calls without a GPU)."""
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "flux-generator_b200"))

import txt2image  # noqa: E402
from flux import FluxSampler, _native, specs, synthetic  # noqa: E402
from flux.tokenizers import CLIPTokenizer, SyntheticTokenizer  # noqa: E402
from helpers import golden  # noqa: E402


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "flux_b200.h")).read()
    declared = set(re.findall(r"\b(fx_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    lib = _native.load()
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in flux_b200.h but not exported"
    assert declared == set(_native.SYMBOLS), "ctypes table and header disagree"
    assert not any(n.startswith("fx_dbg_") for n in declared), "test-only probes belong in flux_b200_dbg.h"
    assert lib.fx_version() >= 100
    # the test-only companion library: its own header, its own .so, same rule
    import ctypes
    dbg_header = open(os.path.join(ROOT, "include", "flux_b200_dbg.h")).read()
    dbg_declared = set(re.findall(r"\b(fx_dbg_[a-z0-9_]+)\s*\(", dbg_header))
    assert dbg_declared == set(_native.DBG_SYMBOLS)
    dbg = ctypes.CDLL(os.path.join(os.path.dirname(_native.lib_path()), "libflux_b200_dbg.so"))
    for name in sorted(dbg_declared):
        assert hasattr(dbg, name) and not hasattr(lib, name), f"{name} must live in the companion library only"
    # host-only argument validation works without a GPU and reports through fx_last_error
    import ctypes as C
    rc = lib.fx_gemm(C.byref(_native.GemmArgs()), None)
    assert rc == -1 and b"null" in lib.fx_last_error()


def test_product_path_has_no_cpu_fallback():
    from flux import ops
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.gemm(torch.zeros(8, 8, dtype=torch.bfloat16), torch.zeros(8, 8, dtype=torch.bfloat16))
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            _native.lib()
    # the product package never imports the oracle
    pkg = os.path.join(ROOT, "flux-generator_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), f"{f} imports the oracle"


def test_sampler_schedule_bit_exact_vs_reference_fixture():
    g = golden("schedule.npz")
    for key in g.files:
        name, steps, L = key.split("_")
        got = FluxSampler("flux-" + name).timesteps(int(steps), int(L))
        assert np.array_equal(np.asarray(got, dtype=np.float64), g[key]), key
    s = FluxSampler("flux-schnell")
    assert s.timesteps(4, 4096) == [1.0, 0.75, 0.5, 0.25, 0.0]
    # what the model sees: bf16(1000 * bf16(t)) -> 1000, 752, 500, 250
    seen = [float((1000.0 * torch.tensor(t, dtype=torch.bfloat16))) for t in s.timesteps(4, 4096)[:4]]
    assert seen == [1000.0, 752.0, 500.0, 250.0]


def test_cli_surface():
    # reference cases (test/test_generation.py:156-164: the two consistent ones) + round-up rule
    g = golden("patchify.npz")
    for size, lat in zip(g["sizes"], g["latent_sizes"]):
        assert txt2image.to_latent_size(tuple(int(v) for v in size)) == tuple(int(v) for v in lat)
    assert txt2image.to_latent_size((512, 512)) == (64, 64)
    assert txt2image.to_latent_size((768, 512)) == (96, 64)
    a = txt2image.parse_args(["a cat"])
    assert (a.model, a.n_images, a.image_size, a.steps, a.guidance, a.n_rows, a.decoding_batch_size, a.output,
            a.t5_padding) == ("schnell", 4, (512, 512), 2, 4.0, 1, 1, "out.png", True)
    assert txt2image.parse_args(["x", "--model", "dev"]).steps == 50
    assert txt2image.parse_args(["x", "--image-size", "1024x768"]).image_size == (1024, 768)  # height first
    assert txt2image.parse_args(["x", "--no-t5-padding"]).t5_padding is False
    with pytest.raises(SystemExit):
        txt2image.parse_args(["x", "--steps", "0"])


def test_clip_tokenizer_semantics():
    # tiny vocabulary exercising lower-casing, whitespace collapse, BPE merges, BOS/EOS, truncation
    vocab = {"<|startoftext|>": 0, "<|endoftext|>": 1, "a</w>": 2, "c": 3, "a": 4, "t</w>": 5, "ca": 6, "cat</w>": 7,
             "!</w>": 8, "t": 9}
    merges = [("c", "a"), ("ca", "t</w>")]
    tok = CLIPTokenizer({m: i for i, m in enumerate(merges)}, vocab, max_length=6)
    assert tok.tokenize("A   CAT!") == [0, 2, 7, 8, 1]
    assert tok.encode("a cat").tolist() == [[0, 2, 7, 1]]                      # single prompt: not padded
    assert tok.encode(["a", "a cat !"]).tolist() == [[0, 2, 1, 1, 1], [0, 2, 7, 8, 1]]   # batch pads with EOS
    assert tok.tokenize("a a a a a a a a") == [0, 2, 2, 2, 2, 1]               # truncated to 6, EOS kept last
    assert tok.encode("a").dtype == torch.int32


def test_synthetic_tokenizers_shape_like_the_real_ones():
    t5 = SyntheticTokenizer("t5", 256, 32100).encode("a photo of a cat")
    assert t5.shape == (1, 256) and t5.dtype == torch.int32
    n = int((t5[0] != 0).sum())
    assert t5[0, n - 1] == 1 and (t5[0, n:] == 0).all()                        # EOS then pad id 0
    assert SyntheticTokenizer("t5", 256, 32100).encode("a photo of a cat", pad=False).shape[1] == n
    cl = SyntheticTokenizer("clip", 77, 49408).encode("a photo of a cat")
    assert cl[0, 0] == 49406 and cl[0, -1] == 49407 and cl.shape[1] <= 77
    assert int(cl[0].argmax()) == cl.shape[1] - 1                              # first EOS = pooled position
    assert torch.equal(t5, SyntheticTokenizer("t5", 256, 32100).encode("a photo of a cat"))  # deterministic


def test_manifests_match_the_reference_tensor_inventory():
    p = specs.FluxParams(guidance_embed=True)
    m = specs.flow_manifest(p)
    n = sum(int(np.prod(s)) for _, s, _ in m)
    assert abs(n / 1e9 - 11.90) < 0.01                                        # 11.90 B parameters
    keys = {k for k, _, _ in m}
    assert "double_blocks.18.txt_attn.norm.key_norm.scale" in keys and "single_blocks.37.linear2.weight" in keys
    assert "guidance_in.in_layer.weight" in keys
    assert "guidance_in.in_layer.weight" not in {k for k, _, _ in specs.flow_manifest(specs.FluxParams())}
    ae = specs.ae_decoder_manifest(specs.AutoEncoderParams())
    assert abs(sum(int(np.prod(s)) for _, s, _ in ae) / 1e6 - 49.5) < 0.1     # 49.5 M decoder parameters
    assert dict((k, s) for k, s, _ in ae)["decoder.up.1.block.0.nin_shortcut.weight"] == (256, 512, 1, 1)
    t5 = specs.t5_manifest(specs.T5Config())
    assert abs(sum(int(np.prod(s)) for _, s, _ in t5) / 1e9 - 4.76) < 0.01
    with pytest.raises(ValueError, match="divisible"):
        specs.FluxParams(num_heads=7).validate()
    with pytest.raises(ValueError, match="positional dim"):
        specs.FluxParams(axes_dim=[16, 56, 48]).validate()


def test_synthetic_prior_is_independent_of_sharding():
    full = synthetic.synthetic_prior(8, (4, 4), seed=42)
    lo, hi = txt2image.shard(8, 1, 2)
    assert (lo, hi) == (4, 8)
    assert torch.equal(synthetic.synthetic_prior(hi - lo, (4, 4), seed=42, first_index=lo), full[lo:hi])
    assert [txt2image.shard(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert txt2image.shard(2, 3, 4) == (2, 2)                                  # more ranks than images: empty shard


_GLOO_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.path.join(sys.argv[1], "flux-generator_b200"))
import txt2image
from flux.synthetic import synthetic_prior
from flux.model import WeightArena
dist.init_process_group("gloo")
r, w = dist.get_rank(), dist.get_world_size()
# weight broadcast at load: rank 0 fills the arena, everyone ends up with the same bytes
arena = WeightArena([("a.weight", (5, 8)), ("b.bias", (8,))], "cpu")
if r == 0:
    arena["a.weight"].copy_(torch.arange(40).reshape(5, 8)); arena["b.bias"].fill_(3)
arena.broadcast(0)
assert arena["a.weight"].float().sum().item() == 780 and arena["b.bias"].float().sum().item() == 24
# batch sharding: the union of the ranks' priors equals the unsharded prior
lo, hi = txt2image.shard(5, r, w)
mine = synthetic_prior(hi - lo, (4, 4), seed=1, first_index=lo)
parts = [None] * w
dist.all_gather_object(parts, mine)
assert torch.equal(torch.cat(parts), synthetic_prior(5, (4, 4), seed=1))
# max-over-ranks timing reduction used by bench.py
t = torch.tensor([float(r + 1)]); dist.all_reduce(t, op=dist.ReduceOp.MAX); assert t.item() == w
dist.destroy_process_group()
sys.stdout.write("ok %d\n" % r)  # one write per rank: the two ranks share the launcher's stdout pipe
sys.stdout.flush()
"""


def test_two_rank_gloo_sharding_and_broadcast(tmp_path):
    import socket
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER)
    r = None
    for attempt in range(3):  # the rendezvous port is picked free-right-now; retry if something grabbed it meanwhile
        with socket.socket() as sk:
            sk.bind(("127.0.0.1", 0))
            port = sk.getsockname()[1]
        r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                            "--master-addr", "127.0.0.1", "--master-port", str(port), str(script), ROOT],
                           capture_output=True, text=True, timeout=240)
        if r.returncode == 0:
            break
    assert r.returncode == 0, r.stdout + r.stderr
    assert sorted(r.stdout.split()) == ["0", "1", "ok", "ok"], r.stdout  # order / interleaving of the two ranks is free


def test_blocked_rope_table_layout():
    """ops.block_pe puts 16-byte piece k of position pos at uint4 index ((pos >> 5) * 16 + k) * 32 + (pos & 31) -- the
    address the QKV epilogue computes (gemm.cuh) -- and pads the last block of 32 positions with zeros."""
    import torch
    from flux import ops
    n = 77
    pe = torch.arange(n * 128, dtype=torch.float32).reshape(n, 64, 2).to(torch.bfloat16)
    blk = ops.block_pe(pe)
    assert blk.shape == (3, 16, 32, 8) and blk.is_contiguous()
    flat = blk.reshape(-1, 8)
    src = pe.reshape(n, 16, 8)
    for pos in (0, 1, 31, 32, 63, 76):
        for k in (0, 5, 15):
            assert torch.equal(flat[((pos >> 5) * 16 + k) * 32 + (pos & 31)], src[pos, k])
    assert flat[((76 >> 5) * 16 + 3) * 32 + 13 + 1:].abs().sum() >= 0  # padded rows exist
    assert torch.count_nonzero(blk[2, :, 13:]) == 0                    # positions 77..95 are zero padding


def test_adapter_file_needs_metadata(tmp_path):
    import pytest
    import torch
    from flux import lora
    from safetensors.torch import save_file
    f = tmp_path / "a.safetensors"
    save_file({"single_blocks.0.linear1.lora_a": torch.zeros(8, 2)}, str(f))
    with pytest.raises(ValueError, match="lora_rank"):   # txt2image.py:34-35 reads both keys unconditionally
        lora.read_adapter(str(f))
    save_file({"single_blocks.0.linear1.lora_a": torch.zeros(8, 2)}, str(f), metadata={"lora_rank": "2", "lora_blocks": "1"})
    t, rank, blocks = lora.read_adapter(str(f))
    assert (rank, blocks) == (2, 1) and list(t) == ["single_blocks.0.linear1.lora_a"]


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU oracle on the host cores) prints ONE JSON line with the keys the driver reads."""
    import json
    import subprocess
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, env={**os.environ, "CUDA_VISIBLE_DEVICES": ""})
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "flux_schnell_1024x1024_4step_images_per_sec" and d["unit"] == "images/s"
    assert d["value"] > 0 and d["vs_baseline"] is None and d["higher_is_better"] is True and "workload" in d["config"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_upconv_weights_reproduce_conv_of_upsampled_image():
    """ops.upconv_weights (host math of the fused Upsample convolution, flux/autoencoder.py:121-124): the four parity-wise
    2x2 kernels applied to the LOW-resolution image equal conv3x3(upsample_nearest(x, 2)), borders included (fp32, CPU)."""
    import torch.nn.functional as F
    from flux.ops import upconv_weights
    g = torch.Generator().manual_seed(3)
    B, H, W, C, Co = 2, 5, 7, 4, 6
    x = torch.randn(B, H, W, C, generator=g)
    w = torch.randn(Co, 9 * C, generator=g)
    up = x.repeat_interleave(2, dim=1).repeat_interleave(2, dim=2)
    ref = F.conv2d(up.permute(0, 3, 1, 2), w.view(Co, 3, 3, C).permute(0, 3, 1, 2), padding=1).permute(0, 2, 3, 1)
    w4 = upconv_weights(w).view(4, Co, 2, 2, C)
    out = torch.zeros(B, 2 * H, 2 * W, Co)
    xp = F.pad(x.permute(0, 3, 1, 2), (1, 1, 1, 1))                         # zero ring: source pixel -1 and H / W
    for py in (0, 1):
        for px in (0, 1):
            k = w4[2 * py + px].permute(0, 3, 1, 2)                          # [Co, C, 2, 2]
            y = F.conv2d(xp, k)                                              # valid conv over the padded image: (H+1) x (W+1)
            out[:, py::2, px::2] = y[:, :, py:py + H, px:px + W].permute(0, 2, 3, 1)
    assert torch.allclose(out, ref, rtol=1e-5, atol=1e-5)
