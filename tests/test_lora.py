"""LoRA adapters at inference: the oracle's restatement of flux/lora.py + flux/flux.py:228-246 and the
product's host-side adapter handling against the golden fixture written by the reference's own LoRA code
(oracle/gen_golden.py::golden_lora: linear_to_lora_layers -> load_weights(strict=False) -> forward -> fuse -> forward)."""
import json
import os

import numpy as np
import pytest
import torch

from flux import lora as plora, specs, synthetic
from oracle import flux_oracle as O

G = os.path.join(os.path.dirname(__file__), "golden")


def fixture():
    g = np.load(os.path.join(G, "lora.npz"), allow_pickle=False)
    cfg = json.loads(str(g["config"]))
    sp = specs.FluxParams(**cfg, guidance_embed=True)
    sd = synthetic.synthetic_state_dict(specs.flow_manifest(sp))
    assert synthetic.state_dict_checksum(sd) == int(g["weights_crc"])
    tensors, rank, blocks = plora.read_adapter(os.path.join(G, "lora_adapter.safetensors"))
    assert (rank, blocks) == (int(g["rank"]), int(g["blocks"]))
    return g, cfg, sp, sd, tensors, rank, blocks


def test_block_selection_matches_reference():
    g, cfg, sp, sd, tensors, rank, blocks = fixture()
    # flux/flux.py:230-233: double + single blocks reversed, first `blocks`
    assert plora.lora_blocks(2, 2, 3) == O.lora_blocks(2, 2, 3) == ["single_blocks.1", "single_blocks.0", "double_blocks.1"]
    assert plora.lora_blocks(19, 38, -1)[0] == "single_blocks.37" and len(plora.lora_blocks(19, 38, -1)) == 57
    mods = sorted(str(m) for m in g["lora_modules"])
    assert sorted({k.rsplit(".", 1)[0] for k in tensors}) == mods  # the file holds exactly the wrapped Linears
    pre = tuple(p + "." for p in plora.lora_blocks(2, 2, blocks))
    assert all(m.startswith(pre) for m in mods)


def test_fuse_matches_reference_weights_and_forward():
    g, cfg, sp, sd, tensors, rank, blocks = fixture()
    sd32 = {k: v.float() for k, v in sd.items()}
    fused = O.lora_fuse(sd32, tensors, blocks, cfg["depth"], cfg["depth_single_blocks"])
    for k in g.files:
        if k.startswith("fused."):
            ck = O.lora_checkpoint_key(k[len("fused."):])
            np.testing.assert_allclose(fused[ck][:48].numpy(), g[k], rtol=1e-6, atol=1e-6)
    assert torch.equal(fused["double_blocks.0.img_attn.qkv.weight"], sd32["double_blocks.0.img_attn.qkv.weight"])
    # the product computes the same deltas (checkpoint-side keys, fp32)
    deltas = plora.adapter_deltas(tensors, rank, blocks, cfg["depth"], cfg["depth_single_blocks"])
    assert len(deltas) == len(g["lora_modules"])
    for k, d in deltas.items():
        assert torch.allclose(sd32[k] + d, fused[k], rtol=0, atol=1e-6), k
    # forward of the fused oracle model == the reference's fused AND unfused forwards
    op = O.FluxParams(**cfg, guidance_embed=True)
    B = g["img"].shape[0]
    tt = torch.full((B,), float(g["t"]), dtype=torch.bfloat16)
    gd = torch.full((B,), float(g["guidance"]), dtype=torch.bfloat16)
    out = O.flux_forward(fused, op, torch.from_numpy(g["img"]), torch.from_numpy(g["img_ids"]), torch.from_numpy(g["txt"]),
                         torch.from_numpy(g["txt_ids"]), tt, torch.from_numpy(g["y"]), gd)
    np.testing.assert_allclose(out.numpy(), g["out_fused"], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(out.numpy(), g["out_unfused"], rtol=1e-4, atol=1e-4)
    base = O.flux_forward(sd32, op, torch.from_numpy(g["img"]), torch.from_numpy(g["img_ids"]), torch.from_numpy(g["txt"]),
                          torch.from_numpy(g["txt_ids"]), tt, torch.from_numpy(g["y"]), gd)
    assert (base - out).abs().max().item() > 1e-2  # the adapter does change the model


def test_adapter_errors():
    g, cfg, sp, sd, tensors, rank, blocks = fixture()
    with pytest.raises(ValueError, match="rank"):
        plora.adapter_deltas(tensors, rank + 1, blocks, cfg["depth"], cfg["depth_single_blocks"])
    t2 = {k: v for k, v in tensors.items() if k != "single_blocks.1.linear1.lora_b"}
    with pytest.raises(ValueError, match="lora_b"):
        plora.adapter_deltas(t2, rank, blocks, cfg["depth"], cfg["depth_single_blocks"])
    # entries outside the wrapped blocks are ignored (load_weights(strict=False), txt2image.py:37)
    assert len(plora.adapter_deltas(tensors, rank, 1, cfg["depth"], cfg["depth_single_blocks"])) == 3


@pytest.mark.gpu
def test_gpu_adapter_fused_forward_vs_reference(tmp_path):
    from flux.model import Flux
    from helpers import cosine, rel_l2
    g, cfg, sp, sd, tensors, rank, blocks = fixture()
    dev, bf = "cuda", torch.bfloat16
    model = Flux(sp, device=dev).load_weights(list(sd.items()))
    B = g["img"].shape[0]
    tt = torch.full((B,), float(g["t"]), dtype=bf, device=dev)
    gd = torch.full((B,), float(g["guidance"]), dtype=bf, device=dev)
    a = [torch.from_numpy(np.asarray(g[k])).to(dev) for k in ("img", "img_ids", "txt", "txt_ids")]
    y = torch.from_numpy(g["y"]).to(dev)
    base = model(a[0].to(bf), a[1], a[2].to(bf), a[3], tt, y.to(bf), gd)
    model.enable_lora(rank, blocks)                      # FluxPipeline.linear_to_lora_layers
    model.load_weights(list(tensors.items()), strict=False)  # txt2image.py:37
    out = model(a[0].to(bf), a[1], a[2].to(bf), a[3], tt, y.to(bf), gd)  # fused on first use
    assert not model._lora_pending
    assert rel_l2(out, g["out_fused"]) <= 2e-2 and cosine(out, g["out_fused"]) >= 0.9995
    assert rel_l2(base, g["out_fused"]) > 2 * rel_l2(out, g["out_fused"])
    with pytest.raises(ValueError):  # without enable_lora the adapter keys are unknown parameters (strict)
        Flux(sp, device=dev).load_weights(list(tensors.items()))
