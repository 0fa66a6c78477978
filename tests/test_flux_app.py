"""The A1111-compatible API (reference flux_app.py:47-62,64-362) with a mocked pipeline -- the way the
reference's own tests exercise it (test/test_api.py:51-129, test/test_generation.py:151-223) -- plus one GPU test that
drives the real (small, synthetic) pipeline through HTTP and compares with a direct pipeline call."""
import base64
import io
from unittest.mock import MagicMock, patch

import numpy as np
import pytest
from fastapi.testclient import TestClient
from PIL import Image

import flux_app
from flux_app import FluxAPI, SDAPIRequest, get_app, to_latent_size

TEST_PARAMS = {"prompt": "test prompt", "width": 128, "height": 128, "steps": 1, "seed": 42, "model": "schnell"}


def mock_pipeline(n=1, h=128, w=128):
    p = MagicMock()
    p.generate_latents.side_effect = lambda *a, **k: iter([np.zeros((n, 16, 16, 4)), np.zeros((n, 64, 64))])
    p.decode.return_value = np.full((1, h, w, 3), 0.999)
    return p


def test_to_latent_size_cases():
    # test/test_generation.py:156-164 (the consistent cases) and the round-UP rule of flux_app.py:333-345
    assert to_latent_size((512, 512)) == (64, 64)
    assert to_latent_size((768, 512)) == (96, 64)
    assert to_latent_size((513, 513)) == (66, 66)
    assert to_latent_size((100, 60)) == (14, 8)


def test_request_defaults():
    r = SDAPIRequest(prompt="x")  # flux_app.py:47-57
    assert (r.negative_prompt, r.width, r.height, r.steps, r.cfg_scale, r.batch_size, r.n_iter, r.seed, r.model) == \
        (None, 512, 512, None, 4.0, 1, 1, -1, "schnell")


def test_txt2img_endpoint_with_mock_pipeline():
    inst = FluxAPI()
    client = TestClient(get_app(inst))
    p = mock_pipeline()
    with patch.object(FluxAPI, "init_pipeline", return_value=p):
        resp = client.post("/sdapi/v1/txt2img", json=TEST_PARAMS)
    assert resp.status_code == 200
    data = resp.json()
    assert set(data) == {"images", "parameters", "info"} and len(data["images"]) == 1
    assert not data["images"][0].startswith("data:")           # bare base64 (flux_app.py:201-202)
    im = Image.open(io.BytesIO(base64.b64decode(data["images"][0])))
    assert im.size == (128, 128) and im.format == "PNG"
    assert np.asarray(im).max() == 254                          # 0.999 * 255 truncates to 254 (flux_app.py:192)
    assert data["parameters"]["seed"] == 42 and data["info"] == "Generated with Flux schnell model"
    p.generate_latents.assert_called_once()
    kw = p.generate_latents.call_args.kwargs
    assert kw["latent_size"] == (16, 16) and kw["num_steps"] == 1 and kw["seed"] == 42 and kw["n_images"] == 1
    p.decode.assert_called_once()


def test_generation_defaults_and_quirks():
    inst = FluxAPI()
    for model, want_steps in (("schnell", 2), ("dev", 2), ("flux-dev", 50), ("flux-schnell", 2)):  # flux_app.py:158
        p = mock_pipeline(n=6, h=104, w=200)
        with patch.object(FluxAPI, "init_pipeline", return_value=p):
            out = inst.generate_images("a", model=model, width=200, height=104, batch_size=2, n_iter=3, return_pil=True)
        kw = p.generate_latents.call_args.kwargs
        assert kw["num_steps"] == want_steps and kw["seed"] is None and kw["n_images"] == 6
        assert kw["latent_size"] == (13, 25) and kw["guidance"] == 4.0        # no /16 rounding (flux_app.py:141)
        assert len(out) == 6 and all(isinstance(i, Image.Image) for i in out) and p.decode.call_count == 6
    client = TestClient(get_app(inst))
    p = mock_pipeline()
    with patch.object(FluxAPI, "init_pipeline", return_value=p):
        client.post("/sdapi/v1/txt2img", json={"prompt": "x", "width": 128, "height": 128})    # seed -1 -> None
    assert p.generate_latents.call_args.kwargs["seed"] is None


def test_pipeline_cache_and_model_names():
    inst = FluxAPI(synthetic=True)
    with patch("flux.FluxPipeline") as cls:
        cls.side_effect = lambda name, **kw: MagicMock(name=name)
        a = inst.init_pipeline("schnell")
        assert cls.call_args.args == ("flux-schnell",) and cls.call_args.kwargs == {"synthetic": True}
        assert inst.init_pipeline("flux-schnell") is a and cls.call_count == 1     # cached (flux_app.py:84-88)
        b = inst.init_pipeline("dev")
        assert b is not a and cls.call_args.args == ("flux-dev",) and inst.current_model == "flux-dev"


def test_quantize_option_reaches_the_flow_model():
    inst = FluxAPI(synthetic=True, quantize=True)
    with patch("flux.FluxPipeline") as cls:
        pipe = MagicMock()
        cls.return_value = pipe
        assert inst.init_pipeline("schnell") is pipe
        pipe.flow.quantize.assert_called_once()          # --quantize: FP8 block Linears (Flux.quantize)
        inst.init_pipeline("flux-schnell")
        pipe.flow.quantize.assert_called_once()          # cached pipeline: not re-quantised


def test_errors_become_http_500():
    client = TestClient(get_app(FluxAPI()))
    with patch.object(FluxAPI, "init_pipeline", side_effect=RuntimeError("boom")):
        resp = client.post("/sdapi/v1/txt2img", json=TEST_PARAMS)
    assert resp.status_code == 500 and resp.json()["detail"] == "boom"         # flux_app.py:120-121
    resp = client.post("/sdapi/v1/txt2img", json={**TEST_PARAMS, "model": "stabilityai/sdxl-turbo"})
    assert resp.status_code == 500 and "Stable Diffusion" in resp.json()["detail"]
    assert client.post("/sdapi/v1/txt2img", json={"width": 64}).status_code == 422   # prompt is required


def test_models_options_progress_endpoints():
    client = TestClient(get_app(FluxAPI()))
    models = client.get("/sdapi/v1/sd-models").json()
    assert [m["model_name"] for m in models] == ["flux-schnell", "flux-dev"]
    for m in models:  # test/test_api.py:88-102
        assert set(m) == {"title", "name", "model_name", "hash", "sha256", "filename", "config"}
        assert m["filename"].endswith(".safetensors")
    opts = client.get("/sdapi/v1/options").json()
    assert "sd_model_checkpoint" in opts and opts["sd_backend"] == "Flux B200" and len(opts["sd_model_list"]) == 2
    assert client.post("/sdapi/v1/options", json={"test": "value"}).json() == {"success": True}
    prog = client.get("/sdapi/v1/progress").json()
    assert prog["progress"] == 0 and prog["textinfo"] == "Idle" and prog["state"]["job_count"] == 0


def test_port_helpers():
    import socket
    with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
        assert not flux_app.check_port_available("127.0.0.1", port)
        assert flux_app.find_available_port("127.0.0.1", port) != port
    assert flux_app.check_port_available("127.0.0.1", port)


@pytest.mark.gpu
def test_gpu_api_matches_direct_pipeline(monkeypatch):
    import flux
    import torch
    from flux import specs
    from helpers import small_configs
    fcfg, acfg, t5c, clc = small_configs()
    real = flux.FluxPipeline

    def small(name, **kw):
        return real(name, flow_params=specs.FluxParams(**fcfg, guidance_embed="dev" in name),
                    ae_params=specs.AutoEncoderParams(**acfg), t5_config=specs.T5Config(**t5c),
                    clip_config=specs.CLIPTextModelConfig(**clc), **kw)

    monkeypatch.setattr(flux, "FluxPipeline", small)
    inst = FluxAPI(synthetic=True)
    client = TestClient(get_app(inst))
    resp = client.post("/sdapi/v1/txt2img", json={"prompt": "a cat", "width": 96, "height": 64, "steps": 2, "seed": 5,
                                                 "batch_size": 2, "model": "schnell"})
    assert resp.status_code == 200, resp.text
    imgs = [np.asarray(Image.open(io.BytesIO(base64.b64decode(s)))) for s in resp.json()["images"]]
    assert len(imgs) == 2 and imgs[0].shape == (64, 96, 3) and not np.array_equal(imgs[0], imgs[1])
    pipe = inst.pipeline
    gen = pipe.generate_latents("a cat", n_images=2, num_steps=2, latent_size=(8, 12), guidance=4.0, seed=5)
    next(gen)
    x_t = list(gen)[-1]
    direct = (pipe.decode(x_t, (8, 12)) * 255).to(torch.uint8).cpu().numpy()   # the reference's per-image conversion
    assert np.array_equal(np.stack(imgs), direct)


@pytest.mark.gpu
def test_gpu_concurrent_requests_are_coalesced_bit_identically(monkeypatch):
    """Two clients with DIFFERENT prompts and seeds, same shape, at the same time: one coalesced batch (flux_serve.py),
    every image bit-identical to what its request produces alone."""
    import threading
    import flux
    from flux import specs
    from helpers import small_configs
    fcfg, acfg, t5c, clc = small_configs()
    real = flux.FluxPipeline

    def small(name, **kw):
        return real(name, flow_params=specs.FluxParams(**fcfg, guidance_embed="dev" in name),
                    ae_params=specs.AutoEncoderParams(**acfg), t5_config=specs.T5Config(**t5c),
                    clip_config=specs.CLIPTextModelConfig(**clc), **kw)

    monkeypatch.setattr(flux, "FluxPipeline", small)
    inst = FluxAPI(synthetic=True)
    base = {"width": 96, "height": 64, "steps": 2, "model": "schnell"}
    reqs = [{**base, "prompt": "a red fox", "seed": 5, "batch_size": 2}, {**base, "prompt": "two blue birds", "seed": 9, "batch_size": 3}]

    def fetch(pl):
        return [np.asarray(i) for i in inst.generate_images(pl["prompt"], model=pl["model"], width=pl["width"], height=pl["height"],
                                                            steps=pl["steps"], seed=pl["seed"], batch_size=pl["batch_size"], return_pil=True)]

    alone = [fetch(pl) for pl in reqs]                   # one request at a time
    inst.scheduler.window_s = 0.5                         # make sure both concurrent requests land in one window
    jobs_before = inst.scheduler.stats["jobs"]
    together = [None, None]
    ts = [threading.Thread(target=lambda i=i: together.__setitem__(i, fetch(reqs[i]))) for i in range(2)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert inst.scheduler.stats["jobs"] == jobs_before + 1 and inst.scheduler.stats["batches"][-1][1] == 5
    for a, b in zip(alone, together):
        assert len(a) == len(b) and all(np.array_equal(x, y) for x, y in zip(a, b))
    assert not np.array_equal(alone[0][0], alone[1][0])
