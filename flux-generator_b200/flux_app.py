#!/usr/bin/env python
"""flux_app.py -- the A1111-compatible HTTP API of the reference (flux_app.py:47-62,64-321), B200 back end.

The second caller of the hot path: `POST /sdapi/v1/txt2img`, `GET /sdapi/v1/sd-models`,
`GET|POST /sdapi/v1/options`, `GET /sdapi/v1/progress`, with the reference's request / response models, defaults
and quirks kept on purpose:
  * latent size is (height // 8, width // 8) -- no rounding to 16 here, unlike the CLI (flux_app.py:141);
  * steps default: `steps or (50 if model == "flux-dev" else 2)` -- compared against the literal "flux-dev", so
    model="dev" gets 2 (flux_app.py:158);
  * seed -1 means "no seed" (flux_app.py:99); negative_prompt is accepted and ignored (flux_app.py:49,110);
  * images are returned as BARE base64 PNG strings, uint8 by truncation (flux_app.py:192-202);
  * any exception of the generation becomes HTTP 500 with detail=str(e) (flux_app.py:120-121);
  * one pipeline is cached, keyed by the model name with the "flux-" prefix added when missing (flux_app.py:71-88).
Not part of this build (reference product shell, SURVEY 2.1 #15-18: OUT OF SCOPE): the Gradio UI, the Stable Diffusion
and MusicGen back ends (requests naming a "stabilityai/..." model are answered with HTTP 500 and a clear message) and
the macOS compatibility check.  B200 specifics: the whole batch is decoded in one call on the GPU; concurrent requests
of the same (model, size, steps, guidance) are COALESCED into batches of up to 8 images per GPU and spread over
`--gpus N` worker processes, one per GPU (flux_serve.py: every image is bit-identical to what its request would have
produced alone); `--synthetic` serves seeded random weights when no checkpoints exist offline, `--quantize` selects the
FP8 path.
"""
from __future__ import annotations

import argparse
import base64
import io
import os
import sys
import threading
from typing import List, Optional, Tuple, Union

from fastapi import FastAPI, HTTPException
from fastapi.middleware.cors import CORSMiddleware
from pydantic import BaseModel

HERE = os.path.dirname(os.path.abspath(__file__))
if HERE not in sys.path:
    sys.path.insert(0, HERE)

import flux  # noqa: E402  (FluxPipeline is looked up on the package at call time so tests can patch flux.FluxPipeline)
import flux_serve  # noqa: E402


# ---------------------------------------------------------------------------------------------
# API models (flux_app.py:47-62)
# ---------------------------------------------------------------------------------------------
class SDAPIRequest(BaseModel):
    prompt: str
    negative_prompt: Optional[str] = None
    width: int = 512
    height: int = 512
    steps: Optional[int] = None
    cfg_scale: float = 4.0
    batch_size: int = 1
    n_iter: int = 1
    seed: int = -1
    model: str = "schnell"  # "schnell", "dev", "flux-schnell", "flux-dev"


class SDAPIResponse(BaseModel):
    images: List[str]
    parameters: dict
    info: str


class FluxAPI:
    """Unified API for external callers (flux_app.py:64-295)."""

    def __init__(self, synthetic: Optional[bool] = None, quantize: bool = False, device: Optional[str] = None,
                 gpus: int = 1, coalesce: bool = True):
        self.pipeline = None
        self.current_model = None
        self.synthetic = synthetic
        self.quantize = quantize
        self.device = device
        self.gpus = gpus
        self.coalesce = coalesce
        self._lock = threading.Lock()
        self._scheduler: Optional[flux_serve.NodeScheduler] = None

    # ------------------------------------------------------------------ coalescing / multi-GPU dispatch (flux_serve.py)
    @property
    def scheduler(self) -> flux_serve.NodeScheduler:
        """One worker per GPU: this process for a single device, one spawned process per device for --gpus N."""
        if self._scheduler is None:
            if self.gpus <= 1:
                workers = [self.run_job]
            else:
                import functools
                make = functools.partial(_make_runner, self.synthetic, self.quantize)
                workers = [flux_serve.ProcessWorker(f"cuda:{i}", make) for i in range(self.gpus)]
            self._scheduler = flux_serve.NodeScheduler(workers)
        return self._scheduler

    def run_job(self, job: "flux_serve.Job") -> list:
        """One coalesced batch on this API's device: per-image prompts, per-image priors keyed by (the request's seed,
        the image's index inside its request) -> uint8 arrays, one per job item."""
        import numpy as np
        import torch
        from flux import ops
        model, height, width, steps, guidance = job.key
        with self._lock:
            pipeline = self.init_pipeline(model)
            latent_size = (height // 8, width // 8)                 # flux_app.py:141
            steps = steps or (50 if model == "flux-dev" else 2)     # flux_app.py:158
            prompts = [it[2] for it in job.items]
            x = torch.cat([ops.prior_packed(1, latent_size, 16, it[3], first_index=it[1], device=pipeline.device)
                           for it in job.items], 0)
            text = prompts[0] if all(p == prompts[0] for p in prompts) else prompts
            latents = pipeline.generate_latents(text, n_images=len(prompts), num_steps=steps, latent_size=latent_size,
                                                guidance=guidance, x_T_packed=x)
            next(latents)
            x_t = None
            for x_t in latents:
                pass
            u8 = pipeline.decode_uint8(x_t, latent_size).cpu()
        return [np.asarray(u8[i]) for i in range(len(prompts))]

    def init_pipeline(self, model: str):
        """flux_app.py:71-88: one cached pipeline, re-created when the model name changes."""
        if model.startswith("stabilityai/"):
            raise ValueError(f"model {model!r}: the Stable Diffusion back ends are not part of the B200 build "
                             "(Flux models only: schnell, dev)")
        flux_model = model if model.startswith("flux-") else f"flux-{model}"
        if self.pipeline is None or self.current_model != flux_model:
            kw = {}
            if self.synthetic is not None:
                kw["synthetic"] = self.synthetic
            if self.device is not None:
                kw["device"] = self.device
            self.pipeline = flux.FluxPipeline(flux_model, **kw)
            if self.quantize:
                self.pipeline.flow.quantize(bits=4)   # like txt2image.py --quantize (the reference quantises to 4 bits)
            self.current_model = flux_model
        return self.pipeline

    async def txt2img(self, request: SDAPIRequest) -> SDAPIResponse:
        """flux_app.py:90-121."""
        try:
            # off the event loop: other requests keep arriving while this one runs, which is what lets them coalesce
            import asyncio
            import functools
            images = await asyncio.to_thread(functools.partial(
                self.generate_images,
                prompt=request.prompt, model=request.model, width=request.width, height=request.height,
                steps=request.steps, guidance=request.cfg_scale, seed=request.seed if request.seed >= 0 else None,
                batch_size=request.batch_size, n_iter=request.n_iter, return_pil=False))
            return SDAPIResponse(
                images=images,
                parameters={"prompt": request.prompt, "negative_prompt": request.negative_prompt, "width": request.width,
                            "height": request.height, "steps": request.steps, "cfg_scale": request.cfg_scale,
                            "seed": request.seed, "model": request.model},
                info=f"Generated with Flux {request.model} model")
        except Exception as e:  # noqa: BLE001  (the reference maps everything to 500)
            raise HTTPException(status_code=500, detail=str(e))

    def generate_images(self, prompt: str, model: str = "schnell", width: int = 512, height: int = 512,
                        steps: Optional[int] = None, guidance: float = 4.0, seed: Optional[int] = None,
                        batch_size: int = 1, n_iter: int = 1, return_pil: bool = False) -> List[Union[str, "Image.Image"]]:
        """flux_app.py:123-204: conditioning -> denoise loop -> decode -> uint8 (truncation) -> PNG -> base64."""
        import numpy as np
        from PIL import Image
        n = batch_size * n_iter
        native = self.gpus > 1
        if not native:
            with self._lock:
                native = getattr(self.init_pipeline(model), "_b200_native", False) is True
        if native and self.coalesce:
            # B200 pipeline(s): the request joins whatever compatible requests are in flight (flux_serve.coalesce).  A
            # request without a seed draws one here, so its images stay consistent when they span several jobs / GPUs.
            if seed is None:
                seed = int.from_bytes(os.urandom(8), "little") >> 1
            req = flux_serve.ImageRequest(prompt=prompt, model=model, height=height, width=width, steps=steps,
                                          guidance=guidance, seed=seed, n_images=n)
            arrays = self.scheduler.submit(req)
            return self._encode(arrays, return_pil)
        with self._lock:
            pipeline = self.init_pipeline(model)
            latent_size = (height // 8, width // 8)                 # flux_app.py:141 (no /16 rounding here)
            steps = steps or (50 if model == "flux-dev" else 2)     # flux_app.py:158
            latents = pipeline.generate_latents(prompt, n_images=n, num_steps=steps, latent_size=latent_size,
                                                guidance=guidance, seed=seed)
            next(latents)                                           # conditioning (T5 / CLIP run here, cached per prompt)
            x_t = None
            for x_t in latents:
                pass
            if getattr(pipeline, "_b200_native", False) is True:    # B200 pipeline: whole batch, uint8 on the GPU
                u8 = pipeline.decode_uint8(x_t, latent_size)
                arrays = [np.asarray(u8[i].cpu()) for i in range(n)]
            else:                                                   # any object with the reference's decode()
                arrays = []
                for i in range(n):
                    img = np.asarray(pipeline.decode(x_t[i:i + 1], latent_size))
                    arrays.append((img[0] * 255).astype(np.uint8))  # flux_app.py:192: truncation
        return self._encode(arrays, return_pil)

    @staticmethod
    def _encode(arrays, return_pil: bool):
        from PIL import Image
        images = []
        for arr in arrays:
            pil_image = Image.fromarray(arr)
            if return_pil:
                images.append(pil_image)
            else:
                buffered = io.BytesIO()
                pil_image.save(buffered, format="PNG")
                images.append(base64.b64encode(buffered.getvalue()).decode())   # bare base64 (flux_app.py:201-202)
        return images

    def list_models(self):
        """flux_app.py:206-245 (the two Flux entries; the Stable Diffusion ones are not served by this build)."""
        return [
            {"title": "flux-schnell", "name": "Flux Schnell (Fast)", "model_name": "flux-schnell", "hash": None,
             "sha256": None, "filename": "flux-schnell.safetensors", "config": None},
            {"title": "flux-dev", "name": "Flux Dev (High Quality)", "model_name": "flux-dev", "hash": None,
             "sha256": None, "filename": "flux-dev.safetensors", "config": None},
        ]

    def get_options(self):
        """flux_app.py:247-274."""
        return {
            "sd_model_checkpoint": self.current_model or "flux-schnell",
            "sd_backend": "Flux B200",
            "sd_model_list": [
                {"title": "Flux Schnell (Fast)", "name": "flux-schnell", "model_name": "flux-schnell"},
                {"title": "Flux Dev (High Quality)", "name": "flux-dev", "model_name": "flux-dev"},
            ],
        }

    def set_options(self, options: dict):
        """flux_app.py:276-278."""
        return {"success": True}

    def get_progress(self):
        """flux_app.py:280-295."""
        return {"progress": 0, "eta_relative": 0,
                "state": {"skipped": False, "interrupted": False, "job": "", "job_count": 0, "job_timestamp": ""},
                "current_image": None, "textinfo": "Idle"}


def _make_runner(synthetic, quantize, device: str):
    """Built inside a worker process (flux_serve.ProcessWorker): a FluxAPI pinned to `device`, serving jobs."""
    return FluxAPI(synthetic=synthetic, quantize=quantize, device=device).run_job


api = FluxAPI()


def create_api(app, instance: Optional[FluxAPI] = None):
    """Mount the endpoints on a FastAPI app (flux_app.py:299-321)."""
    global api
    if instance is not None:
        api = instance

    @app.post("/sdapi/v1/txt2img")
    async def txt2img(request: SDAPIRequest):
        return await api.txt2img(request)

    @app.get("/sdapi/v1/sd-models")
    async def list_models():
        return api.list_models()

    @app.get("/sdapi/v1/options")
    async def get_options():
        return api.get_options()

    @app.post("/sdapi/v1/options")
    async def set_options(options: dict):
        return api.set_options(options)

    @app.get("/sdapi/v1/progress")
    async def get_progress():
        return api.get_progress()

    return api


def to_latent_size(size: Tuple[int, int]) -> Tuple[int, int]:
    """flux_app.py:333-345: round UP to multiples of 16, then /8."""
    h, w = size
    h = ((h + 15) // 16) * 16
    w = ((w + 15) // 16) * 16
    if (h, w) != size:
        print("Warning: The image dimensions need to be divisible by 16px. " f"Changing size to {h}x{w}.")
    return (h // 8, w // 8)


def check_port_available(host: str, port: int) -> bool:
    """flux_app.py:347-355."""
    import socket
    with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as s:
        try:
            s.bind((host, port))
            return True
        except OSError:
            return False


def find_available_port(host: str, start_port: int, max_attempts: int = 10) -> int:
    """flux_app.py:357-362."""
    for port in range(start_port, start_port + max_attempts):
        if check_port_available(host, port):
            return port
    raise RuntimeError(f"Could not find an available port in range {start_port}-{start_port + max_attempts - 1}")


def get_app(instance: Optional[FluxAPI] = None) -> FastAPI:
    """FastAPI app with CORS + the API endpoints (flux_app.py:856-880, without the Gradio mount)."""
    app = FastAPI()
    app.add_middleware(CORSMiddleware, allow_origins=["*"], allow_credentials=True, allow_methods=["*"], allow_headers=["*"])
    create_api(app, instance)
    return app


def main(argv=None):
    """flux_app.py:780-853: --port (default 7860, next free port when taken), --listen-all (0.0.0.0 instead of localhost)."""
    parser = argparse.ArgumentParser(description="FLUX Image Generator (B200 back end)")
    parser.add_argument("--port", type=int, default=7860, help="Port to run the server on")
    listen_group = parser.add_mutually_exclusive_group()
    listen_group.add_argument("--listen-all", action="store_true", help="Listen on all network interfaces (0.0.0.0)")
    parser.add_argument("--synthetic", action="store_true", help="seeded random weights / tokenizers (no checkpoints offline)")
    parser.add_argument("--quantize", "-q", action="store_true", help="NVFP4 (W4A4) block Linears + e4m3 attention (Flux.quantize(bits=4))")
    parser.add_argument("--gpus", type=int, default=1, help="worker processes, one per GPU; coalesced batches are spread over them")
    args = parser.parse_args(argv)
    host = "0.0.0.0" if args.listen_all else "127.0.0.1"
    if args.listen_all:
        print("\nWarning: Server is listening on all network interfaces (0.0.0.0)")
    port = args.port if check_port_available(host, args.port) else find_available_port(host, args.port)
    if port != args.port:
        print(f"\nWarning: Port {args.port} is in use, using port {port} instead")
    app = get_app(FluxAPI(synthetic=True if args.synthetic else None, quantize=args.quantize, gpus=args.gpus))
    print(f"\nStarting Flux server on {host}:{port}")
    import uvicorn
    uvicorn.Server(uvicorn.Config(app, host=host, port=port, log_level="info")).run()


if __name__ == "__main__":
    main()
