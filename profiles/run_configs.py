"""Times the other BASELINE.json configurations (parity-test cases, not bench lines) on one GPU:
ms per denoise step (CUDA-graph replay vs eager) and the VAE decode, synthetic full-size weights."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "flux-generator_b200"))
from flux import FluxPipeline  # noqa: E402

CONFIGS = [  # (model, H, W, batch, steps run, steps nominal, guidance)
    ("schnell", 256, 256, 1, 2, 2, 4.0), ("schnell", 512, 512, 1, 2, 2, 4.0), ("schnell", 1024, 1024, 1, 4, 4, 4.0),
    ("dev", 1024, 1024, 1, 6, 50, 7.0), ("dev", 1536, 1536, 1, 4, 50, 4.0)]


def flops_step(L, S):
    N = L + S
    return N * 12.910e9 + N * N * 700416.0


def main():
    pipes = {}
    for model, H, W, B, steps, nominal, guid in CONFIGS:
        if model not in pipes:
            pipes.clear()
            torch.cuda.empty_cache()
            pipes[model] = FluxPipeline("flux-" + model, synthetic=True, device="cuda")
        pipe = pipes[model]
        if os.environ.get("RUN_QUANT") == "1" and not pipe.flow.quantized:
            pipe.flow.quantize()  # the --quantize (FP8) path
        lat = (H // 8, W // 8)
        L, S = lat[0] * lat[1] // 4, (256 if model == "schnell" else 512)
        for graph in (True, False):
            pipe.use_graph = graph
            res = []
            for rep in range(2):
                gen = pipe.generate_latents("a prompt", n_images=B, num_steps=steps, guidance=guid, latent_size=lat, seed=1)
                next(gen)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                x = None
                for x in gen:
                    pass
                torch.cuda.synchronize()
                t1 = time.perf_counter()
                img = pipe.decode_uint8(x, lat)
                torch.cuda.synchronize()
                t2 = time.perf_counter()
                res.append(((t1 - t0) / steps * 1e3, (t2 - t1) * 1e3))
            ms_step, ms_vae = res[-1]
            ok = bool(torch.isfinite(x.float()).all()) and img.shape == (B, H, W, 3)
            print(f"{model} {H}x{W} B={B} N={L + S} {'graph' if graph else 'eager'}: {ms_step:8.2f} ms/step "
                  f"({flops_step(L, S) * B / ms_step / 1e9:6.0f} TFLOP/s), VAE decode {ms_vae:7.2f} ms, "
                  f"image(s) in {ms_step * nominal + ms_vae:9.1f} ms, finite={ok}", flush=True)


if __name__ == "__main__":
    main()
