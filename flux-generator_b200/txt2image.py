#!/usr/bin/env python
"""txt2image.py -- same command line as the reference's CLI (txt2image.py:42-154), B200 back end.

Flags, defaults and output conventions are the reference's: positional prompt, --model {schnell,dev},
--n-images 4, --image-size HxW (height first, rounded UP to multiples of 16 with a warning),
--steps (>= 1; default 2 for schnell / 50 for dev), --guidance 4.0, --n-rows, --decoding-batch-size,
--output, --save-raw (name.i.suffix), --seed, --verbose, --no-t5-padding; images are written as
truncated uint8 and the grid has a 4-px zero border per image.

Platform notes: --quantize (the reference: MLX 4-bit group quantisation of nn.Linear, txt2image.py:79-82)
selects the Blackwell-native reduced-precision path instead: the block Linears of the MMDiT run as FP8 e4m3
tcgen05 GEMMs (per-row scales, fp32 accumulate; Flux.quantize); --adapter loads a LoRA adapter and always fuses it
(flux/lora.py; --fuse-adapter is accepted for parity).  Extra flags: --synthetic (seeded random weights / tokenizers when no
checkpoints exist offline), --gpus N is handled by launching under torchrun (one process per GPU, the
image batch sharded contiguously, weights broadcast over NCCL).
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def to_latent_size(image_size):
    # txt2image.py:14-25
    h, w = image_size
    h = ((h + 15) // 16) * 16
    w = ((w + 15) // 16) * 16
    if (h, w) != image_size:
        print("Warning: The image dimensions need to be divisible by 16px. " f"Changing size to {h}x{w}.")
    return (h // 8, w // 8)


def build_parser():
    parser = argparse.ArgumentParser(description="Generate images from a textual prompt using stable diffusion")
    parser.add_argument("prompt")
    parser.add_argument("--model", choices=["schnell", "dev"], default="schnell")
    parser.add_argument("--n-images", type=int, default=4)
    parser.add_argument("--image-size", type=lambda x: tuple(map(int, x.split("x"))), default=(512, 512))
    parser.add_argument("--steps", type=int, help="Number of steps (min: 1, default: 2 for schnell, 50 for dev)")
    parser.add_argument("--guidance", type=float, default=4.0)
    parser.add_argument("--n-rows", type=int, default=1)
    parser.add_argument("--decoding-batch-size", type=int, default=1)
    parser.add_argument("--quantize", "-q", action="store_true")
    parser.add_argument("--quantize-bits", type=int, choices=[8, 4], default=4,
                        help="with --quantize: 4 = NVFP4 (W4A4) block Linears + e4m3 attention (default: the reference's --quantize is "
                             "4-bit too), 8 = FP8 e4m3 Linears + attention")
    parser.add_argument("--preload-models", action="store_true")
    parser.add_argument("--output", default="out.png")
    parser.add_argument("--save-raw", action="store_true")
    parser.add_argument("--seed", type=int)
    parser.add_argument("--verbose", "-v", action="store_true")
    parser.add_argument("--adapter")
    parser.add_argument("--fuse-adapter", action="store_true")
    parser.add_argument("--no-t5-padding", dest="t5_padding", action="store_false")
    parser.add_argument("--synthetic", action="store_true", help="seeded random weights/tokenizers (no checkpoints offline)")
    return parser


def parse_args(argv=None):
    parser = build_parser()
    args = parser.parse_args(argv)
    if args.steps is not None and args.steps < 1:
        parser.error("Number of steps must be at least 1")
    args.steps = args.steps or (50 if args.model == "dev" else 2)
    return args


def shard(n_images: int, rank: int, world: int):
    """Contiguous split of the image batch over ranks: rank r gets [lo, hi)."""
    base, rem = divmod(n_images, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def main(argv=None):
    args = parse_args(argv)
    import numpy as np
    import torch
    from PIL import Image
    from tqdm import tqdm

    from flux import FluxPipeline

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    lo, hi = shard(args.n_images, rank, world)

    flux = FluxPipeline("flux-" + args.model, t5_padding=args.t5_padding, synthetic=args.synthetic or None,
                        device=f"cuda:{local}", first_image_index=lo)
    if args.adapter:  # txt2image.py:76-77
        from flux.lora import load_adapter
        load_adapter(flux, args.adapter, fuse=args.fuse_adapter)
    if args.quantize:  # txt2image.py:79-82 quantises flow / t5 / clip to 4 bits; here: the flow model's block Linears -> NVFP4 / FP8
        flux.flow.quantize(bits=args.quantize_bits)
    if args.preload_models:
        flux.ensure_models_are_loaded()

    latent_size = to_latent_size(args.image_size)
    torch.cuda.reset_peak_memory_stats()
    t0 = time.time()
    latents = flux.generate_latents(args.prompt, n_images=hi - lo, num_steps=args.steps, latent_size=latent_size,
                                    guidance=args.guidance, seed=args.seed)
    conditioning = next(latents)
    torch.cuda.synchronize()
    peak_mem_conditioning = torch.cuda.max_memory_allocated() / 1024 ** 3
    torch.cuda.reset_peak_memory_stats()

    x_t = conditioning[0]
    for x_t in tqdm(latents, total=args.steps, disable=rank != 0):
        pass
    torch.cuda.synchronize()
    t_gen = time.time() - t0
    peak_mem_generation = torch.cuda.max_memory_allocated() / 1024 ** 3
    torch.cuda.reset_peak_memory_stats()

    decoded = []
    for i in tqdm(range(0, hi - lo, args.decoding_batch_size), disable=rank != 0):
        decoded.append(flux.decode_uint8(x_t[i: i + args.decoding_batch_size], latent_size))
    x = torch.cat(decoded, dim=0) if decoded else torch.empty((0, latent_size[0] * 8, latent_size[1] * 8, 3), dtype=torch.uint8, device=x_t.device)
    torch.cuda.synchronize()
    peak_mem_decoding = torch.cuda.max_memory_allocated() / 1024 ** 3
    peak_mem_overall = max(peak_mem_conditioning, peak_mem_generation, peak_mem_decoding)

    if world > 1:  # gather the uint8 images on rank 0 (3 MB per 1024^2 image)
        import torch.distributed as dist
        parts = [None] * world
        dist.all_gather_object(parts, x.cpu())
        x = torch.cat(parts, dim=0)
        if rank != 0:
            dist.destroy_process_group()
            return
    x = x.cpu().numpy()

    if args.save_raw:
        *name, suffix = args.output.split(".")
        name = ".".join(name)
        for i in range(len(x)):
            Image.fromarray(x[i]).save(".".join([name, str(i), suffix]))
    else:
        x = np.pad(x, [(0, 0), (4, 4), (4, 4), (0, 0)])
        B, H, W, C = x.shape
        x = x.reshape(args.n_rows, B // args.n_rows, H, W, C).transpose(0, 2, 1, 3, 4)
        x = x.reshape(args.n_rows * H, B // args.n_rows * W, C)
        Image.fromarray(x).save(args.output)

    if args.verbose:
        print(f"Peak memory used for the text:       {peak_mem_conditioning:.3f}GB")
        print(f"Peak memory used for the generation: {peak_mem_generation:.3f}GB")
        print(f"Peak memory used for the decoding:   {peak_mem_decoding:.3f}GB")
        print(f"Peak memory used overall:            {peak_mem_overall:.3f}GB")
        print(f"Denoising: {t_gen:.3f}s for {hi - lo} image(s) x {args.steps} step(s)")
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
