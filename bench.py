#!/usr/bin/env python
"""Headline benchmark: Flux-schnell 1024x1024 4-step images/sec (BASELINE.json), one process per GPU.

    python bench.py --gpus N --steps K --warmup W            # this repo's B200 path
    python bench.py --impl reference --gpus N --steps K ...   # the reference algorithm on the host CPU

A "step" is one pass of the hot path over one per-GPU batch: 4 Euler steps of the MMDiT over 8 images
of 1024x1024 (BASELINE.json configs[3]: 64 images / 8 GPUs = 8 per GPU) + VAE decode + uint8.
Weak scaling: every rank processes its own 8 images; no data-path collective.  Weights are
materialised on rank 0 and broadcast once over NCCL at load (outside the timed region).

`value`   : images/s with prior and conditioning already resident in HBM (CUDA events, max over ranks).
`e2e`     : the same metric through the public API (FluxPipeline.generate_latents/decode_uint8) with
            HOST buffers: pinned x_T and token ids H2D every step, uint8 images D2H every step;
            T5/CLIP run on the first call for the prompt and are cached afterwards (north star).
`roofline`: the tcgen05 GEMM family (81% of the step's FLOPs), timed live with CUDA events around every
            launch inside the timed region; peak = MEASURED_PEAKS.json bf16_tflops_sustained.
`cpu_baseline`: the CPU oracle (torch restatement of the reference; MLX is not installable) timed on this
            box's host cores on a bounded sample of the same workload, extrapolated to images/s.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(ROOT, "flux-generator_b200"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

METRIC = "flux_schnell_1024x1024_4step_images_per_sec"
UNIT = "images/s"
H_IMG = W_IMG = 1024
STEPS_DENOISE = 4
PER_GPU_BATCH = 8
PROMPT = "a synthetic benchmark prompt for the flux denoising hot path"


def workload_config(n_gpus: int) -> dict:
    return {"workload": "flux-schnell 1024x1024, 4 Euler steps + VAE decode, 8 images per GPU (BASELINE configs[3])",
            "global_batch": PER_GPU_BATCH * n_gpus, "per_gpu_batch": PER_GPU_BATCH, "image": [H_IMG, W_IMG],
            "denoise_steps": STEPS_DENOISE, "seq_len": 256 + (H_IMG // 16) * (W_IMG // 16), "parallelism": f"dp{n_gpus}",
            "l2": "inputs larger than L2: 23.8 GB of weights + 2 GB of activations stream per denoise step (126 MB L2)"}


def measured_traffic() -> dict:
    """DRAM bytes per launch of the dominant GEMM members, read from the committed ncu capture of THIS build
    (profiles/gemm_traffic.json, written by profiles/parse_traffic.py from `ncu --metrics dram__bytes_*` over
    tests/gpu_l2_probe.py).  None when no capture has been committed for the build."""
    try:
        with open(os.path.join(ROOT, "profiles", "gemm_traffic.json")) as f:
            return json.load(f)
    except Exception:
        return {}


def peaks() -> dict:
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            d = json.load(f)
        return {"tflops": d["bf16_tflops_sustained"], "hbm": d["hbm_gbs"], "src": "measured (MEASURED_PEAKS.json, sustained)"}
    except Exception:
        return {"tflops": 1400.0, "hbm": 6650.0, "src": "fallback (B200_PROFILING.md sustained)"}


# ---------------------------------------------------------------------------------------------
# CPU oracle timing (cpu_baseline leg and --impl reference)
# ---------------------------------------------------------------------------------------------
def cpu_sample(threads: int):
    """Bounded sample of the same workload on the host cores: one 1024x1024 image (N = 4352 tokens)
    through 1 double + 2 single MMDiT blocks of the 19 + 38, plus the VAE decode of a 256x256 image.
    Returns (seconds per image extrapolated, description)."""
    from flux import specs, synthetic
    from oracle import flux_oracle as O
    torch.set_num_threads(threads)
    p = specs.FluxParams(depth=1, depth_single_blocks=2)
    sd = {k: v.float() for k, v in synthetic.synthetic_state_dict(specs.flow_manifest(p)).items()}
    op = O.FluxParams(depth=1, depth_single_blocks=2)
    g = torch.Generator().manual_seed(0)
    h = w = H_IMG // 8
    x = torch.randn(1, h, w, 16, generator=g)
    img, ids = O.prepare_latent_images(x)
    txt = torch.randn(1, 256, 4096, generator=g)
    y = torch.randn(1, 768, generator=g)
    tids = torch.zeros(1, 256, 3, dtype=torch.int32)
    ts = torch.full((1,), 0.5, dtype=torch.bfloat16)
    ap = specs.AutoEncoderParams()
    ae_sd = {k: v.float() for k, v in synthetic.synthetic_state_dict(specs.ae_decoder_manifest(ap)).items()}
    lat = torch.randn(1, 256, 64, generator=g)

    def run():
        t0 = time.perf_counter()
        O.flux_forward(sd, op, img, ids, txt, tids, ts, y)
        t1 = time.perf_counter()
        O.decode(ae_sd, O.AutoEncoderParams(), lat, (32, 32))
        t2 = time.perf_counter()
        return t1 - t0, t2 - t1

    def to_seconds_per_image(t_blocks, t_vae):
        # 3 of 57 blocks measured (1 of 19 double, 2 of 38 single: the same 1:2 mix) -> x19; 4 steps;
        # VAE FLOPs scale with pixel count (256^2 -> 1024^2 = x16)
        return t_blocks * 19 * STEPS_DENOISE + t_vae * 16

    desc = ("1 image, 1024x1024 (N=4352), 1 double + 2 single MMDiT blocks of 19+38 (x19, x4 steps) + VAE decode "
            "at 256x256 (x16); fp32 torch restatement of the reference (MLX not installable)")
    return run, to_seconds_per_image, desc


def reference_arm(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    run, to_spi, desc = cpu_sample(threads)
    for _ in range(max(args.warmup, 0) and 1):  # one warm-up pass is enough on the CPU
        run()
    t0 = time.perf_counter()
    spis = []
    for _ in range(args.steps):
        tb, tv = run()
        spis.append(to_spi(tb, tv))
    wall = time.perf_counter() - t0
    spi = sum(spis) / len(spis)
    value = 1.0 / spi  # one host serves all N "GPUs' worth" of work: whole-job rate is the host's rate
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": wall / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args.gpus),
            "ms_per_step_note": "wall time of ONE bounded CPU sample (3 of 57 blocks for one image + a 256x256 VAE decode), "
                                "not of the 8-image workload `value` is extrapolated to",
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": desc},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = sorted(int(r[1]) for r in self.rows if len(r) >= 8 and r[1].isdigit())
        mx = [int(r[2]) for r in self.rows if len(r) >= 8 and r[2].isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 8:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
# live per-kernel timing (roofline)
# ---------------------------------------------------------------------------------------------
class KernelTimer:
    """Wraps flux.ops entry points with CUDA events on the launching stream; flops are algorithmic."""

    def __init__(self, ops):
        self.ops, self.records, self.enabled = ops, [], False
        self._orig = {}

    def install(self):
        ops = self.ops

        def wrap(name, flops_fn):
            orig = getattr(ops, name)
            self._orig[name] = orig

            def f(*a, **k):
                if not self.enabled:
                    return orig(*a, **k)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                r = orig(*a, **k)
                e1.record()
                self.records.append((name, flops_fn(*a, **k), e0, e1))
                return r
            setattr(ops, name, f)

        def gemm_flops(a, w, *_, **__):
            rows = a.numel() // a.shape[-1]
            return 2.0 * rows * w.shape[0] * w.shape[1]

        def attn_flops(q, k, v, out, scale, **__):
            B, H, S, D = q.shape
            return 4.0 * B * H * S * S * D

        def conv_flops(x, w, *_, **__):
            B, H, W, _c = x.shape
            return 2.0 * B * H * W * w.shape[0] * w.shape[1]

        wrap("gemm", gemm_flops)
        wrap("gemm_qkv", gemm_flops)
        wrap("attention", attn_flops)
        wrap("conv3x3", conv_flops)
        for name in ("rownorm", "gemv", "groupnorm", "upsample2x", "softmax_rows", "transpose", "finish_image",
                     "unpatchify_scale", "patchify", "euler_step", "timestep_embedding"):
            wrap(name, lambda *a, **k: 0.0)  # HBM-bound helpers: time only

    def summary(self) -> dict:
        out = {}
        for name, fl, e0, e1 in self.records:
            d = out.setdefault(name, {"launches": 0, "ms": 0.0, "tflop": 0.0})
            d["launches"] += 1
            d["ms"] += e0.elapsed_time(e1)
            d["tflop"] += fl / 1e12
        for d in out.values():
            d["tflops"] = d["tflop"] / (d["ms"] * 1e-3) if d["ms"] > 0 else 0.0
        return out


# ---------------------------------------------------------------------------------------------
# the B200 arm
# ---------------------------------------------------------------------------------------------
def main_arm(args) -> None:
    import torch.distributed as dist
    from flux import FluxPipeline, _native, ops
    from flux.synthetic import synthetic_prior

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))
    timer = KernelTimer(ops)
    timer.install()
    use_graph = not args.no_graph

    B = PER_GPU_BATCH
    latent = (H_IMG // 8, W_IMG // 8)
    pipe = FluxPipeline("flux-schnell", synthetic=True, device=dev, first_image_index=rank * B)
    pipe.use_graph = use_graph
    torch.cuda.synchronize()

    # host-side inputs of one step (e2e) and their device-resident copies (value)
    x_T_host = synthetic_prior(B, latent, seed=42, first_index=rank * B).pin_memory()
    out_host = torch.empty((B, H_IMG, W_IMG, 3), dtype=torch.uint8).pin_memory()
    x_T_dev = x_T_host.to(dev)

    split_events = []  # (before denoise loop, after it, after decode) per step: ms_per_denoise_step

    def step_resident():
        gen = pipe.generate_latents(PROMPT, n_images=B, num_steps=STEPS_DENOISE, guidance=4.0, latent_size=latent,
                                    x_T=x_T_dev)
        next(gen)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        ev[0].record()
        x_t = None
        for x_t in gen:
            pass
        ev[1].record()
        out = pipe.decode_uint8(x_t, latent)
        ev[2].record()
        split_events.append(ev)
        return out

    cold = {"on": False, "n": 0}

    def step_e2e():
        prompt = PROMPT
        if cold["on"]:  # a prompt never seen before: tokenise + T5 + CLIP + txt_in + conditioning inside the step
            cold["n"] += 1
            prompt = f"{PROMPT} number {cold['n']} of rank {rank}"
        x = x_T_host.to(dev, non_blocking=True)                      # H2D: prior (+ token ids inside tokenize)
        gen = pipe.generate_latents(prompt, n_images=B, num_steps=STEPS_DENOISE, guidance=4.0, latent_size=latent, x_T=x)
        next(gen)
        x_t = None
        for x_t in gen:
            pass
        out_host.copy_(pipe.decode_uint8(x_t, latent), non_blocking=True)   # D2H: uint8 images
        torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k, w, with_clocks=False):
        for _ in range(w):
            fn()
        clocks = ClockSampler(local) if with_clocks else None
        barrier()
        if clocks:
            clocks.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        barrier()
        wall = time.perf_counter() - t0
        ms = e0.elapsed_time(e1)
        t = torch.tensor([ms, wall * 1e3], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t[0].item(), t[1].item(), (clocks.stop() if clocks else None)

    # ---- value: device-resident inputs; the MMDiT forward is replayed from a CUDA graph
    for _ in range(args.warmup):
        step_resident()
    split_events.clear()
    ms, _, clocks = timed(step_resident, args.steps, 0, with_clocks=True)
    value = B * world * args.steps / (ms * 1e-3)
    denoise_ms = sum(e[0].elapsed_time(e[1]) for e in split_events) / len(split_events) / STEPS_DENOISE
    decode_ms = sum(e[1].elapsed_time(e[2]) for e in split_events) / len(split_events)

    # ---- roofline: the same step once more in eager mode with CUDA events around every launch of the library
    # (graph replay hides the individual launches from the host; the kernels and their order are identical)
    pipe.use_graph = False
    step_resident()
    timer.enabled = True
    timer.records.clear()
    launches0 = _native.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step_resident()
    e1.record()
    torch.cuda.synchronize()
    eager_ms = e0.elapsed_time(e1)
    launches = (_native.launch_count() - launches0)
    timer.enabled = False
    ksum = timer.summary()
    pipe.use_graph = use_graph

    # ---- e2e: public API with host buffers
    e2e_ms, e2e_wall, _ = timed(step_e2e, args.steps, max(1, min(args.warmup, 2)))
    e2e_value = B * world * args.steps / (max(e2e_ms, e2e_wall) * 1e-3)

    # ---- cold prompts: the same call with a NEW prompt every step (text encoders, txt_in, modulation table are not
    # cached; the CUDA graph is keyed on shapes only and is replayed as before)
    cold["on"] = True
    graphs_before = len(pipe.flow._graphs)
    c_ms, c_wall, _ = timed(step_e2e, args.steps, 1)
    cold["on"] = False
    cold_value = B * world * args.steps / (max(c_ms, c_wall) * 1e-3)
    recaptured = len(pipe.flow._graphs) != graphs_before
    # text encoders alone (T5-XXL + CLIP-L shapes, synthetic weights) for one fresh prompt
    te = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    t5_tok, clip_tok = pipe.tokenize(PROMPT + " text encoder timing")
    torch.cuda.synchronize()
    te[0].record()
    pipe.t5(t5_tok)
    pipe.clip(clip_tok)
    te[1].record()
    torch.cuda.synchronize()
    text_ms = te[0].elapsed_time(te[1])

    # ---- --quantize legs (reported beside the bf16 headline, never in its place); same workload, device-resident inputs.
    # `quantized` = the CLI's --quantize (4-bit like the reference's: NVFP4 W4A4 block Linears + e4m3 attention);
    # `quantized.fp8` = --quantize --quantize-bits 8 (FP8 e4m3 W8A8 block Linears + e4m3 attention)
    quant = None
    if not args.no_quantized:
        try:  # a failure of an extra leg must not take the headline line with it (it would fail on every rank alike)
            pipe.flow.quantize()
            split_events.clear()
            q_ms, _, q_clocks = timed(step_resident, args.steps, 2, with_clocks=True)
            quant = {"dtype": "fp8_e4m3 (block Linears W8A8 with per-row scales, fp32 accumulate; attention Q K^T and P V in e4m3 with fp32 "
                              "softmax; embedders / final layer / VAE bf16)",
                     "value": B * world * args.steps / (q_ms * 1e-3), "unit": UNIT, "ms_per_step": q_ms / args.steps,
                     "ms_per_denoise_step": sum(e[0].elapsed_time(e[1]) for e in split_events[-args.steps:]) / args.steps / STEPS_DENOISE,
                     "clocks": q_clocks, "flag": "txt2image.py --quantize / Flux.quantize()",
                     "parity": "tests/test_gpu_fp8.py, tests/test_gpu_fullsize.py::test_fp8_full_depth_four_steps (19+38 blocks, N=4352, 4 steps: "
                               "latents rel-L2 vs the fp32 oracle 4.7e-3 .. 8.7e-3 per step, bf16 3.8e-3 .. 6.3e-3; image mean |diff| 0.76/255 vs "
                               "0.67/255; own tolerance -- the bf16 line above is the headline)"}
            # --quantize --quantize-bits 4: every block Linear as NVFP4 W4A4 (tcgen05 kind::mxf4nvf4.block_scale), FP8 attention
            pipe.flow.quantize(bits=4)
            split_events.clear()
            q4_ms, _, q4_clocks = timed(step_resident, args.steps, 2, with_clocks=True)
            quant4 = {
                "dtype": "nvfp4 (block Linears W4A4: e2m1 + UE4M3 scale per 16 + fp32 row / channel scales, fp32 accumulate; attention e4m3; "
                         "embedders / final layer / VAE bf16)",
                "value": B * world * args.steps / (q4_ms * 1e-3), "unit": UNIT, "ms_per_step": q4_ms / args.steps,
                "ms_per_denoise_step": sum(e[0].elapsed_time(e[1]) for e in split_events[-args.steps:]) / args.steps / STEPS_DENOISE,
                "clocks": q4_clocks, "flag": "txt2image.py --quantize / Flux.quantize(bits=4)",
                "parity": "tests/test_gpu_fp4.py (quantiser bit-exact vs the oracle, GEMM / QKV epilogue vs dequantised fp32 matmul), "
                          "tests/test_gpu_fullsize.py::test_fp8_full_depth_four_steps (latents rel-L2 vs the fp32 oracle 1.1e-2 .. 2.3e-2 per "
                          "step, image mean |diff| 1.47/255; operands of proj / mlp.2 / linear2 emitted by the attention / GELU epilogues; the reference's own --quantize is 4-bit weights, txt2image.py:79-82)"}
            quant["flag"] = "txt2image.py --quantize --quantize-bits 8 / Flux.quantize()"
            quant4["fp8"] = quant
            quant = quant4
        except Exception as exc:  # noqa: BLE001
            quant = {"error": repr(exc)[:300]}
            try:
                pipe.flow.dequantize()
            except Exception:  # noqa: BLE001
                pass

    if rank == 0:
        pk = peaks()
        gem = {"launches": 0, "ms": 0.0, "tflop": 0.0}
        for name in ("gemm", "gemm_qkv"):
            if name in ksum:
                for key in gem:
                    gem[key] += ksum[name][key]
        achieved = gem["tflop"] / (gem["ms"] * 1e-3) if gem["ms"] > 0 else 0.0
        step_ms = ms / args.steps
        traffic = measured_traffic()
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic", "config": workload_config(world),
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": pk["tflops"], "unit": "TFLOP/s",
                         "frac": achieved / pk["tflops"], "traffic": traffic.get("linear1", {}).get("dram_bytes"),
                         "peak_source": pk["src"],
                         "traffic_note": "DRAM bytes (read + write) per launch of the dominant member (single-block linear1, "
                                         "34816x21504x3072; algorithmic 1.84e9), from the committed ncu capture of this build: "
                                         "profiles/gemm_traffic.json (all members under `traffic_members`)",
                         "traffic_members": traffic,
                         "kernel": "fx::gemm_kernel<BN,EPI,CONV> (tcgen05 GEMM family: all Linear layers of the MMDiT + VAE 1x1)",
                         "launches_timed": gem["launches"], "share_of_step": gem["ms"] / eager_ms,
                         "note": "per-launch CUDA events over an eager (non-graph) repeat of the timed steps"},
            "kernels": {k: {"launches": v["launches"], "ms_per_step": v["ms"] / args.steps, "tflops": v["tflops"],
                            "frac_of_peak": v["tflops"] / pk["tflops"]} for k, v in ksum.items()},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(x_T_host.numel() * 2),
                    "d2h_bytes_per_step": int(out_host.numel()), "ms_per_step": max(e2e_ms, e2e_wall) / args.steps},
            "gpu_launches": int(launches), "cuda_graph": bool(use_graph), "eager_ms_per_step": eager_ms / args.steps, "clocks": clocks,
            # BASELINE metric's second half: per-denoise-step ms (one MMDiT forward + Euler update over the 8 images)
            "ms_per_denoise_step": denoise_ms, "ms_vae_decode_batch": decode_ms,
            # a new prompt every step (tokenise + T5 + CLIP + txt_in + modulation table + 4 steps + decode, host buffers)
            "e2e_cold_prompt": {"value": cold_value, "unit": UNIT, "ms_per_step": max(c_ms, c_wall) / args.steps,
                                "vs_warm": cold_value / e2e_value, "graph_recaptured": bool(recaptured)},
            "ms_text_encode": text_ms,
            "quantized": quant,
        }
        if not args.no_cpu:
            threads = os.cpu_count() or 1
            run, to_spi, desc = cpu_sample(threads)
            tb, tv = run()
            line["cpu_baseline"] = {"value": 1.0 / to_spi(tb, tv), "unit": UNIT, "cores": threads, "kind": "port", "sample": desc}
        else:
            line["cpu_baseline"] = None
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-graph", action="store_true", help="do not replay the MMDiT forward from a CUDA graph")
    ap.add_argument("--no-quantized", action="store_true", help="skip the FP8 (--quantize) leg")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
    else:
        main_arm(args)
