"""B200-native drop-in for the reference's ``flux`` package (flux/__init__.py:3-16).

Same import surface (``from flux import FluxPipeline, FluxSampler, load_* ...``); the denoising hot
path runs in ``libflux_b200.so`` (hand-written sm_100a CUDA behind the C ABI in include/flux_b200.h),
which is loaded on first use and fails loudly when missing -- there is no CPU fallback.
"""
from .flux import FluxPipeline  # noqa: F401
from .lora import load_adapter  # noqa: F401
from .sampler import FluxSampler  # noqa: F401
from .specs import AutoEncoderParams, CLIPTextModelConfig, FluxParams, T5Config  # noqa: F401
from .utils import (load_ae, load_clip, load_clip_tokenizer, load_flow_model, load_t5,  # noqa: F401
                    load_t5_tokenizer, save_config)
