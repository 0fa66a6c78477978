"""ctypes binding of libflux_b200.so (C ABI: include/flux_b200.h).

The library is loaded on first use and there is NO fallback: a missing library, a missing symbol or
a non-sm_100 device raises immediately.  Argument errors reported by the library surface as
ValueError (the conditions the reference raises ValueError for), CUDA errors as RuntimeError.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

_LIB_NAME = "libflux_b200.so"
_lib = None

c_i32, c_i64, c_f32, c_vp = C.c_int32, C.c_int64, C.c_float, C.c_void_p


class GemmArgs(C.Structure):
    _fields_ = [("A", c_vp), ("lda", c_i64), ("a_bs", c_i64), ("W", c_vp), ("ldw", c_i64), ("bias", c_vp),
                ("out", c_vp), ("ldo", c_i64), ("out_bs", c_i64), ("out_f32", c_i32), ("act", c_i32),
                ("gate", c_vp), ("gate_bs", c_i64), ("resid", c_vp), ("ldr", c_i64), ("resid_bs", c_i64),
                ("batch", c_i32), ("rows", c_i32), ("N", c_i32), ("K", c_i32),
                ("fp8", c_i32), ("a_scale", c_vp), ("a_scale_bs", c_i64), ("w_scale", c_vp)]


class QkvArgs(C.Structure):
    _fields_ = [("A", c_vp), ("lda", c_i64), ("a_bs", c_i64), ("W", c_vp), ("ldw", c_i64), ("bias", c_vp),
                ("q_scale", c_vp), ("k_scale", c_vp), ("pe", c_vp), ("q", c_vp), ("k", c_vp), ("v", c_vp),
                ("mlp_out", c_vp), ("ld_mlp", c_i64), ("mlp_bs", c_i64), ("rms_eps", c_f32),
                ("batch", c_i32), ("rows", c_i32), ("N", c_i32), ("K", c_i32), ("heads", c_i32),
                ("seq_total", c_i32), ("seq_off", c_i32),
                ("fp8", c_i32), ("a_scale", c_vp), ("a_scale_bs", c_i64), ("w_scale", c_vp), ("pe_blocked", c_i32),
                ("qkv_fp8", c_i32)]


class ConvArgs(C.Structure):
    _fields_ = [("x", c_vp), ("W", c_vp), ("bias", c_vp), ("out", c_vp), ("out_f32", c_i32), ("resid", c_vp),
                ("batch", c_i32), ("H", c_i32), ("Wd", c_i32), ("Cin", c_i32), ("Cout", c_i32), ("gn_partials", c_vp), ("upsample2x", c_i32)]


class AttnArgs(C.Structure):
    _fields_ = [("q", c_vp), ("k", c_vp), ("v", c_vp), ("out", c_vp), ("ld_out", c_i64), ("out_bs", c_i64),
                ("scale", c_f32), ("batch", c_i32), ("heads", c_i32), ("seq", c_i32), ("variant", c_i32), ("fp8", c_i32),
                ("q_out", c_vp), ("sf_out", c_vp), ("e_out", c_vp), ("out_kc", c_i32), ("out_col0", c_i32),
                ("q_out2", c_vp), ("sf_out2", c_vp), ("e_out2", c_vp), ("out_kc2", c_i32), ("out_split", c_i32)]


class AttnSmallArgs(C.Structure):
    _fields_ = [("q", c_vp), ("k", c_vp), ("v", c_vp), ("ld", c_i64), ("bs", c_i64), ("bias", c_vp),
                ("out", c_vp), ("ld_out", c_i64), ("out_bs", c_i64), ("scale", c_f32), ("batch", c_i32),
                ("heads", c_i32), ("seq", c_i32), ("causal", c_i32)]


class RowNormArgs(C.Structure):
    _fields_ = [("x", c_vp), ("ldx", c_i64), ("x_bs", c_i64), ("out", c_vp), ("ldo", c_i64), ("out_bs", c_i64),
                ("p0", c_vp), ("p1", c_vp), ("p_bs", c_i64), ("eps", c_f32), ("mode", c_i32), ("batch", c_i32),
                ("rows", c_i32), ("D", c_i32), ("out_fp8", c_i32), ("scale_out", c_vp), ("scale_bs", c_i64), ("sf_out", c_vp)]


class QuantArgs(C.Structure):
    _fields_ = [("x", c_vp), ("ldx", c_i64), ("x_bs", c_i64), ("q", c_vp), ("ldq", c_i64), ("q_bs", c_i64),
                ("scale", c_vp), ("scale_bs", c_i64), ("batch", c_i32), ("rows", c_i32), ("K", c_i32)]


class Quant4Args(C.Structure):
    _fields_ = [("x", c_vp), ("ldx", c_i64), ("x_bs", c_i64), ("q", c_vp), ("sf", c_vp), ("scale", c_vp),
                ("batch", c_i32), ("rows", c_i32), ("K", c_i32)]


class Gemm4Args(C.Structure):
    _fields_ = [("A", c_vp), ("sfa", c_vp), ("a_scale", c_vp), ("W", c_vp), ("sfw", c_vp), ("w_scale", c_vp), ("bias", c_vp),
                ("out", c_vp), ("ldo", c_i64), ("out_bs", c_i64), ("out_f32", c_i32), ("act", c_i32),
                ("gate", c_vp), ("gate_bs", c_i64), ("resid", c_vp), ("ldr", c_i64), ("resid_bs", c_i64),
                ("batch", c_i32), ("rows", c_i32), ("N", c_i32), ("K", c_i32),
                ("q_out", c_vp), ("sf_out", c_vp), ("e_out", c_vp), ("out_kc", c_i32), ("out_col0", c_i32)]


class Quant4cArgs(C.Structure):
    _fields_ = [("x", c_vp), ("ldx", c_i64), ("x_bs", c_i64), ("batch", c_i32), ("rows", c_i32), ("C", c_i32),
                ("q", c_vp), ("sf", c_vp), ("e", c_vp), ("kc", c_i32), ("col0", c_i32)]


class Gemm4QkvArgs(C.Structure):
    _fields_ = [("A", c_vp), ("sfa", c_vp), ("a_scale", c_vp), ("W", c_vp), ("sfw", c_vp), ("w_scale", c_vp), ("bias", c_vp),
                ("q_scale", c_vp), ("k_scale", c_vp), ("pe", c_vp), ("pe_blocked", c_i32),
                ("q", c_vp), ("k", c_vp), ("v", c_vp), ("qkv_fp8", c_i32), ("rms_eps", C.c_float),
                ("batch", c_i32), ("rows", c_i32), ("K", c_i32), ("heads", c_i32), ("seq_total", c_i32), ("seq_off", c_i32)]


class GemvArgs(C.Structure):
    _fields_ = [("in_", c_vp), ("ld_in", c_i64), ("W", c_vp), ("ldw", c_i64), ("bias", c_vp), ("add", c_vp),
                ("ld_add", c_i64), ("out", c_vp), ("ld_out", c_i64), ("batch", c_i32), ("N", c_i32), ("K", c_i32),
                ("silu_in", c_i32), ("silu_out", c_i32)]


# every symbol include/flux_b200.h declares: (name, restype, argtypes)
SYMBOLS = {
    "fx_version": (C.c_int, []),
    "fx_last_error": (C.c_char_p, []),
    "fx_launch_count": (C.c_uint64, []),
    "fx_check_device": (C.c_int, [C.c_int]),
    "fx_gemm": (C.c_int, [C.POINTER(GemmArgs), c_vp]),
    "fx_gemm_qkv": (C.c_int, [C.POINTER(QkvArgs), c_vp]),
    "fx_quantize_rows_fp4": (C.c_int, [C.POINTER(Quant4Args), c_vp]),
    "fx_gemm_fp4": (C.c_int, [C.POINTER(Gemm4Args), c_vp]),
    "fx_gemm_fp4_qkv": (C.c_int, [C.POINTER(Gemm4QkvArgs), c_vp]),
    "fx_quantize_chunks_fp4": (C.c_int, [C.POINTER(Quant4cArgs), c_vp]),
    "fx_fp4_finalize": (C.c_int, [c_vp, c_vp, c_vp, c_i64, c_i32, c_vp]),
    "fx_conv3x3": (C.c_int, [C.POINTER(ConvArgs), c_vp]),
    "fx_conv3x3_gn_blocks": (C.c_int64, [c_i32, c_i32, c_i32, c_i32]),
    "fx_attention": (C.c_int, [C.POINTER(AttnArgs), c_vp]),
    "fx_attention_small": (C.c_int, [C.POINTER(AttnSmallArgs), c_vp]),
    "fx_rownorm": (C.c_int, [C.POINTER(RowNormArgs), c_vp]),
    "fx_quantize_rows": (C.c_int, [C.POINTER(QuantArgs), c_vp]),
    "fx_gemv": (C.c_int, [C.POINTER(GemvArgs), c_vp]),
    "fx_timestep_embedding": (C.c_int, [c_vp, c_vp, c_i32, c_i32, c_vp]),
    "fx_euler_step": (C.c_int, [c_vp, c_vp, c_f32, c_i64, c_vp]),
    "fx_patchify": (C.c_int, [c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_vp]),
    "fx_prior_packed": (C.c_int, [c_vp, c_i32, c_i32, c_i32, c_i32, C.c_uint64, c_i32, c_vp]),
    "fx_unpatchify_scale": (C.c_int, [c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_i32, c_f32, c_f32, c_vp]),
    "fx_groupnorm_partials_count": (C.c_int64, [c_i32, c_i64]),
    "fx_groupnorm_stats": (C.c_int, [c_vp, c_vp, c_i32, c_i64, c_i32, c_vp]),
    "fx_groupnorm_finalize": (C.c_int, [c_vp, c_vp, c_i32, c_i64, c_i32, c_f32, c_vp]),
    "fx_groupnorm_finalize_blocks": (C.c_int, [c_vp, c_vp, c_i32, c_i64, c_i64, c_i32, c_f32, c_vp]),
    "fx_groupnorm_apply": (C.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_i32, c_i64, c_i32, c_i32, c_vp]),
    "fx_upsample2x": (C.c_int, [c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_vp]),
    "fx_softmax_rows": (C.c_int, [c_vp, c_i64, c_vp, c_i64, c_i64, c_i32, c_f32, c_vp]),
    "fx_transpose": (C.c_int, [c_vp, c_i64, c_vp, c_i64, c_i32, c_i32, c_vp]),
    "fx_finish_image": (C.c_int, [c_vp, c_vp, c_vp, c_i64, c_vp]),
    "fx_embedding": (C.c_int, [c_vp, c_vp, c_vp, c_vp, c_i64, c_i32, c_i32, c_vp]),
    "fx_act_mul": (C.c_int, [c_vp, c_vp, c_vp, c_i64, c_i32, c_vp]),
}


# include/flux_b200_dbg.h: the test-only companion library (bring-up probes, CUDA-core reference GEMM)
DBG_SYMBOLS = {
    "fx_dbg_gemm_ref": (C.c_int, [c_vp, c_i64, c_vp, c_i64, c_vp, c_i64, c_i32, c_i32, c_i32, c_vp]),
    "fx_dbg_umma_tile": (C.c_int, [c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, C.c_uint32, C.c_uint32,
                                   C.c_uint32, c_vp]),
    "fx_dbg_mma_pattern": (C.c_int, [c_i32, c_i32, c_vp, c_vp]),
    "fx_dbg_bs2_tile": (C.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_vp]),
    "fx_dbg_bs_tile": (C.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, C.c_uint32, C.c_uint32,
                                 C.c_uint32, C.c_uint32, C.c_uint32, c_vp]),
}
_dbg = None


def dbg_lib():
    """libflux_b200_dbg.so (tests / profiling scripts only; the product path never loads it)."""
    global _dbg
    if _dbg is None:
        lib()  # the product library first: the companion links against it
        path = os.path.join(os.path.dirname(lib_path()), "libflux_b200_dbg.so")
        if not os.path.exists(path):
            raise RuntimeError(f"{path} not found: build it with `make -C flux-generator_b200/csrc`")
        d = C.CDLL(path)
        for name, (res, args) in DBG_SYMBOLS.items():
            fn = getattr(d, name)
            fn.restype = res
            fn.argtypes = args
        _dbg = d
    return _dbg


def lib_path() -> str:
    # FLUX_B200_LIB: an alternative build of the same library (A/B micro-benchmarks only)
    return os.environ.get("FLUX_B200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), _LIB_NAME)


def load():
    """dlopen the library and bind every declared symbol (raises if anything is missing)."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise RuntimeError(f"{path} not found: build it with `make -C flux-generator_b200/csrc` "
                           "(or __graft_entry__.build()); there is no CPU fallback")
    lib = C.CDLL(path)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


_device_checked = False


def lib():
    """Library handle for compute calls: additionally requires an sm_100 CUDA device."""
    global _device_checked
    l = load()
    if not _device_checked:
        if not torch.cuda.is_available():
            raise RuntimeError("flux_b200: no CUDA device; the hot path has no CPU fallback")
        check(l.fx_check_device(torch.cuda.current_device()))
        _device_checked = True
    return l


def check(rc: int) -> None:
    if rc == 0:
        return
    msg = (load().fx_last_error() or b"").decode()
    if rc == -1:
        raise ValueError(msg)
    raise RuntimeError(f"flux_b200 error {rc}: {msg}")


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def launch_count() -> int:
    return int(load().fx_launch_count())
