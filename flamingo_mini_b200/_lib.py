"""ctypes binding of libflamingo_b200.so (C ABI declared in include/flamingo_b200.h).

The library is the only implementation of the hot path: if it cannot be loaded, importing the
modules that need it raises — there is no eager / CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

from . import _build

_lock = threading.Lock()
_lib = None

c_ll = C.c_longlong
c_vp = C.c_void_p


class GemmDesc(C.Structure):
    _fields_ = [("M", C.c_int), ("N", C.c_int), ("K", C.c_int),
                ("A", c_vp), ("lda", c_ll), ("a_mn", C.c_int),
                ("B", c_vp), ("ldb", c_ll), ("b_mn", C.c_int),
                ("epi", C.c_int),
                ("out", c_vp), ("ldo", c_ll), ("out_f32", C.c_int),
                ("out2", c_vp), ("ldo2", c_ll),
                ("aux", c_vp), ("ldaux", c_ll), ("aux_f32", C.c_int), ("aux2", c_vp), ("ldaux2", c_ll),
                ("col_bias", c_vp), ("gate", c_vp), ("red_out", c_vp),
                ("scale", C.c_float), ("act", C.c_int), ("bn", C.c_int), ("splits", C.c_int), ("splitk_flags", c_vp), ("trace", c_vp)]


class XattnCfg(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("B", "S", "D", "Dv", "n_media", "heads", "dim_head", "ff_inner", "act",
                                       "y_f32", "training")]


class XattnLayout(C.Structure):
    _fields_ = [(n, c_ll) for n in ("attn_norm_w", "attn_norm_b", "to_q", "to_kv", "to_out", "ffw_norm_w",
                                    "ffw_norm_b", "ffw_w1", "ffw_w2", "alpha_attn", "alpha_ffw", "total")]


class ResamplerCfg(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("BN", "T", "F", "Dv", "depth", "heads", "dim_head", "n_latents",
                                       "n_time_embeds", "ff_inner", "act", "x_f32", "training")]


class ResamplerLayout(C.Structure):
    _fields_ = [(n, c_ll) for n in ("latents", "time_pos_emb", "layer0", "layer_stride", "norm_media_w",
                                    "norm_media_b", "norm_latents_w", "norm_latents_b", "to_q", "to_k", "to_v",
                                    "to_out", "ffw_norm_w", "ffw_norm_b", "ffw_w1", "ffw_w2", "norm_w", "norm_b",
                                    "total")]


ACT_IDS = {"gelu": 0, "sqrelu": 1, "relu": 2}

# name -> (restype, argtypes); must list every symbol include/flamingo_b200.h declares
_P = C.POINTER
PROTOTYPES = {
    "fm_version": (C.c_int, []),
    "fm_last_error": (C.c_char_p, []),
    "fm_device_error": (C.c_uint, []),
    "fm_abi_sizes": (C.c_int, [_P(C.c_int)]),
    "fm_set_option": (C.c_int, [C.c_int, C.c_int]),
    "fm_launch_count": (C.c_ulonglong, []),
    "fm_profile_enable": (C.c_int, [C.c_int]),
    "fm_profile_report": (C.c_int, [C.c_char_p, C.c_size_t]),
    "fm_profile_log": (C.c_int, [C.c_char_p, C.c_size_t]),
    "fm_gemm_bf16": (C.c_int, [_P(GemmDesc), c_vp]),
    "fm_gemm_bf16_group": (C.c_int, [_P(GemmDesc), C.c_int, c_vp]),
    "fm_gemm_splitk_flag_ints": (C.c_size_t, [C.c_int, C.c_int]),
    "fm_layernorm_fwd": (C.c_int, [c_vp, C.c_int, c_vp, c_vp, c_vp, C.c_int, c_vp, c_vp, C.c_int, C.c_int, c_vp]),
    "fm_layernorm_bwd_scratch_bytes": (C.c_size_t, [C.c_int]),
    "fm_layernorm_bwd": (C.c_int, [c_vp, c_vp, C.c_int, c_vp, c_vp, c_vp, c_vp, C.c_int, c_vp, C.c_int, c_vp, c_vp,
                                   c_vp, C.c_int, C.c_int, c_vp]),
    "fm_text_time": (C.c_int, [c_vp, c_vp, C.c_int, C.c_int, c_vp]),
    "fm_cast_f32_to_bf16": (C.c_int, [c_vp, c_vp, c_ll, c_vp]),
    "fm_xattn_layout_of": (C.c_int, [_P(XattnCfg), _P(XattnLayout)]),
    "fm_xattn_saved_bytes": (C.c_size_t, [_P(XattnCfg)]),
    "fm_xattn_scratch_bytes": (C.c_size_t, [_P(XattnCfg)]),
    "fm_xattn_fwd": (C.c_int, [_P(XattnCfg), c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, C.c_int, c_vp, c_vp, c_vp]),
    "fm_xattn_bwd": (C.c_int, [_P(XattnCfg), c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp,
                               c_vp]),
    "fm_resampler_layout_of": (C.c_int, [_P(ResamplerCfg), _P(ResamplerLayout)]),
    "fm_resampler_saved_bytes": (C.c_size_t, [_P(ResamplerCfg)]),
    "fm_resampler_scratch_bytes": (C.c_size_t, [_P(ResamplerCfg)]),
    "fm_resampler_fwd": (C.c_int, [_P(ResamplerCfg), c_vp, c_vp, c_vp, c_vp, C.c_int, c_vp, c_vp]),
    "fm_resampler_bwd": (C.c_int, [_P(ResamplerCfg), c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
}


# entry points added in round 2 (validated on hardware)
LAYER_CB = C.CFUNCTYPE(None, c_vp, C.c_int)
PROTOTYPES.update({
    "fm_resampler_bwd_notify": (C.c_int, [_P(ResamplerCfg), c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, LAYER_CB, c_vp, c_vp]),
    "fm_side_join": (C.c_int, [c_vp]),
    "fm_get_option": (C.c_int, [C.c_int]),
    "fm_xattn_core_fwd": (C.c_int, [c_vp, c_vp, c_vp, c_vp, C.c_int, C.c_int, C.c_int, C.c_int, c_vp]),
    "fm_resampler_core_fwd": (C.c_int, [c_vp, c_vp, c_vp, c_vp, C.c_int, C.c_int, C.c_int, c_vp]),
    "fm_xattn_core_bwd": (C.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, c_vp]),
    "fm_resampler_core_bwd": (C.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, C.c_int, C.c_int, C.c_int, C.c_float, c_vp]),
    "fm_cross_entropy_fwd": (C.c_int, [c_vp, c_ll, C.c_int, C.c_int, c_vp, c_ll, c_vp, c_vp, c_vp]),
    "fm_cross_entropy_bwd": (C.c_int, [c_vp, c_ll, C.c_int, C.c_int, c_vp, c_ll, c_vp, c_vp, c_vp, c_vp]),
    "fm_adamw_step": (C.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_ll, C.c_float, C.c_float, C.c_float, C.c_float,
                                C.c_float, C.c_int, c_vp]),
})
STAGING_PROTOTYPES: dict = {}      # (kept for callers that enumerate both tables; empty since the round-2 promotion)


class FlamingoB200Error(RuntimeError):
    pass


def lib_path() -> str:
    """The library this process loads: libflamingo_b200.so."""
    return _build.lib_path()


def type_library(lib, staging: bool):
    """Attach the C-ABI prototypes to a loaded library and check the struct mirrors against fm_abi_sizes()."""
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)       # AttributeError here == ABI mismatch, fail loudly
        fn.restype = res
        fn.argtypes = args
    for name, (res, args) in STAGING_PROTOTYPES.items():
        fn = getattr(lib, name, None)
        if fn is not None:
            fn.restype = res
            fn.argtypes = args
        elif staging:
            raise FlamingoB200Error(f"libflamingo_b200.so does not export {name} (stale build?)")
    sizes = (C.c_int * 5)()
    lib.fm_abi_sizes(sizes)
    mirror = [C.sizeof(GemmDesc), C.sizeof(XattnCfg), C.sizeof(XattnLayout), C.sizeof(ResamplerCfg),
              C.sizeof(ResamplerLayout)]
    if list(sizes) != mirror:
        raise FlamingoB200Error(f"ABI struct size mismatch: library {list(sizes)} vs python mirror {mirror}")
    return lib


def load():
    """Load (building first if the in-tree .so is missing or stale) and type the C ABI."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        path = _build.lib_path()
        if _build.is_stale():
            try:
                path = _build.build()
            except Exception as e:  # keep a stale-but-present library usable on boxes without nvcc
                if not os.path.exists(path):
                    raise FlamingoB200Error(
                        f"{os.path.basename(path)} is missing and could not be built ({e}); "
                        "the sm_100a CUDA library is the only implementation of this path") from e
        lib = type_library(C.CDLL(path), staging=True)
        _apply_env_options(lib)
        _lib = lib
    return _lib


OPTION_KEYS = {"side_stream": 0, "gemm_group": 1, "epi_prefetch": 2, "alpha_from_dw2": 3, "pdl": 4, "ln_reduce_side": 5,
               "sm_reserve": 6, "dattn_from_gemm": 7, "attn_tmem_compact": 8, "defer_join": 9, "dw_splitk": 10}


def _apply_env_options(lib) -> None:
    """FM_B200_OPTS="pdl=1,gemm_group=0": scheduling switches (fm_set_option) for A/B runs.
    An option the loaded build does not know is an error, not a silent no-op."""
    spec = os.environ.get("FM_B200_OPTS", "").strip()
    if not spec:
        return
    for item in spec.split(","):
        name, _, val = item.partition("=")
        name = name.strip()
        if name not in OPTION_KEYS:
            raise FlamingoB200Error(f"FM_B200_OPTS: unknown option {name!r} (known: {sorted(OPTION_KEYS)})")
        rc = lib.fm_set_option(OPTION_KEYS[name], int(val))
        if rc != 0:
            raise FlamingoB200Error(f"FM_B200_OPTS: {os.path.basename(lib._name)} rejects option {name!r}: "
                                    f"{lib.fm_last_error().decode(errors='replace')}")


def has(name: str) -> bool:
    """True when the loaded build exports `name`."""
    return hasattr(load(), name) and (name in PROTOTYPES or name in STAGING_PROTOTYPES)


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().fm_last_error().decode(errors="replace")
        raise FlamingoB200Error(f"{what or 'libflamingo_b200'} failed (code {rc}): {msg}")
