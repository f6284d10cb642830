"""ctypes binding of libkwb200.so (include/kwb200.h).  No fallback: if the library is missing or a call fails, raise."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# KW_LIB_VARIANT=timing selects an instrumented build (libkwb200_timing.so, built by `build.py --variant timing ...`)
LIB_PATH = os.path.join(_HERE, "libkwb200" + ("_" + os.environ["KW_LIB_VARIANT"] if os.environ.get("KW_LIB_VARIANT") else "") + ".so")

KW_F32, KW_BF16 = 0, 1
PROF_ENC_GEMM, PROF_ENC_ATTN, PROF_XKV_GEMM, PROF_DEC_GEMM, PROF_DEC_CROSS, PROF_LOGMEL, PROF_DEC_PASS = range(7)
vp, i32, i64, f32p = C.c_void_p, C.c_int32, C.c_int64, C.c_void_p


class KwError(RuntimeError):
    pass


class kw_config(C.Structure):
    _fields_ = [(n, i32) for n in ("vocab_size", "n_mels", "d_model", "n_heads", "ffn_dim", "enc_layers", "dec_layers",
                                   "max_source_pos", "max_target_pos", "dtype", "max_batch")]


class kw_enc_layer_weights(C.Structure):
    _fields_ = [(n, vp) for n in ("ln1_w", "ln1_b", "wqkv", "bqkv", "wo", "bo", "ln2_w", "ln2_b", "w1", "b1", "w2", "b2")]


class kw_dec_layer_weights(C.Structure):
    _fields_ = [(n, vp) for n in ("ln1_w", "ln1_b", "wqkv", "bqkv", "wo", "bo", "lnx_w", "lnx_b", "wq_x", "bq_x", "wkv_x",
                                  "bkv_x", "wo_x", "bo_x", "ln3_w", "ln3_b", "w1", "b1", "w2", "b2")]


class kw_weights(C.Structure):
    _fields_ = [(n, vp) for n in ("conv1_w", "conv1_b", "conv2_w", "conv2_b", "enc_pos", "enc_ln_w", "enc_ln_b",
                                  "tok_embed", "dec_pos", "dec_ln_w", "dec_ln_b")] + \
               [("enc", C.POINTER(kw_enc_layer_weights)), ("dec", C.POINTER(kw_dec_layer_weights))]


class kw_token_rules(C.Structure):
    _fields_ = [("eos_token_id", i32), ("pad_token_id", i32), ("no_timestamps_token_id", i32),
                ("max_initial_timestamp_index", i32), ("suppress_tokens", C.POINTER(i32)), ("n_suppress", i32),
                ("begin_suppress_tokens", C.POINTER(i32)), ("n_begin_suppress", i32)]


# name -> (restype, argtypes); every symbol include/kwb200.h declares
SIGNATURES = {
    "kw_last_error": (C.c_char_p, []),
    "kw_version": (C.c_char_p, []),
    "kw_logmel": (i32, [vp, vp, i32, i32, i32, vp, vp, vp]),
    "kw_logmel_windows": (i32, [vp, vp, vp, i32, i32, i32, vp, vp, vp]),
    "kw_mel_filterbank": (i32, [i32, vp]),
    "kw_model_create": (i32, [C.POINTER(kw_config), C.POINTER(kw_weights), C.POINTER(kw_token_rules), C.POINTER(vp)]),
    "kw_model_destroy": (None, [vp]),
    "kw_model_workspace_bytes": (i64, [vp]),
    "kw_encode": (i32, [vp, vp, i32, vp, vp]),
    "kw_set_encoder_output": (i32, [vp, vp, i32, vp]),
    "kw_cross_kv": (i32, [vp, i32, vp]),
    "kw_decode_step": (i32, [vp, vp, i32, i32, i32, i32, i32, i32, vp, vp, vp]),
    "kw_greedy_pass": (i32, [vp, i32, C.POINTER(i32), i32, i32, i32, i32, vp, vp]),
    "kw_encode_decode": (i32, [vp, vp, i32, vp, i32, C.POINTER(i32), i32, i32, i32, i32, vp, vp]),
    "kw_decoder_forward": (i32, [vp, vp, i32, i32, vp, vp]),
    "kw_attention": (i32, [vp, vp, vp, vp, i32, i32, i32, i32, i64, i64, i64, i64, i64, i64, i32, vp]),
    "kw_linear": (i32, [vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, vp]),
    "kw_layernorm": (i32, [vp, vp, vp, vp, i32, i32, i32, vp]),
    "kw_sample": (i32, [vp, vp, vp, i32, i32, i32, i32, i32, vp, vp]),
    "kw_set_gemm_impl": (None, [i32]),
    "kw_launch_count": (i64, [i32]),
    "kw_set_gemm_2cta": (None, [i32]),
    "kw_set_decode_impl": (None, [i32]),
    "kw_set_sample_fused": (None, [i32]),
    "kw_set_decode_graph": (None, [i32]),
    "kw_debug_attention_desc": (None, [i32, i32, i32]),
    "kw_debug_gemm_stamps": (None, [vp]),
    "kw_profile_enable": (None, [C.c_uint32]),
    "kw_profile_read": (i32, [i32, C.POINTER(C.c_double), C.POINTER(i64), C.POINTER(C.c_double), i32]),
}

_lib = None


def load() -> C.CDLL:
    """Load libkwb200.so once.  Raises KwError (never falls back) if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise KwError(f"{LIB_PATH} not found: build it with `python -m kotoba_whisper_b200.build` "
                          "(there is no CPU / PyTorch fallback for this path)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def profile_read(category: int, reset: bool = True):
    """-> (device ms, launches, algorithmic work) of one kernel category since the last reset."""
    ms, n, work = C.c_double(), i64(), C.c_double()
    check(load().kw_profile_read(category, C.byref(ms), C.byref(n), C.byref(work), int(reset)), "kw_profile_read")
    return ms.value, n.value, work.value


def check(rc: int, what: str = "") -> int:
    if rc < 0:
        msg = load().kw_last_error().decode(errors="replace")
        raise KwError(f"{what or 'kwb200'} failed ({rc}): {msg}")
    return rc
