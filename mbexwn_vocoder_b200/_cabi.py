"""ctypes binding of include/mbexwn.h (the C-ABI of libmbexwn_b200.so).

There is no fallback: if the shared library is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os

_LIB = None
LIB_NAME = "libmbexwn_b200.so"
ABI_VERSION = 3
MAX_LAYERS = 64
MAX_OPS = 32
MAX_BLOCKS = 4
N_STAGES = 7
STAGE_NAMES = ("f0_net", "excitation", "cond_conv", "wavenet", "post_pqmf", "vtf_net", "stft_ola")

PREC_FP32_SIMT, PREC_BF16X3, PREC_BF16, PREC_F16F8 = 0, 1, 2, 3
PRECISIONS = {"fp32": PREC_FP32_SIMT, "fp32_simt": PREC_FP32_SIMT, "bf16x3": PREC_BF16X3, "bf16": PREC_BF16,
              "f16f8": PREC_F16F8}

OK, ERR_INVALID, ERR_MISSING, ERR_CUDA, ERR_UNSUPPORTED = 0, -1, -2, -3, -4


class Op(C.Structure):
    _fields_ = [("kind", C.c_int32), ("k", C.c_int32), ("cin", C.c_int32), ("cout", C.c_int32),
                ("dilation", C.c_int32), ("pad_l", C.c_int32), ("pad_r", C.c_int32), ("pad_mode", C.c_int32),
                ("subpixel", C.c_int32), ("up", C.c_int32), ("act", C.c_int32), ("act_channels", C.c_int32),
                ("rate_in", C.c_int32), ("rate_out", C.c_int32), ("ch_out", C.c_int32),
                ("name", C.c_char * 96), ("act_name", C.c_char * 96)]


class WnBlock(C.Structure):
    _fields_ = [("c", C.c_int32), ("cond_conv_up", C.c_int32), ("up", C.c_int32), ("reserved", C.c_int32),
                ("name", C.c_char * 96), ("up_name", C.c_char * 96)]


class Config(C.Structure):
    _fields_ = [("abi_version", C.c_int32),
                ("sample_rate", C.c_int32), ("hop", C.c_int32), ("mel_channels", C.c_int32),
                ("pulse_per_frame", C.c_int32), ("steps_per_frame", C.c_int32),
                ("pulse_channels", C.c_int32), ("subbands", C.c_int32),
                ("pulse_rate", C.c_float), ("f0_min", C.c_float), ("f0_max", C.c_float), ("f0_span", C.c_float),
                ("noise_sigma", C.c_float), ("leaky_alpha", C.c_float),
                ("n_pp_ops", C.c_int32), ("n_ps_ops", C.c_int32),
                ("pp_ops", Op * MAX_OPS), ("ps_ops", Op * MAX_OPS),
                ("wn_c", C.c_int32), ("wn_cin", C.c_int32), ("wn_cout", C.c_int32), ("wn_layers", C.c_int32),
                ("wn_k", C.c_int32), ("wn_gate", C.c_int32),
                ("wn_cond_k", C.c_int32), ("wn_cond_conv_up", C.c_int32), ("wn_cond_lin_up", C.c_int32),
                ("wn_dilations", C.c_int32 * MAX_LAYERS),
                ("wn_name", C.c_char * 96), ("post_name", C.c_char * 96),
                ("n_ceps", C.c_int32), ("stft_win", C.c_int32), ("fft_size", C.c_int32),
                ("n_lifters", C.c_int32), ("n_smooth", C.c_int32),
                ("filter_max_log_range", C.c_float),
                ("wt_n_period", C.c_int32), ("wt_n_tables", C.c_int32),
                ("wt_nominal_f0", C.c_float), ("wt_min_transposition", C.c_float),
                ("wt_max_transposition", C.c_float), ("wt_grid_norm", C.c_float),
                ("cumsum_chunk", C.c_int32),
                ("pqmf_q", C.c_int32), ("pqmf_back", C.c_int32),
                ("halo_frames", C.c_int32),
                ("norm_enable", C.c_int32), ("norm_iters", C.c_int32), ("norm_win", C.c_int32),
                ("norm_smooth_win", C.c_int32), ("norm_proj_cols", C.c_int32), ("norm_use_max_limit", C.c_int32),
                ("norm_fact", C.c_float), ("norm_floor", C.c_float), ("norm_compress_exp", C.c_float),
                ("norm_proj_scale", C.c_float), ("norm_lin_scale", C.c_float), ("norm_lin_off", C.c_float),
                ("norm_mel_scale", C.c_float),
                ("ps_mode", C.c_int32), ("ps_preserve_energy", C.c_int32), ("wt_subharm", C.c_int32),
                ("pulse_pqmf_taps", C.c_int32), ("wn_causal", C.c_int32),
                ("wn_n_blocks", C.c_int32), ("wn_blocks", WnBlock * MAX_BLOCKS)]


class Batch(C.Structure):
    _fields_ = [("n_utt", C.c_int32), ("n_frames", C.c_int32), ("n_chunks", C.c_int32),
                ("frame_utt", C.c_void_p), ("utt_begin", C.c_void_p), ("utt_end", C.c_void_p),
                ("chunk_first", C.c_void_p), ("mel", C.c_void_p), ("noise", C.c_void_p),
                ("f0_override", C.c_void_p), ("utt_ids", C.c_void_p), ("seed", C.c_uint64), ("out", C.c_void_p),
                ("phase_carry", C.c_void_p)]


class AnalysisConfig(C.Structure):
    _fields_ = [("hop", C.c_int32), ("win", C.c_int32), ("fft_size", C.c_int32), ("n_mel", C.c_int32),
                ("mode", C.c_int32), ("lin_scale", C.c_float), ("lin_off", C.c_float), ("log_scale", C.c_float),
                ("floor", C.c_float), ("window", C.c_void_p), ("twiddle", C.c_void_p), ("mel_lo", C.c_void_p),
                ("mel_cnt", C.c_void_p), ("mel_off", C.c_void_p), ("mel_w", C.c_void_p)]


class AnalysisBatch(C.Structure):
    _fields_ = [("n_utt", C.c_int32), ("n_pairs", C.c_int32), ("n_frames", C.c_int32), ("n_samples_total", C.c_int64),
                ("sample_begin", C.c_void_p), ("n_samples", C.c_void_p), ("frame_begin", C.c_void_p),
                ("pair_first", C.c_void_p), ("audio", C.c_void_p), ("mel", C.c_void_p), ("mag_tap", C.c_void_p)]


# every symbol include/mbexwn.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "mbexwn_abi_version": (C.c_int, []),
    "mbexwn_create": (C.c_int, [C.POINTER(Config), C.POINTER(C.c_void_p)]),
    "mbexwn_destroy": (None, [C.c_void_p]),
    "mbexwn_last_error": (C.c_char_p, [C.c_void_p]),
    "mbexwn_set_tensor": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_size_t]),
    "mbexwn_set_scalar": (C.c_int, [C.c_void_p, C.c_char_p, C.c_float]),
    "mbexwn_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32]),
    "mbexwn_forward": (C.c_int, [C.c_void_p, C.POINTER(Batch), C.c_int32, C.c_void_p, C.c_size_t, C.c_void_p]),
    "mbexwn_forward_host": (C.c_int, [C.c_void_p, C.POINTER(Batch), C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_size_t, C.c_void_p]),
    "mbexwn_forward_host_begin": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(Batch), C.c_int32, C.c_void_p, C.c_void_p,
                                            C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "mbexwn_forward_host_wait": (C.c_int, [C.c_void_p, C.c_int32]),
    "mbexwn_tap": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int32, C.c_int32, C.c_int32,
                             C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "mbexwn_last_launch_count": (C.c_int, [C.c_void_p]),
    "mbexwn_set_option": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int32]),
    "mbexwn_get_info": (C.c_int, [C.c_void_p, C.c_char_p, C.POINTER(C.c_int32)]),
    "mbexwn_stage_ms": (C.c_int, [C.c_void_p, C.POINTER(C.c_float)]),
    "mbexwn_wavenet_launch_ms": (C.c_int, [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_int32)]),
    "mbexwn_phase_carry": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "mbexwn_gather_rows": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]),
    "mbexwn_range_status": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32), C.c_int32]),
    "mbexwn_tc_trace_read": (C.c_int64, [C.c_void_p, C.c_void_p, C.c_int64]),
    "mbexwn_k_conv1d": (C.c_int, [C.c_void_p, C.POINTER(Batch), C.POINTER(Op), C.c_int32, C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mbexwn_k_tc_gemm": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int32, C.c_int32,
                                   C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "mbexwn_k_tc_gemm_f16f8": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int32, C.c_int32,
                                         C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "mbexwn_mel_analysis": (C.c_int, [C.POINTER(AnalysisConfig), C.POINTER(AnalysisBatch), C.c_void_p]),
    "mbexwn_mel_analysis_host": (C.c_int, [C.POINTER(AnalysisConfig), C.POINTER(AnalysisBatch), C.c_void_p, C.c_void_p,
                                           C.c_void_p]),
    "mbexwn_global_error": (C.c_char_p, []),
    "mbexwn_k_lininterp": (C.c_int, [C.c_void_p, C.POINTER(Batch), C.POINTER(Op), C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p]),
}


def lib_path() -> str:
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), LIB_NAME)


def load():
    """dlopen the in-tree shared library and type every entry point.  Raises if it has not been built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        raise RuntimeError(f"{path} is missing: build it with mbexwn_vocoder_b200/csrc/build.sh "
                           f"(or __graft_entry__.build()); there is no CPU fallback")
    lib = C.CDLL(path)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.mbexwn_abi_version() != ABI_VERSION:
        raise RuntimeError("libmbexwn_b200.so ABI version mismatch; rebuild")
    _LIB = lib
    return lib


_EXC = {ERR_INVALID: RuntimeError, ERR_MISSING: KeyError, ERR_CUDA: RuntimeError, ERR_UNSUPPORTED: NotImplementedError}


def check(lib, handle, rc: int, what: str):
    if rc == OK:
        return
    msg = lib.mbexwn_last_error(handle).decode() if handle else lib.mbexwn_global_error().decode()
    raise _EXC.get(rc, RuntimeError)(f"{what} failed ({rc}): {msg}")
