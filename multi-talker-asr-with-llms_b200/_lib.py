"""ctypes loader for libmtasr.so (the C ABI of include/mtasr.h).  Fails loudly when the library is missing."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmtasr.so")


class MtasrError(RuntimeError):
    pass


class GemmDesc(C.Structure):
    """Mirror of `struct mtasr_gemm_desc` (include/mtasr.h) -- field order and types must match exactly."""
    _fields_ = [
        ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32),
        ("batch0", C.c_int32), ("batch1", C.c_int32),
        ("a_major", C.c_int32), ("b_major", C.c_int32),
        ("block_n", C.c_int32),
        ("a", C.c_void_p),
        ("a_ld", C.c_int64), ("a_sb0", C.c_int64), ("a_sb1", C.c_int64),
        ("a_inner", C.c_int32), ("a_phase", C.c_int32),
        ("a_rows", C.c_int64),
        ("b", C.c_void_p),
        ("b_ld", C.c_int64), ("b_sb0", C.c_int64), ("b_sb1", C.c_int64),
        ("b_rows", C.c_int64),
        ("c", C.c_void_p),
        ("c_dtype", C.c_int32),
        ("c_ld", C.c_int64), ("c_sb0", C.c_int64), ("c_sb1", C.c_int64),
        ("aux", C.c_void_p),
        ("bias", C.c_void_p),
        ("bias_sb0", C.c_int64),
        ("residual", C.c_void_p),
        ("res_dtype", C.c_int32),
        ("r_ld", C.c_int64), ("r_sb0", C.c_int64), ("r_sb1", C.c_int64),
        ("act", C.c_int32),
        ("alpha", C.c_float),
        ("accumulate", C.c_int32),
        ("mode", C.c_int32),
        ("row_vec", C.c_void_p),
        ("row_scale", C.c_void_p),
        ("lse_part", C.c_void_p),
        ("drop_seed", C.c_void_p),
        ("drop_site", C.c_uint32),
        ("drop_keep16", C.c_uint32),
    ]


# name -> (restype, argtypes); every symbol declared in include/mtasr.h
_P, _I32, _I64, _F = C.c_void_p, C.c_int32, C.c_int64, C.c_float
SIGNATURES = {
    "mtasr_version": (C.c_int, []),
    "mtasr_set_sm_budget": (C.c_int, [_I32]),
    "mtasr_last_error_string": (C.c_char_p, []),
    "mtasr_launch_count": (_I64, []),
    "mtasr_profile_begin": (C.c_int, []),
    "mtasr_profile_end": (C.c_int, [C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "mtasr_gemm_bf16": (C.c_int, [C.POINTER(GemmDesc), _P]),
    "mtasr_gemm_n_tiles": (C.c_int, [_I32, _I32]),
    "mtasr_ctc_state_pad": (C.c_int, [_I32]),
    "mtasr_ctc_alpha_fwd": (C.c_int, [_P, _P, _P, _P, _P, _I32, _I32, _I32, _I32, _I32, _P, _P, _P, _P, _P]),
    "mtasr_ctc_beta_bwd": (C.c_int, [_P, _P, _P, _P, _P, _I32, _I32, _I32, _I32, _I32, _P, _P, _P, _P, _P, _P, _P]),
    "mtasr_lse_finalize": (C.c_int, [_P, _I64, _I32, _P, _P, _P]),
    "mtasr_ctc_collapse": (C.c_int, [_P, _I32, _I32, _I64, _I64, _P, _P, _P]),
    "mtasr_ctc_segments": (C.c_int, [_P, _P, _I32, _I32, _I64, _P, _P, _P, _P]),
    "mtasr_segment_mean_fwd": (C.c_int, [_P, _P, _P, _P, _P, _I32, _I32, _I32, _I32, _P, _P, _P]),
    "mtasr_segment_mean_bwd": (C.c_int, [_P, _P, _P, _P, _I32, _I32, _I32, _I32, _P, _P]),
    "mtasr_ctc_gather_cols": (C.c_int, [_P, _P, _P, _I32, _I32, _I32, _I32, _I32, _I64, _P, _P]),
    "mtasr_ctc_scatter_cols": (C.c_int, [_P, _P, _P, _I32, _I32, _I32, _I32, _I32, _I64, _P, _P]),
    "mtasr_ctc_gather_rows": (C.c_int, [_P, _P, _P, _P, _I32, _I32, _I32, _I32, _I64, _I64, _P, _P, _P]),
    "mtasr_ctc_scatter_rows": (C.c_int, [_P, _P, _P, _P, _I32, _I32, _I32, _I32, _I64, _I64, _P, _P, _P]),
    "mtasr_layernorm_fwd": (C.c_int, [_P, _I32, _P, _P, _F, _I64, _I32, _I32, _P, _P, _P, _P, _P]),
    "mtasr_layernorm_bwd": (C.c_int, [_P, _I32, _P, _I32, _P, _P, _P, _P, _I64, _I32, _P, _P, _P, _P, _P]),
    "mtasr_layernorm_bwd_sums": (C.c_int, [_P, _I32, _P, _I32, _P, _P, _P, _P, _I64, _I32, _P, _P, _P, _P, _P, _P, _P, _P]),
    "mtasr_cast_f32_bf16": (C.c_int, [_P, _P, _I64, _P]),
    "mtasr_softmax_from_logits": (C.c_int, [_P, _P, _P, _I64, _I32, _I64, _P, _P, _P]),
    "mtasr_weightnorm_fwd": (C.c_int, [_P, _P, _I64, _I32, _P, _P, _P]),
    "mtasr_weightnorm_bwd": (C.c_int, [_P, _P, _P, _P, _I64, _I32, _P, _P, _P, _P]),
    "mtasr_colsum": (C.c_int, [_P, _I32, _I64, _I32, _I64, _P, _P]),
    "mtasr_relpos_gate_fwd": (C.c_int, [_P, _I32, _P, _P, _P, _I32, _I32, _I32, _P, _P]),
    "mtasr_relpos_gate_bwd": (C.c_int, [_P, _I32, _P, _P, _P, _P, _I32, _I32, _I32, _P, _P, _P, _P, _P, _P]),
    "mtasr_attn_fwd": (C.c_int, [_P, _P, _P, _P, _I32, _I32, _I32, _F, _P, _P, _P, C.c_uint32, C.c_uint32, _P]),
    "mtasr_attn_bwd": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _I32, _I32, _I32, _F, _P, _P, _P, _P, _P, _P, C.c_uint32, C.c_uint32, _P]),
    "mtasr_attn_softmax_fwd": (C.c_int, [_P, _P, _P, _P, _I32, _I32, _I32, _I32, _F, _P, _P]),
    "mtasr_attn_softmax_bwd": (C.c_int, [_P, _P, _P, _P, _I32, _I32, _I32, _I32, _F, _P, _P, _P, _P]),
    "mtasr_pad_cast": (C.c_int, [_P, _I32, _I32, _I32, _I32, _I32, _I32, _P, _P, _P]),
    "mtasr_glu_fwd": (C.c_int, [_P, _I32, _I64, _I32, _P, _P, _P]),
    "mtasr_conv0_fwd": (C.c_int, [_P, _P, _P, _P, _P, _F, _I32, _I32, _I32, _I32, _I32, _I32, _P, _P, _P]),
    "mtasr_groupnorm_gelu": (C.c_int, [_P, _P, _P, _F, _I32, _I32, _I32, _P, _P, _P, _P]),
    "mtasr_groupnorm_gelu_f32": (C.c_int, [_P, _P, _P, _F, _I32, _I32, _I32, _P, _P, _P, _P]),
    "mtasr_split_labels": (C.c_int, [_P, _I32, _I32, _I64, _I32, _I64, _I64, _I32, _I64, _I32, _I64, _I32, _I32, _P, _P, _P, _P]),
    "mtasr_pcgrad_dots": (C.c_int, [_P, _P, _I64, _P, _P]),
    "mtasr_pcgrad_project": (C.c_int, [_P, _P, _I64, _P, _P]),
    "mtasr_dropout": (C.c_int, [_P, _I32, _I64, _I64, _P, C.c_uint32, C.c_uint32, _P, _I32, _P]),
    "mtasr_split_bf16": (C.c_int, [_P, _I64, _I64, _I32, _I32, _P, _P]),
    "mtasr_attn_softmax_fwd_split": (C.c_int, [_P, _P, _P, _P, _I32, _I32, _I32, _I32, _F, _I32, _P, _P]),
    "mtasr_act_bwd": (C.c_int, [_P, _I32, _P, _I32, _I64, _P, _P]),
    "mtasr_lstm_scratch_bytes": (C.c_int64, [_I32, _I32, _I32]),
    "mtasr_lstm_fwd": (C.c_int, [_P, _P, _I32, _I32, _I32, _I32, _P, _P, _P, _P, _P, _P]),
    "mtasr_lstm_bwd": (C.c_int, [_P, _P, _P, _P, _I32, _I32, _I32, _I32, _P, _P, _P]),
    "mtasr_lstm_fwd_f32": (C.c_int, [_P, _P, _I32, _I32, _I32, _I32, _P, _P, _P, _P]),
    "mtasr_lstm_bwd_f32": (C.c_int, [_P, _P, _I32, _I32, _I32, _I32, _P, _P, _P, _P, _P]),
    "mtasr_relu_bwd_f32": (C.c_int, [_P, _P, _I64, _P, _P]),
    "mtasr_softmax_scale_f32": (C.c_int, [_P, _P, _P, _I64, _I32, _I64, _P, _P]),
    "mtasr_glu_bwd": (C.c_int, [_P, _I32, _P, _I32, _I64, _I32, _P, _P]),
}

_lib = None


def load():
    """dlopen libmtasr.so and bind every entry point.  Raises MtasrError if the library has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MtasrError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  There is no CPU or PyTorch fallback for this path.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)           # AttributeError if the .so is stale
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().mtasr_last_error_string()
        raise MtasrError(f"{what} failed (code {rc}): {msg.decode() if msg else '?'}")
