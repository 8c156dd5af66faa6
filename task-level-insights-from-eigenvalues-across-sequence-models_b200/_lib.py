"""ctypes binding of libeigb200.so (include/eigb200.h).  There is no fallback: if the library is missing or a
call fails, this raises -- the CUDA path is the product."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("EIGB200_LIB") or os.path.join(HERE, "libeigb200.so")   # EIGB200_LIB: A/B-test another build of the same ABI

NSLOT = 8
OK, EINVAL, ECUDA, EUNSUPPORTED = 0, -1, -2, -3
CMP_F64, CMP_F32 = 0, 1
NORM_FN = {"exp": 0, "elu": 1, "softplus": 2, "sigmoid": 3}
RATIO_NONE, RATIO_NEXT_OVER_CUR, RATIO_CUR_OVER_NEXT = 0, 1, 2
F32, F64, BF16 = 0, 1, 2
EPI_NONE, EPI_GELU, EPI_GLU_RESIDUAL, EPI_RESIDUAL = 0, 1, 2, 3
GEMM_AUTO, GEMM_SIMT_F32, GEMM_TC_3XTF32, GEMM_TC_TF32, GEMM_TC_F16X3 = 0, 1, 2, 3, 4

_vp, _i, _i64, _f, _sz = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_size_t
_dp = C.POINTER(C.c_double)

# name -> argtypes (restype is int unless listed in _RESTYPES); mirrors include/eigb200.h one to one
SIGNATURES = {
    "eigb200_version": [],
    "eigb200_last_error": [],
    "eigb200_device_info": [_i, C.POINTER(_i), C.POINTER(_i), C.POINTER(_i)],
    "eigb200_set_device": [_i],
    "eigb200_zero_i32": [_vp, _vp, _sz],
    "eigb200_mamba2_eig": [_vp, _vp, _i, _i64, _i64, _i, _vp, _vp, _vp, _i, _vp, _i64, _vp, _dp, _i, _i, _vp, _f],
    "eigb200_mamba2_eig_partials": [_vp, _vp, _i, _i64, _i64, _vp, _vp, _vp, _i64, _vp, _dp, _i, _i, _vp, _f],
    "eigb200_linear_glu_extract": [_vp, _vp, _i64, _vp, _vp, _vp, _i64, _vp, _i64, _i64, _i, _i, _vp, _vp, _vp, _sz],
    "eigb200_mamba2_lti_eig": [_vp, _vp, _vp, _i64, _i64, _i, _vp, _i64, _vp, _dp, _i, _i],
    "eigb200_normattn_gate": [_vp, _vp, _i, _i64, _i64, _i, _vp, _vp, _vp, _i, _i, _vp],
    "eigb200_linattn_nu": [_vp, _vp, _vp, _i64, _i64, _i64, _i, _i, _vp],
    "eigb200_softmax_nu": [_vp, _vp, _vp, _i64, _i64, _i64, _i, _i, _vp, _vp],
    "eigb200_softmax_eta": [_vp, _vp, _vp, _i64, _i64, _i, _vp, _vp, _dp, _i],
    "eigb200_softmax_attn_forward": [_vp, _vp, _vp, _vp, _i64, _f, _vp, _i64, _i64, _i64, _i, _i, _i],
    "eigb200_ratio_hist": [_vp, _vp, _i, _i, _i64, _i64, _i64, _vp, _i64, _vp, _dp, _i, _i],
    "eigb200_count_moments": [_vp, _vp, _i64, _i64, _vp, _vp],
    "eigb200_count_moments_layers": [_vp, _vp, _i64, _i64, _i64, _vp, _vp],
    "eigb200_log_hist": [_vp, _vp, _i, _i64, _i64, _i64, C.c_double, C.c_double, _i, _vp],
    "eigb200_hist_quantiles": [_vp, _vp, _i64, C.c_double, C.c_double, _i, _vp, _i, _vp],
    "eigb200_stats_available": [],
    "eigb200_stats_comm_init_all": [_i, C.POINTER(_i), C.POINTER(_vp)],
    "eigb200_stats_comm_destroy": [_vp],
    "eigb200_stats_group_start": [],
    "eigb200_stats_group_end": [],
    "eigb200_stats_allreduce": [_vp, _vp, _vp, _sz],
    "eigb200_diag_scan": [_vp, _vp, _vp, _vp, _i64, _i64, _i, _i],
    "eigb200_ssd_scan": [_vp, _vp, _i64, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _i64, _vp, _i64, _i64, _i, _i, _i, _i],
    "eigb200_mamba_conv_ssd": [_vp, _vp, _i64, _vp, _vp, _i, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _i, _i, _i, _i],
    "eigb200_dplr_abar": [_vp, _vp, _vp, _vp, _vp, _i64, _i, _vp],
    "eigb200_eigvals_c64": [_vp, _vp, _i64, _i, _vp, _vp],
    "eigb200_linear_workspace_bytes": [_i, _i],
    "eigb200_linear_workspace_bytes_m": [_i64, _i, _i],
    "eigb200_linear": [_vp, _vp, _i64, _vp, _vp, _vp, _i64, _vp, _i64, _i64, _i, _i, _i, _i, _vp, _sz],
    "eigb200_out_glu_fused_supported": [_i, _i],
    "eigb200_out_glu_fused": [_vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _i64, _i64, _i, _i, _vp, _vp],
    "eigb200_mamba_front_fused_supported": [_i, _i, _i, _i, _i, _i],
    "eigb200_mamba_front_fused": [_vp, _vp, _i64, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _i, _i, _i],
    "eigb200_gemm_precision": [],
    "eigb200_set_gemm_precision": [_i],
    "eigb200_gemm_overflow": [_vp, _i, C.POINTER(_i)],
    "eigb200_embedding": [_vp, _vp, _vp, _vp, _vp, _i64, _i64, _i, _i64],
    "eigb200_embedding_stats": [_vp, _vp, _vp, _vp, _vp, _i64, _i64, _i, _i64, _vp, _f],
    "eigb200_rowstats": [_vp, _vp, _i64, _i, _f, _vp],
    "eigb200_linear_prepare": [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp, _sz],
    "eigb200_linear_ln": [_vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _i64, _i64, _i, _i, _i, _vp, _sz],
    "eigb200_layernorm": [_vp, _vp, _vp, _vp, _f, _vp, _i64, _i],
    "eigb200_conv_silu": [_vp, _vp, _i64, _vp, _vp, _i, _vp, _i64, _i64, _i64, _i],
    "eigb200_linattn_forward": [_vp, _vp, _vp, _vp, _i64, _vp, _i, _i, _f, _vp, _i64, _i64, _i64, _i, _i, _i],
    "eigb200_linattn_conv_fusable": [_vp, _vp, _vp, _i64, _vp, _i64, _i, _i, _i],
    "eigb200_linattn_forward_conv": [_vp, _vp, _vp, _vp, _i64, _vp, _i, _i, _f, _vp, _vp, _i, _i, _i, _i, _vp, _i64, _i64, _i64, _i, _i, _i],
    "eigb200_add": [_vp, _vp, _vp, _vp, _i64],
    "eigb200_mul_silu": [_vp, _vp, _vp, _vp, _i64],
    "eigb200_gelu": [_vp, _vp, _vp, _i64],
    "eigb200_scale_cols": [_vp, _vp, _vp, _vp, _i64, _i],
    "eigb200_lti_scale_b": [_vp, _vp, _i64, _i, _i, _vp, _i64, _i, _i],
    "eigb200_s4_kernel_workspace_bytes": [_i, _i],
    "eigb200_s4_kernel": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _sz],
    "eigb200_s4_causal_conv": [_vp, _vp, _vp, _vp, _vp, _i64, _i64, _i],
    "eigb200_ssm_lambda": [_vp, _i, _vp, _vp, _vp, _i, _vp],
}
_RESTYPES = {"eigb200_last_error": C.c_char_p, "eigb200_linear_workspace_bytes": C.c_size_t, "eigb200_linear_workspace_bytes_m": C.c_size_t,
             "eigb200_s4_kernel_workspace_bytes": C.c_size_t}


class Eigb200Error(RuntimeError):
    pass


_lib = None


def load() -> C.CDLL:
    """dlopen libeigb200.so and bind every symbol of the header.  Fails loudly when the library is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise Eigb200Error(
            "libeigb200.so is not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'` -- "
            "there is no CPU or PyTorch fallback for this path." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the header and the library disagree
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, C.c_int)
    _lib = lib
    return lib


def check(rc: int, what: str = ""):
    if rc != OK:
        msg = load().eigb200_last_error().decode(errors="replace")
        raise Eigb200Error("%s failed (%d): %s" % (what or "eigb200 call", rc, msg))


def thresholds_arg(thr):
    arr = (C.c_double * len(thr))(*[float(t) for t in thr])
    return arr, len(thr)
