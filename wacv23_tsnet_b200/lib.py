"""ctypes binding of libtsnet_sm100.so (include/tsnet_b200.h).

The library is built in-tree by `__graft_entry__.build()` / `make -C wacv23_tsnet_b200/csrc`.
There is no fallback: a missing library or a non-sm_100 device raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TSNET_LIB_PATH", os.path.join(_HERE, "libtsnet_sm100.so"))  # override: experiments only

FMT_FP16, FMT_BF16 = 0, 1
TAPS_SAME, TAPS_REFLECT1, TAPS_S2ZERO, TAPS_UP2REFLECT1, TAPS_WINO = 0, 1, 2, 3, 4
MAX_TAPS = 49
ABI_VERSION = 3
CONV_NO_VR, CONV_NO_TAIL_SPLIT, CONV_ONE_CTA, CONV_SMALL_FIRST = 1, 2, 4, 8   # tsnet_conv_desc.flags (launch-plan switches, tests only)
TAPS_GENERIC_UP2 = 1                                     # tsnet_taps_desc.flags

vp = C.c_void_p


class ConvDesc(C.Structure):
    _fields_ = [("B", C.c_int), ("H", C.c_int), ("W", C.c_int), ("Cout", C.c_int), ("Cout_pad", C.c_int),
                ("Cp", C.c_int), ("Hp", C.c_int), ("Wp", C.c_int), ("planes", C.c_int), ("num_taps", C.c_int),
                ("tap_dy", C.c_int8 * MAX_TAPS), ("tap_dx", C.c_int8 * MAX_TAPS), ("tap_plane", C.c_int8 * MAX_TAPS),
                ("block_n", C.c_int), ("split", C.c_int), ("fmt", C.c_int), ("out_scale", C.c_float),
                ("addend", C.c_void_p), ("addend_rows", C.c_int),
                ("fuse_in", C.c_int), ("fuse_relu", C.c_int), ("fuse_mode", C.c_int),
                ("fuse_residual", C.c_void_p), ("fuse_act_out", C.c_void_p),
                ("fuse_act_C_total", C.c_int), ("fuse_act_c_off", C.c_int),
                ("fuse_taps_hi", C.c_void_p), ("fuse_taps_lo", C.c_void_p),
                ("fuse_taps_Cp", C.c_int), ("fuse_taps_c_off", C.c_int),
                ("fuse_act_scale", C.c_float), ("fuse_eps", C.c_float), ("flags", C.c_int)]


class TapsDesc(C.Structure):
    _fields_ = [("B", C.c_int), ("H", C.c_int), ("W", C.c_int), ("C", C.c_int), ("mode", C.c_int),
                ("relu", C.c_int), ("Cp_total", C.c_int), ("c_off", C.c_int), ("fmt", C.c_int),
                ("scale", C.c_float), ("act_C_total", C.c_int), ("act_c_off", C.c_int), ("avg_n", C.c_int),
                ("flags", C.c_int), ("residual2", C.c_void_p), ("res_split", C.c_int), ("res2_batch", C.c_int)]


class CorrDesc(C.Structure):
    _fields_ = [("B", C.c_int), ("n_src", C.c_int), ("C", C.c_int), ("h", C.c_int), ("w", C.c_int),
                ("bbox_h", C.c_int), ("bbox_w", C.c_int), ("bbox_dtype", C.c_int), ("temperature", C.c_float),
                ("split", C.c_int), ("fmt", C.c_int), ("operand_scale", C.c_float), ("sort", C.c_int),
                ("one_cta", C.c_int), ("chunk_kb", C.c_int)]


class WinoGemmDesc(C.Structure):
    _fields_ = [("B", C.c_int), ("TH", C.c_int), ("TW", C.c_int), ("C", C.c_int), ("Cout", C.c_int),
                ("split", C.c_int), ("fmt", C.c_int), ("out_scale", C.c_float), ("chunk_kb", C.c_int),
                ("flags", C.c_int)]


class WinoBridgeDesc(C.Structure):
    _fields_ = [("B", C.c_int), ("H", C.c_int), ("W", C.c_int), ("C", C.c_int), ("relu", C.c_int),
                ("Cp_total", C.c_int), ("c_off", C.c_int), ("fmt", C.c_int), ("scale", C.c_float), ("eps", C.c_float),
                ("act_C_total", C.c_int), ("act_c_off", C.c_int), ("addend_rows", C.c_longlong),
                ("corr_hi", C.c_void_p), ("corr_lo", C.c_void_p), ("corr_rank", C.c_void_p), ("corr_ssq", C.c_void_p),
                ("corr_scale", C.c_float), ("variant", C.c_int)]


class StemConvDesc(C.Structure):
    _fields_ = [("B", C.c_int), ("H", C.c_int), ("W", C.c_int), ("Cimg", C.c_int), ("Clbl", C.c_int),
                ("img_kind", C.c_int), ("lbl_kind", C.c_int), ("img_mean", C.c_float * 3), ("img_div", C.c_float),
                ("Cout", C.c_int), ("split", C.c_int), ("fmt", C.c_int), ("act_scale", C.c_float),
                ("out_scale", C.c_float)]


_SIGNATURES = {
    "tsnet_abi_version": (C.c_int, []),
    "tsnet_last_error": (C.c_char_p, []),
    "tsnet_device_ok": (C.c_int, []),
    "tsnet_launch_count": (C.c_longlong, []),
    "tsnet_pack_conv_weight": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                         C.c_float, C.c_int, vp, vp, vp]),
    "tsnet_conv_gemm_fwd": (C.c_int, [C.POINTER(ConvDesc), vp, vp, vp, vp, vp, vp, vp, vp]),
    "tsnet_wino_weight_transform": (C.c_int, [vp, C.c_int, C.c_int, vp, vp]),
    "tsnet_wino_gemm_fwd": (C.c_int, [C.POINTER(WinoGemmDesc), vp, vp, vp, vp, vp, vp]),
    "tsnet_wino_output": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, C.c_longlong, vp, vp, vp]),
    "tsnet_wino_bridge": (C.c_int, [C.POINTER(WinoBridgeDesc), vp, vp, vp, vp, vp, vp, vp, vp, vp]),
    "tsnet_stem_conv_fwd": (C.c_int, [C.POINTER(StemConvDesc), vp, vp, vp, vp, vp, vp, vp, vp]),
    "tsnet_instnorm_reduce": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_float, vp, vp]),
    "tsnet_build_taps": (C.c_int, [C.POINTER(TapsDesc), vp, vp, vp, vp, vp, vp, vp]),
    "tsnet_stem_taps": (C.c_int, [vp, C.c_int, C.c_int, C.POINTER(C.c_float), C.c_float, vp, C.c_int, C.c_int, C.c_int,
                                  C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, vp, vp, vp]),
    "tsnet_l2norm_split": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, vp, vp, vp, vp]),
    "tsnet_corr_workspace_bytes": (C.c_size_t, [C.POINTER(CorrDesc)]),
    "tsnet_corr_prepare": (C.c_int, [C.POINTER(CorrDesc), vp, C.POINTER(vp), vp, vp, C.c_size_t, vp]),
    "tsnet_corr_rank_table": (vp, [C.POINTER(CorrDesc), vp, C.c_int]),
    "tsnet_corr_warp_fwd": (C.c_int, [C.POINTER(CorrDesc), vp, vp, vp, vp, vp, vp, C.POINTER(vp), vp, vp, vp, vp,
                                      C.c_int, C.c_int, C.c_float, vp, C.c_size_t, vp]),
    "tsnet_corr_tiles": (C.c_int, [C.POINTER(CorrDesc), vp, vp, vp, vp, vp, vp, vp, C.c_size_t, vp]),
    "tsnet_corr_operands": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, vp, vp, vp, vp, vp]),
    "tsnet_corr_norms": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, vp, vp, vp]),
    "tsnet_corr_finish": (C.c_int, [C.POINTER(CorrDesc), C.POINTER(vp), vp, vp, vp, vp, C.c_int, C.c_int, C.c_float,
                                    vp, C.c_size_t, vp]),
    "tsnet_warp_mean_taps": (C.c_int, [C.POINTER(vp), C.c_int, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp,
                                       C.c_int, C.c_int, C.c_int, C.c_float, vp]),
    "tsnet_train_extras_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "tsnet_train_extras_fwd": (C.c_int, [C.POINTER(vp), C.POINTER(C.c_float), C.c_int, vp, C.c_float, vp, C.c_int,
                                         C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, C.c_int, C.c_int, C.c_int,
                                         C.POINTER(C.c_float), vp, vp, vp, C.c_size_t, vp]),
    "tsnet_plane_stats": (C.c_int, [vp, C.c_int, C.c_int, C.c_float, vp, vp]),
    "tsnet_head_conv_tanh": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, C.c_int, C.c_int,
                                       C.POINTER(C.c_float), vp, vp]),
    "tsnet_postprocess_u8": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, vp, vp, vp, C.POINTER(C.c_float), vp, vp]),
    "tsnet_direct_conv_fp32": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, C.c_int, C.c_int,
                                         C.c_int, C.c_int, C.c_int, vp, vp]),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None


class TSNetLibraryError(RuntimeError):
    pass


def load():
    """Load (once) and return the ctypes handle.  Raises if the library has not been built."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise TSNetLibraryError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU / PyTorch fallback for the TS-Net hot path)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        if lib.tsnet_abi_version() != ABI_VERSION:
            raise TSNetLibraryError("libtsnet_sm100.so ABI version mismatch")
        _lib = lib
    return _lib


def check(rc):
    if rc != 0:
        raise TSNetLibraryError(f"libtsnet_sm100 error {rc}: {load().tsnet_last_error().decode()}")


def require_device():
    import torch
    if not torch.cuda.is_available():
        raise TSNetLibraryError("TS-Net B200 forward needs a CUDA device (sm_100); none is visible")
    if not load().tsnet_device_ok():
        raise TSNetLibraryError("TS-Net B200 kernels are built for sm_100a only; current device is not CC 10.x")
