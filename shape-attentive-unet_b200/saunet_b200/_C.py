"""ctypes binding of libsaunet_b200.so (see include/saunet_b200.h).

The product path FAILS LOUDLY when the CUDA library is missing: there is no
CPU / eager fallback anywhere in this package.
"""
import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_double, c_float, c_int, c_longlong, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "libsaunet_b200.so")

ACT_NONE, ACT_RELU, ACT_SIGMOID = 0, 1, 2


class ConvDesc(Structure):
    _fields_ = [("x", c_void_p), ("x_ld", c_int), ("B", c_int), ("Hin", c_int), ("Win", c_int), ("Cin", c_int),
                ("w", c_void_p), ("Cout", c_int), ("KH", c_int), ("KW", c_int),
                ("Hg", c_int), ("Wg", c_int), ("sy", c_int), ("sx", c_int), ("offy", c_int), ("offx", c_int),
                ("y", c_void_p), ("y_ld", c_int), ("Hout", c_int), ("Wout", c_int), ("osy", c_int), ("osx", c_int),
                ("oy0", c_int), ("ox0", c_int),
                ("in_scale", c_void_p), ("in_shift", c_void_p), ("in_relu", c_int),
                ("bias", c_void_p), ("row_scale", c_void_p), ("row_scale_add", c_float),
                ("act", c_int), ("accumulate", c_int), ("stat_sum", c_void_p), ("stat_sumsq", c_void_p),
                ("w_tc", c_void_p), ("tc_bn", c_int), ("tc_passes", c_int), ("tc_cm", c_int),
                ("epi_x", c_void_p), ("epi_x_ld", c_int), ("epi_scale", c_void_p), ("epi_shift", c_void_p),
                ("epi_mean", c_void_p), ("epi_relu", c_int)]


class WgradDesc(Structure):
    _fields_ = [("p", c_void_p), ("p_ld", c_int), ("Ca", c_int),
                ("q", c_void_p), ("q_ld", c_int), ("Cb", c_int), ("B", c_int), ("Hq", c_int), ("Wq", c_int),
                ("KH", c_int), ("KW", c_int), ("Hg", c_int), ("Wg", c_int), ("sy", c_int), ("sx", c_int),
                ("offy", c_int), ("offx", c_int),
                ("q_scale", c_void_p), ("q_shift", c_void_p), ("q_relu", c_int), ("dw", c_void_p), ("precision", c_int)]


class UnpackEntry(Structure):
    _fields_ = [("first", c_longlong), ("packed_off", c_longlong), ("grad_off", c_longlong),
                ("A", c_int), ("Bc", c_int), ("T", c_int), ("pad_", c_int)]


_P, _I, _L, _F, _D = c_void_p, c_int, c_longlong, c_float, c_double

# name -> argtypes (restype is int unless listed in _SPECIAL)
SIGNATURES = {
    "saunet_conv2d_fwd": [POINTER(ConvDesc), _P],
    "saunet_conv2d_wgrad": [POINTER(WgradDesc), _P],
    "saunet_pack_weights": [_P, _P, _I, _I, _I, _I, _I, _P],
    "saunet_unpack_wgrad": [_P, _P, _I, _I, _I, _I, _I, _P],
    "saunet_unpack_wgrad_multi": [_P, _I, _L, _P, _P, _P],
    "saunet_pack_weights_tc": [_P, _I, _I, _I, _I, _I, _P, _P],
    "saunet_pack_weights_tc_cm": [_P, _I, _I, _I, _I, _I, _P, _P],
    "saunet_channel_stats": [_P, _I, _I, _L, _P, _P, _P],
    "saunet_add_d2f": [_P, _P, _I, _P],
    "saunet_bn_finalize": [_P, _P, _D, _P, _P, _P, _P, _F, _F, _I, _I, _P, _P],
    "saunet_affine_act": [_P, _I, _P, _P, _P, _I, _P, _I, _I, _L, _I, _P],
    "saunet_bn_fused_finish": [_P, _P, _D, _I, _I, _P, _P, _P, _I, _P],
    "saunet_bn_fixup": [_P, _I, _P, _I, _P, _I, _I, _L, _P],
    "saunet_bn_bwd_reduce": [_P, _I, _P, _I, _P, _I, _P, _I, _L, _I, _P, _P],
    "saunet_bn_bwd_apply": [_P, _I, _P, _I, _P, _I, _P, _P, _P, _I, _L, _I, _I, _P, _I, _I, _P, _I, _I, _P, _P, _P],
    "saunet_bilinear_fwd": [_P, _I, _I, _I, _I, _I, _P, _I, _I, _I, _P],
    "saunet_bilinear_bwd": [_P, _I, _I, _I, _I, _I, _P, _I, _I, _I, _I, _P],
    "saunet_avgpool2_fwd": [_P, _I, _I, _I, _I, _I, _P, _I, _P],
    "saunet_avgpool2_bwd": [_P, _I, _I, _I, _I, _I, _P, _I, _I, _P],
    "saunet_maxpool2_fwd": [_P, _I, _I, _I, _I, _I, _P, _I, _P],
    "saunet_maxpool2_bwd": [_P, _I, _P, _I, _I, _I, _I, _I, _P, _I, _I, _P],
    "saunet_gap_fwd": [_P, _I, _I, _L, _I, _P, _P],
    "saunet_gap_bwd": [_P, _I, _L, _I, _P, _I, _I, _P],
    "saunet_act_bwd": [_P, _I, _P, _I, _P, _I, _I, _L, _I, _P],
    "saunet_dualatt_combine_fwd": [_P, _I, _P, _P, _I, _L, _I, _P, _I, _P],
    "saunet_dualatt_combine_bwd": [_P, _I, _P, _I, _P, _P, _I, _L, _I, _P, _I, _I, _P, _P, _P],
    "saunet_rowscale_bwd": [_P, _I, _P, _I, _P, _I, _L, _P, _I, _P, _I, _P],
    "saunet_copy_slice": [_P, _I, _P, _I, _I, _L, _I, _P],
    "saunet_nchw_to_nhwc": [_P, _P, _I, _I, _I, _L, _P],
    "saunet_nhwc_to_nchw": [_P, _I, _P, _I, _I, _L, _P],
    "saunet_dual_loss_fwd": [_P, _I, _P, _P, _P, _L, _I, _P, _I, _P, _P, _P, _P],
    "saunet_dual_loss_bwd": [_P, _I, _P, _P, _P, _L, _I, _P, _P, _P, _P, _I, _P, _I, _P],
    "saunet_optimizer_step": [_I, _P, _P, _P, _P, _L, _P, _I, _P, _F, _F, _F, _F, _P],
    "saunet_argmax_u8": [_P, _I, _I, _L, _P, _P],
    "saunet_edge_gt": [_P, _I, _I, _I, _I, _I, _P, _P],
    "saunet_canny_fwd": [_P, _I, _I, _I, _I, _I, _I, _P, _P, _L, _P],
}
_SPECIAL = {
    "saunet_version": ([], c_int),
    "saunet_last_error": ([], c_char_p),
    "saunet_launch_count": ([], c_longlong),
    "saunet_last_kernel": ([], c_char_p),
    "saunet_canny_workspace_bytes": ([_I, _I, _I], c_longlong),
    "saunet_tc_tile_n": ([_I], c_int),
    "saunet_tc_chunk_major": ([_I, _I], c_int),
    "saunet_tc_packed_floats": ([_I, _I, _I, _I], c_longlong),
    "saunet_tc_packed_floats_cm": ([_I, _I, _I, _I, _I], c_longlong),
}
ALL_SYMBOLS = sorted(list(SIGNATURES) + list(_SPECIAL))

_lib = None


class SaunetError(RuntimeError):
    pass


def load():
    """Load the shared library (raises if it has not been built)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "libsaunet_b200.so not found at %s -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or shape-attentive-unet_b200/csrc/build.sh; this package has no CPU fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, args in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = c_int
    for name, (args, res) in _SPECIAL.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = res
    _lib = lib
    return lib


def check(rc, name="saunet"):
    if rc != 0:
        msg = load().saunet_last_error()
        raise SaunetError("%s failed (%d): %s" % (name, rc, msg.decode() if msg else "?"))


# When set to a list, every call is bracketed by CUDA events on the current stream and appended as
# (name, start_event, end_event, flops, bytes): bench.py's per-kernel roofline pass.  None on the normal path.
PROFILE = None
SCOPE = ""         # bench.py --workload blocks: label (e.g. "tail") prefixed to the tag of every call made while it is set


def call(name, *args, flops=0, nbytes=0, tag=""):
    if PROFILE is not None:
        import torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = getattr(load(), name)(*args)
        e1.record()
        kern = load().saunet_last_kernel()
        PROFILE.append((name, e0, e1, flops, nbytes, (SCOPE + ":" + tag) if SCOPE else tag, kern.decode() if kern else name))
    else:
        rc = getattr(load(), name)(*args)
    if rc != 0:
        check(rc, name)


def launch_count():
    return int(load().saunet_launch_count())
