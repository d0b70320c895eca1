"""ctypes loader for oracle/canny_oracle.c -- TEST INFRASTRUCTURE ONLY."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "_build", "libcanny_oracle.so")
        if not os.path.exists(path):
            build()
        _LIB = ctypes.CDLL(path)
        _LIB.canny_oracle_u8.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                         ctypes.c_int, ctypes.c_void_p]
        _LIB.canny_oracle_u8.restype = ctypes.c_int
    return _LIB


def canny_u8(img, low=10, high=100):
    """img: uint8 [H,W] -> uint8 [H,W] in {0,255} (== cv2.Canny(img, low, high))."""
    img = np.ascontiguousarray(img, dtype=np.uint8)
    out = np.empty_like(img)
    rc = _lib().canny_oracle_u8(img.ctypes.data, img.shape[0], img.shape[1], int(low), int(high), out.ctypes.data)
    if rc != 0:
        raise RuntimeError("canny_oracle_u8 failed: %d" % rc)
    return out
