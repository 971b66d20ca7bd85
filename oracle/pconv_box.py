"""ctypes wrapper over oracle/pconv_box.c -- TEST INFRASTRUCTURE ONLY."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    subprocess.run(["make", "-s", "-C", _HERE], check=True)


def _lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "libpconv_box.so")
        if not os.path.exists(so):
            build()
        _LIB = ctypes.CDLL(so)
        _LIB.pconv_box_forward.restype = ctypes.c_int
    return _LIB


def pconv_box_forward(x, mask, w, bias, stride=1, pad=0, dil=1, groups=1, same_holes=False):
    """numpy fp32 NCHW in -> (y, msum, new_mask), each [N,Cout,Ho,Wo]."""
    x = np.ascontiguousarray(x, np.float32); mask = np.ascontiguousarray(mask, np.float32)
    w = np.ascontiguousarray(w, np.float32)
    N, Cin, H, W = x.shape
    Cout, _, kh, kw = w.shape
    Ho = (H + 2 * pad - dil * (kh - 1) - 1) // stride + 1
    Wo = (W + 2 * pad - dil * (kw - 1) - 1) // stride + 1
    y = np.empty((N, Cout, Ho, Wo), np.float32); ms = np.empty_like(y); nm = np.empty_like(y)
    fp = ctypes.POINTER(ctypes.c_float)
    bp = None
    if bias is not None:
        bias = np.ascontiguousarray(bias, np.float32); bp = bias.ctypes.data_as(fp)
    rc = _lib().pconv_box_forward(x.ctypes.data_as(fp), mask.ctypes.data_as(fp), w.ctypes.data_as(fp), bp,
                                  N, Cin, H, W, Cout, kh, kw, stride, pad, dil, groups, int(same_holes),
                                  mask.shape[1], y.ctypes.data_as(fp), ms.ctypes.data_as(fp), nm.ctypes.data_as(fp))
    if rc:
        raise ValueError(f"pconv_box_forward: bad arguments (code {rc})")
    return y, ms, nm
