"""ctypes loader for the C restatement in oracle/nms_c.c (TEST INFRASTRUCTURE ONLY)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "liboracle_nms.so")
_lib = None


def build(force=False):
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(os.path.join(_HERE, "nms_c.c")):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.oracle_nms.restype = ctypes.c_int64
        _lib.oracle_nms.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_double,
                                    ctypes.c_void_p]
        _lib.oracle_batched_nms.restype = ctypes.c_int64
        _lib.oracle_batched_nms.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
                                            ctypes.c_double, ctypes.c_void_p]
    return _lib


def nms(boxes, scores, iou_threshold):
    b = np.ascontiguousarray(boxes, dtype=np.float32).reshape(-1, 4)
    s = np.ascontiguousarray(scores, dtype=np.float32).reshape(-1)
    out = np.empty(s.shape[0], dtype=np.int64)
    n = lib().oracle_nms(b.ctypes.data, s.ctypes.data, s.shape[0], float(iou_threshold), out.ctypes.data)
    return out[:n].copy()


def batched_nms(boxes, scores, idxs, iou_threshold):
    b = np.ascontiguousarray(boxes, dtype=np.float32).reshape(-1, 4)
    s = np.ascontiguousarray(scores, dtype=np.float32).reshape(-1)
    i = np.ascontiguousarray(idxs, dtype=np.int64).reshape(-1)
    out = np.empty(s.shape[0], dtype=np.int64)
    n = lib().oracle_batched_nms(b.ctypes.data, s.ctypes.data, i.ctypes.data, s.shape[0],
                                 float(iou_threshold), out.ctypes.data)
    return out[:n].copy()
