"""ctypes access to oracle/libref_port.so (the C restatement of the reference kernels).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))
_LIB = None
INVALID_COST = -2.0


def build(force: bool = False) -> str:
    so = os.path.join(_DIR, "libref_port.so")
    src = os.path.join(_DIR, "ref_port.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _DIR, "-s", "-B"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
    return _LIB


def _f(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def forward_full(cam, proj, k):
    cam, pc = _f(cam); proj, pp = _f(proj)
    H, W = cam.shape
    out = np.empty((H, W, W), np.float32)
    lib().ref_forward_full(pc, pp, H, W, k, out.ctypes.data_as(ctypes.POINTER(ctypes.c_float)))
    return out


def forward_banded(cam, proj, D, k, invalid=INVALID_COST):
    cam, pc = _f(cam); proj, pp = _f(proj)
    H, W = cam.shape
    out = np.empty((H, W, D), np.float32)
    lib().ref_forward_banded(pc, pp, H, W, D, k, ctypes.c_float(invalid),
                             out.ctypes.data_as(ctypes.POINTER(ctypes.c_float)))
    return out


def backward_full(g, cam, proj, k):
    g, pg = _f(g); cam, pc = _f(cam); proj, pp = _f(proj)
    H, W = cam.shape
    assert g.shape == (H, W, W)
    out = np.empty((H, W), np.float32)
    lib().ref_backward_full(pg, pc, pp, H, W, k, out.ctypes.data_as(ctypes.POINTER(ctypes.c_float)))
    return out


def backward_banded(g, cam, proj, k):
    g, pg = _f(g); cam, pc = _f(cam); proj, pp = _f(proj)
    H, W = cam.shape
    D = g.shape[2]
    assert g.shape == (H, W, D)
    out = np.empty((H, W), np.float32)
    lib().ref_backward_banded(pg, pc, pp, H, W, D, k, out.ctypes.data_as(ctypes.POINTER(ctypes.c_float)))
    return out


def wta_full(vol):
    vol, pv = _f(vol)
    H, W, _ = vol.shape
    best = np.empty((H, W), np.float32); corr = np.empty((H, W), np.int32)
    lib().ref_wta_full(pv, H, W, best.ctypes.data_as(ctypes.POINTER(ctypes.c_float)),
                       corr.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)))
    return best, corr


def wta_banded(vol):
    vol, pv = _f(vol)
    H, W, D = vol.shape
    best = np.empty((H, W), np.float32); disp = np.empty((H, W), np.int32)
    lib().ref_wta_banded(pv, H, W, D, best.ctypes.data_as(ctypes.POINTER(ctypes.c_float)),
                         disp.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)))
    return best, disp
