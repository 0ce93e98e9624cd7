"""ctypes binding of libcustma_b200.so (the C ABI declared in include/custma_b200.h).

This is the only place that touches the shared library.  There is NO fallback: if the library is missing or a call
fails, a RuntimeError is raised (the reference raises RuntimeError through TORCH_CHECK,
custma/include/stereo_matching.hpp:20-25).  PyTorch is used for device memory and streams only; every pointer
crossing the boundary is a plain address.
"""
from __future__ import annotations

import ctypes
import os
import threading

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
# CUSTMA_LIB: another build of the same library (kernel experiments, tools/build_variants.py); never a different backend
LIB_PATH = os.environ.get("CUSTMA_LIB") or os.path.join(PKG_DIR, "libcustma_b200.so")

OK = 0
ERR_INVALID_ARGUMENT = 1
ERR_WORKSPACE = 2
ERR_CUDA = 3
ERR_UNSUPPORTED = 4
FLAG_DIRECT = 1
FLAG_TENSOR = 2
FLAG_PREPARED = 4
INVALID_COST = -2.0
ABI_VERSION = 3

# every symbol include/custma_b200.h declares (tests/test_abi.py checks the header against this list)
SYMBOLS = (
    "custma_abi_version",
    "custma_last_error",
    "custma_launch_count",
    "custma_debug_validate_layout",
    "custma_debug_verdict_info",
    "custma_forward_workspace_bytes",
    "custma_backward_workspace_bytes",
    "custma_forward",
    "custma_forward_wta",
    "custma_backward",
    "custma_backward_rows",
    "custma_backward_prepare",
    "custma_forward_prepare",
    "custma_ingest_u8",
    "custma_backward_projector_workspace_bytes",
    "custma_backward_projector",
    "custma_head_workspace_bytes",
    "custma_forward_head",
    "custma_backward_head",
    "custma_host_step",
    "custma_host_submit",
    "custma_host_submit_u8",
    "custma_host_wait",
    "custma_host_release",
)

_lib = None
_lock = threading.Lock()

_i32 = ctypes.c_int32
_u32 = ctypes.c_uint32
_ptr = ctypes.c_void_p
_size = ctypes.c_size_t


def _declare(lib):
    lib.custma_abi_version.restype = ctypes.c_int
    lib.custma_abi_version.argtypes = []
    lib.custma_last_error.restype = ctypes.c_char_p
    lib.custma_last_error.argtypes = []
    lib.custma_launch_count.restype = ctypes.c_uint64
    lib.custma_launch_count.argtypes = []
    for name in ("custma_forward_workspace_bytes", "custma_backward_workspace_bytes", "custma_head_workspace_bytes",
                 "custma_backward_projector_workspace_bytes"):
        fn = getattr(lib, name)
        fn.restype = _size
        fn.argtypes = [_i32, _i32, _i32, _i32, _i32, _u32]
    lib.custma_debug_validate_layout.restype = ctypes.c_int
    lib.custma_debug_validate_layout.argtypes = [_i32, _i32, _i32, _i32, _i32]
    lib.custma_debug_verdict_info.restype = ctypes.c_int
    lib.custma_debug_verdict_info.argtypes = [_i32, _i32, _i32, _i32, _i32, ctypes.POINTER(_size),
                                              ctypes.POINTER(_u32), ctypes.POINTER(_i32)]
    lib.custma_forward.restype = ctypes.c_int
    lib.custma_forward.argtypes = [_ptr, _ptr, _ptr, _ptr, _ptr, _i32, _i32, _i32, _i32, _i32, _u32, _ptr, _size, _ptr]
    lib.custma_forward_wta.restype = ctypes.c_int
    lib.custma_forward_wta.argtypes = [_ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, ctypes.c_float, _i32, _i32, _i32, _i32,
                                       _i32, _u32, _ptr, _size, _ptr]
    lib.custma_backward.restype = ctypes.c_int
    lib.custma_backward.argtypes = [_ptr, _ptr, _ptr, _ptr, _i32, _i32, _i32, _i32, _i32, _u32, _ptr, _size, _ptr]
    lib.custma_forward_prepare.restype = ctypes.c_int
    lib.custma_forward_prepare.argtypes = [_ptr, _ptr, _i32, _i32, _i32, _i32, _i32, _u32, _ptr, _size, _ptr]
    lib.custma_backward_prepare.restype = ctypes.c_int
    lib.custma_backward_prepare.argtypes = [_ptr, _ptr, _i32, _i32, _i32, _i32, _i32, _u32, _ptr, _size, _ptr]
    lib.custma_backward_rows.restype = ctypes.c_int
    lib.custma_backward_rows.argtypes = [_ptr, _ptr, _ptr, _ptr, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _u32, _ptr,
                                         _size, _ptr]
    lib.custma_backward_projector.restype = ctypes.c_int
    lib.custma_backward_projector.argtypes = [_ptr, _ptr, _ptr, _ptr, _i32, _i32, _i32, _i32, _i32, _u32, _ptr, _size, _ptr]
    lib.custma_forward_head.restype = ctypes.c_int
    lib.custma_forward_head.argtypes = [_ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, ctypes.c_float, ctypes.c_float, _i32, _i32,
                                        _i32, _i32, _i32, _u32, _ptr, _size, _ptr]
    lib.custma_backward_head.restype = ctypes.c_int
    lib.custma_backward_head.argtypes = [_ptr, _ptr, _ptr, _ptr, ctypes.c_float, _ptr, _i32, _i32, _i32, _i32, _i32, _u32,
                                         _ptr, _size, _ptr]
    lib.custma_ingest_u8.restype = ctypes.c_int
    lib.custma_ingest_u8.argtypes = [_ptr, _ptr, _i32, _i32, _i32, _i32, _i32, ctypes.c_float, _ptr]
    lib.custma_host_step.restype = ctypes.c_int
    lib.custma_host_step.argtypes = [_ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _i32, _i32, _i32, _i32, _i32, _u32]
    lib.custma_host_submit.restype = ctypes.c_int
    lib.custma_host_submit.argtypes = [_ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _i32, _i32, _i32, _i32, _i32, _u32,
                                       ctypes.POINTER(ctypes.c_uint64)]
    lib.custma_host_submit_u8.restype = ctypes.c_int
    lib.custma_host_submit_u8.argtypes = [_ptr, _i32, _i32, _ptr, _i32, _i32, ctypes.c_float, _ptr, _ptr, _ptr, _ptr, _ptr,
                                          _i32, _i32, _i32, _i32, _i32, _u32, ctypes.POINTER(ctypes.c_uint64)]
    lib.custma_host_wait.restype = ctypes.c_int
    lib.custma_host_wait.argtypes = [ctypes.c_uint64]
    lib.custma_host_release.restype = ctypes.c_int
    lib.custma_host_release.argtypes = []


def load():
    """Returns the loaded library; raises RuntimeError if it has not been built (no CPU or eager fallback exists)."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise RuntimeError(
                    f"{LIB_PATH} is missing: build it with `python setup.py build_ext --inplace` or "
                    "`python -m custereomatching_b200.build`; custma has no CPU or eager fallback")
            lib = ctypes.CDLL(LIB_PATH)
            _declare(lib)
            got = lib.custma_abi_version()
            if got != ABI_VERSION:
                raise RuntimeError(f"libcustma_b200.so has ABI version {got}, this binding expects {ABI_VERSION}")
            _lib = lib
    return _lib


def last_error() -> str:
    msg = load().custma_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def launch_count() -> int:
    return int(load().custma_launch_count())


def check(rc: int, what: str) -> None:
    if rc != OK:
        raise RuntimeError(f"{what} failed (code {rc}): {last_error()}")


def validate_layout(B, H, W, D, k) -> None:
    check(load().custma_debug_validate_layout(B, H, W, D, k), "custma_debug_validate_layout")


def debug_layout_info(B, H, W, D, k):
    """Where custma_forward leaves its conditioning verdict in the workspace (None if the shape has no fast path)."""
    off, cap, tc = _size(0), _u32(0), _i32(0)
    rc = load().custma_debug_verdict_info(B, H, W, D, k, ctypes.byref(off), ctypes.byref(cap), ctypes.byref(tc))
    if rc != OK:
        return None
    return {"fb_count_offset": int(off.value), "fb_capacity": int(cap.value), "tc_supported": bool(tc.value)}


def forward_workspace_bytes(B, H, W, D, k, flags=0) -> int:
    return int(load().custma_forward_workspace_bytes(B, H, W, D, k, flags))


def backward_workspace_bytes(B, H, W, D, k, flags=0) -> int:
    return int(load().custma_backward_workspace_bytes(B, H, W, D, k, flags))


def forward(camera_ptr, projector_ptr, cost_ptr, best_ptr, index_ptr, B, H, W, D, k, flags, ws_ptr, ws_bytes, stream):
    rc = load().custma_forward(camera_ptr, projector_ptr, cost_ptr or None, best_ptr or None, index_ptr or None,
                               B, H, W, D, k, flags, ws_ptr or None, ws_bytes, stream or None)
    check(rc, "custma_forward")


def forward_wta(camera_ptr, projector_ptr, cost_ptr, best_ptr, index_ptr, mask_ptr, masked_disparity_ptr, threshold,
                B, H, W, D, k, flags, ws_ptr, ws_bytes, stream):
    rc = load().custma_forward_wta(camera_ptr, projector_ptr, cost_ptr or None, best_ptr or None, index_ptr or None,
                                   mask_ptr or None, masked_disparity_ptr or None, float(threshold), B, H, W, D, k, flags,
                                   ws_ptr or None, ws_bytes, stream or None)
    check(rc, "custma_forward_wta")


def backward_rows(grad_ptr, camera_ptr, projector_ptr, camera_grad_ptr, B, H, W, D, k, row_begin, row_end, flags, ws_ptr,
                  ws_bytes, stream):
    rc = load().custma_backward_rows(grad_ptr, camera_ptr, projector_ptr, camera_grad_ptr, B, H, W, D, k, row_begin,
                                     row_end, flags, ws_ptr or None, ws_bytes, stream or None)
    check(rc, "custma_backward_rows")


def forward_prepare(camera_ptr, projector_ptr, B, H, W, D, k, flags, ws_ptr, ws_bytes, stream):
    """The image-dependent part of forward(), ahead of time; then forward(..., flags | FLAG_PREPARED) on the same workspace."""
    rc = load().custma_forward_prepare(camera_ptr, projector_ptr, B, H, W, D, k, flags, ws_ptr or None, ws_bytes, stream or None)
    check(rc, "custma_forward_prepare")


def backward_prepare(camera_ptr, projector_ptr, B, H, W, D, k, flags, ws_ptr, ws_bytes, stream):
    """The image-dependent part of backward(), ahead of time; then backward(..., flags | FLAG_PREPARED) on the same workspace."""
    rc = load().custma_backward_prepare(camera_ptr, projector_ptr, B, H, W, D, k, flags, ws_ptr or None, ws_bytes, stream or None)
    check(rc, "custma_backward_prepare")


def backward_projector_workspace_bytes(B, H, W, D, k, flags=0) -> int:
    return int(load().custma_backward_projector_workspace_bytes(B, H, W, D, k, flags))


def backward_projector(grad_ptr, camera_ptr, projector_ptr, projector_grad_ptr, B, H, W, D, k, flags, ws_ptr, ws_bytes, stream):
    rc = load().custma_backward_projector(grad_ptr, camera_ptr, projector_ptr, projector_grad_ptr, B, H, W, D, k, flags,
                                          ws_ptr or None, ws_bytes, stream or None)
    check(rc, "custma_backward_projector")


def head_workspace_bytes(B, H, W, D, k, flags=0) -> int:
    return int(load().custma_head_workspace_bytes(B, H, W, D, k, flags))


def forward_head(camera_ptr, projector_ptr, soft_ptr, best_ptr, index_ptr, mask_ptr, state_ptr, beta, threshold,
                 B, H, W, D, k, flags, ws_ptr, ws_bytes, stream):
    rc = load().custma_forward_head(camera_ptr, projector_ptr, soft_ptr, best_ptr or None, index_ptr or None,
                                    mask_ptr or None, state_ptr or None, float(beta), float(threshold), B, H, W, D, k,
                                    flags, ws_ptr or None, ws_bytes, stream or None)
    check(rc, "custma_forward_head")


def backward_head(soft_grad_ptr, camera_ptr, projector_ptr, state_ptr, beta, camera_grad_ptr, B, H, W, D, k, flags, ws_ptr,
                  ws_bytes, stream):
    rc = load().custma_backward_head(soft_grad_ptr, camera_ptr, projector_ptr, state_ptr, float(beta), camera_grad_ptr,
                                     B, H, W, D, k, flags, ws_ptr or None, ws_bytes, stream or None)
    check(rc, "custma_backward_head")


def ingest_u8(src_ptr, dst_ptr, B, H, W, channels, channel, scale, stream):
    check(load().custma_ingest_u8(src_ptr, dst_ptr, B, H, W, channels, channel, float(scale), stream or None),
          "custma_ingest_u8")


def backward(grad_ptr, camera_ptr, projector_ptr, camera_grad_ptr, B, H, W, D, k, flags, ws_ptr, ws_bytes, stream):
    rc = load().custma_backward(grad_ptr, camera_ptr, projector_ptr, camera_grad_ptr, B, H, W, D, k, flags,
                                ws_ptr or None, ws_bytes, stream or None)
    check(rc, "custma_backward")


def host_step(h_camera, h_projector, h_best, h_index, h_camera_grad, cost_volume_dev, cost_volume_grad_dev,
              B, H, W, D, k, flags=0):
    rc = load().custma_host_step(h_camera, h_projector, h_best, h_index, h_camera_grad or None,
                                 cost_volume_dev or None, cost_volume_grad_dev or None, B, H, W, D, k, flags)
    check(rc, "custma_host_step")


def host_submit(h_camera, h_projector, h_best, h_index, h_camera_grad, cost_volume_dev, cost_volume_grad_dev,
                B, H, W, D, k, flags=0) -> int:
    """custma_host_step without the final wait; returns the ticket to pass to host_wait."""
    ticket = ctypes.c_uint64(0)
    rc = load().custma_host_submit(h_camera, h_projector, h_best, h_index, h_camera_grad or None,
                                   cost_volume_dev or None, cost_volume_grad_dev or None, B, H, W, D, k, flags,
                                   ctypes.byref(ticket))
    check(rc, "custma_host_submit")
    return int(ticket.value)


def host_submit_u8(h_camera_u8, camera_channels, camera_channel, h_projector_u8, projector_channels, projector_channel,
                   scale, h_best, h_index, h_camera_grad, cost_volume_dev, cost_volume_grad_dev, B, H, W, D, k, flags=0) -> int:
    """custma_host_submit for interleaved uint8 host images; returns the ticket."""
    ticket = ctypes.c_uint64(0)
    rc = load().custma_host_submit_u8(h_camera_u8, camera_channels, camera_channel, h_projector_u8, projector_channels,
                                      projector_channel, float(scale), h_best, h_index, h_camera_grad or None,
                                      cost_volume_dev or None, cost_volume_grad_dev or None, B, H, W, D, k, flags,
                                      ctypes.byref(ticket))
    check(rc, "custma_host_submit_u8")
    return int(ticket.value)


def host_wait(ticket: int = 0):
    check(load().custma_host_wait(ticket), "custma_host_wait")


def host_release():
    check(load().custma_host_release(), "custma_host_release")
