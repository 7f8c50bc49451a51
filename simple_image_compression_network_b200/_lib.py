"""ctypes binding of libfinnconv_b200.so (include/finnconv_b200.h).  Fails loudly when the CUDA
library is missing: there is no CPU path behind this package."""
from __future__ import annotations

import ctypes
import os

from .desc import CLayerDesc

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG, "libfinnconv_b200.so")

_lib = None


class FcbError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"finnconv_b200 error {code}: {msg}")
        self.code = code


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `python -m simple_image_compression_network_b200.build` "
                          "(nvcc, sm_100a). There is no CPU fallback.")
    L = ctypes.CDLL(LIB_PATH)
    vp, u32, u64, sz = ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_size_t
    L.fcb_version.restype = ctypes.c_char_p
    L.fcb_last_error.restype = ctypes.c_char_p
    L.fcb_device_count.restype = ctypes.c_int
    L.fcb_word_bytes.restype = sz
    L.fcb_word_bytes.argtypes = [u32]
    L.fcb_layer_query.argtypes = [ctypes.POINTER(CLayerDesc)] + [ctypes.POINTER(sz)] * 5
    L.fcb_layer_create.argtypes = [ctypes.POINTER(CLayerDesc), vp, vp, vp, ctypes.c_int, ctypes.POINTER(vp)]
    L.fcb_layer_set_params.argtypes = [vp, vp, vp, vp]
    L.fcb_layer_set_param_stream.argtypes = [vp, vp, vp, vp]
    L.fcb_layer_destroy.argtypes = [vp]
    L.fcb_layer_destroy.restype = None
    L.fcb_layer_run.argtypes = [vp, vp, vp, u32]
    L.fcb_layer_run_device.argtypes = [vp, vp, vp, u32, vp]
    L.fcb_layer_engine.argtypes = [vp]
    L.fcb_layer_engine.restype = ctypes.c_char_p
    L.fcb_layer_plan.argtypes = [vp]
    L.fcb_layer_plan.restype = ctypes.c_char_p
    L.fcb_layer_launches.argtypes = [vp]
    L.fcb_layer_launches.restype = u64
    L.fcb_net_create.argtypes = [ctypes.POINTER(vp), u32, ctypes.POINTER(vp)]
    L.fcb_net_destroy.argtypes = [vp]
    L.fcb_net_destroy.restype = None
    L.fcb_net_run.argtypes = [vp, vp, vp, u32]
    L.fcb_net_run_device.argtypes = [vp, vp, vp, u32, vp]
    L.fcb_net_launches.argtypes = [vp]
    L.fcb_net_launches.restype = u64
    L.fcb_synth_fill.argtypes = [vp, sz, u64, u32, u64, vp]
    _lib = L
    return L


def check(rc: int) -> None:
    if rc != 0:
        raise FcbError(rc, lib().fcb_last_error().decode())
