"""ctypes binding of libfinnconv_b200.so (include/finnconv_b200.h).  Fails loudly when the CUDA
library is missing: there is no CPU path behind this package."""
from __future__ import annotations

import ctypes
import os

from .desc import CAddDesc, CLayerDesc

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG, "libfinnconv_b200.so")

EXP_LIB_PATH = os.path.join(os.path.dirname(PKG), "tools", "libfinnconv_exp.so")

_lib = None


class FcbError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"finnconv_b200 error {code}: {msg}")
        self.code = code


def load(path: str) -> ctypes.CDLL:
    """dlopen one build of the library and declare its ABI (include/finnconv_b200.h)."""
    if not os.path.exists(path):
        raise ImportError(f"{path} is missing: build it with `python -m simple_image_compression_network_b200.build` "
                          "(nvcc, sm_100a). There is no CPU fallback.")
    L = ctypes.CDLL(path)
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
    L.fcb_layer_set_host_chunk.argtypes = [vp, u32]
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
    L.fcb_net_set_host_chunk.argtypes = [vp, u32]
    L.fcb_net_set_device_chunk.argtypes = [vp, u32]
    L.fcb_net_launches.argtypes = [vp]
    L.fcb_net_launches.restype = u64
    L.fcb_synth_fill.argtypes = [vp, sz, u64, u32, u64, vp]
    pvp = ctypes.POINTER(vp)
    L.fcb_pool_create.argtypes = [ctypes.POINTER(CLayerDesc), pvp, pvp, pvp, u32, ctypes.POINTER(ctypes.c_int), u32, pvp]
    L.fcb_pool_destroy.argtypes = [vp]
    L.fcb_pool_destroy.restype = None
    L.fcb_pool_replicas.argtypes = [vp]
    L.fcb_pool_replicas.restype = u32
    L.fcb_pool_run.argtypes = [vp, vp, vp, u32]
    L.fcb_shard_range.argtypes = [u32, u32, u32, ctypes.POINTER(u32), ctypes.POINTER(u32)]
    L.fcb_add_streams.argtypes = [ctypes.POINTER(CAddDesc), vp, vp, vp, u64, ctypes.c_int]
    L.fcb_add_streams_device.argtypes = [ctypes.POINTER(CAddDesc), vp, vp, vp, u64, ctypes.c_int, vp]
    L.fcb_host_alloc.argtypes = [pvp, sz]
    L.fcb_host_free.argtypes = [vp]
    L.fcb_host_free.restype = None
    return L


def lib() -> ctypes.CDLL:
    """The library new handles are created from: the product build unless set_default() installed another one."""
    global _lib
    if _lib is None:
        _lib = load(LIB_PATH)
    return _lib


def set_default(L) -> None:
    """Tests / tools: make `L` (e.g. load(EXP_LIB_PATH), the experiment build) the library of handles created from now on;
    None restores the product build.  Existing handles keep the library they were created from."""
    global _lib
    _lib = L


def check(rc: int, L=None) -> None:
    if rc != 0:
        raise FcbError(rc, (L or lib()).fcb_last_error().decode())
