"""oracle/oracle.py -- TEST INFRASTRUCTURE ONLY.

ctypes front-ends for (a) the plain-C restatement libfinn_oracle.so and (b), when it has been
built in a container that has /root/reference, the reference's own templates in
_ref/libref_layers.so.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs import this module.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


class FoSizes(ctypes.Structure):
    _fields_ = [(n, ctypes.c_uint32) for n in ("k_total", "sf", "nf", "out_x", "out_y")] + [
        (n, ctypes.c_size_t) for n in ("in_word_bytes", "out_word_bytes", "in_bytes_per_image", "out_bytes_per_image",
                                       "weight_word_bytes", "weight_bytes", "threshold_bytes", "bias_bytes")]


_lib = None


def build(force: bool = False) -> None:
    """Compile the C restatement (and the reference-linked libraries when /root/reference exists)."""
    subprocess.check_call(["make", "-s", "-C", HERE, "oracle"] + (["-B"] if force else []))
    if os.path.isdir("/root/reference"):
        subprocess.check_call(["make", "-s", "-C", HERE, "ref"])


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        path = os.path.join(HERE, "libfinn_oracle.so")
        if not os.path.exists(path):
            build()
        _lib = ctypes.CDLL(path)
        _lib.fo_layer_query.argtypes = [ctypes.c_void_p, ctypes.POINTER(FoSizes)]
        _lib.fo_layer_run.argtypes = [ctypes.c_void_p] + [ctypes.c_void_p] * 5 + [ctypes.c_uint32]
        _lib.fo_maxpool.argtypes = [ctypes.c_void_p, ctypes.c_void_p] + [ctypes.c_uint32] * 5
        _lib.fo_word_bytes.restype = ctypes.c_size_t
        _lib.fo_word_bytes.argtypes = [ctypes.c_uint32]
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def set_threads(n: int) -> None:
    lib().fo_set_threads(int(n))


def query(desc) -> FoSizes:
    s = FoSizes()
    c = desc.to_c()
    rc = lib().fo_layer_query(ctypes.byref(c), ctypes.byref(s))
    if rc:
        raise ValueError(f"oracle rejects descriptor: rc={rc}")
    return s


def run_layer(desc, in_words, weights, thresholds=None, bias=None, num_reps: int = 1) -> np.ndarray:
    """CPU restatement of one layer on packed byte images; returns the packed output stream (uint8)."""
    s = query(desc)
    in_words = np.ascontiguousarray(in_words, dtype=np.uint8)
    assert in_words.size == s.in_bytes_per_image * num_reps, (in_words.size, s.in_bytes_per_image, num_reps)
    weights = np.ascontiguousarray(weights if weights is not None else np.zeros(0, np.uint8), dtype=np.uint8)
    assert weights.size == s.weight_bytes, (weights.size, s.weight_bytes)
    if weights.size == 0:
        weights = np.zeros(1, np.uint8)  # (a valid pointer; Pool_batch has no parameters)
    if thresholds is not None:
        thresholds = np.ascontiguousarray(thresholds, dtype=np.uint8)
        assert thresholds.size == s.threshold_bytes
    if bias is not None:
        bias = np.ascontiguousarray(bias, dtype=np.uint8)
        assert bias.size == s.bias_bytes
    out = np.zeros(s.out_bytes_per_image * num_reps, dtype=np.uint8)
    c = desc.to_c()
    rc = lib().fo_layer_run(ctypes.byref(c), _ptr(in_words), _ptr(weights), _ptr(thresholds), _ptr(bias), _ptr(out), num_reps)
    if rc:
        raise RuntimeError(f"fo_layer_run rc={rc}")
    return out


def add_streams(in1, in2, n_words, channels, in1_bits, in1_signed, in2_bits, in2_signed, out_bits, offset=0) -> np.ndarray:
    """Restatement of AddStreams_Batch (streamtools.h:669-720) on packed word images."""
    L = lib()
    L.fo_add_streams.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_uint64, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_int, ctypes.c_uint32,
                                                         ctypes.c_int, ctypes.c_uint32, ctypes.c_int32]
    in1, in2 = np.ascontiguousarray(in1, dtype=np.uint8), np.ascontiguousarray(in2, dtype=np.uint8)
    out = np.zeros(int(L.fo_word_bytes(channels * out_bits)) * n_words, dtype=np.uint8)
    rc = L.fo_add_streams(_ptr(in1), _ptr(in2), _ptr(out), n_words, channels, in1_bits, in1_signed, in2_bits, in2_signed, out_bits, offset)
    if rc:
        raise RuntimeError(f"fo_add_streams rc={rc}")
    return out


def ref_add(name: str, in1, in2, out_bytes: int) -> np.ndarray:
    """Reference AddStreams_Batch instantiation `name` of ref_layers.cpp."""
    fn = getattr(ref_lib(), "ref_" + name)
    fn.argtypes = [ctypes.c_void_p] * 3
    in1, in2 = np.ascontiguousarray(in1, dtype=np.uint8), np.ascontiguousarray(in2, dtype=np.uint8)
    out = np.zeros(out_bytes, dtype=np.uint8)
    rc = fn(_ptr(in1), _ptr(in2), _ptr(out))
    if rc:
        raise RuntimeError(f"ref_{name} rc={rc}")
    return out


def maxpool(in_words, dim_x, dim_y, pool, ch, bits) -> np.ndarray:
    in_words = np.ascontiguousarray(in_words, dtype=np.uint8)
    wb = lib().fo_word_bytes(ch * bits)
    out = np.zeros(wb * (dim_x // pool) * (dim_y // pool), dtype=np.uint8)
    rc = lib().fo_maxpool(_ptr(in_words), _ptr(out), dim_x, dim_y, pool, ch, bits)
    if rc:
        raise RuntimeError(f"fo_maxpool rc={rc}")
    return out


# ---------------------------------------------------------------------------------------------
# the reference's own templates (oracle/_ref/libref_layers.so) -- present only where it was built
# ---------------------------------------------------------------------------------------------
_ref = None


def ref_available() -> bool:
    return os.path.exists(os.path.join(HERE, "_ref", "libref_layers.so"))


def ref_lib() -> ctypes.CDLL:
    global _ref
    if _ref is None:
        _ref = ctypes.CDLL(os.path.join(HERE, "_ref", "libref_layers.so"))
    return _ref


def ref_run(case: str, in_words, weights, third, out_bytes: int):
    """Run reference case `case` (see ref_layers.cpp); `third` is the bias or threshold image.
    Returns (packed output, seconds spent inside the reference's layer function)."""
    fn = getattr(ref_lib(), "ref_" + case)
    fn.argtypes = [ctypes.c_void_p] * 4 + [ctypes.POINTER(ctypes.c_double)]
    in_words = np.ascontiguousarray(in_words, dtype=np.uint8)
    weights = np.ascontiguousarray(weights, dtype=np.uint8)
    if weights.size == 0:
        weights = np.zeros(1, np.uint8)
    third = np.ascontiguousarray(third, dtype=np.uint8)
    out = np.zeros(out_bytes, dtype=np.uint8)
    secs = ctypes.c_double(0.0)
    rc = fn(_ptr(in_words), _ptr(weights), _ptr(third), _ptr(out), ctypes.byref(secs))
    if rc:
        raise RuntimeError(f"ref_{case} rc={rc}")
    return out, secs.value


def ref_pool(name: str, in_words, out_bytes: int) -> np.ndarray:
    fn = getattr(ref_lib(), "ref_" + name)
    fn.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    in_words = np.ascontiguousarray(in_words, dtype=np.uint8)
    out = np.zeros(out_bytes, dtype=np.uint8)
    rc = fn(_ptr(in_words), _ptr(out))
    if rc:
        raise RuntimeError(f"ref_{name} rc={rc}")
    return out


def gen_param_stream(desc, weights) -> np.ndarray:
    """Restatement of GenParamStream (dma.h:214-236) for one repetition: image of m_weights[PE][TILES] -> TILES stream words
    of SIMD*PE*WP bits, strMem((SIMD*WP)*(pe+1)-1, (SIMD*WP)*pe) = m_weights[pe][tile] (dma.h:229)."""
    s = query(desc)
    tiles, pe, lane_bits = s.sf * s.nf, desc.pe, desc.simd * desc.w_bits
    wb, sw = int(s.weight_word_bytes), int(lib().fo_word_bytes(lane_bits * pe))
    img = np.ascontiguousarray(weights, dtype=np.uint8).reshape(pe, tiles, wb)
    bits = np.unpackbits(img, axis=2, bitorder="little")[:, :, :lane_bits]          # [pe][tile][bit]
    word = np.zeros((tiles, sw * 8), dtype=np.uint8)
    word[:, : pe * lane_bits] = bits.transpose(1, 0, 2).reshape(tiles, pe * lane_bits)  # PE little-endian
    return np.packbits(word, axis=1, bitorder="little").reshape(-1)


def ref_stream_run(case: str, in_words, weights, thresholds, out_bytes: int, param_bytes: int):
    """Reference case with streamed weights (GenParamStream -> Matrix_Vector_Activate_Stream_Batch, ref_layers.cpp).
    Returns (packed output, first period of the parameter stream)."""
    fn = getattr(ref_lib(), "ref_stream_" + case)
    fn.argtypes = [ctypes.c_void_p] * 5
    in_words = np.ascontiguousarray(in_words, dtype=np.uint8)
    weights = np.ascontiguousarray(weights, dtype=np.uint8)
    thresholds = np.ascontiguousarray(thresholds, dtype=np.uint8)
    out = np.zeros(out_bytes, dtype=np.uint8)
    pw = np.zeros(param_bytes, dtype=np.uint8)
    rc = fn(_ptr(in_words), _ptr(weights), _ptr(thresholds), _ptr(out), _ptr(pw))
    if rc:
        raise RuntimeError(f"ref_stream_{case} rc={rc}")
    return out, pw
