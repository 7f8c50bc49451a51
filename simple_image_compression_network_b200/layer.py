"""Host-side mirror of the reference's layer interface over the C ABI.

`ConvLayer` stands where an instantiation of conv2d<> / deconv522<> (conv_nonsquare_top.cpp:71-280)
or ConvLayer_Batch (convlayer.h:89-125) stands in the reference: it is built from the same
parameter set and the same packed weight / threshold / bias images, and `run(in_words, numReps)`
has the meaning of `top(in_stream, out_stream, numReps)`.  `Net` chains layers the way
eight_layers_net does (conv_nonsquare_top.cpp:295-357).  All compute happens in
libfinnconv_b200.so on the GPU.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib, pack
from .desc import LayerDesc


def _ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


class ConvLayer:
    def __init__(self, desc: LayerDesc, weights, thresholds=None, bias=None, device: int = 0):
        L = self._L = _lib.lib()  # the handle stays with the library build that created it
        self.desc = desc
        self.device = device
        c = desc.to_c()
        sizes = [ctypes.c_size_t() for _ in range(5)]
        _lib.check(L.fcb_layer_query(ctypes.byref(c), *[ctypes.byref(s) for s in sizes]), L)
        self.in_bytes, self.out_bytes, self.weight_bytes, self.threshold_bytes, self.bias_bytes = [s.value for s in sizes]
        w = np.ascontiguousarray(weights if weights is not None else np.zeros(0, np.uint8), dtype=np.uint8)  # (Pool_batch: none)
        if w.size != self.weight_bytes:
            raise ValueError(f"weight image is {w.size} bytes, expected {self.weight_bytes}")
        t = None if thresholds is None else np.ascontiguousarray(thresholds, dtype=np.uint8)
        b = None if bias is None else np.ascontiguousarray(bias, dtype=np.uint8)
        if t is not None and t.size != self.threshold_bytes:
            raise ValueError(f"threshold image is {t.size} bytes, expected {self.threshold_bytes}")
        if b is not None and b.size != self.bias_bytes:
            raise ValueError(f"bias image is {b.size} bytes, expected {self.bias_bytes}")
        h = ctypes.c_void_p()
        _lib.check(L.fcb_layer_create(ctypes.byref(c), _ptr(w), _ptr(t), _ptr(b), device, ctypes.byref(h)), L)
        self._h = h

    def set_params(self, weights, thresholds=None, bias=None) -> None:
        """Swap the layer's weights / thresholds / bias in place (the reference's run-time-writable weight memories,
        dma.h:214-236 + mvau.hpp:209-307); the handle, and any Net built on it, stay valid."""
        w = np.ascontiguousarray(weights, dtype=np.uint8)
        t = None if thresholds is None else np.ascontiguousarray(thresholds, dtype=np.uint8)
        b = None if bias is None else np.ascontiguousarray(bias, dtype=np.uint8)
        if w.size != self.weight_bytes or (t is not None and t.size != self.threshold_bytes) or (b is not None and b.size != self.bias_bytes):
            raise ValueError("parameter image size does not match the layer")
        _lib.check(self._L.fcb_layer_set_params(self._h, _ptr(w), _ptr(t), _ptr(b)), self._L)

    def set_param_stream(self, param_words, thresholds=None, bias=None) -> None:
        """set_params with the weights as one period of the reference's parameter stream (GenParamStream, dma.h:214-236:
        TILES words of SIMD*PE*WP bits, as Matrix_Vector_Activate_Stream_Batch consumes them, mvau.hpp:262-266)."""
        d = self.desc
        tiles = (d.kernel_x * d.kernel_y * d.ifm_ch // d.simd) * (d.ofm_ch // d.pe)
        w = np.ascontiguousarray(param_words, dtype=np.uint8)
        t = None if thresholds is None else np.ascontiguousarray(thresholds, dtype=np.uint8)
        b = None if bias is None else np.ascontiguousarray(bias, dtype=np.uint8)
        if w.size != tiles * pack.word_bytes(d.simd * d.pe * d.w_bits) or (t is not None and t.size != self.threshold_bytes) or \
                (b is not None and b.size != self.bias_bytes):
            raise ValueError("parameter image size does not match the layer")
        _lib.check(self._L.fcb_layer_set_param_stream(self._h, _ptr(w), _ptr(t), _ptr(b)), self._L)

    @property
    def engine(self) -> str:
        return self._L.fcb_layer_engine(self._h).decode()

    @property
    def plan(self) -> str:
        return self._L.fcb_layer_plan(self._h).decode()

    @property
    def launches(self) -> int:
        return int(self._L.fcb_layer_launches(self._h))

    def run(self, in_words, num_reps: int = 1) -> np.ndarray:
        """Host buffers in, host buffers out (H2D + kernels + D2H inside)."""
        x = np.ascontiguousarray(in_words, dtype=np.uint8).reshape(-1)
        if x.size != self.in_bytes * num_reps:
            raise ValueError(f"input stream is {x.size} bytes, expected {self.in_bytes * num_reps}")
        out = np.empty(self.out_bytes * num_reps, dtype=np.uint8)
        _lib.check(self._L.fcb_layer_run(self._h, _ptr(x), _ptr(out), num_reps), self._L)
        return out

    def run_raw(self, in_ptr: int, out_ptr: int, num_reps: int) -> None:
        """Host pointers (e.g. pinned memory) -- the same call as run() without numpy."""
        _lib.check(self._L.fcb_layer_run(self._h, ctypes.c_void_p(in_ptr), ctypes.c_void_p(out_ptr), num_reps), self._L)

    def run_device(self, d_in: int, d_out: int, num_reps: int, stream: int = 0) -> None:
        """Device pointers, asynchronous on `stream` (a cudaStream_t handle as int)."""
        _lib.check(self._L.fcb_layer_run_device(self._h, ctypes.c_void_p(d_in), ctypes.c_void_p(d_out), num_reps,
                                                ctypes.c_void_p(stream)), self._L)

    def set_host_chunk(self, images: int) -> None:
        """Images per staging slot of run() / run_raw() (0 = default) -- the burst length of the reference's
        Mem2Stream_Batch / Stream2Mem_Batch (dma.h:166-176)."""
        _lib.check(self._L.fcb_layer_set_host_chunk(self._h, images), self._L)

    def close(self):
        if getattr(self, "_h", None):
            self._L.fcb_layer_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Net:
    """A chain of layers whose intermediate streams stay in device memory."""

    def __init__(self, layers):
        self.layers = list(layers)
        L = self._L = self.layers[0]._L
        if any(l._L is not L for l in self.layers):
            raise ValueError("layers of a Net must come from the same library build")
        arr = (ctypes.c_void_p * len(self.layers))(*[l._h for l in self.layers])
        h = ctypes.c_void_p()
        _lib.check(L.fcb_net_create(arr, len(self.layers), ctypes.byref(h)), L)
        self._h = h
        self.in_bytes = self.layers[0].in_bytes
        self.out_bytes = self.layers[-1].out_bytes

    @property
    def launches(self) -> int:
        return int(self._L.fcb_net_launches(self._h))

    def run(self, in_words, num_reps: int = 1) -> np.ndarray:
        x = np.ascontiguousarray(in_words, dtype=np.uint8).reshape(-1)
        if x.size != self.in_bytes * num_reps:
            raise ValueError(f"input stream is {x.size} bytes, expected {self.in_bytes * num_reps}")
        out = np.empty(self.out_bytes * num_reps, dtype=np.uint8)
        _lib.check(self._L.fcb_net_run(self._h, _ptr(x), _ptr(out), num_reps), self._L)
        return out

    def run_raw(self, in_ptr: int, out_ptr: int, num_reps: int) -> None:
        """Host pointers (e.g. pinned memory): H2D, the layers and D2H of consecutive chunks overlap on three streams."""
        _lib.check(self._L.fcb_net_run(self._h, ctypes.c_void_p(in_ptr), ctypes.c_void_p(out_ptr), num_reps), self._L)

    def run_device(self, d_in: int, d_out: int, num_reps: int, stream: int = 0) -> None:
        _lib.check(self._L.fcb_net_run_device(self._h, ctypes.c_void_p(d_in), ctypes.c_void_p(d_out), num_reps,
                                              ctypes.c_void_p(stream)), self._L)

    def set_host_chunk(self, images: int) -> None:
        """Images per staging slot of run() / run_raw() (0 = default)."""
        _lib.check(self._L.fcb_net_set_host_chunk(self._h, images), self._L)

    def set_device_chunk(self, images: int) -> None:
        """Images per pass of the layer chain inside run_device() (0 = default); small values keep the intermediates in L2."""
        _lib.check(self._L.fcb_net_set_device_chunk(self._h, images), self._L)

    def close(self):
        if getattr(self, "_h", None):
            self._L.fcb_net_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Pool:
    """The chain replicated on several GPUs behind one handle (fcb_pool_*): `run(in_words, numReps)` has the meaning of
    eight_layers_net(in, out, numReps) (conv_nonsquare_top.cpp:295) and uses every listed device from one call; the batch is split
    into contiguous image ranges (shard.shard_range == fcb_shard_range), no device talks to another."""

    def __init__(self, descs, weights, thresholds=None, biases=None, devices=None):
        L = self._L = _lib.lib()
        n = len(descs)
        self.descs = list(descs)
        keep = []

        def images(seq):
            arr = (ctypes.c_void_p * n)()
            for i in range(n):
                a = None if seq is None or seq[i] is None else np.ascontiguousarray(seq[i], dtype=np.uint8)
                keep.append(a)
                arr[i] = None if a is None else a.ctypes.data
            return arr
        cd = (type(descs[0].to_c()) * n)(*[d.to_c() for d in descs])
        dev = None if not devices else (ctypes.c_int * len(devices))(*devices)
        h = ctypes.c_void_p()
        _lib.check(L.fcb_pool_create(cd, images(weights), images(thresholds), images(biases), n, dev, len(devices or []), ctypes.byref(h)), L)
        self._h = h
        sizes = [ctypes.c_size_t() for _ in range(5)]
        c0, c1 = descs[0].to_c(), descs[-1].to_c()
        _lib.check(L.fcb_layer_query(ctypes.byref(c0), *[ctypes.byref(s) for s in sizes]), L)
        self.in_bytes = sizes[0].value
        _lib.check(L.fcb_layer_query(ctypes.byref(c1), *[ctypes.byref(s) for s in sizes]), L)
        self.out_bytes = sizes[1].value

    @property
    def replicas(self) -> int:
        return int(self._L.fcb_pool_replicas(self._h))

    def run(self, in_words, num_reps: int = 1) -> np.ndarray:
        x = np.ascontiguousarray(in_words, dtype=np.uint8).reshape(-1)
        if x.size != self.in_bytes * num_reps:
            raise ValueError(f"input stream is {x.size} bytes, expected {self.in_bytes * num_reps}")
        out = np.empty(self.out_bytes * num_reps, dtype=np.uint8)
        _lib.check(self._L.fcb_pool_run(self._h, _ptr(x), _ptr(out), num_reps), self._L)
        return out

    def run_raw(self, in_ptr: int, out_ptr: int, num_reps: int) -> None:
        _lib.check(self._L.fcb_pool_run(self._h, ctypes.c_void_p(in_ptr), ctypes.c_void_p(out_ptr), num_reps), self._L)

    def close(self):
        if getattr(self, "_h", None):
            self._L.fcb_pool_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def add_streams(in1, in2, n_words: int, channels: int, in1_bits: int, in1_signed: bool, in2_bits: int, in2_signed: bool, out_bits: int,
                offset: int = 0, device: int = 0) -> np.ndarray:
    """AddStreams_Batch (streamtools.h:669-720) on packed word images (host buffers): per word and channel
    Out_t(op1 + op2 + offset), the residual add of the library."""
    from .desc import CAddDesc
    L = _lib.lib()
    d = CAddDesc(ctypes.sizeof(CAddDesc), channels, in1_bits, int(in1_signed), in2_bits, int(in2_signed), out_bits, offset)
    a, b = np.ascontiguousarray(in1, dtype=np.uint8), np.ascontiguousarray(in2, dtype=np.uint8)
    if a.size != pack.word_bytes(channels * in1_bits) * n_words or b.size != pack.word_bytes(channels * in2_bits) * n_words:
        raise ValueError("stream image size does not match channels x bits x n_words")
    out = np.empty(pack.word_bytes(channels * out_bits) * n_words, dtype=np.uint8)
    _lib.check(L.fcb_add_streams(ctypes.byref(d), _ptr(a), _ptr(b), _ptr(out), n_words, device), L)
    return out


class FCLayer:
    """StreamingFCLayer_Batch (fclayer.h:83-111): a thin wrapper of the same MVAU, i.e. a 1x1 layer on one-pixel frames.  `reps`
    repetitions are presented to the library as frames of `tile` pixels (a 1x1 layer treats every pixel on its own, and the stream
    image -- input vectors back to back -- is the same), the remainder as one-pixel frames, so large batches fill the tensor tiles."""

    def __init__(self, matrix_w: int, matrix_h: int, simd: int, pe: int, weights, thresholds=None, bias=None, tile: int = 256, device: int = 0, **numerics):
        from .desc import KIND_CONV
        mk = lambda x: LayerDesc(kind=KIND_CONV, kernel_x=1, kernel_y=1, ifm_ch=matrix_w, ofm_ch=matrix_h, ifm_x=x, ifm_y=1, stride_x=1,
                                 stride_y=1, pad=0, simd=simd, pe=pe, **numerics)
        self.tile = tile
        self.big = ConvLayer(mk(tile), weights, thresholds=thresholds, bias=bias, device=device)
        self.one = ConvLayer(mk(1), weights, thresholds=thresholds, bias=bias, device=device)
        self.in_bytes, self.out_bytes = self.one.in_bytes, self.one.out_bytes

    def run(self, in_words, reps: int) -> np.ndarray:
        x = np.ascontiguousarray(in_words, dtype=np.uint8).reshape(-1)
        if x.size != self.in_bytes * reps:
            raise ValueError(f"input stream is {x.size} bytes, expected {self.in_bytes * reps}")
        nbig = reps // self.tile
        parts = []
        if nbig:
            parts.append(self.big.run(x[: nbig * self.tile * self.in_bytes], nbig))
        if reps - nbig * self.tile:
            parts.append(self.one.run(x[nbig * self.tile * self.in_bytes:], reps - nbig * self.tile))
        return np.concatenate(parts)


def synth_fill(d_ptr: int, n_bytes: int, seed: int, mask: int = 0xFF, offset: int = 0, stream: int = 0) -> None:
    _lib.check(_lib.lib().fcb_synth_fill(ctypes.c_void_p(d_ptr), n_bytes, seed, mask, offset, ctypes.c_void_p(stream)))
