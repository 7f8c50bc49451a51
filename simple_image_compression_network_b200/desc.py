"""Run-time mirror of the reference's compile-time layer parameters (ctypes view of fcb_layer_desc).

Field meaning follows conv2d<> (conv_nonsquare_top.cpp:198-215), deconv522<> (:71-81),
ConvLayer_Batch (convlayer.h:89-105) and ThresholdsActivation (activations.hpp:168-169).
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass

# conv2d<> | deconv522<> | depth-wise conv (dws sliding window + Vector_Vector_Activate_Batch, vvau.hpp:80-154) |
# generic pooling (dws sliding window + Pool_batch, maxpool.h:525-577)
KIND_CONV, KIND_DECONV522, KIND_DWCONV, KIND_POOL = 0, 1, 2, 3
# pool.hpp:94-226 function objects of Pool_batch (carried in `weight_kind` of a KIND_POOL descriptor; `act_val` = their `size`)
POOLFN_MAX, POOLFN_AVG, POOLFN_ACC, POOLFN_QUANTAVG = 0, 1, 2, 3
W_FIXED, W_BINARY_XNOR, W_BINARY_PM1 = 0, 1, 2
ACT_PASSTHROUGH, ACT_BIAS_RELU, ACT_THRESHOLDS = 0, 1, 2
CMP_LESS, CMP_GREATER, CMP_LESS_EQUAL, CMP_GREATER_EQUAL = 0, 1, 2, 3
# which arithmetic unit multiplies -- the reference's resource argument R (mvau.hpp:87-98); never changes the result
ENGINE_AUTO, ENGINE_IMAD, ENGINE_XNOR_POPC, ENGINE_TENSOR = 0, 1, 2, 3


class CLayerDesc(ctypes.Structure):
    _fields_ = [(n, ctypes.c_uint32) for n in (
        "struct_size", "kind", "kernel_x", "kernel_y", "ifm_ch", "ofm_ch", "ifm_x", "ifm_y", "ofm_x", "ofm_y",
        "stride_x", "stride_y", "pad", "simd", "pe", "in_bits", "in_signed", "w_bits", "weight_kind",
        "acc_bits", "acc_signed", "act_kind", "out_bits", "num_th")] + [
        ("act_val", ctypes.c_int32), ("cmp", ctypes.c_uint32), ("pool", ctypes.c_uint32), ("engine_hint", ctypes.c_uint32),
        ("pad_x_total", ctypes.c_uint32), ("pad_y_total", ctypes.c_uint32), ("pad_style", ctypes.c_uint32),
        ("pool_signed", ctypes.c_uint32), ("pool_min_value", ctypes.c_int32), ("dilation_x", ctypes.c_uint32), ("dilation_y", ctypes.c_uint32)]


class CAddDesc(ctypes.Structure):
    """fcb_add_desc: AddStreams_Batch's template parameters (streamtools.h:669-720)."""
    _fields_ = [(n, ctypes.c_uint32) for n in ("struct_size", "channels", "in1_bits", "in1_signed", "in2_bits", "in2_signed", "out_bits")] + [
        ("offset", ctypes.c_int32)]


@dataclass(frozen=True)
class LayerDesc:
    kind: int = KIND_CONV
    kernel_x: int = 5
    kernel_y: int = 5
    ifm_ch: int = 128
    ofm_ch: int = 128
    ifm_x: int = 384
    ifm_y: int = 256
    stride_x: int = 2
    stride_y: int = 2
    pad: int = 2
    simd: int = 8
    pe: int = 16
    in_bits: int = 8
    in_signed: int = 0
    w_bits: int = 4
    weight_kind: int = W_FIXED
    acc_bits: int = 8
    acc_signed: int = 0
    act_kind: int = ACT_BIAS_RELU
    out_bits: int = 8
    num_th: int = 0
    act_val: int = 0
    cmp: int = CMP_LESS
    pool: int = 0
    engine_hint: int = ENGINE_AUTO
    # FMPadding_nonsquare's Padding_x / Padding_y / PaddingStyle (streamtools.h:361-379), used when pad_style != 0 (then pad = 0)
    pad_x_total: int = 0
    pad_y_total: int = 0
    pad_style: int = 0
    # StreamingMaxPool_Precision's ActType signedness and min_value (maxpool.h:137-170)
    pool_signed: int = 0
    pool_min_value: int = 0
    # ConvolutionInputGenerator_NonSquare_Dilated (slidingwindow.h:1515-1631); 1 = none
    dilation_x: int = 1
    dilation_y: int = 1

    # ---- derived geometry (conv_nonsquare_top.cpp:238-259 / :109-169) ----
    @property
    def pads(self):
        """(left, right, up, down) zeros, split as FMPadding_nonsquare does (streamtools.h:374-379)."""
        if not self.pad_style:
            return (self.pad,) * 4
        extra = 1 if self.pad_style == 2 else 0
        left = self.pad_x_total // 2 + extra * (self.pad_x_total % 2)
        up = self.pad_y_total // 2 + extra * (self.pad_y_total % 2)
        return left, self.pad_x_total - left, up, self.pad_y_total - up

    @property
    def ofm_x(self) -> int:
        if self.kind == KIND_DECONV522:
            return 2 * self.ifm_x
        left, right, _, _ = self.pads
        return (self.ifm_x + left + right - ((self.kernel_x - 1) * self.dilation_x + 1)) // self.stride_x + 1

    @property
    def ofm_y(self) -> int:
        if self.kind == KIND_DECONV522:
            return 2 * self.ifm_y
        _, _, up, down = self.pads
        return (self.ifm_y + up + down - ((self.kernel_y - 1) * self.dilation_y + 1)) // self.stride_y + 1

    @property
    def out_x(self) -> int:
        return self.ofm_x // max(self.pool, 1)

    @property
    def out_y(self) -> int:
        return self.ofm_y // max(self.pool, 1)

    @property
    def k_total(self) -> int:
        """Inputs per output lane: the MVAU's MatrixW, or Kernel_2 for the channel-wise units."""
        if self.kind in (KIND_DWCONV, KIND_POOL):
            return self.kernel_x * self.kernel_y
        return self.kernel_x * self.kernel_y * self.ifm_ch

    @property
    def weight_simd(self) -> int:
        """Lanes per weight word: SIMD, or 1 for Vector_Vector_Activate_Batch (vvau.hpp:128-134)."""
        return 1 if self.kind == KIND_DWCONV else self.simd

    @property
    def macs_per_image(self) -> int:
        """MACs as the reference executes them (dense, structural zeros of deconv included)."""
        return self.ofm_x * self.ofm_y * self.ofm_ch * self.k_total

    def to_c(self) -> CLayerDesc:
        c = CLayerDesc()
        c.struct_size = ctypes.sizeof(CLayerDesc)
        for f in ("kind", "kernel_x", "kernel_y", "ifm_ch", "ofm_ch", "ifm_x", "ifm_y", "stride_x", "stride_y", "pad",
                  "simd", "pe", "in_bits", "in_signed", "w_bits", "weight_kind", "acc_bits", "acc_signed", "act_kind",
                  "out_bits", "num_th", "act_val", "cmp", "pool", "engine_hint", "pad_x_total", "pad_y_total", "pad_style",
                  "pool_signed", "pool_min_value", "dilation_x", "dilation_y"):
            setattr(c, f, getattr(self, f))
        c.ofm_x, c.ofm_y = self.ofm_x, self.ofm_y
        return c
