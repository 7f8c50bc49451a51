"""oracle/cases.py -- TEST INFRASTRUCTURE ONLY.

The parity cases: one LayerDesc per instantiation in oracle/ref_layers.cpp (same names), plus the
seeded synthetic tensors each case is fed (SURVEY.md 8(d) seeding rule).
"""
from __future__ import annotations

import dataclasses

import numpy as np

from simple_image_compression_network_b200 import pack, synth
from simple_image_compression_network_b200.desc import (ACT_BIAS_RELU, ACT_PASSTHROUGH, ACT_THRESHOLDS, KIND_CONV, KIND_DECONV522,
                                                        KIND_DWCONV, KIND_POOL, POOLFN_ACC, POOLFN_AVG, POOLFN_MAX, POOLFN_QUANTAVG,
                                                        W_BINARY_XNOR, W_FIXED, LayerDesc)


def _c2d(kx, ky, simd, pe, wb, c, ofm, ix, iy, s, p, inb, actb):
    return LayerDesc(kind=KIND_CONV, kernel_x=kx, kernel_y=ky, ifm_ch=c, ofm_ch=ofm, ifm_x=ix, ifm_y=iy, stride_x=s,
                     stride_y=s, pad=p, simd=simd, pe=pe, in_bits=inb, w_bits=wb, acc_bits=actb, acc_signed=0,
                     act_kind=ACT_BIAS_RELU, out_bits=actb)


def _dc(ix, iy, c, ofm, simd, pe, wb):
    return LayerDesc(kind=KIND_DECONV522, kernel_x=5, kernel_y=5, ifm_ch=c, ofm_ch=ofm, ifm_x=ix, ifm_y=iy, stride_x=2,
                     stride_y=2, pad=2, simd=simd, pe=pe, in_bits=8, w_bits=wb, acc_bits=8, acc_signed=0,
                     act_kind=ACT_BIAS_RELU, out_bits=8)


def _th(k, simd, pe, wb, c, ofm, ix, iy, p, inb, nth, tab, trb, av, pool=0):
    return LayerDesc(kind=KIND_CONV, kernel_x=k, kernel_y=k, ifm_ch=c, ofm_ch=ofm, ifm_x=ix, ifm_y=iy, stride_x=1,
                     stride_y=1, pad=p, simd=simd, pe=pe, in_bits=inb, w_bits=wb, acc_bits=tab, acc_signed=1,
                     act_kind=ACT_THRESHOLDS, out_bits=trb, num_th=nth, act_val=av, pool=pool)


def _xn(k, simd, pe, c, ofm, ix, iy, tab):
    return LayerDesc(kind=KIND_CONV, kernel_x=k, kernel_y=k, ifm_ch=c, ofm_ch=ofm, ifm_x=ix, ifm_y=iy, stride_x=1,
                     stride_y=1, pad=0, simd=simd, pe=pe, in_bits=1, w_bits=1, weight_kind=W_BINARY_XNOR, acc_bits=tab,
                     acc_signed=1, act_kind=ACT_THRESHOLDS, out_bits=1, num_th=1, act_val=0)


def _pl(k, c, pe, ix, iy, s, p, inb, ins, tab, tas, outb, fn, size=0):
    """Pool_batch behind the depth-wise sliding window (maxpool.h:525-577, pool.hpp:94-226)."""
    return LayerDesc(kind=KIND_POOL, kernel_x=k, kernel_y=k, ifm_ch=c, ofm_ch=c, ifm_x=ix, ifm_y=iy, stride_x=s, stride_y=s, pad=p,
                     simd=pe, pe=pe, in_bits=inb, in_signed=ins, w_bits=0, weight_kind=fn, acc_bits=tab, acc_signed=tas,
                     act_kind=ACT_PASSTHROUGH, out_bits=outb, act_val=size)


def _dw(k, c, pe, ix, iy, s, p, wb, tab, outb, nth=0):
    """Depth-wise convolution: dws sliding window + Vector_Vector_Activate_Batch (vvau.hpp:80-154)."""
    return LayerDesc(kind=KIND_DWCONV, kernel_x=k, kernel_y=k, ifm_ch=c, ofm_ch=c, ifm_x=ix, ifm_y=iy, stride_x=s, stride_y=s, pad=p,
                     simd=pe, pe=pe, in_bits=8, in_signed=0, w_bits=wb, acc_bits=tab, acc_signed=1,
                     act_kind=ACT_THRESHOLDS if nth else ACT_PASSTHROUGH, out_bits=outb, num_th=nth)


# name -> LayerDesc ; names and parameters mirror the X-macro tables of ref_layers.cpp
CASES = {
    "c2d_a": _c2d(5, 5, 2, 3, 4, 4, 6, 12, 8, 2, 2, 8, 8),
    "c2d_b": _c2d(5, 5, 8, 8, 4, 16, 32, 40, 24, 2, 2, 8, 8),
    "c2d_c": _c2d(5, 5, 3, 8, 4, 3, 16, 32, 20, 2, 2, 8, 8),
    "c2d_d": _c2d(3, 3, 16, 4, 8, 32, 32, 20, 12, 1, 1, 8, 8),
    "c2d_e": _c2d(5, 5, 8, 16, 4, 128, 128, 48, 32, 2, 2, 8, 8),
    "c2d_f": _c2d(3, 3, 4, 2, 4, 8, 8, 10, 6, 1, 1, 8, 16),
    "c2d_g": _c2d(5, 5, 8, 24, 4, 128, 192, 24, 16, 2, 2, 8, 8),
    "c2d_L1band": _c2d(5, 5, 8, 16, 4, 128, 128, 384, 32, 2, 2, 8, 8),
    "c2d_L1": _c2d(5, 5, 8, 16, 4, 128, 128, 384, 256, 2, 2, 8, 8),
    "dc_a": _dc(6, 4, 4, 6, 2, 3, 4),
    "dc_b": _dc(12, 8, 16, 16, 8, 8, 4),
    "dc_c": _dc(24, 16, 128, 128, 8, 16, 4),
    "dc_d": _dc(24, 16, 128, 3, 8, 3, 4),
    "dc_e": _dc(12, 8, 192, 128, 12, 16, 4),
    "dc_L4": _dc(48, 32, 192, 128, 12, 16, 4),
    "th_a": _th(3, 4, 2, 4, 8, 8, 10, 6, 1, 8, 15, 24, 4, 0),
    "th_b": _th(3, 16, 8, 4, 32, 32, 16, 12, 1, 8, 255, 24, 8, 0),
    "th_c": _th(3, 8, 4, 4, 16, 16, 9, 7, 0, 8, 3, 16, 2, 0),
    "th_d": _th(3, 3, 8, 4, 3, 16, 16, 12, 1, 8, 255, 24, 8, 0),
    "th_cfg4": _th(3, 32, 32, 4, 256, 256, 64, 48, 1, 8, 255, 24, 8, 0),
    "xn_a": _xn(3, 8, 4, 8, 8, 12, 10, 16),
    "xn_b": _xn(3, 64, 16, 64, 64, 16, 12, 16),
    "xn_c": _xn(3, 32, 8, 64, 32, 20, 9, 16),
    # FMPadding_nonsquare totals + style, StreamingMaxPool_Precision forms, odd lane widths (ref_layers.cpp: run_thresh_pad_pool)
    "px_odd2": dataclasses.replace(_th(3, 4, 2, 4, 8, 8, 10, 6, 0, 8, 15, 24, 4, 0), pad_style=2, pad_x_total=3, pad_y_total=1),
    "px_odd1": dataclasses.replace(_th(3, 4, 2, 4, 8, 8, 10, 6, 0, 8, 15, 24, 4, 0), pad_style=1, pad_x_total=3, pad_y_total=1),
    "pk3_signed": dataclasses.replace(_th(3, 4, 2, 4, 8, 8, 12, 12, 1, 8, 15, 24, 4, 0, pool=3), pool_signed=1, pool_min_value=-8),
    "pk2_min5": dataclasses.replace(_th(3, 4, 2, 4, 8, 8, 12, 12, 1, 8, 15, 24, 4, 0, pool=2), pool_min_value=5),
    "lw3": _th(3, 4, 2, 3, 8, 8, 10, 6, 1, 3, 7, 12, 3, 0),
    "lw5x12": _th(3, 3, 4, 5, 6, 12, 9, 7, 0, 5, 40, 16, 6, 0),
    "acc40": LayerDesc(kind=KIND_CONV, kernel_x=3, kernel_y=3, ifm_ch=8, ofm_ch=8, ifm_x=10, ifm_y=6, stride_x=1, stride_y=1, pad=1, simd=4,
                       pe=2, in_bits=16, in_signed=1, w_bits=16, acc_bits=40, acc_signed=1, act_kind=ACT_PASSTHROUGH, out_bits=32),
    # sliding-window variants (ref_layers.cpp: run_swg_variant): dilated (slidingwindow.h:1515-1631), kernel_stride (K % S != 0, :447-575)
    "dil_x2": dataclasses.replace(_th(3, 4, 2, 4, 8, 8, 14, 8, 0, 8, 15, 24, 4, 0), dilation_x=2),
    "dil_x3_k2": dataclasses.replace(_th(3, 8, 4, 4, 16, 8, 13, 7, 0, 8, 15, 24, 4, 0), kernel_x=2, kernel_y=3, dilation_x=3),
    "ks_k3s2": dataclasses.replace(_th(3, 4, 2, 4, 8, 8, 11, 11, 0, 8, 15, 24, 4, 0), stride_x=2, stride_y=2),
    # StreamingFCLayer_Batch (fclayer.h:83-111) = a 1x1 layer over `reps` one-pixel frames, here laid out as one 7-pixel row
    "fc_a": LayerDesc(kind=KIND_CONV, kernel_x=1, kernel_y=1, ifm_ch=64, ofm_ch=32, ifm_x=7, ifm_y=1, stride_x=1, stride_y=1, pad=0, simd=8,
                      pe=4, in_bits=8, in_signed=0, w_bits=4, acc_bits=16, acc_signed=1, act_kind=ACT_PASSTHROUGH, out_bits=16),
    # channel-wise units (ref_layers.cpp: run_pool_batch / run_vvau)
    "pl_max_a": _pl(2, 8, 4, 12, 12, 2, 0, 8, 0, 8, 0, 8, POOLFN_MAX),
    "pl_max_s": _pl(3, 4, 2, 10, 6, 1, 1, 8, 1, 8, 1, 8, POOLFN_MAX),
    "pl_avg": _pl(2, 8, 8, 8, 8, 2, 0, 8, 0, 10, 0, 8, POOLFN_AVG, 4),
    "pl_qavg": _pl(4, 4, 4, 16, 16, 4, 0, 8, 1, 12, 1, 8, POOLFN_QUANTAVG, 4),
    "pl_acc": _pl(3, 6, 3, 9, 7, 1, 0, 4, 0, 8, 0, 8, POOLFN_ACC, 9),
    "dw_a": _dw(3, 8, 4, 10, 6, 1, 1, 4, 16, 16),
    "dw_b": _dw(3, 16, 8, 12, 12, 1, 1, 4, 16, 4, nth=15),
    "dw_c": _dw(2, 4, 2, 8, 8, 2, 0, 4, 12, 12),
}

# cases that take long in the reference C-simulation (seconds): excluded from the quick sets
SLOW = {"c2d_L1": 30, "c2d_L1band": 4, "dc_L4": 12, "th_cfg4": 8, "xn_b": 30, "dc_c": 5}


def make_inputs(d: LayerDesc, seed_shift: int = 0, num_reps: int = 1, relu_range: bool = False):
    """Seeded synthetic tensors for a case -> dict of logical arrays and packed images."""
    from simple_image_compression_network_b200 import configs
    out = configs.synthetic_params(d, seed_shift)
    out["x"], out["in_words"] = configs.synthetic_input(d, seed_shift, num_reps, relu_range)
    return out


def third_image(inp):
    """The bias-or-threshold image the ref_* entry points take as their third argument."""
    if inp["bias"] is not None:
        return inp["bias"]
    return inp["thresholds"] if inp["thresholds"] is not None else np.zeros(1, np.uint8)  # (unused by pass-through / pool cases)
