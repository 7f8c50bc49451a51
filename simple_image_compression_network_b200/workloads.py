"""The measured workloads of BASELINE.json as run-time descriptors (host-side data only).

  config 2   CONV_1 of config_nonsquare.h on 4096 images            -> configs.net_layer(1)           (the headline)
  config 3   1-bit W/A XNOR-popcount ConvLayer_Batch 64->64, 3x3, 128x96                       -> config3()
  config 4   4-bit W / 8-bit A, 255 thresholds + fused 2x2 max pool, 256->256, 64x48           -> config4()
  config 5a  the reference's own analysis transform: conv2d layers 0-3 of eight_layers_net     -> [configs.net_layer(i) for i in 0..3]
  config 5b  analysis-transform-shaped stack [K3 S1 P1 conv -> 255 thresholds -> 2x2 pool] x 4 -> stack5b()
  wide lanes 16-bit activations x 8-bit weights (north_star: "IMAD for wider types")           -> imad16()
"""
from __future__ import annotations

from .desc import ACT_PASSTHROUGH, ACT_THRESHOLDS, KIND_CONV, W_BINARY_XNOR, LayerDesc

# roofline denominators that are not in MEASURED_PEAKS.json
INT8_SPEC_TOPS = 4500.0        # dense INT8, B200 data sheet (the sparse figure is twice that and is not used)
INT8_MMA_ONLY_TOPS = 4380.0    # tools/umma_peak.cu, profiles/r01_umma_peak_int8.log (MMA issue only, operands resident)
POPC_WORDS_PER_S = 4.58e12     # tools/popc_peak.cu, profiles/r01_popc_peak.log (15.7 lane-popc / clk / SM at 1965 MHz)
# integer multiply-add and the packed dot products issue on the heavy half of the FMA pipe: 64 lanes / clk / SM at 1965 MHz
# (profiles/r02_imad_ncu_full_summary.txt: sm__pipe_fmaheavy_cycles_active); MACs per instruction: IMAD 1, IDP.2A 2, IDP.4A 4
INT_DOT_INSTR_PER_S = 148 * 64 * 1.965e9


def direct_macs_per_s(d: LayerDesc) -> float:
    """Ceiling of the universal direct engine for a layer with weights of at most 8 bits (csrc/fcb_direct.cu, dot_conv_kernel)."""
    return INT_DOT_INSTR_PER_S * (4 if d.in_bits <= 8 else 2)


def config3() -> LayerDesc:
    return LayerDesc(kind=KIND_CONV, kernel_x=3, kernel_y=3, ifm_ch=64, ofm_ch=64, ifm_x=128, ifm_y=96, stride_x=1, stride_y=1, pad=0,
                     simd=64, pe=16, in_bits=1, w_bits=1, weight_kind=W_BINARY_XNOR, acc_bits=16, acc_signed=1,
                     act_kind=ACT_THRESHOLDS, out_bits=1, num_th=1)


def config4() -> LayerDesc:
    return LayerDesc(kind=KIND_CONV, kernel_x=3, kernel_y=3, ifm_ch=256, ofm_ch=256, ifm_x=64, ifm_y=48, stride_x=1, stride_y=1, pad=1,
                     simd=32, pe=32, in_bits=8, w_bits=4, acc_bits=24, acc_signed=1, act_kind=ACT_THRESHOLDS, out_bits=8, num_th=255,
                     pool=2)


def _stage(c, ofm, x, y, simd, pe) -> LayerDesc:
    return LayerDesc(kind=KIND_CONV, kernel_x=3, kernel_y=3, ifm_ch=c, ofm_ch=ofm, ifm_x=x, ifm_y=y, stride_x=1, stride_y=1, pad=1,
                     simd=simd, pe=pe, in_bits=8, w_bits=4, acc_bits=24, acc_signed=1, act_kind=ACT_THRESHOLDS, out_bits=8,
                     num_th=255, pool=2)


def stack5b() -> list[LayerDesc]:
    """768x512x3 -> 384x256x128 -> 192x128x128 -> 96x64x128 -> 48x32x192."""
    return [_stage(3, 128, 768, 512, 3, 16), _stage(128, 128, 384, 256, 32, 16), _stage(128, 128, 192, 128, 32, 16),
            _stage(128, 192, 96, 64, 32, 24)]


def imad16() -> LayerDesc:
    """Wide lanes: ap_int<16> activations x ap_int<8> weights, 32-bit pass-through accumulators (no tensor-core form: kind::i8
    takes 8-bit operands), 64 -> 64 channels, 3x3, 96x64."""
    return LayerDesc(kind=KIND_CONV, kernel_x=3, kernel_y=3, ifm_ch=64, ofm_ch=64, ifm_x=96, ifm_y=64, stride_x=1, stride_y=1, pad=1,
                     simd=16, pe=16, in_bits=16, in_signed=1, w_bits=8, acc_bits=32, acc_signed=1, act_kind=ACT_PASSTHROUGH, out_bits=32)


def nonzero_macs(d: LayerDesc) -> float:
    """MACs that are not structural zeros (deconv522 multiplies 75 % inserted zeros in the reference: SURVEY.md A.6)."""
    return d.macs_per_image / 4 if d.kind == 1 else d.macs_per_image
