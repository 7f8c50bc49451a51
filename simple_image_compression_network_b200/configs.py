"""The reference network's layer parameters (config_nonsquare.h:1-135) as run-time descriptors, and the
seeded synthetic parameters/benchmark tensors of SURVEY.md 8(d).  Host-side data only."""
from __future__ import annotations

from . import pack, synth
from .desc import ACT_BIAS_RELU, ACT_THRESHOLDS, KIND_CONV, KIND_DECONV522, KIND_POOL, W_BINARY_XNOR, LayerDesc

# (kind, ifm_ch, ifm_x("ROW"), ifm_y("COL"), ofm_ch, simd, pe) -- CONV_n_* of config_nonsquare.h; K5 S2 P2, 8-bit
# activations, 4-bit weights everywhere; layers 0-3 are conv2d<>, 4-7 deconv522<> (conv_nonsquare_top.cpp:295-357)
_NET = [
    (KIND_CONV, 3, 768, 512, 128, 3, 8),
    (KIND_CONV, 128, 384, 256, 128, 8, 16),
    (KIND_CONV, 128, 192, 128, 128, 8, 16),
    (KIND_CONV, 128, 96, 64, 192, 8, 24),
    (KIND_DECONV522, 192, 48, 32, 128, 12, 16),
    (KIND_DECONV522, 128, 96, 64, 128, 8, 16),
    (KIND_DECONV522, 128, 192, 128, 128, 8, 16),
    (KIND_DECONV522, 128, 384, 256, 3, 8, 3),
]


def net_layer(i: int) -> LayerDesc:
    """Descriptor of layer i (0..7) of eight_layers_net."""
    kind, c, x, y, ofm, simd, pe = _NET[i]
    return LayerDesc(kind=kind, kernel_x=5, kernel_y=5, ifm_ch=c, ofm_ch=ofm, ifm_x=x, ifm_y=y, stride_x=2, stride_y=2, pad=2,
                     simd=simd, pe=pe, in_bits=8, in_signed=0, w_bits=4, acc_bits=8, acc_signed=0, act_kind=ACT_BIAS_RELU,
                     out_bits=8)


def synthetic_params(d: LayerDesc, seed_shift: int = 0):
    """Seeded weights (+ bias or thresholds) for a descriptor -> dict of logical arrays and packed images."""
    import numpy as np
    k = d.k_total
    if d.kind == KIND_POOL:  # Pool_batch has no parameters
        return {"w": None, "weights": np.zeros(0, np.uint8), "bias": None, "thresholds": None, "b": None, "t": None}
    w = synth.weights(synth.SEED_WEIGHTS + seed_shift, d.ofm_ch, k, d.w_bits)
    out = {"w": w, "weights": pack.pack_weights(w, d.weight_simd, d.pe, d.w_bits), "bias": None, "thresholds": None, "b": None, "t": None}
    if d.act_kind == ACT_BIAS_RELU:
        b = synth.bias(synth.SEED_BIAS + seed_shift, d.ofm_ch)
        out["b"], out["bias"] = b, pack.pack_bias(b)
    if d.act_kind == ACT_THRESHOLDS:
        if d.weight_kind == W_BINARY_XNOR:
            # matches ~ Binomial(K, 1/2): thresholds around K/2 +- 2 sigma discriminate
            lo, hi = int(k / 2 - k ** 0.5), int(k / 2 + k ** 0.5)
        else:
            # thresholds where they discriminate: mean +- 2.5 sigma of sum_k w_k*a_k for uniform lanes
            av = np.arange(1 << d.in_bits, dtype=np.float64) - ((1 << (d.in_bits - 1)) if d.in_signed else 0)
            wv = np.arange(1 << d.w_bits, dtype=np.float64) - (1 << (d.w_bits - 1))
            mean = k * wv.mean() * av.mean()
            sigma = (k * ((wv ** 2).mean() * (av ** 2).mean() - (wv.mean() * av.mean()) ** 2)) ** 0.5
            lo, hi = int(mean - 2.5 * sigma), int(mean + 2.5 * sigma)
        lim = (1 << (d.acc_bits - 1)) - 1
        lo, hi = max(lo, -lim - 1), min(hi, lim)
        t = synth.thresholds(synth.SEED_THRESH + seed_shift, d.ofm_ch, d.num_th, lo, hi)
        out["t"], out["thresholds"] = t, pack.pack_thresholds(t, d.pe, d.acc_bits)
    return out


def synthetic_input(d: LayerDesc, seed_shift: int = 0, num_reps: int = 1, relu_range: bool = False):
    x = synth.lanes(synth.SEED_INPUT + seed_shift, (num_reps, d.ifm_y, d.ifm_x, d.ifm_ch), d.in_bits, signed=bool(d.in_signed),
                    mask=(0x7F if (relu_range and d.in_bits == 8) else None))
    return x, pack.pack_stream(x, d.in_bits)
