"""Counter-based synthetic data: v(idx) = splitmix64(seed ^ idx).

The same generator exists as a CUDA kernel (csrc/fcb_synth.cu) so that host-made and
device-made tensors agree bit for bit without shipping data.
"""
from __future__ import annotations

import numpy as np

SEED_INPUT, SEED_WEIGHTS, SEED_BIAS, SEED_THRESH = 0x1001, 0x2001, 0x3001, 0x4001
_M = (1 << 64) - 1


def splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        x = (x.astype(np.uint64) + np.uint64(0x9E3779B97F4A7C15))
        z = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def rand_u64(seed: int, n: int, offset: int = 0) -> np.ndarray:
    idx = np.arange(offset, offset + n, dtype=np.uint64)
    return splitmix64(np.uint64(seed & _M) ^ idx)


def lanes(seed: int, shape, bits: int, signed: bool = False, mask: int | None = None, offset: int = 0) -> np.ndarray:
    """Uniform `bits`-bit lanes; `mask` narrows the range (0x7F = reachable range after bias+ReLU wrap8)."""
    n = int(np.prod(shape))
    v = rand_u64(seed, n, offset) & np.uint64((1 << bits) - 1 if mask is None else mask)
    v = v.astype(np.int64)
    if signed:
        v = v - (1 << (bits - 1))
    return v.reshape(shape)


def weights(seed: int, ofm: int, k: int, w_bits: int) -> np.ndarray:
    """s<w_bits> weights W[OFM, K]: (v & (2^b-1)) - 2^(b-1); 1-bit -> {0,1}."""
    if w_bits == 1:
        return lanes(seed, (ofm, k), 1)
    return lanes(seed, (ofm, k), w_bits, signed=True)


def bias(seed: int, ofm: int) -> np.ndarray:
    return lanes(seed, (ofm,), 8, signed=True)


def thresholds(seed: int, ofm: int, num_th: int, lo: int, hi: int) -> np.ndarray:
    """num_th values per channel uniform in [lo, hi], sorted ascending (order is irrelevant to the reference)."""
    v = rand_u64(seed, ofm * num_th).astype(np.float64) / float(1 << 64)
    t = (lo + np.floor(v * (hi - lo + 1))).astype(np.int64).reshape(ofm, num_th)
    return np.sort(t, axis=1)
