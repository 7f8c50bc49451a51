"""Host-side packing of logical tensors into the reference's memory images.

The reference moves data as `ap_uint<W>` stream words (dma.h:135-199, streamtools.h:463-526)
and keeps parameters as `m_weights[PE][TILES]` (weights.hpp:69,113) and
`m_thresholds[PE][NF][NumTH]` (activations.hpp:172).  These helpers build exactly those byte
images ("ap-word containers": 1/2/4/8 bytes for W <= 8/16/32/64, else 8*ceil(W/64), little
endian, lane 0 at the LSB -- interpret.hpp:211) from plain numpy arrays,
and cut them apart again.  Pure host code; no device work happens here.
"""
from __future__ import annotations

import numpy as np


def word_bytes(bits: int) -> int:
    """Size of the container of one ap_uint<bits>."""
    if bits <= 8:
        return 1
    if bits <= 16:
        return 2
    if bits <= 32:
        return 4
    if bits <= 64:
        return 8
    return 8 * ((bits + 63) // 64)


def pack_words(lanes: np.ndarray, bits: int) -> np.ndarray:
    """lanes[..., L] integers -> uint8[..., word_bytes(L*bits)]; lane l at bits [l*bits,(l+1)*bits)."""
    lanes = np.asarray(lanes)
    nl = lanes.shape[-1]
    wb = word_bytes(nl * bits)
    lead = lanes.shape[:-1]
    if bits == 8 and wb == nl:  # dense byte lanes: the word image IS the byte tensor
        return np.ascontiguousarray(lanes.astype(np.int64) & 0xFF).astype(np.uint8).reshape(lead + (wb,))
    flat = (lanes.reshape(-1, nl).astype(np.int64)) & ((1 << bits) - 1)
    shifts = np.arange(bits, dtype=np.int64)
    b = ((flat[:, :, None] >> shifts) & 1).astype(np.uint8).reshape(flat.shape[0], nl * bits)
    padded = np.zeros((flat.shape[0], wb * 8), dtype=np.uint8)
    padded[:, : nl * bits] = b
    return np.packbits(padded, axis=1, bitorder="little").reshape(lead + (wb,))


def unpack_words(words: np.ndarray, n_lanes: int, bits: int, signed: bool = False) -> np.ndarray:
    """uint8[..., word_bytes] -> int64[..., n_lanes]."""
    words = np.asarray(words, dtype=np.uint8)
    wb = word_bytes(n_lanes * bits)
    assert words.shape[-1] == wb, (words.shape, wb)
    lead = words.shape[:-1]
    if bits == 8 and wb == n_lanes:
        out = words.astype(np.int64)
    else:
        b = np.unpackbits(words.reshape(-1, wb), axis=1, bitorder="little")[:, : n_lanes * bits]
        b = b.reshape(-1, n_lanes, bits).astype(np.int64)
        out = (b << np.arange(bits, dtype=np.int64)).sum(axis=2).reshape(lead + (n_lanes,))
    if signed:
        out = (out ^ (1 << (bits - 1))) - (1 << (bits - 1))
    return out


def pack_stream(x: np.ndarray, bits: int) -> np.ndarray:
    """x[N, Y, X, C] -> flat uint8 stream image (image-major, y-major, x fastest)."""
    return pack_words(x, bits).reshape(-1)


def unpack_stream(buf: np.ndarray, n: int, y: int, x: int, c: int, bits: int, signed: bool = False) -> np.ndarray:
    wb = word_bytes(c * bits)
    return unpack_words(np.asarray(buf, dtype=np.uint8).reshape(n, y, x, wb), c, bits, signed)


def pack_weights(w: np.ndarray, simd: int, pe: int, w_bits: int) -> np.ndarray:
    """W[OFM, K] (k = (ky*Kx+kx)*C + c) -> image of m_weights[PE][TILES], tile = nf*SF + sf (mvau.hpp:117,148)."""
    ofm, k = w.shape
    assert ofm % pe == 0 and k % simd == 0
    nf, sf = ofm // pe, k // simd
    # [nf, pe, sf, simd] -> [pe, nf, sf, simd]
    t = np.asarray(w).reshape(nf, pe, sf, simd).transpose(1, 0, 2, 3)
    return pack_words(t, w_bits).reshape(-1)


def pack_param_stream(w: np.ndarray, simd: int, pe: int, w_bits: int) -> np.ndarray:
    """W[OFM, K] -> one period (TILES words) of the parameter stream GenParamStream writes (dma.h:214-236): word `tile`
    = ap_uint<SIMD*PE*WP>, PE-row pe at bits [pe*SIMD*WP, (pe+1)*SIMD*WP), SIMD lane l of it at l*WP."""
    ofm, k = w.shape
    assert ofm % pe == 0 and k % simd == 0
    nf, sf = ofm // pe, k // simd
    # [nf, pe, sf, simd] -> [nf, sf, pe*simd]
    t = np.asarray(w).reshape(nf, pe, sf, simd).transpose(0, 2, 1, 3).reshape(nf * sf, pe * simd)
    return pack_words(t, w_bits).reshape(-1)


def unpack_weights(buf: np.ndarray, ofm: int, k: int, simd: int, pe: int, w_bits: int, signed: bool = True) -> np.ndarray:
    nf, sf = ofm // pe, k // simd
    wb = word_bytes(simd * w_bits)
    t = unpack_words(np.asarray(buf, dtype=np.uint8).reshape(pe, nf, sf, wb), simd, w_bits, signed)
    return t.transpose(1, 0, 2, 3).reshape(ofm, k)


def pack_thresholds(t: np.ndarray, pe: int, acc_bits: int) -> np.ndarray:
    """T[OFM, NumTH] -> image of m_thresholds[PE][NF][NumTH] (activations.hpp:172)."""
    ofm, nth = t.shape
    nf = ofm // pe
    a = np.asarray(t).reshape(nf, pe, nth).transpose(1, 0, 2).reshape(-1, 1)
    # one scalar per container; sign-extend into the container like the vendor library stores ap_int
    cb = word_bytes(acc_bits)
    v = a.astype(np.int64).reshape(-1)
    out = np.zeros((v.size, cb), dtype=np.uint8)
    for i in range(min(cb, 8)):
        out[:, i] = (v >> (8 * i)) & 0xFF
    return out.reshape(-1)


def pack_bias(b: np.ndarray) -> np.ndarray:
    """bias[OFM] (s8) -> image of FixedPointWeights<1,ap_int<8>,1,OFM>::m_weights[1][OFM]."""
    return (np.asarray(b).astype(np.int64) & 0xFF).astype(np.uint8)


def stream_to_axi_memory(stream: np.ndarray, word_bits: int, mem_bits: int = 64) -> np.ndarray:
    """Stream image (ap-word containers of `word_bits`-bit words) -> the memory image Mem2Stream_Batch reads / Stream2Mem_Batch writes
    (dma.h:135-199) when a StreamingDataWidthConverter_Batch (streamtools.h:463-526) sits between the `mem_bits`-bit AXI words and the
    layer: both move bits LSB-first, so the image is the dense bit string of the words' low `word_bits` bits (the 16-image bursts of
    the *_Batch blocks do not change the image).  The total number of bits must be a whole number of memory words."""
    wb = word_bytes(word_bits)
    words = np.asarray(stream, dtype=np.uint8).reshape(-1, wb)
    bits = np.unpackbits(words, axis=1, bitorder="little")[:, :word_bits].reshape(-1)
    if bits.size % mem_bits:
        raise ValueError(f"{bits.size} stream bits are not a whole number of {mem_bits}-bit memory words")
    return np.packbits(bits, bitorder="little")


def axi_memory_to_stream(mem: np.ndarray, word_bits: int, n_words: int) -> np.ndarray:
    """Inverse of stream_to_axi_memory: `n_words` stream words back into their ap-word containers."""
    bits = np.unpackbits(np.asarray(mem, dtype=np.uint8), bitorder="little")[: n_words * word_bits].reshape(n_words, word_bits)
    out = np.zeros((n_words, word_bytes(word_bits) * 8), dtype=np.uint8)
    out[:, :word_bits] = bits
    return np.packbits(out, axis=1, bitorder="little").reshape(-1)
