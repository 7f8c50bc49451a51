"""Host packing (simple_image_compression_network_b200/pack.py) against the container rule of SURVEY.md A.1."""
import numpy as np
import pytest

from simple_image_compression_network_b200 import pack, synth


@pytest.mark.parametrize("bits,expect", [(1, 1), (8, 1), (9, 2), (16, 2), (24, 4), (32, 4), (48, 8), (64, 8), (65, 16), (96, 16),
                                         (128, 16), (1024, 128), (1536, 192)])
def test_word_bytes(bits, expect):
    assert pack.word_bytes(bits) == expect


@pytest.mark.parametrize("lanes,bits", [(3, 8), (128, 8), (8, 4), (12, 4), (64, 1), (6, 1), (8, 16), (5, 2), (3, 4)])
def test_roundtrip(lanes, bits):
    x = synth.lanes(7, (11, lanes), bits)
    w = pack.pack_words(x, bits)
    assert w.shape == (11, pack.word_bytes(lanes * bits))
    assert np.array_equal(pack.unpack_words(w, lanes, bits), x)


def test_lane0_is_lsb():
    # lane 0 occupies the low bits (interpret.hpp:211); ap_uint<24> sits in a 4-byte container
    w = pack.pack_words(np.array([[0x11, 0x22, 0x33]]), 8)
    assert w.tolist() == [[0x11, 0x22, 0x33, 0x00]]
    w = pack.pack_words(np.array([[0x1, 0x2, 0x3]]), 4)  # ap_uint<12>: 0x321
    assert w.tolist() == [[0x21, 0x03]]


def test_dense_u8_is_identity():
    x = synth.lanes(3, (2, 4, 6, 128), 8)
    assert np.array_equal(pack.pack_stream(x, 8), x.astype(np.uint8).reshape(-1))


def test_weight_image_layout():
    # m_weights[pe][nf*SF+sf] lane simd == W[nf*PE+pe][sf*SIMD+simd]   (mvau.hpp:117,148)
    ofm, k, simd, pe, wb = 6, 8, 2, 3, 4
    w = synth.weights(5, ofm, k, wb)
    img = pack.pack_weights(w, simd, pe, wb)
    nf, sf = ofm // pe, k // simd
    words = img.reshape(pe, nf * sf, pack.word_bytes(simd * wb))
    for p in range(pe):
        for t in range(nf * sf):
            lanes = pack.unpack_words(words[p, t], simd, wb, signed=True)
            n, s = divmod(t, sf)
            assert np.array_equal(lanes, w[n * pe + p, s * simd:(s + 1) * simd])
    assert np.array_equal(pack.unpack_weights(img, ofm, k, simd, pe, wb), w)


def test_threshold_image_layout():
    t = np.arange(4 * 3).reshape(4, 3) - 5
    img = pack.pack_thresholds(t, pe=2, acc_bits=24).reshape(2, 2, 3, 4)  # [pe][nf][i][4 bytes]
    v = img.astype(np.int64)
    val = v[..., 0] | (v[..., 1] << 8) | (v[..., 2] << 16)
    val = (val ^ (1 << 23)) - (1 << 23)
    for pe in range(2):
        for nf in range(2):
            assert np.array_equal(val[pe, nf], t[nf * 2 + pe])


def test_splitmix_known_answer():
    # splitmix64 reference values (seed 0 stream: first output for state 0 after one increment)
    assert int(synth.splitmix64(np.array([0], dtype=np.uint64))[0]) == 0xE220A8397B1DCDAF


@pytest.mark.parametrize("lanes,bits,mem", [(3, 8, 64), (16, 8, 64), (12, 4, 32), (128, 8, 64), (5, 3, 8)])
def test_axi_memory_image_roundtrip(lanes, bits, mem):
    """Mem2Stream_Batch / Stream2Mem_Batch memory image (dma.h:135-199 + the width converter): dense LSB-first bit string."""
    n = mem * 3  # any count that makes a whole number of memory words
    x = synth.lanes(11, (n, lanes), bits)
    stream = pack.pack_words(x, bits).reshape(-1)
    img = pack.stream_to_axi_memory(stream, lanes * bits, mem)
    assert img.size * 8 == n * lanes * bits
    assert np.array_equal(pack.axi_memory_to_stream(img, lanes * bits, n), stream)
    if bits == 8 and pack.word_bytes(lanes * 8) == lanes:  # byte lanes in dense containers: the memory image IS the stream image
        assert np.array_equal(img, stream)
    if lanes == 3 and bits == 8:  # ap_uint<24> words: the container's pad byte is not part of the memory image
        assert np.array_equal(img.reshape(-1, 3), stream.reshape(-1, 4)[:, :3])
