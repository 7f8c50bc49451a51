"""eight_layers_net (conv_nonsquare_top.cpp:295-357) with the reference's own fixture weights
(memdata_nonsquare.h, dumped to tests/golden/params_nonsquare.npz by oracle/gen_golden.py) on
(a) the testbench's constant-1 image (conv3_nonsquare_tb.cpp:801) and (b) a seeded random image.
Golden outputs were produced by the UNMODIFIED reference top (oracle/_ref/libref_net.so)."""
import hashlib
import os

import numpy as np
import pytest

from simple_image_compression_network_b200 import configs, pack, synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _image(tag):
    if tag == "ones":
        return np.ones((1, 512, 768, 3), np.int64)
    return synth.lanes(synth.SEED_INPUT, (1, 512, 768, 3), 8)


@pytest.fixture(scope="module")
def params():
    return np.load(os.path.join(GOLD, "params_nonsquare.npz"))


@pytest.mark.slow
@pytest.mark.parametrize("tag", ["ones", "rand"])
def test_oracle_chain_matches_reference_net(tag, params, oracle_mod):
    g = np.load(os.path.join(GOLD, f"net8_{tag}.npz"))
    s = pack.pack_stream(_image(tag), 8)
    assert _sha(s) == str(g["in_sha"])
    for i in range(8):
        s = oracle_mod.run_layer(configs.net_layer(i), s, params[f"w{i}"], None, params[f"b{i}"])
    assert np.array_equal(s, g["out"])


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["ones", "rand"])
def test_gpu_net_matches_reference_net(tag, params, fcb_lib):
    from simple_image_compression_network_b200.layer import ConvLayer, Net
    g = np.load(os.path.join(GOLD, f"net8_{tag}.npz"))
    layers = [ConvLayer(configs.net_layer(i), params[f"w{i}"], bias=params[f"b{i}"]) for i in range(8)]
    net = Net(layers)
    s = pack.pack_stream(_image(tag), 8)
    out = net.run(s, 1)
    assert np.array_equal(out, g["out"]), f"engines: {[l.engine for l in layers]}"
    assert net.launches >= 8


@pytest.mark.gpu
def test_gpu_net_batch_of_two(params, fcb_lib):
    """Two different images through the net in one call == each alone (numReps semantics)."""
    from simple_image_compression_network_b200.layer import ConvLayer, Net
    layers = [ConvLayer(configs.net_layer(i), params[f"w{i}"], bias=params[f"b{i}"]) for i in range(8)]
    net = Net(layers)
    a, b = pack.pack_stream(_image("ones"), 8), pack.pack_stream(_image("rand"), 8)
    both = net.run(np.concatenate([a, b]), 2)
    ga = np.load(os.path.join(GOLD, "net8_ones.npz"))["out"]
    gb = np.load(os.path.join(GOLD, "net8_rand.npz"))["out"]
    assert np.array_equal(both, np.concatenate([ga, gb]))


@pytest.mark.gpu
def test_gpu_full_size_layers_random_weights(fcb_lib, oracle_mod):
    """Every layer of config_nonsquare.h at FULL size (768x512 image) with seeded random weights and biases -- the fixture
    weights of memdata_nonsquare.h repeat one row per PE -- chained on the GPU, checked layer by layer against
    the oracle fed with the GPU's own previous output (so a wrong layer is named, not just the end of the chain)."""
    from simple_image_compression_network_b200.layer import ConvLayer
    s = None
    for i in range(8):
        d = configs.net_layer(i)
        prm = configs.synthetic_params(d, seed_shift=50 + i)
        if s is None:
            _, s = configs.synthetic_input(d, 50, 1, False)
        L = ConvLayer(d, prm["weights"], bias=prm["bias"])
        got = L.run(s, 1)
        want = oracle_mod.run_layer(d, s, prm["weights"], None, prm["bias"])
        bad = np.flatnonzero(got != want)
        assert bad.size == 0, f"layer {i} [{L.engine}: {L.plan}]: {bad.size}/{got.size} bytes differ, first at {bad[:8].tolist()}"
        s = got
