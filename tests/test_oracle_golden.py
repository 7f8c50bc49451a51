"""The plain-C restatement (oracle/finn_oracle.c) against golden vectors produced by the REFERENCE's own
templates (tests/golden/layer_*.npz, made by oracle/gen_golden.py from /root/reference) -- and, where the
reference library is present (this container), against the reference run live."""
import glob
import hashlib
import os

import numpy as np
import pytest

from oracle import cases

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
QUICK = [n for n in cases.CASES if n not in ("c2d_L1",)]


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_golden_files_present():
    have = {os.path.basename(p)[6:-4] for p in glob.glob(os.path.join(GOLD, "layer_*.npz"))}
    assert set(cases.CASES) <= have, sorted(set(cases.CASES) - have)


@pytest.mark.parametrize("name", QUICK)
def test_restatement_matches_golden(name, oracle_mod):
    g = np.load(os.path.join(GOLD, f"layer_{name}.npz"))
    d = cases.CASES[name]
    inp = cases.make_inputs(d)
    assert _sha(inp["in_words"]) == str(g["in_sha"]), "seeded input drifted from the one the golden was made with"
    assert _sha(inp["weights"]) == str(g["w_sha"])
    out = oracle_mod.run_layer(d, inp["in_words"], inp["weights"], inp["thresholds"], inp["bias"])
    assert out.size == int(g["out_bytes"])
    assert np.array_equal(out[:4096], g["head"])
    assert _sha(out) == str(g["out_sha"])
    if "out" in g.files:
        assert np.array_equal(out, g["out"])


@pytest.mark.parametrize("name", ["th_a", "th_b", "th_c"])
def test_param_stream_matches_golden(name, oracle_mod):
    """The weight stream GenParamStream writes (dma.h:214-236), as recorded from the reference (param_stream_*.npz, whose
    out_sha is the output of Matrix_Vector_Activate_Stream_Batch fed by it == the static-weights layer output): the oracle's
    restatement and the host packer must produce the same bytes."""
    from simple_image_compression_network_b200 import pack
    g = np.load(os.path.join(GOLD, f"param_stream_{name}.npz"))
    d = cases.CASES[name]
    inp = cases.make_inputs(d)
    assert _sha(inp["weights"]) == str(g["w_sha"])
    assert str(g["out_sha"]) == str(np.load(os.path.join(GOLD, f"layer_{name}.npz"))["out_sha"])
    assert np.array_equal(oracle_mod.gen_param_stream(d, inp["weights"]), g["param_words"])
    assert np.array_equal(pack.pack_param_stream(inp["w"], d.simd, d.pe, d.w_bits), g["param_words"])
    if oracle_mod.ref_available():
        s = oracle_mod.query(d)
        inp2 = cases.make_inputs(d, seed_shift=5)
        out, pw = oracle_mod.ref_stream_run(name, inp2["in_words"], inp2["weights"], inp2["thresholds"], s.out_bytes_per_image,
                                            g["param_words"].size)
        assert np.array_equal(pw, oracle_mod.gen_param_stream(d, inp2["weights"]))
        assert np.array_equal(out, oracle_mod.run_layer(d, inp2["in_words"], inp2["weights"], inp2["thresholds"], None))


@pytest.mark.parametrize("name", [n for n in cases.CASES if n not in cases.SLOW])
def test_restatement_matches_live_reference(name, oracle_mod):
    if not oracle_mod.ref_available():
        pytest.skip("oracle/_ref/libref_layers.so not built here (needs /root/reference)")
    d = cases.CASES[name]
    inp = cases.make_inputs(d, seed_shift=17)  # a different draw than the golden's
    s = oracle_mod.query(d)
    ref, _ = oracle_mod.ref_run(name, inp["in_words"], inp["weights"], cases.third_image(inp), s.out_bytes_per_image)
    out = oracle_mod.run_layer(d, inp["in_words"], inp["weights"], inp["thresholds"], inp["bias"])
    assert np.array_equal(out, ref)


def test_pool_restatement_vs_reference(oracle_mod):
    """StreamingMaxPool_Precision / StreamingMaxPool (maxpool.h:137-185, :66-96) on the square cases the
    reference supports; the restatement generalises them to non-square extents."""
    if not oracle_mod.ref_available():
        pytest.skip("reference library not built here")
    from simple_image_compression_network_b200 import pack, synth
    for name, dim, pd, ch, bits in (("pool_prec_8_2_8", 8, 2, 8, 8), ("pool_prec_12_2_16", 12, 2, 16, 8), ("pool_prec_12_3_4", 12, 3, 4, 4),
                                    ("pool_bin_8_2_16", 8, 2, 16, 1), ("pool_bin_12_2_64", 12, 2, 64, 1)):
        x = synth.lanes(99, (1, dim, dim, ch), bits)
        words = pack.pack_stream(x, bits)
        nout = pack.word_bytes(ch * bits) * (dim // pd) ** 2
        ref = oracle_mod.ref_pool(name, words, nout)
        mine = oracle_mod.maxpool(words, dim, dim, pd, ch, bits)
        assert np.array_equal(ref, mine), name


def test_fused_pool_equals_pool_after_layer(oracle_mod):
    """desc.pool is the layer followed by the stand-alone pool (what a StreamingMaxPool after ConvLayer does)."""
    import dataclasses
    d = cases.CASES["th_b"]
    dp = dataclasses.replace(d, pool=2)
    inp = cases.make_inputs(d)
    full = oracle_mod.run_layer(d, inp["in_words"], inp["weights"], inp["thresholds"], None)
    fused = oracle_mod.run_layer(dp, inp["in_words"], inp["weights"], inp["thresholds"], None)
    assert np.array_equal(fused, oracle_mod.maxpool(full, d.ofm_x, d.ofm_y, 2, d.ofm_ch, d.out_bits))


def test_closed_form_conv_small(oracle_mod):
    """Independent numpy evaluation of SURVEY.md A.3 on the tiniest case."""
    from simple_image_compression_network_b200 import pack
    d = cases.CASES["c2d_a"]
    inp = cases.make_inputs(d)
    x, w, b = inp["x"][0], inp["w"], inp["b"]
    xp = np.zeros((d.ifm_y + 4, d.ifm_x + 4, d.ifm_ch), np.int64)
    xp[2:-2, 2:-2] = x
    out = np.zeros((d.ofm_y, d.ofm_x, d.ofm_ch), np.int64)
    for oy in range(d.ofm_y):
        for ox in range(d.ofm_x):
            win = xp[2 * oy:2 * oy + 5, 2 * ox:2 * ox + 5].reshape(-1)  # (ky, kx, c)
            t = (w @ win + b) % 256
            out[oy, ox] = np.where(t >= 128, 0, t)
    got = oracle_mod.run_layer(d, inp["in_words"], inp["weights"], None, inp["bias"])
    assert np.array_equal(pack.unpack_stream(got, 1, d.ofm_y, d.ofm_x, d.ofm_ch, 8)[0], out)


def test_oracle_rejects_bad_shapes(oracle_mod):
    import dataclasses
    d = cases.CASES["c2d_a"]
    with pytest.raises(ValueError):
        oracle_mod.query(dataclasses.replace(d, simd=3))  # IFM_CH % SIMD (slidingwindow.h:1259)
    with pytest.raises(ValueError):
        oracle_mod.query(dataclasses.replace(d, pe=4))    # OFM_CH % PE


def _add_cases():
    from simple_image_compression_network_b200 import pack, synth
    return {
        # name: (in1, in2, n_words, channels, in1_bits, in1_signed, in2_bits, in2_signed, out_bits, offset) -- ref_layers.cpp run_add
        "add_u8": (pack.pack_words(synth.lanes(5, (120, 16), 8), 8).reshape(-1), pack.pack_words(synth.lanes(6, (120, 16), 8), 8).reshape(-1),
                   120, 16, 8, 0, 8, 0, 8, 0),
        "add_s8_off": (pack.pack_words(synth.lanes(7, (66, 6), 8), 8).reshape(-1), pack.pack_words(synth.lanes(8, (66, 6), 4), 4).reshape(-1),
                       66, 6, 8, 1, 4, 0, 10, -7),
    }


@pytest.mark.parametrize("name", ["add_u8", "add_s8_off"])
def test_add_streams_restatement_vs_reference_golden(name, oracle_mod):
    """AddStreams_Batch (streamtools.h:669-720): the C restatement against outputs recorded from the reference's own template."""
    a, b, n, ch, b1, s1, b2, s2, ob, off = _add_cases()[name]
    want = np.load(os.path.join(GOLD, f"{name}.npz"))["out"]
    assert np.array_equal(oracle_mod.add_streams(a, b, n, ch, b1, s1, b2, s2, ob, off), want)
