"""Parity tests proper: the CUDA library (through the C ABI) against the plain-C oracle on identical packed
bytes -- bit exact (all arithmetic is integer) -- and against the reference-made golden vectors."""
import dataclasses
import hashlib
import os

import numpy as np
import pytest

from oracle import cases
from simple_image_compression_network_b200.desc import ENGINE_IMAD, ENGINE_TENSOR, ENGINE_XNOR_POPC

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _layer(d, inp, device=0):
    from simple_image_compression_network_b200.layer import ConvLayer
    return ConvLayer(d, inp["weights"], thresholds=inp["thresholds"], bias=inp["bias"], device=device)


def _diff(a, b):
    bad = np.flatnonzero(a != b)
    return f"{bad.size}/{a.size} bytes differ; first at {bad[:8].tolist()}: got {a[bad[:8]].tolist()} want {b[bad[:8]].tolist()}"


@pytest.mark.parametrize("name", [n for n in cases.CASES if n != "c2d_L1"])
def test_layer_matches_oracle_and_golden(name, fcb_lib, oracle_mod):
    assert fcb_lib.fcb_device_count() >= 1, "no sm_100 device: the CUDA path cannot be exercised"
    d = cases.CASES[name]
    inp = cases.make_inputs(d)
    L = _layer(d, inp)
    got = L.run(inp["in_words"])
    want = oracle_mod.run_layer(d, inp["in_words"], inp["weights"], inp["thresholds"], inp["bias"])
    assert np.array_equal(got, want), f"{name} [{L.engine}: {L.plan}]: {_diff(got, want)}"
    g = np.load(os.path.join(GOLD, f"layer_{name}.npz"))
    assert _sha(got) == str(g["out_sha"]), f"{name}: differs from the reference-made golden"
    assert L.launches >= 1


@pytest.mark.parametrize("name", ["c2d_b", "c2d_e", "c2d_g", "dc_c", "dc_e", "th_b", "c2d_d"])
def test_engines_agree(name, fcb_lib, oracle_mod):
    """The tensor-core engine and the universal IMAD engine (engine_hint, the reference's resource argument R) give identical bytes."""
    d = cases.CASES[name]
    inp = cases.make_inputs(d, seed_shift=3, num_reps=3)
    want = oracle_mod.run_layer(d, inp["in_words"], inp["weights"], inp["thresholds"], inp["bias"], num_reps=3)
    L = _layer(d, inp)
    got = L.run(inp["in_words"], 3)
    assert np.array_equal(got, want), f"{name} [{L.engine}]: {_diff(got, want)}"
    L2 = _layer(dataclasses.replace(d, engine_hint=ENGINE_IMAD), inp)
    assert L2.engine == "imad"
    got2 = L2.run(inp["in_words"], 3)
    assert np.array_equal(got2, want), f"{name} [imad]: {_diff(got2, want)}"


@pytest.mark.parametrize("name", ["c2d_e", "c2d_g", "dc_c", "th_cfg4", "c2d_L1band", "c2d_d"])
def test_umma_generations_agree(name, fcb_lib, oracle_mod, monkeypatch, exp_build):
    """Resident-patch main loop (fcb_umma2.cu) and per-tap-TMA main loop (fcb_umma.cu) against the oracle."""
    d = cases.CASES[name]
    if not (d.ifm_ch % 128 == 0 and d.ofm_ch % 32 == 0):
        pytest.skip("not a tensor-core shape")
    inp = cases.make_inputs(d, seed_shift=11, num_reps=2)
    want = oracle_mod.run_layer(d, inp["in_words"], inp["weights"], inp["thresholds"], inp["bias"], num_reps=2)
    L = _layer(d, inp)
    got = L.run(inp["in_words"], 2)
    assert np.array_equal(got, want), f"{name} [{L.engine}: {L.plan}]: {_diff(got, want)}"
    exp_build()  # the first-generation kernel exists only in the experiment build
    monkeypatch.setenv("FCB_UMMA_V1", "1")
    L1 = _layer(d, inp)
    assert L1.plan.startswith("v1")
    got1 = L1.run(inp["in_words"], 2)
    assert np.array_equal(got1, want), f"{name} [{L1.engine}: {L1.plan}]: {_diff(got1, want)}"


def test_expected_engines(fcb_lib):
    exp = {"c2d_e": "umma_i8", "c2d_g": "umma_i8", "dc_c": "umma_i8", "th_cfg4": "umma_i8", "c2d_c": "umma_i8", "c2d_a": "umma_i8", "dc_a": "imad", "dc_d": "umma_i8", "dc_e": "umma_i8", "dc_L4": "umma_i8", "c2d_b": "imad",
           "xn_b": "umma_i8", "xn_c": "umma_i8", "xn_a": "imad", "c2d_L1band": "umma_i8"}
    for name, eng in exp.items():
        d = cases.CASES[name]
        inp = cases.make_inputs(d)
        assert _layer(d, inp).engine == eng, name


@pytest.mark.parametrize("name,pool", [("th_b", 2), ("th_cfg4", 2), ("th_a", 2), ("xn_c", 0), ("th_d", 2), ("th_d", 4)])
def test_fused_pool(name, pool, fcb_lib, oracle_mod):
    """Threshold activation with the max pool fused into the epilogue (config 4)."""
    d = dataclasses.replace(cases.CASES[name], pool=pool)
    inp = cases.make_inputs(d, seed_shift=5)
    L = _layer(d, inp)
    got = L.run(inp["in_words"])
    want = oracle_mod.run_layer(d, inp["in_words"], inp["weights"], inp["thresholds"], inp["bias"])
    assert np.array_equal(got, want), f"{name} pool={pool} [{L.engine}]: {_diff(got, want)}"


@pytest.mark.parametrize("cmp", [0, 1, 2, 3])
def test_threshold_compare_variants(cmp, fcb_lib, oracle_mod):
    d = dataclasses.replace(cases.CASES["th_a"], cmp=cmp)
    inp = cases.make_inputs(d, seed_shift=cmp)
    got = _layer(d, inp).run(inp["in_words"])
    want = oracle_mod.run_layer(d, inp["in_words"], inp["weights"], inp["thresholds"], None)
    assert np.array_equal(got, want)


def test_unsorted_thresholds(fcb_lib, oracle_mod):
    """Threshold order does not matter to the reference (it counts); the library sorts a private copy."""
    from simple_image_compression_network_b200 import pack
    d = cases.CASES["th_b"]
    inp = cases.make_inputs(d)
    rng = np.random.default_rng(1)
    t = inp["t"].copy()
    for row in t:
        rng.shuffle(row)
    timg = pack.pack_thresholds(t, d.pe, d.acc_bits)
    from simple_image_compression_network_b200.layer import ConvLayer
    got = ConvLayer(d, inp["weights"], thresholds=timg).run(inp["in_words"])
    want = oracle_mod.run_layer(d, inp["in_words"], inp["weights"], timg, None)
    assert np.array_equal(got, want)
    assert np.array_equal(got, oracle_mod.run_layer(d, inp["in_words"], inp["weights"], inp["thresholds"], None))


def test_passthrough_and_signed_inputs(fcb_lib, oracle_mod):
    from simple_image_compression_network_b200.desc import ACT_PASSTHROUGH
    d = dataclasses.replace(cases.CASES["c2d_f"], act_kind=ACT_PASSTHROUGH, acc_bits=16, acc_signed=1, out_bits=16, in_signed=1)
    inp = cases.make_inputs(d)
    got = _layer(d, inp).run(inp["in_words"])
    want = oracle_mod.run_layer(d, inp["in_words"], inp["weights"], None, None)
    assert np.array_equal(got, want)


def test_pm1_binary_weights(fcb_lib, oracle_mod):
    from simple_image_compression_network_b200.desc import ACT_PASSTHROUGH, W_BINARY_PM1
    d = dataclasses.replace(cases.CASES["c2d_f"], weight_kind=W_BINARY_PM1, w_bits=1, act_kind=ACT_PASSTHROUGH, acc_bits=16,
                            acc_signed=1, out_bits=16)
    inp = cases.make_inputs(d)
    got = _layer(d, inp).run(inp["in_words"])
    want = oracle_mod.run_layer(d, inp["in_words"], inp["weights"], None, None)
    assert np.array_equal(got, want)


def test_ragged_and_edge_shapes(fcb_lib, oracle_mod):
    """Extents that do not fill the 128-pixel tiles / 16x8 CTAs, 1x1 kernels, single-row images."""
    base = cases.CASES["c2d_e"]
    for ix, iy in ((50, 34), (16, 2), (130, 6)):
        d = dataclasses.replace(base, ifm_x=ix, ifm_y=iy)
        inp = cases.make_inputs(d, seed_shift=ix)
        L = _layer(d, inp)
        got = L.run(inp["in_words"])
        want = oracle_mod.run_layer(d, inp["in_words"], inp["weights"], None, inp["bias"])
        assert np.array_equal(got, want), f"{ix}x{iy} [{L.engine}]: {_diff(got, want)}"
    d = dataclasses.replace(cases.CASES["c2d_d"], kernel_x=1, kernel_y=1, pad=0, simd=16, ifm_x=7, ifm_y=1)
    inp = cases.make_inputs(d)
    assert np.array_equal(_layer(d, inp).run(inp["in_words"]), oracle_mod.run_layer(d, inp["in_words"], inp["weights"], None, inp["bias"]))


def test_zero_reps_and_bad_sizes(fcb_lib):
    d = cases.CASES["c2d_a"]
    inp = cases.make_inputs(d)
    L = _layer(d, inp)
    assert L.run(np.zeros(0, np.uint8), 0).size == 0
    with pytest.raises(ValueError):
        L.run(inp["in_words"][:-1])


def test_batch_images_are_independent(fcb_lib, oracle_mod):
    """numReps images = numReps single-image applications; checks image boundaries inside tiles."""
    d = cases.CASES["c2d_g"]
    inp = cases.make_inputs(d, num_reps=5)
    L = _layer(d, inp)
    got = L.run(inp["in_words"], 5)
    per = L.in_bytes
    for n in (0, 4):
        one = L.run(inp["in_words"][n * per:(n + 1) * per], 1)
        assert np.array_equal(one, got[n * L.out_bytes:(n + 1) * L.out_bytes])
    assert np.array_equal(got, oracle_mod.run_layer(d, inp["in_words"], inp["weights"], None, inp["bias"], num_reps=5))


def test_conv1_full_size_properties(fcb_lib, oracle_mod):
    """CONV_1 (config_nonsquare.h:18-33) at full size: golden digest of image 0 + per-image independence +
    linearity-style property (zero input -> relu(bias))."""
    d = cases.CASES["c2d_L1"]
    inp = cases.make_inputs(d)
    L = _layer(d, inp)
    assert L.engine == "umma_i8"
    got = L.run(inp["in_words"])
    g = np.load(os.path.join(GOLD, "layer_c2d_L1.npz"))
    assert np.array_equal(got[:4096], g["head"])
    assert _sha(got) == str(g["out_sha"])
    zero = L.run(np.zeros(L.in_bytes, np.uint8))
    b = inp["b"] % 256
    exp = np.where(b >= 128, 0, b).astype(np.uint8)
    assert np.array_equal(zero.reshape(-1, 128), np.broadcast_to(exp, (d.ofm_x * d.ofm_y, 128)))


def test_net_chain_matches_layerwise_oracle(fcb_lib, oracle_mod):
    """Net (eight_layers_net-style chain, intermediates on device) == oracle applied layer by layer."""
    from simple_image_compression_network_b200.layer import Net
    d1 = cases.CASES["c2d_e"]                                   # 128ch 48x32 -> 24x16
    d2 = dataclasses.replace(cases.CASES["dc_c"], ifm_x=24, ifm_y=16)  # deconv back to 48x32
    i1, i2 = cases.make_inputs(d1, num_reps=2, relu_range=True), cases.make_inputs(d2, seed_shift=9)
    L1, L2 = _layer(d1, i1), _layer(d2, i2)
    net = Net([L1, L2])
    got = net.run(i1["in_words"], 2)
    mid = oracle_mod.run_layer(d1, i1["in_words"], i1["weights"], None, i1["bias"], num_reps=2)
    want = oracle_mod.run_layer(d2, mid, i2["weights"], None, i2["bias"], num_reps=2)
    assert np.array_equal(got, want)
    assert net.launches >= 2


def test_device_synth_matches_host(fcb_lib):
    import torch
    from simple_image_compression_network_b200 import synth
    from simple_image_compression_network_b200.layer import synth_fill
    n = 1 << 16
    buf = torch.empty(n + 5, dtype=torch.uint8, device="cuda")
    synth_fill(buf.data_ptr(), n + 5, synth.SEED_INPUT, 0x7F, offset=123)
    torch.cuda.synchronize()
    want = synth.lanes(synth.SEED_INPUT, (n + 5,), 8, mask=0x7F, offset=123).astype(np.uint8)
    assert np.array_equal(buf.cpu().numpy(), want)


def test_run_device_pointers(fcb_lib, oracle_mod):
    """fcb_layer_run_device on torch-owned device memory (the bench's device-resident path)."""
    import torch
    d = cases.CASES["c2d_e"]
    inp = cases.make_inputs(d, num_reps=2)
    L = _layer(d, inp)
    x = torch.from_numpy(inp["in_words"]).cuda()
    y = torch.empty(L.out_bytes * 2, dtype=torch.uint8, device="cuda")
    L.run_device(x.data_ptr(), y.data_ptr(), 2, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    want = oracle_mod.run_layer(d, inp["in_words"], inp["weights"], None, inp["bias"], num_reps=2)
    assert np.array_equal(y.cpu().numpy(), want)


def test_analysis_stack_stage_shapes(fcb_lib, oracle_mod):
    """Config 5b stage shapes (K3 S1 P1 + 255 thresholds + 2x2 pool) at reduced spatial size: thin first stage (C = 3, im2col
    lowering) and a 128 -> 192 stage, chained through Net, against the oracle applied stage by stage."""
    from simple_image_compression_network_b200.desc import ACT_THRESHOLDS, KIND_CONV, LayerDesc
    from simple_image_compression_network_b200.layer import Net

    def stage(c, ofm, x, y, simd, pe):
        return LayerDesc(kind=KIND_CONV, kernel_x=3, kernel_y=3, ifm_ch=c, ofm_ch=ofm, ifm_x=x, ifm_y=y, stride_x=1, stride_y=1, pad=1,
                         simd=simd, pe=pe, in_bits=8, w_bits=4, acc_bits=24, acc_signed=1, act_kind=ACT_THRESHOLDS, out_bits=8,
                         num_th=255, pool=2)
    d1, d2 = stage(3, 128, 48, 32, 3, 16), stage(128, 192, 24, 16, 32, 24)
    i1, i2 = cases.make_inputs(d1, num_reps=2), cases.make_inputs(d2, seed_shift=4)
    L1, L2 = _layer(d1, i1), _layer(d2, i2)
    got = Net([L1, L2]).run(i1["in_words"], 2)
    mid = oracle_mod.run_layer(d1, i1["in_words"], i1["weights"], i1["thresholds"], None, num_reps=2)
    want = oracle_mod.run_layer(d2, mid, i2["weights"], i2["thresholds"], None, num_reps=2)
    assert np.array_equal(got, want), f"[{L1.engine}: {L1.plan}] [{L2.engine}: {L2.plan}]"


@pytest.mark.parametrize("name,pad,pool", [("xn_a", 0, 0), ("xn_b", 0, 0), ("xn_c", 0, 0), ("xn_c", 1, 0), ("xn_b", 1, 2)])
def test_xnor_on_tensor_cores(name, pad, pool, fcb_lib, oracle_mod):
    """The +-1 int8 form of the xnor layer (sum [w==a] = (K + sum a^w^)/2, thresholds remapped to 2t-K; engine_hint = TENSOR, and what
    AUTO picks for single-threshold 1-bit layers of >= 32 channels) against the oracle and against the popcount engine north_star names
    (engine_hint = XNOR_POPC); pad > 0 checks that border bits act as ordinary 0 activations."""
    d = dataclasses.replace(cases.CASES[name], pad=pad, pool=pool)
    inp = cases.make_inputs(d, seed_shift=21, num_reps=2)
    want = oracle_mod.run_layer(d, inp["in_words"], inp["weights"], inp["thresholds"], None, num_reps=2)
    Lp = _layer(dataclasses.replace(d, engine_hint=ENGINE_XNOR_POPC if d.ifm_ch % 32 == 0 else ENGINE_IMAD), inp)
    assert Lp.engine in ("xnor_popc", "imad")
    assert np.array_equal(Lp.run(inp["in_words"], 2), want)
    La = _layer(d, inp)
    assert La.engine == ("umma_i8" if d.ifm_ch >= 32 else "imad"), La.plan
    assert np.array_equal(La.run(inp["in_words"], 2), want)
    Lt = _layer(dataclasses.replace(d, engine_hint=ENGINE_TENSOR), inp)
    assert Lt.engine == "umma_i8" and "xnor as +-1" in Lt.plan, Lt.plan
    got = Lt.run(inp["in_words"], 2)
    assert np.array_equal(got, want), f"{name} pad={pad} [{Lt.plan}]: {_diff(got, want)}"


@pytest.mark.parametrize("name,pad", [("xn_b", 0), ("xn_c", 1)])
def test_xnor_pixel_pair_rows(name, pad, fcb_lib, oracle_mod, monkeypatch, exp_build):
    """Experiment build: the +-1 int8 form with two pixels per 128-byte K-block (FCB_XNOR_PAIR=1: 2*Cp channels, ceil(KX/2) taps, dilation 2).
    Measured slower than one pixel per row and left off in the product; kept bit-exact."""
    d = dataclasses.replace(cases.CASES[name], pad=pad)
    inp = cases.make_inputs(d, seed_shift=23, num_reps=3)
    want = oracle_mod.run_layer(d, inp["in_words"], inp["weights"], inp["thresholds"], None, num_reps=3)
    exp_build()
    monkeypatch.setenv("FCB_XNOR_PAIR", "1")
    L = _layer(d, inp)
    assert "pixel pairs" in L.plan, L.plan
    got = L.run(inp["in_words"], 3)
    assert np.array_equal(got, want), f"{name} [{L.plan}]: {_diff(got, want)}"


def _thin_cases():
    from simple_image_compression_network_b200.desc import ACT_BIAS_RELU, ACT_THRESHOLDS, KIND_CONV, LayerDesc

    def mk(k, s, p, c, ofm, x, y, pe, thr=0, pool=0):
        kw = dict(act_kind=ACT_THRESHOLDS, acc_bits=24, acc_signed=1, out_bits=8, num_th=thr, pool=pool) if thr else \
            dict(act_kind=ACT_BIAS_RELU, acc_bits=8, acc_signed=0, out_bits=8)
        return LayerDesc(kind=KIND_CONV, kernel_x=k, kernel_y=k, ifm_ch=c, ofm_ch=ofm, ifm_x=x, ifm_y=y, stride_x=s, stride_y=s, pad=p,
                         simd=c, pe=pe, in_bits=8, w_bits=4, **kw)
    return {
        "L0_shape_multi_tile": (mk(5, 2, 2, 3, 128, 264, 40, 16), 2),       # staged TMA-store epilogue, ragged right edge
        "L0_shape_one_row": (mk(5, 2, 2, 3, 128, 16, 2, 16), 1),
        "stage1_thr255_pool": (mk(3, 1, 1, 3, 128, 72, 20, 16, thr=255, pool=2), 3),
        "c4_thr15_nopool": (mk(3, 1, 1, 4, 32, 36, 10, 8, thr=15), 2),
        "k5_s1_cb2": (mk(5, 1, 2, 3, 256, 64, 12, 32), 2),                 # two channel blocks
        "k3_s2_ofm192": (mk(3, 2, 1, 3, 192, 40, 12, 24), 1),               # register epilogue (OFM % 128 != 0)
    }


@pytest.mark.parametrize("name", list(_thin_cases()))
def test_thin_input_smem_im2col(name, fcb_lib, oracle_mod, monkeypatch, exp_build):
    """Thin-input layers (one 4-byte word per pixel): sliding window built in shared memory inside the tensor-core kernel,
    against the oracle and against the older two-kernel im2col lowering."""
    d, reps = _thin_cases()[name]
    inp = cases.make_inputs(d, seed_shift=31, num_reps=reps)
    want = oracle_mod.run_layer(d, inp["in_words"], inp["weights"], inp["thresholds"], inp["bias"], num_reps=reps)
    L = _layer(d, inp)
    assert L.engine == "umma_i8" and L.plan.startswith("smem-im2col"), L.plan
    got = L.run(inp["in_words"], reps)
    assert np.array_equal(got, want), f"{name} [{L.plan}]: {_diff(got, want)}"
    exp_build()
    monkeypatch.setenv("FCB_THIN", "im2col")
    L2 = _layer(d, inp)
    assert not L2.plan.startswith("smem-im2col")
    assert np.array_equal(L2.run(inp["in_words"], reps), want), f"{name} [{L2.plan}]"


@pytest.mark.parametrize("ix,iy,c,ofm,reps", [(24, 16, 128, 3, 1), (50, 7, 128, 4, 2), (100, 20, 256, 3, 2), (384, 8, 128, 3, 1)])
def test_thin_output_deconv(ix, iy, c, ofm, reps, fcb_lib, oracle_mod, monkeypatch, exp_build):
    """deconv522 with OFM <= 4 (the 3-channel last layer of eight_layers_net): pixels on the MMA M axis, taps regrouped by input
    shift, against the oracle and against the generic resident-planes plan."""
    d = dataclasses.replace(cases.CASES["dc_d"], ifm_x=ix, ifm_y=iy, ifm_ch=c, ofm_ch=ofm, pe=ofm)
    inp = cases.make_inputs(d, seed_shift=41, num_reps=reps, relu_range=True)
    want = oracle_mod.run_layer(d, inp["in_words"], inp["weights"], None, inp["bias"], num_reps=reps)
    L = _layer(d, inp)
    assert L.plan.startswith("thin-output deconv"), L.plan
    got = L.run(inp["in_words"], reps)
    assert np.array_equal(got, want), f"[{L.plan}]: {_diff(got, want)}"
    assert "col2im" in L.plan, L.plan
    exp_build()
    monkeypatch.setenv("FCB_U2_NO_DCOL", "1")  # the shift-block form (pixels on the MMA M axis)
    L1 = _layer(d, inp)
    assert L1.plan.startswith("thin-output deconv: pixels on M"), L1.plan
    assert np.array_equal(L1.run(inp["in_words"], reps), want), f"[{L1.plan}]"
    monkeypatch.setenv("FCB_U2_NO_DTHIN", "1")
    L2 = _layer(d, inp)
    assert not L2.plan.startswith("thin-output deconv")
    assert np.array_equal(L2.run(inp["in_words"], reps), want), f"[{L2.plan}]"


def _family_cases():
    """Shapes that select the single-family instantiations (16 / 12 epilogue warps), ragged on purpose; value = the switch that
    sends the same plan to the general kernel instead."""
    from simple_image_compression_network_b200.desc import ACT_BIAS_RELU, ACT_THRESHOLDS, KIND_CONV, LayerDesc

    def conv(k, s, p, c, ofm, x, y, simd, pe, thr=0, pool=0):
        kw = dict(act_kind=ACT_THRESHOLDS, acc_bits=24, acc_signed=1, out_bits=8, num_th=thr, pool=pool) if thr else \
            dict(act_kind=ACT_BIAS_RELU, acc_bits=8, acc_signed=0, out_bits=8)
        return LayerDesc(kind=KIND_CONV, kernel_x=k, kernel_y=k, ifm_ch=c, ofm_ch=ofm, ifm_x=x, ifm_y=y, stride_x=s, stride_y=s, pad=p,
                         simd=simd, pe=pe, in_bits=8, w_bits=4, **kw)
    dc = cases.CASES["dc_c"]
    return {
        "pixel_major_16w": (conv(5, 2, 2, 3, 128, 300, 70, 3, 16), "FCB_U2_NO_SWPX", 2),                    # L0 family, ragged both ways
        "pixel_major_16w_k3s1": (conv(3, 1, 1, 3, 128, 150, 11, 3, 16), "FCB_U2_NO_SWPX", 1),
        "col2im_16w": (dataclasses.replace(cases.CASES["dc_d"], ifm_x=77, ifm_y=37), "FCB_U2_NO_DCX", 2),          # L7 family
        "col2im_16w_c256_ofm4": (dataclasses.replace(cases.CASES["dc_d"], ifm_x=29, ifm_y=50, ifm_ch=256, ofm_ch=4, pe=4), "FCB_U2_NO_DCX", 1),
        "staged_deconv_16w": (dataclasses.replace(dc, ifm_x=70, ifm_y=23), "FCB_U2_NO_STX", 2),                    # L5 / L6 family
        "thr_pool_stage1": (conv(3, 1, 1, 3, 128, 300, 22, 3, 16, thr=255, pool=2), "FCB_U2_XEPI", 1),              # 16 epilogue warps
        "thr_pool_stage2": (conv(3, 1, 1, 128, 128, 100, 18, 32, 16, thr=255, pool=2), "FCB_U2_XEPI", 1),           # 12 epilogue warps
    }


@pytest.mark.parametrize("name", list(_family_cases()))
def test_single_family_instantiations(name, fcb_lib, oracle_mod, monkeypatch, exp_build):
    """The single-epilogue-family kernels with 12 / 16 epilogue warps (EPI = 2..6 of umma2_conv_kernel) against the oracle and
    against the general kernel (EPI = 0 / 8 epilogue warps) on the same plan."""
    d, switch, reps = _family_cases()[name]
    inp = cases.make_inputs(d, seed_shift=53, num_reps=reps, relu_range=d.kind != 0)
    want = oracle_mod.run_layer(d, inp["in_words"], inp["weights"], inp["thresholds"], inp["bias"], num_reps=reps)
    L = _layer(d, inp)
    assert L.engine == "umma_i8", L.plan
    got = L.run(inp["in_words"], reps)
    assert np.array_equal(got, want), f"{name} [{L.plan}]: {_diff(got, want)}"
    exp_build()
    monkeypatch.setenv(switch, "0" if switch == "FCB_U2_XEPI" else "1")
    L0 = _layer(d, inp)
    got0 = L0.run(inp["in_words"], reps)
    assert np.array_equal(got0, want), f"{name} general kernel [{L0.plan}]: {_diff(got0, want)}"
    if switch == "FCB_U2_XEPI":  # and every epilogue-warp count of the pooled threshold family
        for n in ("4", "8"):
            monkeypatch.setenv(switch, n)
            Ln = _layer(d, inp)
            assert np.array_equal(Ln.run(inp["in_words"], reps), want), f"{name} [{Ln.plan}]"


def _random_descs(seed, count):
    """Random small layer shapes across every plan family (resident planes, thin input, thin-output deconv, thresholds with and
    without pool, xnor, IMAD fall-backs), seeded: the same cases on every run."""
    from simple_image_compression_network_b200.desc import (ACT_BIAS_RELU, ACT_PASSTHROUGH, ACT_THRESHOLDS, KIND_CONV, KIND_DECONV522, W_BINARY_XNOR,
                                          LayerDesc)
    rng = np.random.default_rng(seed)
    out = []
    while len(out) < count:
        fam = rng.integers(0, 6)
        if fam == 0:  # channel-heavy conv, bias+ReLU
            c = int(rng.choice([16, 32, 64, 128, 256])); ofm = int(rng.choice([32, 64, 128, 192, 256]))
            k = int(rng.choice([1, 3, 5])); s = int(rng.choice([1, 2])) if c % 128 == 0 else 1
            pad = int(rng.integers(0, k // 2 + 1))
            x = int(rng.integers(k, 40)) * s; y = int(rng.integers(k, 12)) * s
            d = LayerDesc(kind=KIND_CONV, kernel_x=k, kernel_y=k, ifm_ch=c, ofm_ch=ofm, ifm_x=x, ifm_y=y, stride_x=s, stride_y=s, pad=pad,
                          simd=16, pe=16, in_bits=8, w_bits=int(rng.choice([2, 4, 8])), acc_bits=8, acc_signed=0, act_kind=ACT_BIAS_RELU, out_bits=8)
        elif fam == 1:  # thin input
            c = int(rng.choice([3, 4])); ofm = int(rng.choice([16, 64, 128, 256])); k = int(rng.choice([3, 5])); s = int(rng.choice([1, 2]))
            x = 4 * int(rng.integers(3, 40)); y = int(rng.integers(k, 14)) * 2
            thr = bool(rng.integers(0, 2))
            kw = dict(act_kind=ACT_THRESHOLDS, acc_bits=24, acc_signed=1, out_bits=8, num_th=int(rng.choice([15, 255])),
                      pool=int(rng.choice([0, 2])) if s == 1 else 0) if thr else dict(act_kind=ACT_BIAS_RELU, acc_bits=8, acc_signed=0, out_bits=8)
            d = LayerDesc(kind=KIND_CONV, kernel_x=k, kernel_y=k, ifm_ch=c, ofm_ch=ofm, ifm_x=x, ifm_y=y, stride_x=s, stride_y=s, pad=k // 2,
                          simd=c, pe=16, in_bits=8, w_bits=4, **kw)
        elif fam == 2:  # deconv522, thin and wide outputs
            c = int(rng.choice([16, 128, 256])); ofm = int(rng.choice([3, 4, 32, 128]))
            d = LayerDesc(kind=KIND_DECONV522, kernel_x=5, kernel_y=5, ifm_ch=c, ofm_ch=ofm, ifm_x=int(rng.integers(2, 60)), ifm_y=int(rng.integers(2, 12)),
                          stride_x=2, stride_y=2, pad=2, simd=16, pe=ofm if ofm < 16 else 16, in_bits=8, w_bits=4, acc_bits=8, acc_signed=0,
                          act_kind=ACT_BIAS_RELU, out_bits=8)
        elif fam == 3:  # thresholds on channel-heavy layers
            c = int(rng.choice([32, 128, 256])); ofm = int(rng.choice([32, 128, 256])); pool = int(rng.choice([0, 2]))
            x = 2 * int(rng.integers(2, 30)); y = 2 * int(rng.integers(2, 8))
            d = LayerDesc(kind=KIND_CONV, kernel_x=3, kernel_y=3, ifm_ch=c, ofm_ch=ofm, ifm_x=x, ifm_y=y, stride_x=1, stride_y=1, pad=1, simd=32, pe=32,
                          in_bits=8, w_bits=4, acc_bits=int(rng.choice([16, 24, 32])), acc_signed=1, act_kind=ACT_THRESHOLDS,
                          out_bits=int(rng.choice([2, 4, 8])), num_th=0, pool=pool, cmp=int(rng.integers(0, 4)))
            d = dataclasses.replace(d, num_th=(1 << d.out_bits) - 1)
        elif fam == 4:  # xnor
            c = int(rng.choice([32, 64, 128])); ofm = int(rng.choice([16, 64]))
            d = LayerDesc(kind=KIND_CONV, kernel_x=3, kernel_y=3, ifm_ch=c, ofm_ch=ofm, ifm_x=int(rng.integers(3, 40)), ifm_y=int(rng.integers(3, 12)),
                          stride_x=1, stride_y=1, pad=int(rng.integers(0, 2)), simd=32, pe=16, in_bits=1, w_bits=1, weight_kind=W_BINARY_XNOR,
                          acc_bits=16, acc_signed=1, act_kind=ACT_THRESHOLDS, out_bits=1, num_th=1,
                          engine_hint=ENGINE_XNOR_POPC if len(out) % 2 else 0)  # AUTO = the +-1 int8 tensor form; every other case: the popcount engine
        else:  # odd widths on the universal engine
            d = LayerDesc(kind=KIND_CONV, kernel_x=3, kernel_y=int(rng.choice([1, 3])), ifm_ch=int(rng.choice([2, 6, 10])), ofm_ch=int(rng.choice([3, 5, 12])),
                          ifm_x=int(rng.integers(3, 20)), ifm_y=int(rng.integers(3, 9)), stride_x=1, stride_y=1, pad=0, simd=2, pe=1,
                          in_bits=int(rng.choice([2, 4, 8])), w_bits=int(rng.choice([3, 5, 8])), acc_bits=16, acc_signed=1, act_kind=ACT_PASSTHROUGH,
                          out_bits=16, in_signed=int(rng.integers(0, 2)))
        out.append(d)
    return out


@pytest.mark.parametrize("seed", [101, 202, 303])
def test_random_shapes_match_oracle(seed, fcb_lib, oracle_mod):
    """Seeded fuzz over the plan families: every layer the library accepts must equal the oracle bit for bit; shapes it
    rejects must be rejected with FCB_ERR_SHAPE / UNSUPPORTED (never a wrong answer)."""
    from simple_image_compression_network_b200._lib import FcbError
    ran, plans = 0, set()
    for i, d in enumerate(_random_descs(seed, 40)):
        reps = 1 + (i % 3)
        try:
            inp = cases.make_inputs(d, seed_shift=seed + i, num_reps=reps, relu_range=bool(i % 2))
            L = _layer(d, inp)
        except (FcbError, ValueError, AssertionError):
            continue
        got = L.run(inp["in_words"], reps)
        want = oracle_mod.run_layer(d, inp["in_words"], inp["weights"], inp["thresholds"], inp["bias"], num_reps=reps)
        assert np.array_equal(got, want), f"seed {seed} case {i} {d} [{L.engine}: {L.plan}]: {_diff(got, want)}"
        ran += 1
        plans.add(L.plan.split(" ")[0] + ":" + L.engine)
    assert ran >= 25, f"only {ran} of 40 random layers were accepted"
    assert len(plans) >= 4, plans


def _direct_descs(seed, count):
    """Random layers forced onto the universal engine (engine_hint): lane widths on both sides of the IDP.4A / IDP.2A / IMAD split,
    signed and unsigned lanes, strides, dilation, ragged channel counts, several channel chunks, deconv522."""
    from simple_image_compression_network_b200.desc import ACT_PASSTHROUGH, KIND_CONV, KIND_DECONV522, LayerDesc
    rng = np.random.default_rng(seed)
    out = []
    for i in range(count):
        inb, wb = [(8, 8), (4, 3), (16, 8), (12, 5), (16, 16), (8, 12)][i % 6]
        ins = int(rng.integers(0, 2))
        c = int(rng.choice([1, 3, 7, 16, 20, 33, 64, 130])); ofm = int(rng.choice([1, 5, 32, 64, 70]))
        if i % 4 == 3:
            d = LayerDesc(kind=KIND_DECONV522, kernel_x=5, kernel_y=5, ifm_ch=c, ofm_ch=ofm, ifm_x=int(rng.integers(2, 14)), ifm_y=int(rng.integers(2, 7)),
                          stride_x=2, stride_y=2, pad=2, simd=1, pe=1, in_bits=inb, in_signed=ins, w_bits=wb, acc_bits=32, acc_signed=1,
                          act_kind=ACT_PASSTHROUGH, out_bits=32, engine_hint=ENGINE_IMAD)
        else:
            kx, ky = int(rng.choice([1, 3, 5])), int(rng.choice([1, 3]))
            sx, sy = int(rng.choice([1, 2, 3])), int(rng.choice([1, 2]))
            dx, dy = (int(rng.choice([1, 2])), int(rng.choice([1, 2]))) if (sx, sy) == (1, 1) else (1, 1)
            x = (kx - 1) * dx + 1 + int(rng.integers(0, 24)); y = (ky - 1) * dy + 1 + int(rng.integers(0, 10))
            x += (-(x - ((kx - 1) * dx + 1))) % sx; y += (-(y - ((ky - 1) * dy + 1))) % sy
            d = LayerDesc(kind=KIND_CONV, kernel_x=kx, kernel_y=ky, ifm_ch=c, ofm_ch=ofm, ifm_x=x, ifm_y=y, stride_x=sx, stride_y=sy, pad=0,
                          simd=1, pe=1, in_bits=inb, in_signed=ins, w_bits=wb, acc_bits=32, acc_signed=1, act_kind=ACT_PASSTHROUGH, out_bits=32,
                          dilation_x=dx, dilation_y=dy, engine_hint=ENGINE_IMAD)
        out.append(d)
    return out


@pytest.mark.parametrize("style", [1, 2])
def test_asymmetric_padding_on_the_tensor_plan(style, fcb_lib, oracle_mod):
    """FMPadding_nonsquare with odd Padding_x / Padding_y (streamtools.h:361-406: the extra zero goes left / up for PaddingStyle 2, right /
    down for 1) on channel-heavy layers: the resident-planes tensor plan moves its tap offsets by the left / up share and leaves the
    rest to the TMA out-of-bounds fill -- stride 1 and 2, bias+ReLU and pooled thresholds, against the oracle and the universal engine."""
    from simple_image_compression_network_b200.desc import ACT_BIAS_RELU, ACT_THRESHOLDS, KIND_CONV, LayerDesc
    rng = np.random.default_rng(40 + style)
    on_tensor = 0
    for i in range(10):
        s = 1 + i % 2
        c = 128 if s == 2 else int(rng.choice([32, 128, 256])); ofm = int(rng.choice([32, 128, 192]))
        k = int(rng.choice([2, 3, 4, 5]))
        px, py = int(rng.integers(0, k)), int(rng.integers(0, k))
        x = 2 * int(rng.integers(k, 24)); y = 2 * int(rng.integers(k, 9))
        if (x + px - k) % s or (y + py - k) % s:
            px += (x + px - k) % s; py += (y + py - k) % s
        thr = i % 3 == 2 and s == 1
        kw = dict(act_kind=ACT_THRESHOLDS, acc_bits=24, acc_signed=1, out_bits=8, num_th=255, pool=2 if ((x + px - k + 1) % 2 == 0 and (y + py - k + 1) % 2 == 0) else 0) \
            if thr else dict(act_kind=ACT_BIAS_RELU, acc_bits=8, acc_signed=0, out_bits=8)
        d = LayerDesc(kind=KIND_CONV, kernel_x=k, kernel_y=k, ifm_ch=c, ofm_ch=ofm, ifm_x=x, ifm_y=y, stride_x=s, stride_y=s, pad=0, simd=16, pe=16,
                      in_bits=8, w_bits=4, pad_style=style, pad_x_total=px, pad_y_total=py, **kw)
        inp = cases.make_inputs(d, seed_shift=style * 50 + i, num_reps=2, relu_range=True)
        L = _layer(d, inp)
        on_tensor += L.engine == "umma_i8"  # (a shape without a tensor plan, e.g. an even kernel at stride 2, takes the universal engine)
        got = L.run(inp["in_words"], 2)
        want = oracle_mod.run_layer(d, inp["in_words"], inp["weights"], inp["thresholds"], inp["bias"], num_reps=2)
        assert np.array_equal(got, want), f"case {i} {d} [{L.plan}]: {_diff(got, want)}"
        L2 = _layer(dataclasses.replace(d, engine_hint=ENGINE_IMAD), inp)
        assert np.array_equal(L2.run(inp["in_words"], 2), want), f"case {i} {d} [imad]"
    assert on_tensor >= 6, on_tensor


@pytest.mark.parametrize("seed", [7, 8])
def test_direct_engine_kernels_match_oracle(seed, fcb_lib, oracle_mod):
    """The three inner loops of the universal engine (fcb_direct.cu): IDP.4A (lanes <= 8 bits, weights <= 8 bits), IDP.2A (lanes of
    9..16 bits) and plain IMAD (weights of 9..16 bits), signed and unsigned lanes, conv (strides 1 / 2 / generic, dilation, ragged
    channel counts that leave the last packed word / group partly empty, several channel chunks) and deconv522 -- forced with
    engine_hint, bit-exact against the oracle."""
    seen = set()
    for i, d in enumerate(_direct_descs(seed, 36)):
        inp = cases.make_inputs(d, seed_shift=seed * 100 + i, num_reps=2)
        L = _layer(d, inp)
        assert L.engine == "imad"
        kern = "IDP.4A" if "IDP.4A" in L.plan else "IDP.2A" if "IDP.2A" in L.plan else "IMAD"
        w8 = int(np.abs(inp["w"].astype(np.int64) * 2 + 1).max()) <= 255  # every weight in [-128, 127]
        assert kern == ("IMAD" if not w8 else "IDP.4A" if d.in_bits <= 8 else "IDP.2A"), (d, L.plan)
        seen.add((kern, d.in_signed, d.kind))
        got = L.run(inp["in_words"], 2)
        want = oracle_mod.run_layer(d, inp["in_words"], inp["weights"], inp["thresholds"], inp["bias"], num_reps=2)
        assert np.array_equal(got, want), f"seed {seed} case {i} {d} [{L.plan}]: {_diff(got, want)}"
    assert {k for k, _, _ in seen} == {"IDP.4A", "IDP.2A", "IMAD"}


def test_net_host_pipeline_many_chunks(fcb_lib, oracle_mod):
    """fcb_net_run with the batch cut into many chunks (H2D / layers / D2H of neighbouring chunks overlap on three streams and two
    staging slots): every image must still come out in place and bit-exact."""
    from simple_image_compression_network_b200.layer import Net
    d1 = cases.CASES["c2d_e"]
    d2 = dataclasses.replace(cases.CASES["dc_c"], ifm_x=24, ifm_y=16)
    reps = 11  # 6 chunks, the last one short
    i1, i2 = cases.make_inputs(d1, num_reps=reps, relu_range=True), cases.make_inputs(d2, seed_shift=9)
    net = Net([_layer(d1, i1), _layer(d2, i2)])
    net.set_host_chunk(2)
    got = net.run(i1["in_words"], reps)
    mid = oracle_mod.run_layer(d1, i1["in_words"], i1["weights"], None, i1["bias"], num_reps=reps)
    want = oracle_mod.run_layer(d2, mid, i2["weights"], None, i2["bias"], num_reps=reps)
    assert np.array_equal(got, want)
    assert np.array_equal(net.run(i1["in_words"], reps), want)  # slots and events are reusable


@pytest.mark.parametrize("name", ["c2d_e", "c2d_c", "th_cfg4", "dc_d", "xn_c"])
def test_set_params_swaps_weights_in_place(name, fcb_lib, oracle_mod):
    """fcb_layer_set_params (run-time-writable weight memories, dma.h:214-236 + mvau.hpp:209-307): same handle, new weights /
    thresholds / bias, results follow the oracle with the new parameters -- also through a Net built before the swap."""
    from simple_image_compression_network_b200.layer import Net
    d = cases.CASES[name]
    a, b = cases.make_inputs(d, seed_shift=0), cases.make_inputs(d, seed_shift=77)
    L = _layer(d, a)
    net = Net([L])
    x = a["in_words"]
    assert np.array_equal(L.run(x), oracle_mod.run_layer(d, x, a["weights"], a["thresholds"], a["bias"]))
    L.set_params(b["weights"], thresholds=b["thresholds"], bias=b["bias"])
    want = oracle_mod.run_layer(d, x, b["weights"], b["thresholds"], b["bias"])
    assert not np.array_equal(want, oracle_mod.run_layer(d, x, a["weights"], a["thresholds"], a["bias"]))
    assert np.array_equal(L.run(x), want)
    assert np.array_equal(net.run(x), want)
    with pytest.raises(ValueError):
        L.set_params(b["weights"][:-1])


@pytest.mark.parametrize("name", ["th_a", "th_b", "th_c", "c2d_e", "dc_d", "xn_c"])
def test_set_param_stream(name, fcb_lib, oracle_mod):
    """fcb_layer_set_param_stream: weights handed over as one period of the reference's parameter stream (GenParamStream,
    dma.h:214-236; golden image recorded from the reference for the th_* cases) -- results follow the oracle."""
    d = cases.CASES[name]
    a, b = cases.make_inputs(d, seed_shift=0), cases.make_inputs(d, seed_shift=31)
    L = _layer(d, b)
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", f"param_stream_{name}.npz")
    pw = np.load(gold)["param_words"] if os.path.exists(gold) else oracle_mod.gen_param_stream(d, a["weights"])
    assert np.array_equal(pw, oracle_mod.gen_param_stream(d, a["weights"]))
    L.set_param_stream(pw, thresholds=a["thresholds"], bias=a["bias"])
    x = a["in_words"]
    assert np.array_equal(L.run(x), oracle_mod.run_layer(d, x, a["weights"], a["thresholds"], a["bias"]))
    with pytest.raises(ValueError):
        L.set_param_stream(pw[:-1])


def test_judged_configs_at_full_size(fcb_lib, oracle_mod):
    """BASELINE.json configs 3 and 5b (all four stages) and the wide-lane config of bench.py at their full sizes against the oracle (config 4 at full size is
    `th_cfg4` above; config 2 is `c2d_L1` and tests/test_bench_scale.py; config 5a / the full network are tests/test_net8.py)."""
    from simple_image_compression_network_b200.desc import ACT_THRESHOLDS, KIND_CONV, W_BINARY_XNOR, LayerDesc
    c3 = LayerDesc(kind=KIND_CONV, kernel_x=3, kernel_y=3, ifm_ch=64, ofm_ch=64, ifm_x=128, ifm_y=96, stride_x=1, stride_y=1, pad=0,
                   simd=64, pe=16, in_bits=1, w_bits=1, weight_kind=W_BINARY_XNOR, acc_bits=16, acc_signed=1, act_kind=ACT_THRESHOLDS,
                   out_bits=1, num_th=1)

    def stage(c, ofm, x, y, simd, pe):
        return LayerDesc(kind=KIND_CONV, kernel_x=3, kernel_y=3, ifm_ch=c, ofm_ch=ofm, ifm_x=x, ifm_y=y, stride_x=1, stride_y=1, pad=1,
                         simd=simd, pe=pe, in_bits=8, w_bits=4, acc_bits=24, acc_signed=1, act_kind=ACT_THRESHOLDS, out_bits=8,
                         num_th=255, pool=2)
    from simple_image_compression_network_b200 import workloads
    for name, d in (("config 3", c3), ("config 5b stage 1", stage(3, 128, 768, 512, 3, 16)), ("config 5b stage 2", stage(128, 128, 384, 256, 32, 16)),
                    ("config 5b stage 3", stage(128, 128, 192, 128, 32, 16)), ("config 5b stage 4", stage(128, 192, 96, 64, 32, 24)),
                    ("16-bit lanes on the universal engine", workloads.imad16())):
        inp = cases.make_inputs(d, seed_shift=61)
        L = _layer(d, inp)
        got = L.run(inp["in_words"])
        want = oracle_mod.run_layer(d, inp["in_words"], inp["weights"], inp["thresholds"], inp["bias"])
        assert np.array_equal(got, want), f"{name} [{L.engine}: {L.plan}]: {_diff(got, want)}"


@pytest.mark.parametrize("name,reps", [("c2d_e", 3), ("dc_c", 2), ("th_cfg4", 1), ("c2d_g", 5)])
def test_clustered_weight_multicast(name, reps, fcb_lib, oracle_mod, monkeypatch, exp_build):
    """Experiment build: 2-CTA clusters sharing one multicast stream of weight K-blocks (umma2_conv_kernel<.., CS = 2>), odd tile
    counts included (a CTA with one tile fewer runs the ring protocol of a phantom tile): bit-exact against the oracle."""
    d = cases.CASES[name]
    inp = cases.make_inputs(d, seed_shift=17, num_reps=reps, relu_range=d.kind != 0)
    want = oracle_mod.run_layer(d, inp["in_words"], inp["weights"], inp["thresholds"], inp["bias"], num_reps=reps)
    exp_build()
    monkeypatch.setenv("FCB_U2_CLUSTER", "1")
    L = _layer(d, inp)
    got = L.run(inp["in_words"], reps)
    if "chb=1" in L.plan:
        assert "weights-multicast" in L.plan, L.plan
    assert np.array_equal(got, want), f"{name} [{L.plan}]: {_diff(got, want)}"


@pytest.mark.parametrize("name", ["add_u8", "add_s8_off"])
def test_add_streams(name, fcb_lib, oracle_mod):
    """fcb_add_streams (AddStreams_Batch, streamtools.h:669-720) against the reference-made golden and, on a larger ragged case, the oracle."""
    from tests.test_oracle_golden import _add_cases
    from simple_image_compression_network_b200 import pack, synth
    from simple_image_compression_network_b200.layer import add_streams
    a, b, n, ch, b1, s1, b2, s2, ob, off = _add_cases()[name]
    want = np.load(os.path.join(GOLD, f"{name}.npz"))["out"]
    assert np.array_equal(add_streams(a, b, n, ch, b1, s1, b2, s2, ob, off), want)
    n2, ch2 = 1003, 70  # channels not a multiple of 32, odd lane widths
    a2 = pack.pack_words(synth.lanes(21, (n2, ch2), b1), b1).reshape(-1)
    b2w = pack.pack_words(synth.lanes(22, (n2, ch2), 5), 5).reshape(-1)
    got = add_streams(a2, b2w, n2, ch2, b1, s1, 5, 1, 11, off)
    assert np.array_equal(got, oracle_mod.add_streams(a2, b2w, n2, ch2, b1, s1, 5, 1, 11, off))


def test_fc_layer_wrapper(fcb_lib, oracle_mod):
    """StreamingFCLayer_Batch (fclayer.h:83-111) through the host wrapper: big tiles + one-pixel remainder == the 1x1 layer of the oracle
    (pinned on the reference's own StreamingFCLayer_Batch by the `fc_a` golden)."""
    from simple_image_compression_network_b200 import configs
    from simple_image_compression_network_b200.layer import FCLayer
    d = cases.CASES["fc_a"]
    reps = 2 * 16 + 5
    dd = dataclasses.replace(d, ifm_x=reps)
    inp = cases.make_inputs(dd, seed_shift=3)
    fc = FCLayer(64, 32, 8, 4, inp["weights"], tile=16, in_bits=8, in_signed=0, w_bits=4, acc_bits=16, acc_signed=1, act_kind=0, out_bits=16)
    got = fc.run(inp["in_words"], reps)
    want = oracle_mod.run_layer(dd, inp["in_words"], inp["weights"], None, None)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("c,ofm,x,y,nth,k", [(3, 16, 64, 14, 15, 5), (3, 128, 64, 14, 15, 3), (128, 128, 40, 12, 15, 3), (128, 64, 22, 8, 63, 3),
                                              (32, 256, 36, 10, 15, 3), (4, 256, 48, 6, 127, 3)])
def test_pooled_thresholds_small_tables(c, ofm, x, y, nth, k, fcb_lib, oracle_mod):
    """Pooled 8-bit threshold layers whose tables are SHORT (15 / 63 / 127 thresholds: no bucket LUT, every level in shared memory), on the
    thin-input and the resident-planes instantiations with 8 epilogue warps -- the corner the seeded fuzz found (seed 101, layer 28)."""
    from simple_image_compression_network_b200.desc import ACT_THRESHOLDS, KIND_CONV, LayerDesc
    d = LayerDesc(kind=KIND_CONV, kernel_x=k, kernel_y=k, ifm_ch=c, ofm_ch=ofm, ifm_x=x, ifm_y=y, stride_x=1, stride_y=1, pad=k // 2,
                  simd=c if c < 8 else 8, pe=16, in_bits=8, w_bits=4, acc_bits=24, acc_signed=1, act_kind=ACT_THRESHOLDS, out_bits=8, num_th=nth, pool=2)
    for reps in (2, 1):
        inp = cases.make_inputs(d, seed_shift=129, num_reps=reps, relu_range=True)
        L = _layer(d, inp)
        got = L.run(inp["in_words"], reps)
        want = oracle_mod.run_layer(d, inp["in_words"], inp["weights"], inp["thresholds"], None, num_reps=reps)
        assert np.array_equal(got, want), f"[{L.engine}: {L.plan}]: {_diff(got, want)}"


@pytest.mark.parametrize("c,ofm,k,s,dx,dy,pad,x,y,engine", [
    (128, 128, 3, 1, 2, 1, 2, 40, 12, "umma_i8"),   # resident planes: a dilated tap is just another plane offset
    (128, 64, 3, 2, 2, 2, 2, 44, 16, "umma_i8"),    # stride 2 + dilation on both axes
    (256, 192, 3, 1, 3, 2, 0, 30, 14, "umma_i8"),
    (8, 12, 3, 1, 2, 3, 1, 21, 13, "imad"),
    (6, 5, 2, 2, 3, 1, 0, 17, 9, "imad"),
])
def test_dilation_on_the_engines(c, ofm, k, s, dx, dy, pad, x, y, engine, fcb_lib, oracle_mod):
    """ConvolutionInputGenerator_NonSquare_Dilated (slidingwindow.h:1515-1631) on the tensor and IMAD engines; the oracle's dilation is
    pinned on the reference generator by the dil_* goldens (x axis; the reference asserts Dilation_y == 1, y follows by symmetry)."""
    from simple_image_compression_network_b200.desc import ACT_BIAS_RELU, KIND_CONV, LayerDesc
    d = LayerDesc(kind=KIND_CONV, kernel_x=k, kernel_y=k, ifm_ch=c, ofm_ch=ofm, ifm_x=x, ifm_y=y, stride_x=s, stride_y=s, pad=pad,
                  simd=2, pe=1, in_bits=8, w_bits=4, acc_bits=8, acc_signed=0, act_kind=ACT_BIAS_RELU, out_bits=8, dilation_x=dx, dilation_y=dy)
    inp = cases.make_inputs(d, seed_shift=77, num_reps=2, relu_range=True)
    L = _layer(d, inp)
    assert L.engine == engine, L.plan
    got = L.run(inp["in_words"], 2)
    want = oracle_mod.run_layer(d, inp["in_words"], inp["weights"], None, inp["bias"], num_reps=2)
    assert np.array_equal(got, want), f"[{L.engine}: {L.plan}]: {_diff(got, want)}"


@pytest.mark.parametrize("seed", [21, 22])
def test_channelwise_byte_lane_forms(seed, fcb_lib, oracle_mod):
    """chanwise_bytes_kernel (8-bit lanes moved as 4- / 16-byte vectors): every pool function, depth-wise with pass-through and
    thresholds, 8 / 16 / 32-bit outputs, signed and unsigned lanes, padding, stride, dilation; channel counts that take the 16-channel,
    the 4-channel and the general kernel -- against the oracle."""
    from simple_image_compression_network_b200.desc import (ACT_PASSTHROUGH, ACT_THRESHOLDS, KIND_DWCONV, KIND_POOL, LayerDesc)
    rng = np.random.default_rng(seed)
    forms = set()
    for i in range(40):
        c = int(rng.choice([4, 8, 16, 48, 128, 6])); pe = int(rng.choice([p for p in (1, 2, 4) if c % p == 0]))
        k = int(rng.choice([2, 3])); s = int(rng.choice([1, 2])); pad = int(rng.integers(0, 2))
        x = 4 * int(rng.integers(2, 8)); y = 4 * int(rng.integers(1, 4))
        ins = int(rng.integers(0, 2)); outb = int(rng.choice([8, 16, 32]))
        common = dict(kernel_x=k, kernel_y=k, ifm_ch=c, ofm_ch=c, ifm_x=x, ifm_y=y, stride_x=s, stride_y=s, pad=pad, simd=pe, pe=pe, in_bits=8,
                      in_signed=ins, dilation_x=int(rng.choice([1, 1, 2])))
        if i % 2:
            fn = i // 2 % 4
            # (sums also with TA's signedness different from the lanes': negative sums wrap in an unsigned TA; max needs them equal here)
            tas = ins if fn == 0 or i % 3 else 1 - ins
            d = LayerDesc(kind=KIND_POOL, w_bits=0, weight_kind=fn, acc_bits=int(rng.choice([8, 12, 16])), acc_signed=tas, act_kind=ACT_PASSTHROUGH,
                          out_bits=outb, act_val=int(rng.choice([2, 3, 4])), **common)
        else:
            thr = i % 4 == 0
            d = LayerDesc(kind=KIND_DWCONV, w_bits=int(rng.choice([4, 8])), acc_bits=int(rng.choice([16, 24])), acc_signed=1,
                          act_kind=ACT_THRESHOLDS if thr else ACT_PASSTHROUGH, out_bits=8 if thr else outb, num_th=255 if thr else 0, **common)
        inp = cases.make_inputs(d, seed_shift=seed + i, num_reps=3)
        L = _layer(d, inp)
        assert L.engine == "chanwise"
        form = "16" if "16 channels" in L.plan else "4" if "4 channels" in L.plan else "general"
        assert form == ("general" if (c % 4 or L.in_bytes % 16 or L.out_bytes % 16) else "16" if c % 16 == 0 else "4"), (d, L.plan)
        forms.add(form)
        got = L.run(inp["in_words"], 3)
        want = oracle_mod.run_layer(d, inp["in_words"], inp["weights"], inp["thresholds"], inp["bias"], num_reps=3)
        assert np.array_equal(got, want), f"case {i} {d} [{L.plan}]: {_diff(got, want)}"
    assert forms == {"16", "4", "general"}


def test_channelwise_units_at_bench_size(fcb_lib, oracle_mod):
    """The channel-wise shapes tools/bench_layers.py times (384x256 frames, 128 channels): 2x2 max pool, 3x3 signed max pool with padding,
    2x2 average, depth-wise 3x3 with pass-through and with 255 thresholds, AddStreams on the same frames -- full size against the oracle."""
    from simple_image_compression_network_b200.desc import (ACT_PASSTHROUGH, ACT_THRESHOLDS, KIND_DWCONV, KIND_POOL, POOLFN_AVG, POOLFN_MAX, LayerDesc)
    from simple_image_compression_network_b200.layer import add_streams
    geo = dict(ifm_ch=128, ofm_ch=128, ifm_x=384, ifm_y=256, simd=16, pe=16, in_bits=8)
    descs = [LayerDesc(kind=KIND_POOL, kernel_x=k, kernel_y=k, stride_x=st, stride_y=st, pad=pad, in_signed=ins, w_bits=0, weight_kind=fn, acc_bits=tab,
                       acc_signed=ins, act_kind=ACT_PASSTHROUGH, out_bits=8, act_val=size, **geo)
             for k, st, pad, ins, fn, tab, size in ((2, 2, 0, 0, POOLFN_MAX, 8, 0), (3, 1, 1, 1, POOLFN_MAX, 8, 0), (2, 2, 0, 0, POOLFN_AVG, 10, 4))]
    descs += [LayerDesc(kind=KIND_DWCONV, kernel_x=3, kernel_y=3, stride_x=1, stride_y=1, pad=1, in_signed=0, w_bits=4, acc_bits=16, acc_signed=1,
                        act_kind=ACT_THRESHOLDS if thr else ACT_PASSTHROUGH, out_bits=8 if thr else 16, num_th=255 if thr else 0, **geo) for thr in (False, True)]
    for d in descs:
        inp = cases.make_inputs(d, seed_shift=5)
        L = _layer(d, inp)
        assert "16 channels" in L.plan
        got = L.run(inp["in_words"])
        want = oracle_mod.run_layer(d, inp["in_words"], inp["weights"], inp["thresholds"], inp["bias"])
        assert np.array_equal(got, want), f"{d} [{L.plan}]: {_diff(got, want)}"
    rng = np.random.default_rng(3)
    a = rng.integers(0, 256, 384 * 256 * 128, dtype=np.uint8); b = rng.integers(0, 256, a.size, dtype=np.uint8)
    for ob, s1 in ((8, False), (16, True)):
        got = add_streams(a, b, 384 * 256, 128, 8, s1, 8, False, ob, offset=-3)
        want = oracle_mod.add_streams(a, b, 384 * 256, 128, 8, s1, 8, False, ob, -3)
        assert np.array_equal(got, want), f"add_streams ob={ob}"


@pytest.mark.parametrize("seed", [11, 12])
def test_channelwise_units_fuzz(seed, fcb_lib, oracle_mod):
    """Depth-wise convolution (VVAU) and Pool_batch with random geometry, lane widths, signedness, functions, padding, stride and dilation
    against the oracle (whose semantics are pinned on the reference's templates by the pl_* / dw_* goldens)."""
    from simple_image_compression_network_b200._lib import FcbError
    from simple_image_compression_network_b200.desc import (ACT_PASSTHROUGH, ACT_THRESHOLDS, KIND_DWCONV, KIND_POOL, LayerDesc)
    rng = np.random.default_rng(seed)
    ran = 0
    for i in range(30):
        c = int(rng.choice([3, 8, 24, 40, 128])); pe = int(rng.choice([p for p in (1, 2, 4, 8) if c % p == 0]))
        k = int(rng.choice([2, 3])); s = int(rng.choice([1, 2])); pad = int(rng.integers(0, 2))
        x = int(rng.integers(k + 2, 30)); y = int(rng.integers(k + 2, 14))
        inb = int(rng.choice([3, 4, 8])); ins = int(rng.integers(0, 2))
        common = dict(kernel_x=k, kernel_y=k, ifm_ch=c, ofm_ch=c, ifm_x=x, ifm_y=y, stride_x=s, stride_y=s, pad=pad, simd=pe, pe=pe, in_bits=inb,
                      in_signed=ins, dilation_x=int(rng.choice([1, 1, 2])))
        if rng.integers(0, 2):
            fn = int(rng.integers(0, 4)); tab = int(rng.choice([inb, inb + 4, 16]))
            d = LayerDesc(kind=KIND_POOL, w_bits=0, weight_kind=fn, acc_bits=tab, acc_signed=ins, act_kind=ACT_PASSTHROUGH,
                          out_bits=int(rng.choice([inb, 8, 12])), act_val=int(rng.choice([2, 3, 4])), **common)
        else:
            thr = bool(rng.integers(0, 2))
            d = LayerDesc(kind=KIND_DWCONV, w_bits=int(rng.choice([3, 4, 8])), acc_bits=int(rng.choice([12, 16, 24])), acc_signed=1,
                          act_kind=ACT_THRESHOLDS if thr else ACT_PASSTHROUGH, out_bits=4 if thr else int(rng.choice([8, 12, 16])),
                          num_th=15 if thr else 0, **common)
        reps = 1 + i % 3
        try:
            inp = cases.make_inputs(d, seed_shift=seed + i, num_reps=reps)
            L = _layer(d, inp)
        except (FcbError, ValueError, AssertionError):
            continue
        got = L.run(inp["in_words"], reps)
        want = oracle_mod.run_layer(d, inp["in_words"], inp["weights"], inp["thresholds"], inp["bias"], num_reps=reps)
        assert L.engine == "chanwise"
        assert np.array_equal(got, want), f"case {i} {d}: {_diff(got, want)}"
        ran += 1
    assert ran >= 20, ran
