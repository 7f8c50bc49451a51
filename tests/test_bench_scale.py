"""Parity at the BENCHMARKED scale: the batch sizes bench.py times (4096 CONV_1 images per launch; eight_layers_net over several
device and host chunks) are compared with the oracle on sampled images, and the reference's UNMODIFIED testbench is run against
the GPU backend (conv3_nonsquare_tb.cpp:1068-1104 is the pass criterion).  The inputs are generated on the device with the same
counter-based rule the host uses (tests/test_gpu_parity.py::test_device_synth_matches_host)."""
import os
import subprocess

import numpy as np
import pytest

from simple_image_compression_network_b200 import configs, synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _host_image(first_byte: int, n_bytes: int, mask: int) -> np.ndarray:
    return synth.lanes(synth.SEED_INPUT, (n_bytes,), 8, mask=mask, offset=first_byte).astype(np.uint8)


def test_conv1_4096_images_sampled_against_oracle(fcb_lib, oracle_mod):
    """The headline workload as bench.py runs it: ONE launch over 4096 device-generated CONV_1 images (TMA image coordinate up to
    4095, 64-bit tile bookkeeping); images 0, 1, middle, last-1, last are compared byte for byte with the oracle."""
    import torch
    from simple_image_compression_network_b200.layer import ConvLayer, synth_fill
    d = configs.net_layer(1)
    prm = configs.synthetic_params(d)
    L = ConvLayer(d, prm["weights"], bias=prm["bias"])
    n = 4096
    free, _ = torch.cuda.mem_get_info()
    if free < n * (L.in_bytes + L.out_bytes) * 1.05:
        n = 1024
    x = torch.empty(n * L.in_bytes, dtype=torch.uint8, device="cuda")
    y = torch.zeros(n * L.out_bytes, dtype=torch.uint8, device="cuda")
    synth_fill(x.data_ptr(), x.numel(), synth.SEED_INPUT, 0x7F)
    L.run_device(x.data_ptr(), y.data_ptr(), n, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert L.engine == "umma_i8"
    for i in (0, 1, n // 2, n - 2, n - 1):
        xi = _host_image(i * L.in_bytes, L.in_bytes, 0x7F)
        assert np.array_equal(x[i * L.in_bytes:(i + 1) * L.in_bytes].cpu().numpy(), xi), f"device-generated image {i} differs from the host rule"
        want = oracle_mod.run_layer(d, xi, prm["weights"], None, prm["bias"])
        got = y[i * L.out_bytes:(i + 1) * L.out_bytes].cpu().numpy()
        bad = np.flatnonzero(got != want)
        assert bad.size == 0, f"image {i} of {n}: {bad.size} bytes differ, first at {bad[:8].tolist()}"
    del x, y
    torch.cuda.empty_cache()


def test_net8_many_chunks_sampled_against_oracle(fcb_lib, oracle_mod):
    """eight_layers_net at full size (768x512) over a batch that spans 4 device chunks (fcb_net_run_device, default chunking) and,
    through the host-buffer call, 8 staging chunks: sampled images against the oracle's 8-layer chain, and the two entry points
    against each other on every byte."""
    import torch
    from simple_image_compression_network_b200.layer import ConvLayer, Net, synth_fill
    descs = [configs.net_layer(i) for i in range(8)]
    prms = [configs.synthetic_params(d, seed_shift=70 + i) for i, d in enumerate(descs)]
    net = Net([ConvLayer(d, p["weights"], bias=p["bias"]) for d, p in zip(descs, prms)])
    n = 300
    x = torch.empty(n * net.in_bytes, dtype=torch.uint8, device="cuda")
    y = torch.zeros(n * net.out_bytes, dtype=torch.uint8, device="cuda")
    synth_fill(x.data_ptr(), x.numel(), synth.SEED_INPUT, 0xFF)  # (the pad byte of the ap_uint<24> container is ignored by readers)
    net.run_device(x.data_ptr(), y.data_ptr(), n, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    got_all = y.cpu().numpy()
    for i in (0, 85, 149, n - 1):  # 85 = first image of the second device chunk
        s = _host_image(i * net.in_bytes, net.in_bytes, 0xFF)
        for d, p in zip(descs, prms):
            s = oracle_mod.run_layer(d, s, p["weights"], None, p["bias"])
        got = got_all[i * net.out_bytes:(i + 1) * net.out_bytes]
        bad = np.flatnonzero(got != s)
        assert bad.size == 0, f"image {i} of {n}: {bad.size} bytes differ, first at {bad[:8].tolist()}"
    host = net.run(x.cpu().numpy(), n)  # H2D | layers | D2H pipelined over 8 chunks of 42 images
    assert np.array_equal(host, got_all)
    net.set_device_chunk(7)  # many small passes of the chain (the L2-resident regime), ragged last pass
    y.zero_()
    net.run_device(x.data_ptr(), y.data_ptr(), n, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert np.array_equal(y.cpu().numpy(), got_all)


def test_layer_host_call_many_chunks_lowered_and_direct(fcb_lib, oracle_mod):
    """fcb_layer_run with the batch cut into many staging chunks on two streams, on a LOWERED layer (im2col rows staged through
    per-slot scratch buffers), a tensor layer and a direct-engine layer: every image in place and bit-exact."""
    import dataclasses
    from oracle import cases
    from simple_image_compression_network_b200.desc import ACT_BIAS_RELU, KIND_CONV, LayerDesc
    from simple_image_compression_network_b200.layer import ConvLayer
    lowered = LayerDesc(kind=KIND_CONV, kernel_x=3, kernel_y=3, ifm_ch=8, ofm_ch=32, ifm_x=30, ifm_y=14, stride_x=1, stride_y=1, pad=1,
                        simd=8, pe=8, in_bits=8, w_bits=4, acc_bits=8, acc_signed=0, act_kind=ACT_BIAS_RELU, out_bits=8)
    small_out = dataclasses.replace(cases.CASES["xn_a"])  # 8 x 1-bit lanes: 1-byte output words (chunk offsets need rounding)
    for d, want_plan in ((lowered, "im2col rows"), (cases.CASES["c2d_e"], "resident-planes"), (cases.CASES["dc_a"], "direct"), (small_out, "direct")):
        reps = 23
        inp = cases.make_inputs(d, seed_shift=5, num_reps=reps)
        L = ConvLayer(d, inp["weights"], thresholds=inp["thresholds"], bias=inp["bias"])
        assert want_plan in L.plan, L.plan
        want = oracle_mod.run_layer(d, inp["in_words"], inp["weights"], inp["thresholds"], inp["bias"], num_reps=reps)
        for chunk in (2, 5, 0):
            L.set_host_chunk(chunk)
            got = L.run(inp["in_words"], reps)
            bad = np.flatnonzero(got != want)
            assert bad.size == 0, f"[{L.plan}] chunk {chunk}: {bad.size} bytes differ, first at {bad[:8].tolist()}"


def test_unmodified_reference_testbench_on_gpu_backend(fcb_lib):
    """oracle/_ref/tb_b200 = the reference's UNMODIFIED conv3_nonsquare_tb.cpp linked against include/finnconv_hls_adapter.hpp +
    libfinnconv_b200.so instead of conv_nonsquare_top.cpp (oracle/Makefile `tb_b200`, built where /root/reference exists).  Its
    own golden chain (conv.hpp:91-123, ~3 min on one host core) judges the GPU network: "passed the testing", 0 ERROR lines."""
    tb = os.path.join(ROOT, "oracle", "_ref", "tb_b200")
    if not os.path.exists(tb):
        pytest.skip("oracle/_ref/tb_b200 was not built (needs /root/reference at build time)")
    env = dict(os.environ)
    env["LD_LIBRARY_PATH"] = os.path.join(ROOT, "simple_image_compression_network_b200") + ":" + env.get("LD_LIBRARY_PATH", "")
    r = subprocess.run(f"ulimit -s unlimited 2>/dev/null || ulimit -s 1048576; exec {tb}", shell=True, capture_output=True, text=True,
                       timeout=900, env=env, cwd=ROOT)
    out = r.stdout + r.stderr
    assert r.returncode == 0, out[-2000:]
    assert "passed the testing" in out, out[-2000:]
    assert "ERROR" not in out, out[-2000:]


def test_pool_splits_the_batch_over_replicas(fcb_lib, oracle_mod):
    """fcb_pool_*: the chain replicated behind one handle, numReps split into contiguous image ranges, one host thread per replica.
    On a one-GPU box the replicas share device 0 (a device may be listed twice); with more GPUs every device is used."""
    import torch
    from oracle import cases
    from simple_image_compression_network_b200.layer import Pool
    d1 = cases.CASES["c2d_e"]
    import dataclasses
    d2 = dataclasses.replace(cases.CASES["dc_c"], ifm_x=24, ifm_y=16)
    reps = 13
    i1, i2 = cases.make_inputs(d1, num_reps=reps, relu_range=True), cases.make_inputs(d2, seed_shift=9)
    mid = oracle_mod.run_layer(d1, i1["in_words"], i1["weights"], None, i1["bias"], num_reps=reps)
    want = oracle_mod.run_layer(d2, mid, i2["weights"], None, i2["bias"], num_reps=reps)
    ndev = torch.cuda.device_count()
    for devices in ([0, 0, 0], None, list(range(ndev)) * 2):
        pool = Pool([d1, d2], [i1["weights"], i2["weights"]], None, [i1["bias"], i2["bias"]], devices=devices)
        assert pool.replicas == (len(devices) if devices else ndev)
        for n in (reps, 2, 1):  # fewer images than replicas: the high ranks get empty ranges
            got = pool.run(i1["in_words"][: n * pool.in_bytes], n)
            assert np.array_equal(got, want[: n * pool.out_bytes]), f"devices={devices} n={n}"
        pool.close()
    one = Pool([d1], [i1["weights"]], None, [i1["bias"]], devices=[0, 0])  # single layer: fcb_layer_run underneath
    assert np.array_equal(one.run(i1["in_words"], reps), mid)


def test_adapter_against_reference_functions(fcb_lib):
    """oracle/_ref/adapter_check: include/finnconv_hls_adapter.hpp (HLS streams, QDMA streams with TKEEP / TLAST, 64-bit AXI memory
    images incl. the 16-image bursts of Mem2Stream_Batch / Stream2Mem_Batch, and the multi-GPU pool) against the reference's own
    conv2d<>, Qdma2Stream_Batch / Stream2Qdma_Batch (streamtools.h:1001-1037), Mem2Stream_Batch / Stream2Mem_Batch (dma.h:135-199) and
    StreamingDataWidthConverter_Batch (streamtools.h:463-526), compiled from /root/reference into the checker binary."""
    exe = os.path.join(ROOT, "oracle", "_ref", "adapter_check")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/adapter_check was not built (needs /root/reference at build time)")
    env = dict(os.environ)
    env["LD_LIBRARY_PATH"] = os.path.join(ROOT, "simple_image_compression_network_b200") + ":" + env.get("LD_LIBRARY_PATH", "")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    out = r.stdout + r.stderr
    assert r.returncode == 0 and "all 6 checks passed" in out, out[-2000:]
