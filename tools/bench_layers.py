"""Per-layer and whole-network device-resident throughput of the reference network (config_nonsquare.h) and of the
judged synthetic configs (BASELINE.json configs 3/4): images/s, TOP/s on the MACs the reference executes and on the
non-zero MACs, achieved HBM GB/s on the algorithmic bytes.  Diagnostics, not the headline bench.
    python tools/bench_layers.py [--images N]"""
import argparse, dataclasses, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from simple_image_compression_network_b200 import configs, synth
from simple_image_compression_network_b200.desc import (ACT_THRESHOLDS, ENGINE_IMAD, ENGINE_TENSOR, KIND_CONV, KIND_DECONV522, W_BINARY_XNOR,
                                                        LayerDesc)
from simple_image_compression_network_b200.layer import ConvLayer, Net, synth_fill
if os.environ.get("FCB_EXP"):  # the experiment build (environment switches that bend plans): measurements only
    from simple_image_compression_network_b200 import _lib
    _lib.set_default(_lib.load(_lib.EXP_LIB_PATH))
if os.environ.get("FCB_LIB"):  # any other build of the same sources (A/B of a compile-time choice)
    from simple_image_compression_network_b200 import _lib
    _lib.set_default(_lib.load(os.environ["FCB_LIB"]))


def timed(fn, steps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


CHECK = False


def bench_layer(name, d, n, mask=0x7F):
    prm = configs.synthetic_params(d)
    L = ConvLayer(d, prm["weights"], thresholds=prm["thresholds"], bias=prm["bias"])
    x = torch.empty(n * L.in_bytes, dtype=torch.uint8, device="cuda")
    y = torch.empty(n * L.out_bytes, dtype=torch.uint8, device="cuda")
    synth_fill(x.data_ptr(), x.numel(), synth.SEED_INPUT, mask)
    ms = timed(lambda: L.run_device(x.data_ptr(), y.data_ptr(), n, torch.cuda.current_stream().cuda_stream))
    if CHECK:  # first and last image against the independent IMAD engine (itself pinned to the oracle by tests/)
        L2 = ConvLayer(dataclasses.replace(d, engine_hint=ENGINE_IMAD), prm["weights"], thresholds=prm["thresholds"], bias=prm["bias"])
        assert L2.engine != L.engine or L.engine == "imad"
        for i in sorted({0, n - 1}):
            y2 = torch.empty(L.out_bytes, dtype=torch.uint8, device="cuda")
            L2.run_device(x.data_ptr() + i * L.in_bytes, y2.data_ptr(), 1, torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            assert torch.equal(y2, y[i * L.out_bytes:(i + 1) * L.out_bytes]), f"{name}: image {i} differs from the IMAD engine"
    macs = d.macs_per_image
    nz = macs / 4 if d.kind == KIND_DECONV522 else macs
    r = dict(layer=name, engine=L.engine, plan=L.plan, images=n, ms=round(ms, 3), img_s=round(n / ms * 1e3),
             TOPs_dense=round(2 * macs * n / ms / 1e9, 1), TOPs_nonzero=round(2 * nz * n / ms / 1e9, 1),
             GBs=round((L.in_bytes + L.out_bytes) * n / ms / 1e6, 1), checked=bool(CHECK))
    print(json.dumps(r), flush=True)
    return L, r


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=64)
    ap.add_argument("--only", default="", help="comma list of net layers to time alone, e.g. L0,L6 (skips the rest)")
    ap.add_argument("--check", action="store_true", help="compare image 0 / n-1 of every timed layer with the IMAD engine")
    a = ap.parse_args()
    CHECK = a.check
    c3 = LayerDesc(kind=KIND_CONV, kernel_x=3, kernel_y=3, ifm_ch=64, ofm_ch=64, ifm_x=128, ifm_y=96, stride_x=1, stride_y=1, pad=0,
                   simd=64, pe=16, in_bits=1, w_bits=1, weight_kind=W_BINARY_XNOR, acc_bits=16, acc_signed=1,
                   act_kind=ACT_THRESHOLDS, out_bits=1, num_th=1)
    c4 = LayerDesc(kind=KIND_CONV, kernel_x=3, kernel_y=3, ifm_ch=256, ofm_ch=256, ifm_x=64, ifm_y=48, stride_x=1, stride_y=1, pad=1,
                   simd=32, pe=32, in_bits=8, w_bits=4, acc_bits=24, acc_signed=1, act_kind=ACT_THRESHOLDS, out_bits=8, num_th=255,
                   pool=2)
    # BASELINE.json config 5b: analysis-transform-shaped stack [K3 S1 P1 conv -> 255 thresholds (u8) -> 2x2 max pool] x 4,
    # channels 3 -> 128 -> 128 -> 128 -> 192 on 768x512; 5a = layers 0-3 of the reference net
    def stage(c, ofm, x, y, simd, pe):
        return LayerDesc(kind=KIND_CONV, kernel_x=3, kernel_y=3, ifm_ch=c, ofm_ch=ofm, ifm_x=x, ifm_y=y, stride_x=1, stride_y=1, pad=1,
                         simd=simd, pe=pe, in_bits=8, w_bits=4, acc_bits=24, acc_signed=1, act_kind=ACT_THRESHOLDS, out_bits=8,
                         num_th=255, pool=2)
    st = [stage(3, 128, 768, 512, 3, 16), stage(128, 128, 384, 256, 32, 16), stage(128, 128, 192, 128, 32, 16), stage(128, 192, 96, 64, 32, 24)]
    if a.only:
        for nm in a.only.split(","):
            if nm.startswith("s5b"):
                i = int(nm[3:]) - 1
                bench_layer(f"stack5b_stage{i + 1}", st[i], a.images if i == 0 else a.images * 4, 0xFF)
                continue
            if nm == "cfg3t":
                bench_layer("cfg3_xnor_tensor", dataclasses.replace(c3, engine_hint=ENGINE_TENSOR), a.images, 0xFF)
            elif nm == "cfg3":
                bench_layer("cfg3_xnor", c3, a.images, 0xFF)
            elif nm == "cfg3p":  # the XNOR/popc warp kernels (FCB_ENGINE_AUTO takes the tensor form for this shape class)
                from simple_image_compression_network_b200.desc import ENGINE_XNOR_POPC
                bench_layer("cfg3_xnor_popc", dataclasses.replace(c3, engine_hint=ENGINE_XNOR_POPC), a.images, 0xFF)
            elif nm == "cfg4n":
                bench_layer("cfg4_thr_nopool", dataclasses.replace(c4, pool=0), a.images, 0xFF)
            elif nm == "cfg4":
                bench_layer("cfg4_thr_pool", c4, a.images, 0xFF)
            elif nm in ("pool", "pool3s", "avg", "dw", "dwthr"):  # channel-wise streaming units (fcb_chanwise.cu), HBM-bound
                from simple_image_compression_network_b200.desc import ACT_PASSTHROUGH, KIND_DWCONV, KIND_POOL, POOLFN_AVG, POOLFN_MAX
                if nm == "dw" or nm == "dwthr":  # depth-wise 3x3, 128 channels, 384x256 (u8 x s4)
                    d = LayerDesc(kind=KIND_DWCONV, kernel_x=3, kernel_y=3, ifm_ch=128, ofm_ch=128, ifm_x=384, ifm_y=256, stride_x=1, stride_y=1, pad=1,
                                  simd=16, pe=16, in_bits=8, in_signed=0, w_bits=4, acc_bits=16, acc_signed=1,
                                  act_kind=ACT_THRESHOLDS if nm == "dwthr" else ACT_PASSTHROUGH, out_bits=8 if nm == "dwthr" else 16,
                                  num_th=255 if nm == "dwthr" else 0)
                else:
                    k, st, pad, ins, fn, tab, size = {"pool": (2, 2, 0, 0, POOLFN_MAX, 8, 0), "pool3s": (3, 1, 1, 1, POOLFN_MAX, 8, 0),
                                                      "avg": (2, 2, 0, 0, POOLFN_AVG, 10, 4)}[nm]
                    d = LayerDesc(kind=KIND_POOL, kernel_x=k, kernel_y=k, ifm_ch=128, ofm_ch=128, ifm_x=384, ifm_y=256, stride_x=st, stride_y=st, pad=pad,
                                  simd=16, pe=16, in_bits=8, in_signed=ins, w_bits=0, weight_kind=fn, acc_bits=tab, acc_signed=ins,
                                  act_kind=ACT_PASSTHROUGH, out_bits=8, act_val=size)
                bench_layer(f"chanwise_{nm}", d, a.images, 0xFF)
            elif nm == "add":  # AddStreams_Batch on two 384x256x128-byte streams per image
                import ctypes
                from simple_image_compression_network_b200 import _lib
                from simple_image_compression_network_b200.desc import CAddDesc
                n = a.images; words = 384 * 256 * n
                x1 = torch.empty(words * 128, dtype=torch.uint8, device="cuda"); x2 = torch.empty_like(x1); y = torch.empty_like(x1)
                synth_fill(x1.data_ptr(), x1.numel(), synth.SEED_INPUT, 0x7F); synth_fill(x2.data_ptr(), x2.numel(), synth.SEED_INPUT + 1, 0x7F)
                cd = CAddDesc(ctypes.sizeof(CAddDesc), 128, 8, 0, 8, 0, 8, 0)
                Lb = _lib.lib()
                run = lambda: _lib.check(Lb.fcb_add_streams_device(ctypes.byref(cd), ctypes.c_void_p(x1.data_ptr()), ctypes.c_void_p(x2.data_ptr()),
                                                                   ctypes.c_void_p(y.data_ptr()), words, 0, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), Lb)
                ms = timed(run)
                print(json.dumps(dict(layer="add_streams_u8_128ch_384x256", images=n, ms=round(ms, 3), img_s=round(n / ms * 1e3),
                                      GBs=round(3 * x1.numel() / ms / 1e6, 1))), flush=True)
            elif nm == "imad16":  # 16-bit lanes x 8-bit weights: the universal engine's own shape (workloads.imad16)
                from simple_image_compression_network_b200 import workloads
                bench_layer("wide16_imad", workloads.imad16(), a.images * 8, 0xFF)
            elif nm.endswith("i"):  # L2i: layer 2 of the reference net forced onto the universal engine
                i = int(nm[1:-1])
                bench_layer(nm, dataclasses.replace(configs.net_layer(i), engine_hint=ENGINE_IMAD), a.images, 0xFF if i == 0 else 0x7F)
            else:
                i = int(nm[1:])
                bench_layer(nm, configs.net_layer(i), a.images, 0xFF if i == 0 else 0x7F)
        sys.exit(0)
    layers = []
    for i in range(8):
        n = a.images * (1 if i in (0, 6, 7) else 4)
        L, _ = bench_layer(f"L{i}", configs.net_layer(i), n, 0xFF if i == 0 else 0x7F)
        layers.append(L)
    net = Net(layers)
    n = a.images
    x = torch.empty(n * net.in_bytes, dtype=torch.uint8, device="cuda")
    y = torch.empty(n * net.out_bytes, dtype=torch.uint8, device="cuda")
    synth_fill(x.data_ptr(), x.numel(), synth.SEED_INPUT, 0xFF)
    ms = timed(lambda: net.run_device(x.data_ptr(), y.data_ptr(), n, torch.cuda.current_stream().cuda_stream))
    print(json.dumps(dict(layer="eight_layers_net", images=n, ms=round(ms, 3), img_s=round(n / ms * 1e3, 1),
                          TOPs_nonzero=round(2 * 28.94e9 * n / ms / 1e9, 1))), flush=True)
    # the same network through the host-buffer entry point (pinned memory; H2D + 8 layers + D2H, chunks pipelined on 3 streams)
    n = a.images * 8
    hx = torch.empty(n * net.in_bytes, dtype=torch.uint8, pin_memory=True)
    hy = torch.empty(n * net.out_bytes, dtype=torch.uint8, pin_memory=True)
    hx.random_(0, 256)
    import time
    net.run_raw(hx.data_ptr(), hy.data_ptr(), n)
    t0 = time.perf_counter()
    for _ in range(3):
        net.run_raw(hx.data_ptr(), hy.data_ptr(), n)
    dt = (time.perf_counter() - t0) / 3
    print(json.dumps(dict(layer="eight_layers_net_e2e_host_buffers", images=n, ms=round(dt * 1e3, 3), img_s=round(n / dt, 1),
                          pcie_GBs=round(n * (net.in_bytes + net.out_bytes) / dt / 1e9, 1))), flush=True)
    sl = []
    for i, d in enumerate(st):
        L, _ = bench_layer(f"stack5b_stage{i + 1}", d, a.images if i == 0 else a.images * 4, 0xFF)
        sl.append(L)
    for name, ls, macs in (("stack5b", sl, 20.84e9), ("stack5a_layers0-3", layers[:4], 14.47e9)):
        net2 = Net(ls)
        n = a.images
        x = torch.empty(n * net2.in_bytes, dtype=torch.uint8, device="cuda")
        y = torch.empty(n * net2.out_bytes, dtype=torch.uint8, device="cuda")
        synth_fill(x.data_ptr(), x.numel(), synth.SEED_INPUT, 0xFF)
        ms = timed(lambda: net2.run_device(x.data_ptr(), y.data_ptr(), n, torch.cuda.current_stream().cuda_stream))
        print(json.dumps(dict(layer=name, images=n, ms=round(ms, 3), img_s=round(n / ms * 1e3, 1), TOPs=round(2 * macs * n / ms / 1e9, 1))),
              flush=True)
    bench_layer("cfg3_xnor", c3, a.images * 16, 0xFF)
    bench_layer("cfg3_xnor_as_pm1_int8_tensor", dataclasses.replace(c3, engine_hint=ENGINE_TENSOR), a.images * 16, 0xFF)
    bench_layer("cfg4_thr_pool", c4, a.images * 16, 0xFF)
    bench_layer("cfg4_thr_nopool", dataclasses.replace(c4, pool=0), a.images * 16, 0xFF)
