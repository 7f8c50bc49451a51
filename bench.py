#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path (BASELINE.json: conv-layer images/sec at 1/2/4/8 B200).

Workload (config[1] of BASELINE.json): the CONV_1 layer of config_nonsquare.h:18-33 (128->128 channels,
384x256 -> 192x128, K5 S2 P2, u8 activations x s4 weights, 8-bit wrap + bias + ReLU) on a batch of 4096
synthetic images per GPU, images sharded over the GPUs with no collective (weak scaling).

  python bench.py [--gpus N] [--steps K] [--warmup W]            our CUDA path (torchrun for N > 1)
  python bench.py --impl reference [...]                         the reference's own CPU implementation

One JSON line on stdout (rank 0).  `value` = images/s with inputs resident in HBM; `e2e` = images/s through
the host-buffer C-ABI call (pinned host memory, H2D + kernel + D2H inside the timed region).  After the headline
(outside its timed region) every other BASELINE.json config is timed the same way and reported under `configs`,
and the first / last image of the timed 4096-image batch is compared with the oracle (`parity_checked_images`).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "conv_layer_images_per_sec"
UNIT = "images/s"
IMAGES_PER_GPU = 4096
E2E_IMAGES = 512
WORKLOAD = ("CONV_1 of config_nonsquare.h (128->128 ch, 384x256 -> 192x128, K5 S2 P2, u8 x s4, wrap8+bias+ReLU), "
            "synthetic images, batch sharded over GPUs")
CONFIG = {"workload": WORKLOAD, "images": "whole 384x256x128 images", "parallelism": "batch sharding, no collective"}


def _peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons; reports the samples that fall into a time window."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.monotonic(), line.strip()))

    def window(self, t0: float, t1: float):
        """Clock statistics of the samples taken in [t0, t1] (monotonic seconds)."""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm, mx, reasons = [], [], set()
        for ts, ln in list(self.lines):
            if ts < t0 or ts > t1:
                continue
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}

    def stop(self):
        if not self.proc:
            return
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()


# ---------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's own conv2d<> (oracle/_ref, built from /root/reference)
# ---------------------------------------------------------------------------------------------------
def _ref_worker(case):
    """One CONV_1 image (`c2d_L1`) or one band of it (`c2d_L1band`: 384x32 input rows -> 192x16 output rows, 1/8 image) through the
    reference's conv2d<> template; without oracle/_ref, through the oracle's C port on one thread."""
    from oracle import cases, oracle
    d = cases.CASES[case]
    inp = cases.make_inputs(d)
    s = oracle.query(d)
    if oracle.ref_available():
        _, secs = oracle.ref_run(case, inp["in_words"], inp["weights"], inp["bias"], s.out_bytes_per_image)
        return secs, "reference"
    t0 = time.perf_counter()
    oracle.set_threads(1)
    oracle.run_layer(d, inp["in_words"], inp["weights"], None, inp["bias"])
    return time.perf_counter() - t0, "port"


def cpu_reference_step(pool, cores: int, case: str):
    """One step = `cores` units in parallel (one process per core). Returns (images/s, wall seconds, kind)."""
    frac = 1.0 if case == "c2d_L1" else 1.0 / 8.0
    t0 = time.perf_counter()
    res = pool.map(_ref_worker, [case] * cores)
    wall = time.perf_counter() - t0
    return cores * frac / wall, wall, res[0][1]


def _ref_sample_text(case, cores, kind):
    what = ("whole CONV_1 images (384x256x128 -> 192x128x128)" if case == "c2d_L1"
            else "bands of CONV_1 (384x32 input rows -> 192x16 output rows, 1/8 image each)")
    return (f"per step: {cores} {what}, one process per core, through the reference's conv2d<> template "
            f"({'oracle/_ref, compiled from /root/reference' if kind == 'reference' else 'oracle C port'})")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = max(1, min(os.cpu_count() or 1, 64))
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        # warm-up = library load and page-in: bands are enough (1/8 of the work of a timed step); the first one also calibrates
        _, band_wall, kind = cpu_reference_step(pool, cores, "c2d_L1band")
        for _ in range(max(args.warmup - 1, 0)):
            cpu_reference_step(pool, cores, "c2d_L1band")
        # timed steps run WHOLE images (the literal CONV_1 instantiation) when K of them fit ~5.5 minutes, else bands
        case = "c2d_L1" if args.steps * 8.0 * band_wall <= 330.0 else "c2d_L1band"
        t0 = time.perf_counter()
        for _ in range(args.steps):
            _, _, kind = cpu_reference_step(pool, cores, case)
        total = time.perf_counter() - t0
    per_step = cores * (1.0 if case == "c2d_L1" else 0.125)
    value = args.steps * per_step / total
    sample = _ref_sample_text(case, cores, kind)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1000.0 * total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": CONFIG,
            "run": {"images_per_step": per_step, "warmup_steps_are": "bands (library load / page-in only)"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------
def _bind_to_gpu_numa_node(index: int):
    """Run this rank on the CPUs next to its GPU (NVML's affinity mask), so the pinned host buffers of the e2e leg are
    first-touched on the NUMA node its PCIe link hangs off; with every rank on node 0 the host-buffer path stops scaling."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
    except Exception:
        pass


class Harness:
    """Device timing shared by the headline and the `configs` block: CUDA events on the launch stream, warm-up >= 3, barrier +
    synchronize on both sides, MAX over ranks, clocks sampled during the timed region."""

    def __init__(self, torch, dist, world, local, sampler):
        self.torch, self.dist, self.world, self.local, self.sampler = torch, dist, world, local, sampler
        self.stream = torch.cuda.Stream()

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, v: float) -> float:
        t = self.torch.tensor([v], dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def time_steps(self, step, steps: int, warmup: int):
        """Returns (ms of this rank, max ms over ranks, clock window)."""
        torch = self.torch
        for _ in range(max(warmup, 3)):
            step()
        self.barrier()
        t0 = time.monotonic()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(self.stream)
        for _ in range(steps):
            step()
        ev1.record(self.stream)
        self.barrier()
        t1 = time.monotonic()
        ms = ev0.elapsed_time(ev1)
        return ms, self.max_over_ranks(ms), self.sampler.window(t0, t1)

    def time_for(self, step, target_s: float = 0.5, max_steps: int = 200):
        """Times enough steps for ~target_s of device time (so that nvidia-smi gets several samples of the sustained clock)."""
        ms1, _, _ = self.time_steps(step, 1, 3)
        steps = int(self.max_over_ranks(float(max(3, min(max_steps, round(target_s * 1000.0 / max(ms1, 1e-3)))))))
        ms, ms_max, clk = self.time_steps(step, steps, 0)
        return steps, ms, ms_max, clk


def _bound_entry(name, bound, img_s, ms_step, n_img, steps, clk, world, peaks, *, ops_per_image=None, bytes_per_image=None,
                 popc_words_per_image=None, alu_macs_per_image=None, alu_macs_per_s=None, extra=None):
    from simple_image_compression_network_b200 import workloads as W
    e = {"name": name, "images_per_gpu": n_img, "steps": steps, "ms_per_step": ms_step, "img_s": img_s, "n_gpus": world, "bound": bound,
         "sm_mhz": clk.get("sm_mhz"), "clock_reasons": clk.get("reasons")}
    per_gpu = img_s / world
    if ops_per_image is not None:
        tops = per_gpu * ops_per_image / 1e12
        e.update({"TOPs_per_gpu": tops, "frac_of_int8_spec_4500": tops / W.INT8_SPEC_TOPS, "frac_of_int8_mma_only_4380": tops / W.INT8_MMA_ONLY_TOPS})
    if bytes_per_image is not None:
        gbs = per_gpu * bytes_per_image / 1e9
        e.update({"GBs_per_gpu": gbs, "frac_of_hbm_measured": gbs / float(peaks["hbm_gbs"]), "algorithmic_bytes_per_image": bytes_per_image})
    if popc_words_per_image is not None:
        e.update({"popc_Twords_s_per_gpu": per_gpu * popc_words_per_image / 1e12, "frac_of_popc_peak": per_gpu * popc_words_per_image / W.POPC_WORDS_PER_S})
    if alu_macs_per_s is not None:  # universal direct engine: the packed dot-product issue rate (workloads.direct_macs_per_s)
        e.update({"TMACs_per_gpu": per_gpu * alu_macs_per_image / 1e12, "alu_peak_TMACs": alu_macs_per_s / 1e12,
                  "frac_of_dot_issue_rate": per_gpu * alu_macs_per_image / alu_macs_per_s})
    e["frac"] = {"tensor": e.get("frac_of_int8_spec_4500"), "hbm": e.get("frac_of_hbm_measured"), "popc": e.get("frac_of_popc_peak"),
                 "alu": e.get("frac_of_dot_issue_rate")}[bound]
    if extra:
        e.update(extra)
    return e


def run_configs(h: Harness, peaks, n_scale: float):
    """Every other BASELINE.json config, device-resident, timed like the headline (outside its timed region)."""
    import numpy as np
    import torch
    import dataclasses
    from simple_image_compression_network_b200 import configs, synth, workloads as W
    from simple_image_compression_network_b200.desc import ENGINE_XNOR_POPC
    from simple_image_compression_network_b200.layer import ConvLayer, Net, synth_fill
    sh = h.stream.cuda_stream
    out = []

    def mk(d, seed_shift=0):
        prm = configs.synthetic_params(d, seed_shift)
        return ConvLayer(d, prm["weights"], thresholds=prm["thresholds"], bias=prm["bias"], device=h.local)

    def run_one(name, runner, in_bytes, out_bytes, n, mask, entry_kw, extra=None):
        n = max(8, int(n * n_scale))
        x = torch.empty(n * in_bytes, dtype=torch.uint8, device="cuda")
        y = torch.empty(n * out_bytes, dtype=torch.uint8, device="cuda")
        synth_fill(x.data_ptr(), x.numel(), synth.SEED_INPUT, mask)
        torch.cuda.synchronize()
        steps, ms, ms_max, clk = h.time_for(lambda: runner(x.data_ptr(), y.data_ptr(), n, sh))
        img_s = h.world * n * steps / (ms_max / 1000.0)
        out.append(_bound_entry(name, entry_kw.pop("bound"), img_s, ms_max / steps, n, steps, clk, h.world, peaks, extra=extra, **entry_kw))
        del x, y
        torch.cuda.empty_cache()

    def layer_entry(name, d, n, mask, bound, seed_shift=0):
        L = mk(d, seed_shift)
        kw = {"bound": bound}
        if bound == "tensor":
            kw["ops_per_image"] = 2.0 * W.nonzero_macs(d)
        if bound in ("hbm", "tensor"):
            kw["bytes_per_image"] = L.in_bytes + L.out_bytes
        if bound == "popc":
            kw["popc_words_per_image"] = d.macs_per_image / 32.0
            kw["bytes_per_image"] = L.in_bytes + L.out_bytes
        if bound == "alu":
            kw["bytes_per_image"] = L.in_bytes + L.out_bytes
            kw["alu_macs_per_image"] = float(d.macs_per_image)
            kw["alu_macs_per_s"] = W.direct_macs_per_s(d)
        run_one(name, L.run_device, L.in_bytes, L.out_bytes, n, mask, kw, extra={"engine": L.engine, "plan": L.plan,
                                                                                 "GMAC_per_image_nonzero": W.nonzero_macs(d) / 1e9})
        return L

    # ---- layers of eight_layers_net (config_nonsquare.h): thin input / dominant deconv / thin output
    net_layers = {}
    for i, n, bound in ((0, 1024, "hbm"), (2, 4096, "tensor"), (3, 8192, "tensor"), (4, 8192, "tensor"), (5, 4096, "tensor"),
                        (6, 1024, "tensor"), (7, 1024, "hbm")):
        net_layers[i] = layer_entry(f"L{i}", configs.net_layer(i), n, 0xFF if i == 0 else 0x7F, bound)
    net_layers[1] = mk(configs.net_layer(1))
    # ---- config 3 (xnor-popcount) and config 4 (thresholds + pool)
    # config 3 twice: what FCB_ENGINE_AUTO runs (the +-1 int8 form on the tensor cores) and the XNOR/popc warp kernels north_star names
    layer_entry("config3_xnor_64x64_3x3_128x96", W.config3(), 16384, 0xFF, "tensor")
    layer_entry("config3_xnor_popc_engine_64x64_3x3_128x96", dataclasses.replace(W.config3(), engine_hint=ENGINE_XNOR_POPC), 16384, 0xFF, "popc")
    layer_entry("config4_thr255_pool_256x256_3x3_64x48", W.config4(), 4096, 0xFF, "tensor")
    layer_entry("wide_lanes_imad_s16xs8_64x64_3x3_96x64", W.imad16(), 2048, 0xFF, "alu")
    # ---- stacks: 5a = layers 0-3 of the reference net, 5b = 4 x [conv3x3 -> 255 thresholds -> pool], and the whole net
    st = W.stack5b()
    stages = [layer_entry(f"config5b_stage{i + 1}", d, (512, 1024, 4096, 8192)[i], 0xFF, "hbm" if i == 0 else "tensor", seed_shift=i)
              for i, d in enumerate(st)]
    for name, ls, n, macs, io_bytes in (
            ("config5a_conv_layers0-3", [net_layers[i] for i in range(4)], 1024, sum(W.nonzero_macs(configs.net_layer(i)) for i in range(4)), None),
            ("config5b_stack", stages, 512, sum(d.macs_per_image for d in st), None),
            ("eight_layers_net", [net_layers[i] for i in range(8)], 1024, sum(W.nonzero_macs(configs.net_layer(i)) for i in range(8)), None)):
        net = Net(ls)
        run_one(name, net.run_device, net.in_bytes, net.out_bytes, n, 0xFF,
                {"bound": "tensor", "ops_per_image": 2.0 * macs, "bytes_per_image": net.in_bytes + net.out_bytes},
                extra={"launches_per_pass": len(ls), "GMAC_per_image_nonzero": macs / 1e9,
                       "layer_io_bytes_per_image": sum(l.in_bytes + l.out_bytes for l in ls)})
        net.close()
    return out


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    from simple_image_compression_network_b200 import configs, synth, workloads as W
    from simple_image_compression_network_b200.shard import weak_range
    from simple_image_compression_network_b200.layer import ConvLayer, synth_fill

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    _bind_to_gpu_numa_node(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    d = configs.net_layer(1)  # CONV_1
    inp = configs.synthetic_params(d)  # weights / bias (the input is generated on the device)
    layer = ConvLayer(d, inp["weights"], bias=inp["bias"], device=local)
    n_img = args.images
    free, _ = torch.cuda.mem_get_info()
    need = n_img * (layer.in_bytes + layer.out_bytes)
    if need > free * 0.9:
        n_img = max(1, int(free * 0.9 // (layer.in_bytes + layer.out_bytes)))
    x = torch.empty(n_img * layer.in_bytes, dtype=torch.uint8, device="cuda")
    y = torch.empty(n_img * layer.out_bytes, dtype=torch.uint8, device="cuda")
    # this rank's images: global indices [rank*n_img, (rank+1)*n_img); lanes 0..127 (post-ReLU range of the previous layer)
    first_img, _ = weak_range(n_img, rank)
    synth_fill(x.data_ptr(), x.numel(), synth.SEED_INPUT, 0x7F, offset=first_img * layer.in_bytes)
    torch.cuda.synchronize()

    sampler = ClockSampler(local)
    sampler.start()  # nvidia-smi takes a while to produce its first line: started before the warm-up, windows select the samples
    h = Harness(torch, dist, world, local, sampler)
    sh = h.stream.cuda_stream

    def step():
        layer.run_device(x.data_ptr(), y.data_ptr(), n_img, sh)

    l0 = layer.launches
    ms, ms_max, clocks = h.time_steps(step, args.steps, args.warmup)
    launches = layer.launches - l0 - max(args.warmup, 3)
    value = world * n_img * args.steps / (ms_max / 1000.0)

    # ---- parity of the TIMED batch (untimed): first and last image of this rank's 4096 against the oracle
    checked = []
    if not args.no_parity_check:
        from oracle import oracle
        oracle.set_threads(min(16, os.cpu_count() or 1))
        for i in sorted({0, n_img - 1}):
            xi = synth.lanes(synth.SEED_INPUT, (layer.in_bytes,), 8, mask=0x7F, offset=(first_img + i) * layer.in_bytes).astype(np.uint8)
            want = oracle.run_layer(d, xi, inp["weights"], None, inp["bias"])
            got = y[i * layer.out_bytes:(i + 1) * layer.out_bytes].cpu().numpy()
            if not np.array_equal(got, want):
                raise SystemExit(f"bench: image {first_img + i} of the timed batch differs from the oracle ({int((got != want).sum())} bytes)")
            checked.append(first_img + i)

    # ---- e2e: host buffers (pinned) -> C-ABI host call -> host buffers, copies inside the timed region
    e_img = min(args.e2e_images, n_img)
    hx = torch.empty(e_img * layer.in_bytes, dtype=torch.uint8, pin_memory=True)
    hy = torch.empty(e_img * layer.out_bytes, dtype=torch.uint8, pin_memory=True)
    hx.copy_(x[: e_img * layer.in_bytes])
    torch.cuda.synchronize()
    e_steps = max(2, min(args.steps, 5))
    layer.run_raw(hx.data_ptr(), hy.data_ptr(), e_img)  # warm-up (allocates the staging slots)
    h.barrier()
    t0 = time.perf_counter()
    for _ in range(e_steps):
        layer.run_raw(hx.data_ptr(), hy.data_ptr(), e_img)
    e_s = time.perf_counter() - t0
    e2e_value = world * e_img * e_steps / h.max_over_ranks(e_s)
    if not torch.equal(hy.cuda(), y[: e_img * layer.out_bytes]):
        raise SystemExit("bench: the host-buffer call and the device-resident call disagree")
    # the denominator of e2e: the same bytes moved with NO kernel -- H2D of the inputs and D2H of the outputs of one step, on two
    # streams at once, all ranks concurrently (what the box's PCIe / host memory sustains for this rank layout)
    s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()
    dx, dy = x[: e_img * layer.in_bytes], y[: e_img * layer.out_bytes]

    def copies():
        with torch.cuda.stream(s_up):
            dx.copy_(hx, non_blocking=True)
        with torch.cuda.stream(s_dn):
            hy.copy_(dy, non_blocking=True)
    copies()
    h.barrier()
    t0 = time.perf_counter()
    for _ in range(e_steps):
        copies()
    torch.cuda.synchronize()
    copy_value = world * e_img * e_steps / h.max_over_ranks(time.perf_counter() - t0)

    # ---- roofline of the dominant kernel (the only kernel in the step)
    peaks, src = _peaks()
    kernel_ms = ms / args.steps  # one launch per step, back to back on one stream
    ops = 2.0 * d.macs_per_image * n_img
    achieved = ops / (kernel_ms / 1000.0) / 1e12
    proxy = 2.0 * float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1400.0)))
    traffic, traffic_src = None, None
    try:  # DRAM bytes per image of this kernel from this round's committed `ncu --set full` capture
        t = json.load(open(os.path.join(ROOT, "profiles", "conv1_dram_traffic.json")))
        traffic, traffic_src = float(t["dram_bytes_per_image"]) * n_img, t["source"]
    except Exception:
        pass
    roofline = {"bound": "tensor", "achieved": achieved, "peak": W.INT8_SPEC_TOPS, "unit": "TFLOP/s", "frac": achieved / W.INT8_SPEC_TOPS,
                "peak_source": "dense INT8 tensor spec of B200, 4500 TOP/s (MEASURED_PEAKS.json has no INT8 figure); ops are int8 MACs x 2",
                "traffic": traffic, "traffic_source": traffic_src, "kernel": layer.plan, "kernel_ms": kernel_ms,
                "frac_of_measured_int8_mma_only_peak_4380_TOPs": achieved / W.INT8_MMA_ONLY_TOPS,
                "frac_of_2x_bf16_sustained_proxy": achieved / proxy,
                "proxy_source": f"2 x bf16_tflops_sustained of MEASURED_PEAKS.json ({src}) = {proxy:.0f}: what cuBLAS bf16 sustains under the same power cap, x2 for int8",
                "algorithmic_bytes_per_image": layer.in_bytes + layer.out_bytes,
                "hbm_GBs_at_this_rate": (layer.in_bytes + layer.out_bytes) * n_img / (kernel_ms / 1000.0) / 1e9}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic", "config": CONFIG,
            "run": {"images_per_gpu": n_img, "engine": layer.engine,
                    "l2": f"inputs ({n_img * layer.in_bytes / 1e9:.1f} GB per GPU) are larger than L2; no flush needed"},
            "clocks": clocks, "gpu_launches": int(launches), "parity_checked_images": checked,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": e_img * layer.in_bytes,
                    "d2h_bytes_per_step": e_img * layer.out_bytes, "images_per_step": e_img, "steps": e_steps,
                    "copy_only_ceiling": copy_value, "frac_of_copy_only_ceiling": e2e_value / copy_value,
                    "ceiling_is": "the step's H2D + D2H bytes with no kernel, both directions at once, all ranks concurrently (pinned memory)"},
            "roofline": roofline}
    del x, y, hx, hy
    torch.cuda.empty_cache()

    if not args.no_configs:
        line["configs"] = run_configs(h, peaks, args.config_scale)
    sampler.stop()

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        import multiprocessing as mp
        cores = max(1, min(os.cpu_count() or 1, 64))
        with mp.get_context("spawn").Pool(cores) as pool:
            _, band_wall, kind = cpu_reference_step(pool, cores, "c2d_L1band")  # warm (library load) + calibration
            case = "c2d_L1" if 8.0 * band_wall <= 30.0 else "c2d_L1band"
            v, wall, kind = cpu_reference_step(pool, cores, case)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": kind,
                                "sample": _ref_sample_text(case, cores, kind) + f"; one step, {wall:.1f} s wall"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--images", type=int, default=IMAGES_PER_GPU, help="images per GPU per step")
    ap.add_argument("--e2e-images", type=int, default=E2E_IMAGES)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the per-config block (profiling runs)")
    ap.add_argument("--no-parity-check", action="store_true", help="skip the oracle check of the timed batch (profiling runs)")
    ap.add_argument("--config-scale", type=float, default=1.0, help="scales the image counts of the `configs` block")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
