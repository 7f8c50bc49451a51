"""Sustained (power-capped) throughput of one layer of eight_layers_net on the EXPERIMENT build, for perf decomposition:
    FCB_U2_DEBUG=<mask> python tools/sustained.py L1 [seconds] [images]
mask bits (fcb_umma2.cu): 1 no weight TMA, 2 no plane TMA, 4 no stores, 8 no epilogue, 128 no MMA.  Prints img/s and the median SM clock."""
import json, os, subprocess, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from simple_image_compression_network_b200 import _lib, configs, synth
_lib.set_default(_lib.load(_lib.EXP_LIB_PATH))
from simple_image_compression_network_b200.layer import ConvLayer, synth_fill
name = sys.argv[1] if len(sys.argv) > 1 else "L1"
secs = float(sys.argv[2]) if len(sys.argv) > 2 else 1.5
n = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
d = configs.net_layer(int(name[1:]))
prm = configs.synthetic_params(d)
L = ConvLayer(d, prm["weights"], bias=prm["bias"])
x = torch.empty(n * L.in_bytes, dtype=torch.uint8, device="cuda"); y = torch.empty(n * L.out_bytes, dtype=torch.uint8, device="cuda")
synth_fill(x.data_ptr(), x.numel(), synth.SEED_INPUT, 0x7F)
st = torch.cuda.current_stream().cuda_stream
for _ in range(3): L.run_device(x.data_ptr(), y.data_ptr(), n, st)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); L.run_device(x.data_ptr(), y.data_ptr(), n, st); e1.record(); torch.cuda.synchronize()
steps = max(3, int(secs * 1000 / e0.elapsed_time(e1)))
clk = []
p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE, text=True)
threading.Thread(target=lambda: [clk.append((time.monotonic(), l)) for l in p.stdout], daemon=True).start()
time.sleep(0.3)
for _ in range(steps): L.run_device(x.data_ptr(), y.data_ptr(), n, st)  # reach the sustained clock
torch.cuda.synchronize()
t0 = time.monotonic(); e0.record()
for _ in range(steps): L.run_device(x.data_ptr(), y.data_ptr(), n, st)
e1.record(); torch.cuda.synchronize(); t1 = time.monotonic()
p.terminate()
mhz = sorted(float(l.split(",")[0]) for t, l in clk if t0 <= t <= t1)
pw = sorted(float(l.split(",")[1]) for t, l in clk if t0 <= t <= t1)
ms = e0.elapsed_time(e1) / steps
print(json.dumps(dict(layer=name, debug=os.environ.get("FCB_U2_DEBUG", "0"), images=n, ms=round(ms, 3), img_s=round(n / ms * 1e3), sm_mhz=mhz[len(mhz) // 2] if mhz else None,
                      watts=pw[len(pw) // 2] if pw else None, clk_per_image=round(ms * 1e-3 * (mhz[len(mhz) // 2] if mhz else 0) * 1e6 / n), plan=L.plan)), flush=True)
