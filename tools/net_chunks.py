"""eight_layers_net device-resident (fcb_net_run_device) against the images per pass of the layer chain: does keeping the
intermediate streams of a few images inside the 126 MB L2 beat long launches?   python tools/net_chunks.py [images] [chunks...]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from simple_image_compression_network_b200 import configs, synth
from simple_image_compression_network_b200.layer import ConvLayer, Net, synth_fill
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
chunks = [int(a) for a in sys.argv[2:]] or [0, 64, 32, 16, 8, 4, 2]
layers = []
for i in range(8):
    d = configs.net_layer(i); prm = configs.synthetic_params(d)
    layers.append(ConvLayer(d, prm["weights"], bias=prm["bias"]))
net = Net(layers)
x = torch.empty(n * net.in_bytes, dtype=torch.uint8, device="cuda"); y = torch.empty(n * net.out_bytes, dtype=torch.uint8, device="cuda")
synth_fill(x.data_ptr(), x.numel(), synth.SEED_INPUT, 0xFF)
st = torch.cuda.current_stream().cuda_stream
ref = None
for c in chunks:
    net.set_device_chunk(c)
    for _ in range(2): net.run_device(x.data_ptr(), y.data_ptr(), n, st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): net.run_device(x.data_ptr(), y.data_ptr(), n, st)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    if ref is None: ref = y.clone()
    print(json.dumps(dict(images=n, images_per_pass=c or "default(85)", ms=round(ms, 3), img_s=round(n / ms * 1e3), same=bool(torch.equal(ref, y)))), flush=True)
