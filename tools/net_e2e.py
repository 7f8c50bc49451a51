"""eight_layers_net through the host-buffer entry point (fcb_net_run: H2D | 8 layers | D2H pipelined over chunks), images/s.
    python tools/net_e2e.py [n_images] [images_per_host_chunk]"""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from simple_image_compression_network_b200 import configs
from simple_image_compression_network_b200.layer import ConvLayer, Net
layers = []
for i in range(8):
    d = configs.net_layer(i); prm = configs.synthetic_params(d)
    layers.append(ConvLayer(d, prm["weights"], thresholds=prm["thresholds"], bias=prm["bias"]))
net = Net(layers)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 0
net.set_host_chunk(chunk)
hx = torch.empty(n * net.in_bytes, dtype=torch.uint8, pin_memory=True); hy = torch.empty(n * net.out_bytes, dtype=torch.uint8, pin_memory=True)
hx.random_(0, 256)
net.run_raw(hx.data_ptr(), hy.data_ptr(), n)
t0 = time.perf_counter()
for _ in range(3): net.run_raw(hx.data_ptr(), hy.data_ptr(), n)
dt = (time.perf_counter() - t0) / 3
print(chunk or "default", n, round(n / dt, 1), "img/s")
