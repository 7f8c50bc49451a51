import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import cases, oracle
from simple_image_compression_network_b200.layer import ConvLayer
name = sys.argv[1] if len(sys.argv) > 1 else "c2d_a"
d = cases.CASES[name]
inp = cases.make_inputs(d)
L = ConvLayer(d, inp["weights"], thresholds=inp["thresholds"], bias=inp["bias"], device=0)
print(L.engine, L.plan, flush=True)
t0 = time.time()
try:
    got = L.run(inp["in_words"])
    print("ran in", time.time() - t0, flush=True)
    want = oracle.run_layer(d, inp["in_words"], inp["weights"], inp["thresholds"], inp["bias"])
    bad = np.flatnonzero(got != want)
    print("mismatches", bad.size, "of", got.size, bad[:16], got[bad[:16]], want[bad[:16]])
except Exception as e:
    print("FAILED after", time.time() - t0, str(e)[:300], flush=True)
