"""Run ONE parity case on the GPU and print where it differs from the oracle (pixels / channels of the mismatches).
    python tests/check_case.py <name>      name = a key of oracle/cases.py CASES or of tests/test_gpu_parity.py::_thin_cases()"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from oracle import cases, oracle
from simple_image_compression_network_b200.layer import ConvLayer

name = sys.argv[1] if len(sys.argv) > 1 else "c2d_a"
reps = 1
if name in cases.CASES:
    d = cases.CASES[name]
else:
    import test_gpu_parity as T
    d, reps = T._thin_cases()[name]
inp = cases.make_inputs(d, seed_shift=31, num_reps=reps)
L = ConvLayer(d, inp["weights"], thresholds=inp["thresholds"], bias=inp["bias"], device=0)
print(L.engine, L.plan, flush=True)
t0 = time.time()
try:
    got = L.run(inp["in_words"], reps)
    want = oracle.run_layer(d, inp["in_words"], inp["weights"], inp["thresholds"], inp["bias"], num_reps=reps)
    bad = np.flatnonzero(got != want)
    print("mismatches", bad.size, "of", got.size, bad[:12], got[bad[:12]], want[bad[:12]])
    if bad.size:
        wb = max(1, got.size // (d.ofm_x * d.ofm_y * reps))
        px = bad // wb
        print("bad pixels: count", np.unique(px).size, "first", np.unique(px)[:20], "last", np.unique(px)[-5:])
        print("bad byte-in-word:", np.unique(bad % wb)[:40])
except Exception as e:
    print("FAILED after", time.time() - t0, str(e)[:300], flush=True)
