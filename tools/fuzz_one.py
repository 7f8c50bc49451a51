"""Run ONE layer of the seeded fuzz of tests/test_gpu_parity.py in its own process (a CUDA fault kills the context): python tools/fuzz_one.py SEED INDEX [imad]"""
import dataclasses, sys, numpy as np
sys.path.insert(0, ".")
from oracle import cases, oracle
from tests.test_gpu_parity import _random_descs, _layer
import os
if os.environ.get("FCB_EXP"):
    from simple_image_compression_network_b200 import _lib
    _lib.set_default(_lib.load(_lib.EXP_LIB_PATH))
if os.environ.get("FCB_LIB"):
    from simple_image_compression_network_b200 import _lib
    _lib.set_default(_lib.load(os.environ["FCB_LIB"]))
seed, i = int(sys.argv[1]), int(sys.argv[2])
d = _random_descs(seed, 40)[i]
if len(sys.argv) > 3: d = dataclasses.replace(d, engine_hint=1)
reps = 1 + (i % 3)
try:
    inp = cases.make_inputs(d, seed_shift=seed + i, num_reps=reps, relu_range=bool(i % 2))
    L = _layer(d, inp)
except Exception as e:
    print(i, "rejected", str(e)[:80]); sys.exit(0)
print(i, L.engine, L.plan, flush=True)
got = L.run(inp["in_words"], reps)
want = oracle.run_layer(d, inp["in_words"], inp["weights"], inp["thresholds"], inp["bias"], num_reps=reps)
print(i, "OK" if np.array_equal(got, want) else "MISMATCH")
