"""Extended seeded fuzz of the CUDA library against the oracle: the shape generator of tests/test_gpu_parity.py with more seeds
(and larger batches) than the test suite runs.  Test infrastructure (uses oracle/).
    python tests/fuzz_extended.py [first_seed] [n_seeds] [cases_per_seed]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))  # (this directory)
import numpy as np
import test_gpu_parity as T
from oracle import cases, oracle
from simple_image_compression_network_b200._lib import FcbError

first, n, per = (int(a) for a in (sys.argv[1:] + ["1000", "10", "40"][len(sys.argv) - 1:]))
ran = bad = 0
plans = {}
for seed in range(first, first + n):
    for i, d in enumerate(T._random_descs(seed, per)):
        reps = 1 + (i % 4)
        try:
            inp = cases.make_inputs(d, seed_shift=seed + i, num_reps=reps, relu_range=bool(i % 2))
            L = T._layer(d, inp)
        except (FcbError, ValueError, AssertionError):
            continue
        got = L.run(inp["in_words"], reps)
        want = oracle.run_layer(d, inp["in_words"], inp["weights"], inp["thresholds"], inp["bias"], num_reps=reps)
        ran += 1
        key = L.plan.split(" ")[0] + ":" + L.engine
        plans[key] = plans.get(key, 0) + 1
        if not np.array_equal(got, want):
            bad += 1
            print(f"MISMATCH seed {seed} case {i} reps {reps} {d} [{L.engine}: {L.plan}]", flush=True)
    for i, d in enumerate(T._direct_descs(seed, per)):  # the universal engine's three inner loops
        inp = cases.make_inputs(d, seed_shift=seed * 100 + i, num_reps=2)
        L = T._layer(d, inp)
        got = L.run(inp["in_words"], 2)
        want = oracle.run_layer(d, inp["in_words"], inp["weights"], inp["thresholds"], inp["bias"], num_reps=2)
        ran += 1
        key = "direct:" + ("IDP.4A" if "IDP.4A" in L.plan else "IDP.2A" if "IDP.2A" in L.plan else "IMAD")
        plans[key] = plans.get(key, 0) + 1
        if not np.array_equal(got, want):
            bad += 1
            print(f"MISMATCH seed {seed} direct case {i} {d} [{L.engine}: {L.plan}]", flush=True)
print(f"fuzz: {ran} layers run, {bad} mismatches; plans: {plans}")
sys.exit(1 if bad else 0)
