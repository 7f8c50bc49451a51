"""oracle/gen_golden.py -- TEST INFRASTRUCTURE ONLY.

Generates tests/golden/*.npz by running the REFERENCE's own templates (oracle/_ref/libref_layers.so,
built from /root/reference by oracle/Makefile) on the seeded synthetic tensors of oracle/cases.py, and
the fixture-weight dumps + eight_layers_net outputs from oracle/_ref/libref_net.so.  Only runs where
/root/reference exists; the vectors travel, the reference does not.

    python -m oracle.gen_golden            # all layer cases
    python -m oracle.gen_golden --net      # also the 8-layer net fixtures (minutes)
    python -m oracle.gen_golden --streams-only   # only the parameter-stream fixtures
"""
from __future__ import annotations

import ctypes
import hashlib
import os
import sys
import time

import numpy as np

from oracle import cases, oracle

GOLD = os.path.join(os.path.dirname(oracle.HERE), "tests", "golden")
FULL_LIMIT = 64 * 1024  # outputs up to this size are stored whole, larger ones as digest + head


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def gen_layers(names):
    os.makedirs(GOLD, exist_ok=True)
    for name in names:
        d = cases.CASES[name]
        inp = cases.make_inputs(d)
        s = oracle.query(d)
        t0 = time.time()
        ref, secs = oracle.ref_run(name, inp["in_words"], inp["weights"], cases.third_image(inp), s.out_bytes_per_image)
        mine = oracle.run_layer(d, inp["in_words"], inp["weights"], inp["thresholds"], inp["bias"])
        ok = np.array_equal(ref, mine)
        rec = dict(out_sha=sha(ref), in_sha=sha(inp["in_words"]), w_sha=sha(inp["weights"]), out_bytes=ref.size,
                   ref_seconds=secs, head=ref[:4096].copy())
        if ref.size <= FULL_LIMIT:
            rec["out"] = ref
        np.savez_compressed(os.path.join(GOLD, f"layer_{name}.npz"), **rec)
        print(f"{name:12s} ref={secs:7.2f}s restatement_matches={ok} bytes={ref.size} wall={time.time()-t0:.1f}s", flush=True)
        if not ok:
            raise SystemExit(f"restatement differs from the reference on {name}")


STREAM_CASES = ("th_a", "th_b", "th_c")  # STREAM_CASES of ref_layers.cpp


def gen_param_streams():
    """Streamed-weights form (GenParamStream -> Matrix_Vector_Activate_Stream_Batch): the parameter-stream image of each case's
    weights as the reference emits it, and proof that the streamed MVAU gives the stored layer output."""
    for name in STREAM_CASES:
        d = cases.CASES[name]
        inp = cases.make_inputs(d)
        s = oracle.query(d)
        pbytes = s.sf * s.nf * oracle.lib().fo_word_bytes(d.simd * d.pe * d.w_bits)
        out, pw = oracle.ref_stream_run(name, inp["in_words"], inp["weights"], inp["thresholds"], s.out_bytes_per_image, pbytes)
        gold = np.load(os.path.join(GOLD, f"layer_{name}.npz"))
        same_out = sha(out) == str(gold["out_sha"])
        mine = oracle.gen_param_stream(d, inp["weights"])
        ok = np.array_equal(mine, pw)
        np.savez_compressed(os.path.join(GOLD, f"param_stream_{name}.npz"), param_words=pw, w_sha=sha(inp["weights"]), out_sha=sha(out))
        print(f"param_stream_{name}: {pw.size} bytes, streamed MVAU == static MVAU: {same_out}, restatement_matches={ok}", flush=True)
        if not (ok and same_out):
            raise SystemExit(f"parameter-stream mismatch on {name}")


def gen_net():
    lib = ctypes.CDLL(os.path.join(oracle.HERE, "_ref", "libref_net.so"))
    lib.ref_dump_weights.restype = ctypes.c_long
    lib.ref_dump_bias.restype = ctypes.c_long
    lib.ref_dump_weights.argtypes = [ctypes.c_int, ctypes.c_void_p]
    lib.ref_dump_bias.argtypes = [ctypes.c_int, ctypes.c_void_p]
    params = {}
    for l in range(8):
        nw, nb = lib.ref_dump_weights(l, None), lib.ref_dump_bias(l, None)
        w, b = np.zeros(nw, np.uint8), np.zeros(nb, np.uint8)
        lib.ref_dump_weights(l, w.ctypes.data_as(ctypes.c_void_p))
        lib.ref_dump_bias(l, b.ctypes.data_as(ctypes.c_void_p))
        params[f"w{l}"], params[f"b{l}"] = w, b
    np.savez_compressed(os.path.join(GOLD, "params_nonsquare.npz"), **params)
    print("params_nonsquare.npz:", os.path.getsize(os.path.join(GOLD, "params_nonsquare.npz")), "bytes", flush=True)
    # eight_layers_net on (a) the testbench's constant-1 image (conv3_nonsquare_tb.cpp:801), (b) a seeded random image
    from simple_image_compression_network_b200 import pack, synth
    fn = lib.ref_eight_layers_net
    fn.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(ctypes.c_double)]
    for tag, img in (("ones", np.ones((1, 512, 768, 3), np.int64)), ("rand", synth.lanes(synth.SEED_INPUT, (1, 512, 768, 3), 8))):
        words = pack.pack_stream(img, 8)
        out = np.zeros(768 * 512 * 4, np.uint8)
        secs = ctypes.c_double(0)
        rc = fn(words.ctypes.data_as(ctypes.c_void_p), out.ctypes.data_as(ctypes.c_void_p), ctypes.byref(secs))
        assert rc == 0, rc
        np.savez_compressed(os.path.join(GOLD, f"net8_{tag}.npz"), out=out, in_sha=sha(words), out_sha=sha(out), ref_seconds=secs.value)
        print(f"net8_{tag}: {secs.value:.1f}s, {os.path.getsize(os.path.join(GOLD, f'net8_{tag}.npz'))} bytes", flush=True)


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    if "--streams-only" not in sys.argv:
        gen_layers(args or list(cases.CASES))
    gen_param_streams()
    if "--net" in sys.argv:
        gen_net()
