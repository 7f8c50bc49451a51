"""The C-ABI library loads on a CPU-only box, exports every symbol include/finnconv_b200.h declares,
validates descriptors like the reference's CASSERTs, and refuses to compute without a GPU."""
import ctypes
import dataclasses
import os
import re
import subprocess

import numpy as np
import pytest

from oracle import cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "finnconv_b200.h")).read()
    return sorted(set(re.findall(r"FCB_API\s+[\w\s\*]+?\b(fcb_\w+)\s*\(", text)))


def test_header_symbols_exported(fcb_lib):
    names = _declared()
    assert len(names) >= 17, names
    for n in names:
        assert hasattr(fcb_lib, n), f"{n} declared in the header but not exported"


def test_version_and_word_bytes(fcb_lib):
    assert b"sm_100a" in fcb_lib.fcb_version()
    from simple_image_compression_network_b200 import pack
    for bits in (1, 8, 12, 24, 32, 48, 64, 96, 1024, 1536):
        assert fcb_lib.fcb_word_bytes(bits) == pack.word_bytes(bits)


def _query(lib, desc):
    c = desc.to_c()
    sizes = [ctypes.c_size_t() for _ in range(5)]
    rc = lib.fcb_layer_query(ctypes.byref(c), *[ctypes.byref(s) for s in sizes])
    return rc, [s.value for s in sizes]


@pytest.mark.parametrize("name", list(cases.CASES))
def test_query_sizes_match_oracle(name, fcb_lib, oracle_mod):
    d = cases.CASES[name]
    rc, (ib, ob, wb, tb, bb) = _query(fcb_lib, d)
    assert rc == 0, fcb_lib.fcb_last_error()
    s = oracle_mod.query(d)
    assert (ib, ob, wb, tb, bb) == (s.in_bytes_per_image, s.out_bytes_per_image, s.weight_bytes, s.threshold_bytes, s.bias_bytes)


def test_query_rejects_like_the_reference(fcb_lib):
    d = cases.CASES["c2d_a"]
    assert _query(fcb_lib, dataclasses.replace(d, simd=3))[0] == -2   # IFMChannels % SIMD (slidingwindow.h:1259)
    assert b"SIMD" in fcb_lib.fcb_last_error()
    assert _query(fcb_lib, dataclasses.replace(d, pe=4))[0] == -2     # OFM % PE (streamtools.h:505)
    c = d.to_c()
    c.ofm_x += 1                                                      # inconsistent OFMDim
    assert fcb_lib.fcb_layer_query(ctypes.byref(c), None, None, None, None, None) == -2
    c = d.to_c()
    c.struct_size = 8
    assert fcb_lib.fcb_layer_query(ctypes.byref(c), None, None, None, None, None) == -1
    dc = cases.CASES["dc_a"]
    assert _query(fcb_lib, dataclasses.replace(dc, kernel_x=3, kernel_y=3))[0] == -2  # deconv522 is k5 s2 p2
    assert _query(fcb_lib, dataclasses.replace(d, in_bits=3))[0] == 0                # any ap_uint<N> lane width (interpret.hpp:191-217)
    assert _query(fcb_lib, dataclasses.replace(d, in_bits=17))[0] == -3              # valid upstream, unsupported here
    assert _query(fcb_lib, dataclasses.replace(cases.CASES["th_b"], pool=3))[0] == -2  # 16 x 12 map: ImgDim % PoolDim (maxpool.h:140)


def test_no_cpu_fallback(fcb_lib):
    """Without a usable sm_100 device the library must refuse, not compute on the host."""
    if fcb_lib.fcb_device_count() > 0:
        pytest.skip("a B200 is present")
    from simple_image_compression_network_b200._lib import FcbError
    from simple_image_compression_network_b200.layer import ConvLayer
    d = cases.CASES["c2d_a"]
    inp = cases.make_inputs(d)
    with pytest.raises(FcbError) as e:
        ConvLayer(d, inp["weights"], bias=inp["bias"])
    assert e.value.code == -4


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under the package may reference it."""
    pkg = os.path.join(ROOT, "simple_image_compression_network_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "finn_oracle" not in text and "import oracle" not in text and "from oracle" not in text, f


def test_product_library_has_no_experiment_switches(fcb_lib):
    """No environment variable may change the kernel a production caller gets: the experiment switches (FCB_U2_*, FCB_FORCE_ENGINE,
    ...), the first-generation kernel and the cross-check instantiations exist only in tools/libfinnconv_exp.so (-DFCB_EXPERIMENT)."""
    from simple_image_compression_network_b200 import _lib
    blob = open(_lib.LIB_PATH, "rb").read()
    for needle in (b"FCB_U2_", b"FCB_FORCE_ENGINE", b"FCB_XNOR_ENGINE", b"FCB_THIN", b"FCB_NET_CHUNK", b"FCB_UMMA_V1", b"umma_conv_kernel"):
        assert needle not in blob, needle
    syms = subprocess.run(["nm", "-D", "--undefined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "getenv" not in syms


def test_engine_hint_and_padding_fields_validated(fcb_lib):
    d = cases.CASES["c2d_a"]
    assert _query(fcb_lib, dataclasses.replace(d, engine_hint=7))[0] == -1
    assert _query(fcb_lib, dataclasses.replace(d, pad_style=2, pad_x_total=3, pad_y_total=3))[0] == -1      # pad must be 0 with a style
    assert _query(fcb_lib, dataclasses.replace(d, pad=0, pad_x_total=4))[0] == -1                           # totals need a style
    ok = dataclasses.replace(d, pad=0, pad_style=2, pad_x_total=4, pad_y_total=4)                           # == pad 2 on every side
    assert (ok.ofm_x, ok.ofm_y) == (d.ofm_x, d.ofm_y)
    assert _query(fcb_lib, ok)[0] == 0, fcb_lib.fcb_last_error()


def test_v1_descriptor_size_still_accepted(fcb_lib):
    """A client compiled against ABI 0.1 passes fcb_layer_desc without the appended dilation fields (FCB_LAYER_DESC_SIZE_V1)."""
    d = cases.CASES["c2d_a"]
    c = d.to_c()
    c.struct_size = ctypes.sizeof(c) - 8
    c.dilation_x = 77  # garbage past the caller's struct must not be read
    sizes = [ctypes.c_size_t() for _ in range(5)]
    assert fcb_lib.fcb_layer_query(ctypes.byref(c), *[ctypes.byref(s) for s in sizes]) == 0, fcb_lib.fcb_last_error()
    c.struct_size = ctypes.sizeof(c) - 4
    assert fcb_lib.fcb_layer_query(ctypes.byref(c), *[ctypes.byref(s) for s in sizes]) == -1
