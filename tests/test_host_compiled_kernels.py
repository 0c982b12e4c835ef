"""CPU: the PRODUCT's encode / decode / layer-out device functions (csrc/slfp_common.cuh), compiled
for the host by nvcc (tests/host_check.cu), swept against the oracle: every float32 mantissa at a
set of exponents, all four formats, bit-exact.  No GPU needed."""
import ctypes

import numpy as np
import pytest

from conftest import bits


def _hq(lib, x, fmt, k=1.0, zz=0):
    x = np.ascontiguousarray(x, np.float32)
    c = np.empty(x.shape, np.uint8)
    q = np.empty_like(x)
    lib.hostcheck_quantize(x.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(x.size), ctypes.c_float(k),
                           ctypes.c_int(fmt), ctypes.c_int(zz), c.ctypes.data_as(ctypes.c_void_p),
                           q.ctypes.data_as(ctypes.c_void_p))
    return c, q


@pytest.mark.parametrize("fmt", [0, 1, 2, 3])
def test_exhaustive_mantissa_sweep(hostcheck, orc, fmt):
    mant = np.arange(1 << 23, dtype=np.uint32)
    exps = [-5, -4, -3, 0, 3, 4] + ([-127, -126, 7, 8] if fmt == 3 else [])
    for e in exps:
        x = (mant | np.uint32(max(e + 127, 0) << 23)).view(np.float32)
        if e == 0:
            x = -x
        c, q = _hq(hostcheck, x, fmt)
        oc, oq = orc.quantize(x, fmt)
        assert (bits(q) == bits(oq)).all(), (fmt, e)
        if fmt < 3:
            assert (c == oc).all(), (fmt, e)


def test_edge_values_and_prescale(hostcheck, orc, g_quant):
    x = g_quant["x"]
    for fmt in range(4):
        for k in (1.0, 0.17032258, 3.0):
            c, q = _hq(hostcheck, x, fmt, k)
            oc, oq = orc.quantize(x, fmt, kdiv=k)
            ok = (bits(q) == bits(oq)) | (np.isnan(q) & np.isnan(oq))
            assert ok.all(), (fmt, k)
    # zero_is_zero switch of the layer-out quantizer
    _, q = _hq(hostcheck, np.array([0.0, -0.0, 1.0], np.float32), 3, zz=1)
    assert q.tolist() == [0.0, 0.0, 1.0]
    _, q = _hq(hostcheck, np.array([0.0], np.float32), 3, zz=0)
    assert np.isnan(q[0])


def test_decode_all_codes(hostcheck, orc):
    codes = np.arange(256, dtype=np.uint8)
    for fmt in (0, 1):
        out = np.empty(256, np.float32)
        hostcheck.hostcheck_decode(codes.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(256), ctypes.c_int(fmt),
                                   out.ctypes.data_as(ctypes.c_void_p))
        ref = orc.decode(codes, fmt)
        assert ((bits(out) == bits(ref)) | (np.isnan(out) & np.isnan(ref))).all()
