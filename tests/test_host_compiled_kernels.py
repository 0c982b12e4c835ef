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


def _relu_codes(lib, q, sfp33):
    q = np.ascontiguousarray(q, np.float32)
    c, c16 = np.empty(q.shape, np.uint8), np.empty(q.shape, np.uint8)
    v = np.empty_like(q)
    lib.hostcheck_relu_codes(q.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(q.size), ctypes.c_int(sfp33),
                             c.ctypes.data_as(ctypes.c_void_p), v.ctypes.data_as(ctypes.c_void_p),
                             c16.ctypes.data_as(ctypes.c_void_p))
    return c, v, c16


@pytest.mark.parametrize("sfp33", [0, 1])
def test_post_relu_codes_match_the_reference_quantizer(hostcheck, orc, sfp33):
    """The fused pipeline's unsigned post-ReLU codes (2-instruction encoder) against the reference quantizer:
    for every float32 mantissa in every octave the decoded value, as the float16 tensor-core operand, equals the
    reference's fake-quant value - except exact ties of the mantissa rounding, which go up instead of to even
    (documented in include/slfp_b200.h) - and negative inputs give code 0 (the ReLU)."""
    fmt = 0 if sfp33 else 1
    mant = np.arange(1 << 23, dtype=np.uint32)
    tie_mask = np.uint32((1 << (20 if sfp33 else 19)) - 1)       # bits below the kept mantissa
    tie_val = np.uint32(1 << (19 if sfp33 else 18))              # exactly one half
    for e in (-9, -5, -4, -3, -1, 0, 2, 3, 4, 9):
        q = (mant | np.uint32((e + 127) << 23)).view(np.float32)
        c, v, c16 = _relu_codes(hostcheck, q, sfp33)
        _, want = orc.quantize(q, fmt)
        same = v.astype(np.float16) == want.astype(np.float16)
        ties = (mant & tie_mask) == tie_val
        assert same[~ties].all(), (sfp33, e, int((~same & ~ties).sum()))
        # a tie rounds up: one grid step above (or equal to) the reference's round-half-even value
        assert (v[ties].astype(np.float16) >= want[ties].astype(np.float16)).all()
        # decode agrees with the oracle's numpy restatement of the code table
        assert (bits(v) == bits(orc.decode_relu(c, bool(sfp33)))).all()
        # the epilogue's clamp(q/16, 0, 1) form yields the same code (or an alias of the top value)
        assert (orc.decode_relu(c16, bool(sfp33)) == v).all(), (sfp33, e)
    neg = -np.abs(np.random.default_rng(0).standard_normal(1000).astype(np.float32))
    c, v, c16 = _relu_codes(hostcheck, np.concatenate([neg, [0.0, -0.0]]).astype(np.float32), sfp33)
    assert (c == 0).all() and (v == 0).all() and (c16 == 0).all()
    # all 256 codes decode monotonically (max-pooling on codes commutes with decoding)
    allv = orc.decode_relu(np.arange(256, dtype=np.uint8), bool(sfp33))
    assert (np.diff(allv) >= 0).all()


@pytest.mark.parametrize("fmt", [0, 1, 2])
def test_fast_encoder_equals_reference_encoder(hostcheck, fmt):
    """encode_inrange<FMT> (the stand-alone quantizer's common path) == encode<FMT> for every float32 mantissa, both
    signs, exponents from denormals to beyond the saturation thresholds, and Inf (only NaN takes the general path)."""
    hostcheck.hostcheck_encode_inrange_mismatches.restype = ctypes.c_size_t
    mant = np.arange(1 << 23, dtype=np.uint32)
    exps = [0, 1, 60, 100, 118, 121, 122, 123, 124, 125, 126, 127, 128, 129, 130]
    for e in exps:
        x = (mant | np.uint32(e << 23)).view(np.float32)
        for sgn in (1.0, -1.0):
            v = np.ascontiguousarray(x * np.float32(sgn))
            bad = hostcheck.hostcheck_encode_inrange_mismatches(v.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(v.size),
                                                                ctypes.c_int(fmt))
            assert bad == 0, (fmt, e, sgn, bad)
    v = np.array([15.0, 15.32165, 15.5, 1e30, np.inf, -np.inf, 3e38, 0.0, -0.0, 1e-45, -1e-45], np.float32)   # everything but NaN
    assert hostcheck.hostcheck_encode_inrange_mismatches(v.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(v.size), fmt) == 0


@pytest.mark.parametrize("fmt", [0, 1])
def test_pipe_balanced_encoder_equals_reference_encoder(hostcheck, fmt):
    """encode_balanced<FMT> (Veltkamp-split rounding on the FMA pipe + IMAD sign / bias arithmetic) == encode<FMT>
    for every float32 mantissa, both signs, exponents from denormals past the saturation thresholds, and Inf."""
    hostcheck.hostcheck_encode_balanced_mismatches.restype = ctypes.c_size_t
    mant = np.arange(1 << 23, dtype=np.uint32)
    for e in [0, 1, 60, 100, 118, 121, 122, 123, 124, 125, 126, 127, 128, 129, 130, 131, 140, 200, 254]:
        x = (mant | np.uint32(e << 23)).view(np.float32)
        for sgn in (1.0, -1.0):
            v = np.ascontiguousarray(x * np.float32(sgn))
            bad = hostcheck.hostcheck_encode_balanced_mismatches(v.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(v.size), ctypes.c_int(fmt))
            assert bad == 0, (fmt, e, sgn, bad)
    v = np.array([15.0, 15.32165, 15.5, 16.0, 1e30, np.inf, -np.inf, 3e38, 0.0, -0.0, 1e-45, -1e-45, 0.0625, 0.125, -0.0625], np.float32)
    assert hostcheck.hostcheck_encode_balanced_mismatches(v.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(v.size), fmt) == 0


@pytest.mark.parametrize("fmt", [0, 1])
def test_table_encoder_equals_reference_encoder(hostcheck, fmt):
    """The stand-alone quantizer's table encoder (class decisions folded into the rounded value by saturating FMA-pipe
    arithmetic, code = one byte from a shared-memory table) == encode<FMT> for every float32 mantissa, both signs,
    exponents from denormal quotients to 3e38.  Its domain excludes only NaN / Inf quotients and -0 (group probe)."""
    hostcheck.hostcheck_encode_lut_mismatches.restype = ctypes.c_size_t
    mant = np.arange(1 << 23, dtype=np.uint32)
    for e in [0, 1, 27, 60, 100, 118, 121, 122, 123, 124, 125, 126, 127, 128, 129, 130, 131, 132, 133, 134, 140, 153, 154, 155, 200, 254]:
        x = (mant | np.uint32(e << 23)).view(np.float32)
        if e == 0:
            x = x[1:]                                  # +-0 are checked below (-0 is outside the domain)
        for sgn in (1.0, -1.0):
            v = np.ascontiguousarray(x * np.float32(sgn))
            bad = hostcheck.hostcheck_encode_lut_mismatches(v.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(v.size), ctypes.c_int(fmt))
            assert bad == 0, (fmt, e, sgn, bad)
    v = np.array([15.0, 15.32165, 15.5, 16.0, 1e30, 3e38, -3e38, 0.0, 1e-45, -1e-45, 0.0625, 0.125, -0.0625, 14.75, 15.25, 15.75], np.float32)
    assert hostcheck.hostcheck_encode_lut_mismatches(v.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(v.size), fmt) == 0


def test_weight_bucket_encoder_equals_reference_encoder(hostcheck):
    """encode_wgt_bucket (the 16-threshold search of quantize_weight(8) as one bucket-table look-up, used by the
    weight-preparation kernels) == encode<SLFP34_WGT> for every float32 mantissa, both signs, all classes, NaN / Inf."""
    hostcheck.hostcheck_encode_wgt_bucket_mismatches.restype = ctypes.c_size_t
    mant = np.arange(1 << 23, dtype=np.uint32)
    for e in [0, 1, 100, 121, 122, 123, 124, 125, 126, 127, 128, 129, 130, 131, 200, 254, 255]:
        x = (mant | np.uint32(e << 23)).view(np.float32)
        for sgn in (0, 1):
            v = np.ascontiguousarray((x.view(np.uint32) | np.uint32(sgn << 31)).view(np.float32))
            bad = hostcheck.hostcheck_encode_wgt_bucket_mismatches(v.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(v.size))
            assert bad == 0, (e, sgn, bad)


def test_weight_lut_encoder_equals_reference_encoder(hostcheck):
    """encode_wgt_lut (ONE table look-up per weight in wprep_rows_kernel: bucket threshold, codes on both sides and their
    float16 images; csrc/slfp_common.cuh) == encode<SLFP34_WGT> and float16(decode(code)) for every float32 mantissa of
    every binade that matters (below, inside and above the code range) plus far-out exponents, both signs."""
    fn = hostcheck.hostcheck_encode_wgt_lut_mismatches
    fn.restype = ctypes.c_size_t
    mant = np.arange(1 << 23, dtype=np.uint32)
    for e in [1, 60, 100, 120, 121, 122, 123, 124, 125, 126, 127, 128, 129, 130, 131, 132, 190, 254]:
        x = (mant | np.uint32(e << 23)).view(np.float32)
        for sgn in (0, 1):
            v = np.ascontiguousarray((x.view(np.uint32) | np.uint32(sgn << 31)).view(np.float32))
            bad = fn(v.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(v.size))
            assert bad == 0, (e, sgn, bad)


def test_layerout_relu_fast_form_equals_reference_form(hostcheck):
    """relu(quantize_layerout(y)) as the fused fast epilogues compute it (Veltkamp split, min / max) is bit-exact with the
    generic routine for every mantissa at a spread of exponents (both signs, the 248 clamp, values around it)."""
    mant = np.arange(1 << 23, dtype=np.uint32)
    fn = hostcheck.hostcheck_layerout_relu_mismatches
    fn.restype = ctypes.c_size_t
    for e in (-20, -8, -4, -1, 0, 3, 6, 7, 8, 9, 20):
        for sign in (0, 1):
            x = (mant | np.uint32((e + 127) << 23) | np.uint32(sign << 31)).view(np.float32)
            assert fn(x.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(x.size)) == 0, (e, sign)
    edge = np.array([0.0, -0.0, 247.9, 248.0, 248.1, 252.0, 1e30, -1e30, 1e-30, np.inf, -np.inf], np.float32)
    assert fn(edge.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(edge.size)) == 0


def test_signed_fast_sfp33_codes(hostcheck):
    """SLFP_FMT_SFP33_SFAST (sign bit + 7-bit magnitude code, written by the fused depthwise kernels that have no ReLU):
    value == sign(q) * value of the post-ReLU code of |q| for every mantissa / octave, i.e. the same grid and rounding
    as the (already swept) unsigned format, and the 7-bit clamp loses nothing."""
    mant = np.arange(0, 1 << 23, 7, dtype=np.uint32)
    for e in (-6, -5, -4, -3, 0, 2, 3, 4, 6):
        for sign in (0, 1):
            q = (mant | np.uint32((e + 127) << 23) | np.uint32(sign << 31)).view(np.float32)
            c = np.empty(q.shape, np.uint8); v = np.empty_like(q)
            hostcheck.hostcheck_sfast_codes(q.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(q.size), c.ctypes.data_as(ctypes.c_void_p),
                                            v.ctypes.data_as(ctypes.c_void_p))
            cu = np.empty(q.shape, np.uint8); vu = np.empty_like(q); c16 = np.empty(q.shape, np.uint8)
            aq = np.abs(q)
            hostcheck.hostcheck_relu_codes(aq.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(q.size), ctypes.c_int(1),
                                           cu.ctypes.data_as(ctypes.c_void_p), vu.ctypes.data_as(ctypes.c_void_p),
                                           c16.ctypes.data_as(ctypes.c_void_p))
            want = np.where(sign == 1, -vu, vu)
            assert (v == want).all(), (e, sign)
            assert ((c >> 7) == sign).all()
