"""GPU: the fused quantizer / de-quantizer / abs-max kernels through the C ABI, bit-exact against the
oracle and against the reference-generated fixtures."""
import ctypes

import numpy as np
import pytest
import torch

from conftest import bits, same_bits

pytestmark = pytest.mark.gpu

FMTS = {"sfp33_act": 0, "slfp34_act": 1, "slfp34_wgt": 2, "sfp44_out": 3}


@pytest.mark.parametrize("name", list(FMTS))
def test_golden_samples_bit_exact(g_quant, name):
    from gpu_util import quantize_gpu
    codes, fq, _ = quantize_gpu(g_quant["x"], FMTS[name])
    assert same_bits(g_quant[name], fq).all()


@pytest.mark.parametrize("name", ["slfp34_act", "slfp34_wgt", "sfp33_act"])
def test_prescale_is_ieee_division(g_quant, name):
    from gpu_util import quantize_gpu
    for i, k in enumerate(g_quant["prescale_k"]):
        _, fq, _ = quantize_gpu(g_quant["prescale_x"], FMTS[name], kdiv=k)
        assert same_bits(g_quant["scaled_" + name][i], fq).all()


@pytest.mark.parametrize("fmt", [0, 1, 2, 3])
def test_exhaustive_mantissas_vs_oracle(orc, fmt):
    """All 2^23 mantissas at several exponents (vector path + ragged tail), codes and fake-quant."""
    from gpu_util import quantize_gpu
    mant = np.arange(1 << 23, dtype=np.uint32)
    for e in ([-5, -4, -1, 3] + ([-127, 7] if fmt == 3 else [])):
        x = (mant | np.uint32(max(e + 127, 0) << 23)).view(np.float32)[: (1 << 23) - 5]   # ragged length
        if e == -1:
            x = -x
        codes, fq, f16 = quantize_gpu(x, fmt, want_f16=True)
        oc, oq = orc.quantize(x, fmt)
        assert same_bits(oq, fq).all(), (fmt, e)
        if fmt != 3:
            assert (codes == oc).all(), (fmt, e)
        assert (f16 == oq.astype(np.float16)).all() or np.isnan(oq).any()


def test_sizes_alignment_and_empty(orc):
    from cnns_slfp_quantization_b200 import _native as nv
    rng = np.random.default_rng(3)
    for n in (0, 1, 3, 4095, 4096, 4097, 12289, 1 << 20):
        x = (rng.standard_normal(n + 3) * 4).astype(np.float32)
        xt = torch.from_numpy(x).cuda()
        for off in (0, 1):                       # off=1: mis-aligned pointer -> scalar kernel
            xs = xt[off:off + n]
            codes = torch.empty(n + 1, dtype=torch.uint8, device="cuda")[off:off + n]
            fq = torch.empty(n + 3, dtype=torch.float32, device="cuda")[off:off + n]
            nv.check(nv.lib().slfp_quantize_f32(xs.data_ptr(), n, 0.37, 1, 0, codes.data_ptr(), fq.data_ptr(), None, nv.stream()))
            torch.cuda.synchronize()
            oc, oq = orc.quantize(x[off:off + n], 1, kdiv=0.37)
            assert (codes.cpu().numpy() == oc).all() and same_bits(oq, fq.cpu().numpy()).all(), (n, off)


def test_dequantize_all_codes(orc):
    from cnns_slfp_quantization_b200 import _native as nv
    codes = torch.arange(256, dtype=torch.uint8, device="cuda").repeat(37)[:-3].contiguous()
    for fmt in (0, 1, 2):
        out = torch.empty(codes.numel(), dtype=torch.float32, device="cuda")
        nv.check(nv.lib().slfp_dequantize(codes.data_ptr(), codes.numel(), fmt, out.data_ptr(), nv.stream()))
        torch.cuda.synchronize()
        assert same_bits(orc.decode(codes.cpu().numpy(), fmt), out.cpu().numpy()).all()


def test_roundtrip_idempotent_at_full_size():
    """Size-independent properties at the BASELINE layer size (256x256x56x56 = 205.5 M elements):
    decode(encode(x)) is a fixed point of the quantizer and |q| never exceeds the format's top."""
    from cnns_slfp_quantization_b200 import _native as nv
    n = 256 * 256 * 56 * 56
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(n, device="cuda", generator=g) * 4
    codes = torch.empty(n, dtype=torch.uint8, device="cuda")
    fq = torch.empty(n, dtype=torch.float32, device="cuda")
    lib = nv.lib()
    nv.check(lib.slfp_quantize_f32(x.data_ptr(), n, 1.0, 1, 0, codes.data_ptr(), fq.data_ptr(), None, nv.stream()))
    dq = torch.empty_like(fq)
    nv.check(lib.slfp_dequantize(codes.data_ptr(), n, 1, dq.data_ptr(), nv.stream()))
    assert torch.equal(dq.view(torch.int32), fq.view(torch.int32))
    codes2 = torch.empty_like(codes)
    nv.check(lib.slfp_quantize_f32(fq.data_ptr(), n, 1.0, 1, 0, codes2.data_ptr(), None, None, nv.stream()))
    # idempotent except at the very top: the top grid value 2^(3+15/16) is 3 ulp ABOVE the
    # saturation literal 15.32165, so the two swap when re-quantized (a reference quirk)
    same = codes2 == codes
    u = codes & 0x7f
    top = (u == 2) | (u == 127)
    assert bool((same | top).all())
    assert bool(((codes2 & 0x7f)[u == 127] == 2).all()) and bool(((codes2 & 0x7f)[u == 2] == 127).all())
    assert float(fq.abs().max()) <= 15.3216553
    assert bool(((fq == 0) == (x == 0)).all())
    assert bool((torch.sign(fq) == torch.sign(x)).all())


def test_absmax(orc):
    from cnns_slfp_quantization_b200 import _native as nv
    rng = np.random.default_rng(5)
    for n in (1, 1000, 1 << 20, (1 << 22) + 7):
        x = rng.standard_normal(n).astype(np.float32) * 3
        x[rng.integers(n)] = -77.25
        xt = torch.from_numpy(x).cuda()
        out = torch.full((1,), 123.0, device="cuda")
        nv.check(nv.lib().slfp_absmax_f32(xt.data_ptr(), n, out.data_ptr(), 1, nv.stream()))
        assert float(out) == orc.absmax(x) == 77.25
        nv.check(nv.lib().slfp_absmax_f32(xt.data_ptr(), n, out.data_ptr(), 0, nv.stream()))   # accumulates
        assert float(out) == 77.25


def test_quantize_nhwc_padded(orc):
    from cnns_slfp_quantization_b200 import _native as nv
    rng = np.random.default_rng(9)
    for C, Cp in ((3, 4), (24, 32), (58, 64), (16, 16)):
        x = (rng.standard_normal((50, C)) * 3).astype(np.float32)
        xt = torch.from_numpy(x).cuda()
        codes = torch.full((50, Cp), 0x77, dtype=torch.uint8, device="cuda")
        nv.check(nv.lib().slfp_quantize_nhwc_f32(xt.data_ptr(), 50, C, Cp, 0.5, 1, codes.data_ptr(), nv.stream()))
        oc, _ = orc.quantize(x, 1, kdiv=0.5)
        got = codes.cpu().numpy()
        assert (got[:, :C] == oc).all() and (got[:, C:] == 0).all()


def test_quantize_nchw_input(orc):
    """Network-input quantizer (NCHW float32 -> NHWC codes, padded channels): the 4-pixels-per-thread path (c_phys = 4,
    HW % 4 == 0) and the generic path, both formats, bit-exact against the oracle on the permuted tensor."""
    from cnns_slfp_quantization_b200 import _native as nv
    rng = np.random.default_rng(21)
    for fmt in (0, 1):
        for N, C, H, W, Cp in ((3, 3, 8, 12, 4), (2, 4, 6, 6, 4), (2, 3, 5, 7, 4), (1, 1, 4, 4, 4), (2, 6, 4, 4, 8), (5, 3, 32, 32, 4)):
            x = (rng.standard_normal((N, C, H, W)) * 2.5).astype(np.float32)
            x.flat[:: 17] = 0.0
            xt = torch.from_numpy(x).cuda()
            codes = torch.full((N, H, W, Cp), 0x77, dtype=torch.uint8, device="cuda")
            nv.check(nv.lib().slfp_quantize_nchw_f32(xt.data_ptr(), N, C, H * W, Cp, 0.7, fmt, codes.data_ptr(), nv.stream()))
            oc, _ = orc.quantize(np.ascontiguousarray(x.transpose(0, 2, 3, 1)), fmt, kdiv=0.7)
            got = codes.cpu().numpy()
            assert (got[..., :C] == oc.reshape(N, H, W, C)).all() and (got[..., C:] == 0).all(), (fmt, N, C, H, W)


def test_host_buffer_entry(orc):
    from cnns_slfp_quantization_b200 import _native as nv
    x = (np.random.default_rng(2).standard_normal(100001) * 5).astype(np.float32)
    codes = np.empty(x.size, np.uint8)
    fq = np.empty_like(x)
    nv.check(nv.lib().slfp_quantize_host_f32(x.ctypes.data_as(ctypes.c_void_p), x.size, 0.9, 1,
                                             codes.ctypes.data_as(ctypes.c_void_p), fq.ctypes.data_as(ctypes.c_void_p)))
    oc, oq = orc.quantize(x, 1, kdiv=0.9)
    assert (codes == oc).all() and same_bits(oq, fq).all()


@pytest.mark.parametrize("qbit,kind", [(8, "act"), (8, "weight"), (7, "act")])
def test_dynamic_max_scaling_quantizer(orc, qbit, kind):
    """abs-max -> max-scaling -> quantize entirely on the device (slfp_quantize_dyn_f32 reads K from device memory):
    codes and fake-quant equal the oracle's quantizer run with K = float32(float64(max|x|) / 15.5), the reference's
    recipe (nets_cifar/mobilenetv1.py:15); the whole sequence is capturable in a CUDA graph and follows the data."""
    import torch
    from cnns_slfp_quantization_b200 import calibration
    rng = np.random.default_rng(21 + qbit)
    x = (rng.standard_normal(300007) * rng.uniform(0.01, 30)).astype(np.float32)
    xt = torch.from_numpy(x).cuda()
    codes, k = calibration.quantize_dynamic(xt, qbit, kind)
    fq, k2 = calibration.quantize_dynamic(xt, qbit, kind, want="fakeq")
    torch.cuda.synchronize()
    k_want = np.float32(np.float64(np.abs(x).max()) / 15.5)
    assert np.float32(k.item()) == k_want and np.float32(k2.item()) == k_want
    want_codes, want_fq = orc.quantize(x, orc.fmt_for(qbit, kind), float(k_want))
    assert (codes.cpu().numpy() == want_codes).all()
    assert (fq.cpu().numpy().view(np.uint32) == want_fq.view(np.uint32)).all()
    # graph capture: replaying after the INPUT changed re-derives K on the device
    amax = torch.zeros((), dtype=torch.float32, device="cuda")
    out = torch.empty(x.size, dtype=torch.uint8, device="cuda")
    kd = torch.empty((), dtype=torch.float32, device="cuda")
    from cnns_slfp_quantization_b200 import _native as nv
    lib = nv.lib()
    fmt = nv.fmt_for(qbit, kind)

    def seq():
        st = nv.stream()
        nv.check(lib.slfp_absmax_f32(xt.data_ptr(), xt.numel(), amax.data_ptr(), 1, st))
        nv.check(lib.slfp_quantize_dyn_f32(xt.data_ptr(), xt.numel(), amax.data_ptr(), 15.5, fmt, 0, out.data_ptr(), None, None, kd.data_ptr(), st))
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        seq()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        seq()
    x2 = (x * 3.7).astype(np.float32)
    xt.copy_(torch.from_numpy(x2))
    g.replay()
    torch.cuda.synchronize()
    k2_want = np.float32(np.float64(np.abs(x2).max()) / 15.5)
    assert np.float32(kd.item()) == k2_want
    assert (out.cpu().numpy() == orc.quantize(x2, orc.fmt_for(qbit, kind), float(k2_want))[0]).all()


@pytest.mark.parametrize("fmt", [0, 1])
def test_s2d_input_quantizer_fast_path_with_special_values(orc, fmt):
    """The stem's NCHW -> space-to-depth quantizer (two folded pixels per thread, table encoder, one group probe per
    thread for the general path): bit-exact codes against the oracle, including threads whose 24 elements contain
    NaN / Inf / -0 / denormal-range values (they take the general encoder)."""
    import torch
    from cnns_slfp_quantization_b200 import _native as nv
    lib = nv.lib()
    rng = np.random.default_rng(77)
    N, C, H, W = 3, 3, 16, 24
    x = (rng.standard_normal((N, C, H, W)) * 3).astype(np.float32)
    flat = x.reshape(-1)
    specials = np.array([np.inf, -np.inf, -0.0, 0.0, 1e-40, -1e-40, 3e-38, 1e30, -1e30, 0.0625 * 0.37, 15.32165 * 0.37], np.float32)
    pos = rng.choice(flat.size, specials.size * 4, replace=False)
    flat[pos] = np.tile(specials, 4)
    k = np.float32(0.37)
    xt = torch.from_numpy(x).cuda()
    out = torch.empty((N, H // 2, W // 2, 16), dtype=torch.uint8, device="cuda")
    nv.check(lib.slfp_quantize_nchw_s2d_f32(xt.data_ptr(), N, C, H, W, 16, float(k), fmt, out.data_ptr(), nv.stream()))
    torch.cuda.synchronize()
    codes, _ = orc.quantize(x, fmt, float(k))
    want = np.zeros((N, H // 2, W // 2, 16), np.uint8)
    for dy in range(2):
        for dx in range(2):
            for c in range(C):
                want[..., (dy * 2 + dx) * C + c] = codes[:, c, dy::2, dx::2]
    assert (out.cpu().numpy() == want).all()
    # SLFP_FMT_F16Q forms of the input quantizers (round 2): float16 images of exactly those codes' values
    _, fq = orc.quantize(x, fmt, float(k), want_codes=False)
    with np.errstate(over="ignore", invalid="ignore"):
        img = fq.astype(np.float16)                                                      # [n, c, h, w]
    # (a) width-folded stem input: zero-padded [n, hp, wp, 16] buffer, interior at (2, 2)
    hp, wp = H // 2 + 3, W // 2 + 3
    o16 = torch.zeros((N, hp, wp, 16), dtype=torch.float16, device="cuda")
    nv.check(lib.slfp_quantize_nchw_s2d_f16q(xt.data_ptr(), N, H, W, float(k), fmt, 2, 2, hp, wp, o16.data_ptr(), nv.stream()))
    torch.cuda.synchronize()
    want16 = np.zeros((N, hp, wp, 16), np.float16)
    for dy in range(2):
        for dx in range(2):
            for c in range(C):
                want16[:, 2:2 + H // 2, 2:2 + W // 2, (dy * 2 + dx) * C + c] = img[:, c, dy::2, dx::2]
    assert (o16.cpu().numpy().view(np.uint16) == want16.view(np.uint16)).all()
    # (b) 3x3 / pad 1 im2col matrix [n, h, w, 64], entry (r * 3 + s) * 4 + c
    oi = torch.full((N, H, W, 64), 7.0, dtype=torch.float16, device="cuda")
    nv.check(lib.slfp_quantize_nchw_im2col3x3_f16q(xt.data_ptr(), N, H, W, float(k), fmt, oi.data_ptr(), nv.stream()))
    torch.cuda.synchronize()
    wanti = np.zeros((N, H, W, 64), np.float16)
    pad = np.pad(img.transpose(0, 2, 3, 1), ((0, 0), (1, 1), (1, 1), (0, 0)))
    for r in range(3):
        for s_ in range(3):
            wanti[..., (r * 3 + s_) * 4:(r * 3 + s_) * 4 + 3] = pad[:, r:r + H, s_:s_ + W]
    assert (oi.cpu().numpy().view(np.uint16) == wanti.view(np.uint16)).all()


def test_gather_quantize_matches_plain_quantizer(orc):
    """slfp_gather_quantize_f16 (split / cat / channel_shuffle as an index map): codes equal the oracle's quantizer of
    the gathered float16 values, pad channels are code 0."""
    import ctypes
    import torch
    from cnns_slfp_quantization_b200 import _native as nv
    lib = nv.lib()
    rng = np.random.default_rng(5)
    npix = 1000
    a = torch.from_numpy((rng.standard_normal((npix, 58)) * 3).astype(np.float16)).cuda()
    b = torch.from_numpy((rng.standard_normal((npix, 24)) * 3).astype(np.float16)).cuda()
    chans = [(a, j // 2) if j % 2 == 0 else (b, (j // 2) % 24) for j in range(58)]
    tab = (nv.SlfpGatherChan * 58)()
    for e, (t, ch) in zip(tab, chans):
        e.src, e.stride, e.ch = t.data_ptr(), t.shape[1], ch
    dtab = torch.frombuffer(bytearray(bytes(tab)), dtype=torch.uint8).cuda()
    for fmt in (0, 1):
        out = torch.full((npix, 64), 9, dtype=torch.uint8, device="cuda")
        nv.check(lib.slfp_gather_quantize_f16(dtab.data_ptr(), npix, 58, 64, 0.21, fmt, out.data_ptr(), nv.stream()))
        torch.cuda.synchronize()
        g = np.stack([(a if t is a else b).cpu().numpy()[:, ch] for t, ch in chans], 1).astype(np.float32)
        want, _ = orc.quantize(g, fmt, 0.21)
        got = out.cpu().numpy()
        assert (got[:, :58] == want).all() and (got[:, 58:] == 0).all()


def test_e4m3_encoder_is_the_sfp33_quantizer(orc):
    """SLFP_FMT_E4M3: the cvt-based device encoder (depthwise kernel epilogue) produces, for every float32 mantissa at every
    octave of the SFP<3,3> range and both signs, the e4m3 byte whose value IS the reference quantizer's value
    (utils/sfp_quant.py:63-78: round-half-even, 0.0625 / 0.125 / 15 clamps; +-1e-10 is stored as +-0)."""
    import ctypes
    import torch
    from cnns_slfp_quantization_b200 import _native as nv
    from gpu_util import decode_e4m3
    lib = nv.lib()
    dev = torch.device("cuda:0")
    C = 16
    mant = np.arange(0, 1 << 23, 3, dtype=np.uint32)
    n = (mant.size // C) * C
    mant = mant[:n]
    wt = torch.zeros((C, 1, 1, 1), dtype=torch.float32, device=dev) + 1.0          # depthwise 1x1 with unit weights: y = x_q
    for e in (-6, -5, -4, -3, -1, 0, 2, 3, 4, 5):
        for sign in (0, 1):
            x = (mant | np.uint32((e + 127) << 23) | np.uint32(sign << 31)).view(np.float32)
            # route the values through the gather-quantizer's e4m3 output (exact table encoder + re-spelling)
            src = torch.from_numpy(x.astype(np.float16)).to(dev).reshape(-1, C)
            tab = (nv.SlfpGatherChan * C)()
            for j, t_ in enumerate(tab):
                t_.src, t_.stride, t_.ch = src.data_ptr(), C, j
            dtab = torch.frombuffer(bytearray(bytes(tab)), dtype=torch.uint8).to(dev)
            out = torch.empty((src.shape[0], C), dtype=torch.uint8, device=dev)
            nv.check(lib.slfp_gather_quantize_f16(dtab.data_ptr(), src.shape[0], C, C, 1.0, nv.FMT_E4M3, out.data_ptr(), nv.stream()))
            torch.cuda.synchronize()
            xh = src.float().cpu().numpy().reshape(-1)
            _, want = orc.quantize(xh, 0, 1.0, want_codes=False)
            want = np.where(np.abs(want) <= 1e-9, 0.0, want)
            got = decode_e4m3(out.cpu().numpy().reshape(-1))
            assert (got == want).all(), (e, sign, int((got != want).sum()))


@pytest.mark.parametrize("fmt", [0, 1, 7])
def test_gather_quantize_runs(orc, fmt):
    """slfp_gather_quantize_runs_f16 (the run-based, shared-memory-staged form the ShuffleNetV2 plan uses): three source
    tensors with runs of different lengths and output steps (1, 2, 4 - the shuffle's interleaving), a ragged pixel count,
    values that include 0, the 248 clamp and out-of-table entries (negative, 300, inf): codes equal the oracle's quantizer."""
    import ctypes
    import torch
    from cnns_slfp_quantization_b200 import _native as nv
    from gpu_util import decode_e4m3
    lib = nv.lib()
    rng = np.random.default_rng(9)
    npix = 1000 + 7
    a = np.abs(rng.standard_normal((npix, 64)) * 3).astype(np.float16)
    b = np.abs(rng.standard_normal((npix, 128)) * 40).astype(np.float16)
    c = np.abs(rng.standard_normal((npix, 64)) * 0.5).astype(np.float16)
    a[::7, 3] = 0; b[::5, 70] = 248; b[1::5, 71] = 300
    if fmt != 7:                       # the e4m3 form folds the ReLU into its low clamp: sources are non-negative by contract
        c[::11, 5] = -2.0; c[3::11, 6] = np.inf
    ta, tb, tc = (torch.from_numpy(t).cuda() for t in (a, b, c))
    # output channel j (58 logical channels, c_phys 64): even j <- a[29 + j/2]; j = 1 mod 4 <- b[64 + j//4]; j = 3 mod 4 <- c[j//4]
    runs_spec = [(ta, 64, 29, 29, 0, 2), (tb, 128, 64, 15, 1, 4), (tc, 64, 0, 14, 3, 4)]
    tab = (nv.SlfpGatherRun * 3)()
    mg, sh = ctypes.c_uint(), ctypes.c_uint()
    want_vals = np.zeros((npix, 64), np.float32)
    covered = np.zeros(64, bool)
    for e, (t, stride, ch0, ln, d0, st_) in zip(tab, runs_spec):
        lib.slfp_magic_u32((((ch0 + ln + 7) & ~7) - (ch0 & ~7)) >> 3, ctypes.byref(mg), ctypes.byref(sh))
        e.src, e.stride, e.ch0, e.len, e.dst_start, e.dst_step, e.magic, e.shift = t.data_ptr(), stride, ch0, ln, d0, st_, mg.value, sh.value
        for i in range(ln):
            want_vals[:, d0 + i * st_] = t.float().cpu().numpy()[:, ch0 + i]
            covered[d0 + i * st_] = True
    assert covered[:58].all() and not covered[58:].any()
    dtab = torch.frombuffer(bytearray(bytes(tab)), dtype=torch.uint8).cuda()
    out = torch.full((npix, 64), 9, dtype=torch.uint8, device="cuda")
    bpp = 16 * sum((((ch0 + ln + 7) & ~7) - (ch0 & ~7)) >> 3 for _, _, ch0, ln, _, _ in runs_spec)
    nv.check(lib.slfp_gather_quantize_runs_f16(dtab.data_ptr(), 3, npix, 58, 64, bpp, 0.21, fmt, out.data_ptr(), nv.stream()))
    torch.cuda.synchronize()
    got = out.cpu().numpy()
    assert (got[:, 58:] == 0).all()
    wc, wq = orc.quantize(want_vals[:, :58], 0 if fmt == 7 else fmt, 0.21)
    if fmt == 7:
        wq = np.where(np.abs(wq) <= 1e-9, 0.0, wq)
        assert (decode_e4m3(got[:, :58]) == wq).all()
    else:
        assert (got[:, :58] == wc).all()
