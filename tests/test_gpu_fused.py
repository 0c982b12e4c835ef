"""GPU: the fused eval pipeline's pieces - post-ReLU code formats written by the conv epilogue, the
space-to-depth stem (asymmetric padding), batched weight re-quantization, max-pool on unsigned codes -
and the stability of the compiled plan (whole-net parity: tests/test_gpu_parity.py)."""
import ctypes
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _h(a):
    return np.asarray(a, np.float32).astype(np.float16)


def _conv_fast(x, w, ka, kw, mul, add, stride, pad, residual=None, want_f16=False, next_ks=(0.21,), qbit=8, pad_extra=0,
               in_codes=None, in_fmt=None):
    """conv through the C ABI with the folded epilogue; returns dict(y32 from a second launch, y16, codes...)."""
    from cnns_slfp_quantization_b200 import _native as nv
    lib = nv.lib()
    dev = torch.device("cuda:0")
    xt = torch.from_numpy(np.ascontiguousarray(x, np.float32)).to(dev).permute(0, 2, 3, 1).contiguous()
    wt = torch.from_numpy(np.ascontiguousarray(w, np.float32)).to(dev)
    N, H, W, C = xt.shape
    K, _, R, S = wt.shape
    Cp = (C + 15) // 16 * 16
    afmt, wfmt = nv.fmt_for(qbit, "act"), nv.fmt_for(qbit, "weight")
    st = nv.stream()
    if in_codes is None:
        xc = torch.empty((N, H, W, Cp), dtype=torch.uint8, device=dev)
        nv.check(lib.slfp_quantize_nhwc_f32(xt.data_ptr(), N * H * W, C, Cp, float(np.float32(ka)), afmt, xc.data_ptr(), st))
        fmt_in = afmt
    else:
        xc, fmt_in = in_codes, in_fmt
    d = nv.SlfpConvDesc(N, H, W, C, Cp, K, R, S, stride, stride, pad, pad, 1, 1, 1, fmt_in, pad_extra, pad_extra)
    Ho = (H + 2 * pad + pad_extra - (R - 1) - 1) // stride + 1
    Wo = (W + 2 * pad + pad_extra - (S - 1) - 1) // stride + 1
    pitch = lib.slfp_conv_wpitch(ctypes.byref(d))
    wh = torch.empty((K * pitch,), dtype=torch.float16, device=dev)
    so, sc, sr, ss = wt.stride()
    nv.check(lib.slfp_prepare_weights(ctypes.byref(d), wt.data_ptr(), so, sc, sr, ss, float(np.float32(kw)), wfmt, wh.data_ptr(),
                                      None, None, st))
    mul_t = torch.from_numpy(np.asarray(mul, np.float32)).to(dev)
    add_t = torch.from_numpy(np.asarray(add, np.float32)).to(dev)
    res_t = None if residual is None else torch.from_numpy(np.ascontiguousarray(residual.transpose(0, 2, 3, 1))).to(dev)
    out = {}
    # launch 1: float32 output of the same folded arithmetic (generic epilogue) = the quantizer's input
    e = nv.SlfpEpilogue()
    e.ch_mul, e.ch_add, e.relu = mul_t.data_ptr(), add_t.data_ptr(), 1
    y32 = torch.empty((N, Ho, Wo, K), dtype=torch.float32, device=dev)
    e.y_f32 = y32.data_ptr()
    if res_t is not None:
        e.residual, e.residual_f16 = res_t.data_ptr(), 1
    nv.check(lib.slfp_conv2d_fwd(ctypes.byref(d), xc.data_ptr(), wh.data_ptr(), ctypes.byref(e), st))
    # launch 2: the fast epilogue (post-ReLU codes, optional float16)
    e2 = nv.SlfpEpilogue()
    e2.ch_mul, e2.ch_add, e2.relu = mul_t.data_ptr(), add_t.data_ptr(), 1
    if res_t is not None:
        e2.residual, e2.residual_f16 = res_t.data_ptr(), 1
    rfmt = nv.relu_fmt(afmt)
    codes = [torch.full((N, Ho, Wo, K), 77, dtype=torch.uint8, device=dev) for _ in next_ks]
    e2.y_codes, e2.next_k_div, e2.next_fmt, e2.k_phys_out = codes[0].data_ptr(), float(np.float32(next_ks[0])), rfmt, K
    if len(next_ks) > 1:
        e2.y_codes2, e2.next_k_div2 = codes[1].data_ptr(), float(np.float32(next_ks[1]))
    y16 = None
    if want_f16:
        y16 = torch.empty((N, Ho, Wo, K), dtype=torch.float16, device=dev)
        e2.y_f16 = y16.data_ptr()
    nv.check(lib.slfp_conv2d_fwd(ctypes.byref(d), xc.data_ptr(), wh.data_ptr(), ctypes.byref(e2), st))
    torch.cuda.synchronize()
    out["y32"] = y32.cpu().numpy()
    out["y16"] = None if y16 is None else y16.cpu().numpy()
    out["codes"] = [c.cpu().numpy() for c in codes]
    out["codes_dev"], out["rfmt"] = codes, rfmt
    return out


def _check_codes(orc, y32, codes, kd, sfp33):
    """decode(code), as the float16 tensor-core operand, equals the reference quantizer of y / Ka - evaluated at
    y/Ka * (1 -+ 2^-21) to admit the epilogue's reciprocal multiply and FMA-folded scale near a class boundary."""
    got = _h(orc.decode_relu(codes, sfp33))
    q = y32.astype(np.float64) / float(np.float32(kd))
    fmt = 0 if sfp33 else 1
    ok = np.zeros(q.shape, bool)
    for eps in (0.0, -5e-7, 5e-7):
        _, want = orc.quantize((q * (1.0 + eps)).astype(np.float32), fmt, want_codes=False)
        ok |= got == _h(want)
    assert ok.all(), f"{(~ok).sum()} of {ok.size} codes off the reference grid value"
    # and away from boundaries it is exact: at most a sliver may need the +-eps alternatives
    _, want0 = orc.quantize(q.astype(np.float32), fmt, want_codes=False)
    assert (got != _h(want0)).mean() < 2e-4


@pytest.mark.parametrize("qbit", [8, 7])
def test_fast_epilogue_codes_only(orc, qbit):
    rng = np.random.default_rng(5)
    for (N, C, H, K, k, st, pad) in [(2, 64, 12, 96, 3, 1, 1), (3, 128, 9, 64, 1, 1, 0), (2, 32, 10, 272, 3, 2, 1)]:
        x = (rng.standard_normal((N, C, H, H)) * 2).astype(np.float32)
        w = (rng.standard_normal((K, C, k, k)) * 0.2).astype(np.float32)
        ka, kw = float(np.abs(x).max() / 15.5), float(np.abs(w).max() / 15.5)
        mul = (rng.uniform(0.5, 1.5, K) * ka * kw).astype(np.float32)
        add = (rng.standard_normal(K) * 0.5).astype(np.float32)
        nk = 0.19
        o = _conv_fast(x, w, ka, kw, mul, add, st, pad, next_ks=(nk,), qbit=qbit)
        assert (o["y32"] >= 0).all()
        _check_codes(orc, o["y32"], o["codes"][0], nk, qbit == 7)


def test_fast_epilogue_residual_f16_two_consumers(orc):
    rng = np.random.default_rng(6)
    N, C, H, K = 2, 64, 11, 256
    x = (rng.standard_normal((N, C, H, H)) * 2).astype(np.float32)
    w = (rng.standard_normal((K, C, 1, 1)) * 0.2).astype(np.float32)
    ka, kw = float(np.abs(x).max() / 15.5), float(np.abs(w).max() / 15.5)
    mul = (rng.uniform(0.1, 0.3, K) * ka * kw).astype(np.float32)
    add = (rng.standard_normal(K) * 0.2).astype(np.float32)
    res = np.abs(rng.standard_normal((N, K, H, H))).astype(np.float16)
    o = _conv_fast(x, w, ka, kw, mul, add, 1, 0, residual=res, want_f16=True, next_ks=(0.17, 0.31))
    y = o["y32"]
    assert (o["y16"] == _h(y)).all()                       # the float16 copy is RN(float32 result)
    _check_codes(orc, y, o["codes"][0], 0.17, False)
    _check_codes(orc, y, o["codes"][1], 0.31, False)
    # a dense layer consumes the post-ReLU codes: its output equals the conv of their decoded values
    from cnns_slfp_quantization_b200 import _native as nv
    xq = orc.decode_relu(o["codes"][0], False).transpose(0, 3, 1, 2)
    w2 = (rng.standard_normal((64, K, 3, 3)) * 0.1).astype(np.float32)
    kw2 = float(np.abs(w2).max() / 15.5)
    o2 = _conv_fast(xq, w2, 0.17, kw2, np.full(64, 0.17 * kw2, np.float32), np.zeros(64, np.float32), 1, 1,
                    in_codes=o["codes_dev"][0], in_fmt=o["rfmt"])
    _, wq = orc.quantize(w2, 2, kw2, want_codes=False)
    x16, w16 = _h(xq).astype(np.float64), _h(wq).astype(np.float64)
    want = torch.nn.functional.conv2d(torch.from_numpy(x16), torch.from_numpy(w16), None, 1, 1).numpy() * (0.17 * kw2)
    l1 = torch.nn.functional.conv2d(torch.from_numpy(np.abs(x16)), torch.from_numpy(np.abs(w16)), None, 1, 1).numpy() * 0.17 * kw2
    got = o2["y32"].transpose(0, 3, 1, 2)
    assert (np.abs(got - np.maximum(want, 0)) <= 3e-6 * l1 + 1e-6).all()


def test_space_to_depth_stem_equals_strided_conv(orc):
    """7x7 / stride 2 / pad 3 on 3 channels through the folded 4x4 / stride 1 form (asymmetric padding, TMA
    im2col kernel) against the same layer through the 4-channel-input kernel."""
    from cnns_slfp_quantization_b200 import _native as nv
    from gpu_util import conv_fwd_gpu
    lib = nv.lib()
    rng = np.random.default_rng(8)
    N, C, H, K = 3, 3, 32, 64
    x = (rng.standard_normal((N, C, H, H)) * 2).astype(np.float32)
    w = (rng.standard_normal((K, C, 7, 7)) * 0.2).astype(np.float32)
    ka, kw = float(np.abs(x).max() / 15.5), float(np.abs(w).max() / 15.5)
    ref = conv_fwd_gpu(x, w, None, ka, kw, 8, 2, 3, 1, 1)["y"]
    dev = torch.device("cuda:0")
    xt = torch.from_numpy(x).to(dev)
    cp = 16
    xs = torch.empty((N, H // 2, H // 2, cp), dtype=torch.uint8, device=dev)
    nv.check(lib.slfp_quantize_nchw_s2d_f32(xt.data_ptr(), N, C, H, H, cp, float(np.float32(ka)), nv.FMT_SLFP34_ACT, xs.data_ptr(),
                                            nv.stream()))
    torch.cuda.synchronize()
    # the folded codes are the plain quantizer's codes, re-arranged
    codes, _ = orc.quantize(x, 1, ka)
    want = np.zeros((N, H // 2, H // 2, cp), np.uint8)
    for dy in range(2):
        for dx in range(2):
            for c in range(C):
                want[..., (dy * 2 + dx) * C + c] = codes[:, c, dy::2, dx::2]
    assert (xs.cpu().numpy() == want).all()
    wp = np.zeros((K, C, 8, 8), np.float32)
    wp[:, :, 1:, 1:] = w
    w2 = wp.reshape(K, C, 4, 2, 4, 2).transpose(0, 3, 5, 1, 2, 4).reshape(K, 4 * C, 4, 4)
    xq = orc.decode(want, 1)[..., :4 * C].transpose(0, 3, 1, 2)
    o = _conv_fast(xq, w2, ka, kw, np.full(K, np.float32(ka) * np.float32(kw), np.float32), np.zeros(K, np.float32), 1, 2,
                   pad_extra=-1, in_codes=xs, in_fmt=nv.FMT_SLFP34_ACT)
    got = o["y32"].transpose(0, 3, 1, 2)
    assert got.shape == ref.shape
    scale = np.abs(ref).max()
    assert np.abs(got - np.maximum(ref, 0)).max() <= 2e-5 * scale


def test_batched_weight_preparation_equals_per_layer(orc):
    from cnns_slfp_quantization_b200 import _native as nv
    lib = nv.lib()
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(9)
    shapes = [(64, 3, 7, 7, 4), (96, 64, 3, 3, 64), (1000, 80, 1, 1, 80), (24, 24, 3, 3, 32)]
    descs, ws, singles, batch_out, strides, kws = [], [], [], [], [], []
    for (K, C, R, S, Cp) in shapes:
        d = nv.SlfpConvDesc(1, 8, 8, C, Cp, K, R, S, 1, 1, 0, 0, 1, 1, 1, nv.FMT_SLFP34_ACT)
        w = torch.from_numpy((rng.standard_normal((K, C, R, S)) * 0.3).astype(np.float32)).to(dev)
        pitch = lib.slfp_conv_wpitch(ctypes.byref(d))
        a = torch.zeros(K * pitch, dtype=torch.float16, device=dev)
        b = torch.ones(K * pitch, dtype=torch.float16, device=dev)
        kw = float(np.float32(0.3 * 3 / 15.5))
        nv.check(lib.slfp_prepare_weights(ctypes.byref(d), w.data_ptr(), *w.stride(), kw, nv.FMT_SLFP34_WGT, a.data_ptr(), None, None,
                                          nv.stream()))
        descs.append(d); ws.append(w); singles.append(a); batch_out.append(b); strides += list(w.stride()); kws.append(kw)
    n = len(shapes)
    jobs = (nv.SlfpWeightJob * n)()
    for i, j in enumerate(jobs):
        j.desc, j.w, j.kw, j.w_f16 = ctypes.pointer(descs[i]), ws[i].data_ptr(), kws[i], batch_out[i].data_ptr()
        j.w_stride[:] = strides[4 * i:4 * i + 4]
    nv.check(lib.slfp_prepare_weights_jobs(n, jobs, nv.FMT_SLFP34_WGT, nv.stream()))
    torch.cuda.synchronize()
    for a, b in zip(singles, batch_out):
        assert torch.equal(a, b)


def test_fused_block_tail_equals_two_convs(orc):
    """conv3 + downsample as one concatenated-K GEMM (slfp_conv2d_fwd_dual) against the two separate launches
    (downsample -> float16, conv3 with that residual): same float16 output up to the ratio-scaled weights'
    extra rounding, and the codes follow the output."""
    from cnns_slfp_quantization_b200 import engine, _native as nv
    import torch.nn as nn
    from cnns_slfp_quantization_b200.utils import conv2d_func as cf
    torch.manual_seed(3)
    dev = torch.device("cuda:0")
    N, C1, C2, K, H = 2, 64, 128, 256, 12
    for stride in (1, 2):
        c3 = cf.conv2d_Q(8, 0.011, 0.4)(C1, K, 1).to(dev)
        ds = cf.conv2d_Q(8, 0.013, 0.5)(C2, K, 1, stride=stride).to(dev)
        bn3, bnd = nn.BatchNorm2d(K).to(dev).eval(), nn.BatchNorm2d(K).to(dev).eval()
        for bn in (bn3, bnd):
            bn.weight.data.uniform_(0.2, 1.5); bn.bias.data.normal_(0, 0.3)
            bn.running_mean.normal_(0, 0.2); bn.running_var.uniform_(0.5, 1.5)
        bn3.weight.data[:5] = 0.0                        # zero-initialised residual BN: the ratio must stay finite
        outs = []
        for fused in (True, False):
            P = engine.Plan(N, dev, 8)
            a2 = P._alloc(N, H, H, C1, "codes", engine._k32(c3.Ka), cp=C1, fmt=nv.FMT_SLFP34_RELU)
            x = P._alloc(N, H * stride, H * stride, C2, "codes", engine._k32(ds.Ka), cp=C2, fmt=nv.FMT_SLFP34_RELU)
            g = torch.Generator(device="cpu").manual_seed(5)
            a2.buf.copy_(torch.randint(0, 200, a2.buf.shape, generator=g, dtype=torch.uint8))
            x.buf.copy_(torch.randint(0, 200, x.buf.shape, generator=g, dtype=torch.uint8))
            if fused:
                o = P.conv_dual(a2, c3, bn3, x, ds, bnd, relu=True, codes=[0.3], f16=True)
            else:
                r = P.conv(x, ds, bn=bnd, relu=False, f16=True)["f16"]
                o = P.conv(a2, c3, bn=bn3, relu=True, residual=r, codes=[0.3], f16=True)
            P.run()
            torch.cuda.synchronize()
            outs.append((o["f16"].buf.float().cpu().numpy(), o["codes"][0.3].buf.cpu().numpy()))
        (yf, cf_), (ys, cs_) = outs
        scale = np.abs(ys).max()
        assert np.abs(yf - ys).max() <= 4e-3 * scale, float(np.abs(yf - ys).max() / scale)
        # codes: both are the quantizer of (their own) float32 y; where the float16 outputs agree the two float32
        # values differ by < 2^-11 relative, which still straddles a code boundary (one per 4.4 %) now and then
        same = yf == ys
        assert same.mean() > 0.5
        assert (orc.decode_relu(cf_, False)[same] == orc.decode_relu(cs_, False)[same]).mean() > 0.995


@pytest.mark.parametrize("qbit,stride,C,H", [(8, 1, 32, 20), (7, 2, 64, 19), (8, 1, 144, 7), (7, 1, 16, 3)])
def test_depthwise_fused_kernels(orc, qbit, stride, C, H):
    """Depthwise 3x3 through the fused-pipeline stencil kernels (strip-mined 3x3 and the per-pixel variant for
    tiny maps): signed input codes -> folded affine + ReLU -> post-ReLU codes, against a float64 depthwise
    convolution of the decoded operands pushed through the reference quantizer."""
    from cnns_slfp_quantization_b200 import _native as nv
    lib = nv.lib()
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(40 + C)
    N = 3
    x = (rng.standard_normal((N, C, H, H)) * 2).astype(np.float32)
    w = (rng.standard_normal((C, 1, 3, 3)) * 0.3).astype(np.float32)
    ka, kw = float(np.abs(x).max() / 15.5), float(np.abs(w).max() / 15.5)
    afmt, wfmt = nv.fmt_for(qbit, "act"), nv.fmt_for(qbit, "weight")
    xt = torch.from_numpy(x).to(dev).permute(0, 2, 3, 1).contiguous()
    xc = torch.empty((N, H, H, C), dtype=torch.uint8, device=dev)
    nv.check(lib.slfp_quantize_nhwc_f32(xt.data_ptr(), N * H * H, C, C, float(np.float32(ka)), afmt, xc.data_ptr(), nv.stream()))
    d = nv.SlfpConvDesc(N, H, H, C, C, C, 3, 3, stride, stride, 1, 1, 1, 1, C, afmt)
    wt = torch.from_numpy(w).to(dev)
    wc = torch.empty((C * 9,), dtype=torch.uint8, device=dev)
    nv.check(lib.slfp_prepare_weights(ctypes.byref(d), wt.data_ptr(), *wt.stride(), float(np.float32(kw)), wfmt, None, wc.data_ptr(), None,
                                      nv.stream()))
    Ho = (H + 2 - 3) // stride + 1
    mul = (rng.uniform(0.5, 1.5, C) * np.float32(ka) * np.float32(kw)).astype(np.float32)
    add = (rng.standard_normal(C) * 0.4).astype(np.float32)
    mul_t, add_t = torch.from_numpy(mul).to(dev), torch.from_numpy(add).to(dev)
    nk = 0.23
    y = torch.full((N, Ho, Ho, C), 99, dtype=torch.uint8, device=dev)
    e = nv.SlfpEpilogue()
    e.ch_mul, e.ch_add, e.relu = mul_t.data_ptr(), add_t.data_ptr(), 1
    e.y_codes, e.next_k_div, e.next_fmt, e.k_phys_out = y.data_ptr(), float(np.float32(nk)), nv.relu_fmt(afmt), C
    nv.check(lib.slfp_conv2d_fwd(ctypes.byref(d), xc.data_ptr(), wc.data_ptr(), ctypes.byref(e), nv.stream()))
    torch.cuda.synchronize()
    _, xq = orc.quantize(x, 0 if qbit == 7 else 1, ka, want_codes=False)
    _, wq = orc.quantize(w, 0 if qbit == 7 else 2, kw, want_codes=False)
    ref = torch.nn.functional.conv2d(torch.from_numpy(xq).double(), torch.from_numpy(wq).double(), None, stride, 1, 1, C).numpy()
    yref = np.maximum(ref * mul[None, :, None, None].astype(np.float64) + add[None, :, None, None], 0).transpose(0, 2, 3, 1)
    got = _h(orc.decode_relu(y.cpu().numpy(), qbit == 7))
    q = yref / float(np.float32(nk))
    ok = np.zeros(q.shape, bool)
    for eps in (0.0, -3e-6, 3e-6):                   # float32 accumulation of 9 products + the epilogue's folded scale
        _, want = orc.quantize((q * (1.0 + eps)).astype(np.float32), 0 if qbit == 7 else 1, want_codes=False)
        ok |= got == _h(want)
    assert ok.all(), int((~ok).sum())


def test_maxpool_on_post_relu_codes(orc):
    from cnns_slfp_quantization_b200 import _native as nv
    lib = nv.lib()
    rng = np.random.default_rng(10)
    N, H, W, C = 2, 13, 14, 32
    codes = rng.integers(0, 256, (N, H, W, C), dtype=np.uint8)
    x = torch.from_numpy(codes).cuda()
    Ho, Wo = (H + 2 - 3) // 2 + 1, (W + 2 - 3) // 2 + 1
    y = torch.empty((N, Ho, Wo, C), dtype=torch.uint8, device="cuda")
    nv.check(lib.slfp_maxpool_codes(x.data_ptr(), N, H, W, C, nv.FMT_SLFP34_RELU, 3, 3, 2, 1, y.data_ptr(), nv.stream()))
    torch.cuda.synchronize()
    vals = torch.from_numpy(orc.decode_relu(codes, False)).permute(0, 3, 1, 2)
    want = torch.nn.functional.max_pool2d(vals, 3, 2, 1).permute(0, 2, 3, 1).numpy()
    assert (orc.decode_relu(y.cpu().numpy(), False) == want).all()    # pooling commutes with decoding (monotone codes)


# Whole-net parity (module-level drop-in, fused engine, CUDA graph) lives in tests/test_gpu_parity.py: decisive
# reference fixtures (identical top-1 on every image), per-layer teacher-forced code parity, pinned logit RMS.


def test_shufflenetv2_calibration_taps():
    """ShuffleNetV2's calibration taps are recorded only after the reset calls (the reference REQUIRES them before a
    forward, nets_cifar/shufflenet_v2.py:175-183, :197; here forward() also works without)."""
    sys.path.insert(0, ROOT)
    from cnns_slfp_quantization_b200 import nets_common as nc
    from cnns_slfp_quantization_b200.nets_cifar import ShuffleNetV2
    m = ShuffleNetV2(7)
    m.load_state_dict(nc.synth_state_dict(m))
    m = m.cuda().eval()
    x = nc.synth_images(4, 32).cuda()
    with torch.no_grad():
        m(x)
    assert not hasattr(m, "layer_inputs")
    m.reset_layer_inputs_outputs(); m.reset_layer_weights()
    with torch.no_grad():
        m(x)
    assert len(m.get_layer_inputs()) == 57 and len(m.get_layer_weights()) == 57 and 55 in m.get_layer_outputs()


def test_plan_is_stable_and_deterministic_over_many_steps():
    """Regression for the code-ring hazard (a decode group one barrier phase ahead of a TMA load that landed late: stale
    codes, then an `unspecified launch failure` about once per 1 000 steps, only with the weights re-quantized between
    steps): 400 graph replays of the ResNet-50 plan - weight preparation + 55 kernels each - must run clean and
    produce bit-identical logits every time (the forward path has no atomics)."""
    sys.path.insert(0, ROOT)
    import bench
    from cnns_slfp_quantization_b200 import engine, nets_common as nc
    dev = torch.device("cuda", 0)
    model = bench.build_model_gpu(224, dev)
    plan = engine.compile_resnet50(model, 64, 224, device=dev)
    plan.input.copy_(nc.synth_images(64, 224, seed=7).to(dev))
    plan.prepare_weights()
    plan.run()
    torch.cuda.synchronize()
    plan.capture()
    plan()
    torch.cuda.synchronize()
    first = plan.output.clone()
    assert torch.isfinite(first).all()
    for i in range(400):
        plan()
        if i % 50 == 49:
            torch.cuda.synchronize()
            assert torch.equal(plan.output, first), f"logits changed at replay {i}"
    torch.cuda.synchronize()
    assert torch.equal(plan.output, first)
    # Plan.refresh(): BatchNorm / bias updates reach the captured graph through the in-place folded vectors ...
    with torch.no_grad():
        model.layer3[2].bn2.bias.add_(0.25)
        model.bn1.weight.mul_(1.1)
    plan.refresh()
    plan()
    torch.cuda.synchronize()
    changed = plan.output.clone()
    fresh = engine.compile_resnet50(model, 64, 224, device=dev)
    fresh.input.copy_(plan.input)
    fresh.run()
    torch.cuda.synchronize()
    assert not torch.equal(changed, first) and torch.equal(changed, fresh.output)
    # ... while a changed scale (baked into descriptors and epilogues) is refused
    model.layer2[0].conv1.Ka = model.layer2[0].conv1.Ka * 1.5
    with pytest.raises(RuntimeError, match="Ka / Kw changed"):
        plan.verify()


def _decode_any(orc, codes, fmt):
    from gpu_util import decode_codes
    return decode_codes(orc, codes, fmt)


@pytest.mark.parametrize("K,stride,qbit,relu,two,e4m3", [(32, 2, 7, True, False, False), (24, 1, 7, False, True, False),
                                                         (32, 1, 8, True, False, False), (24, 1, 8, False, True, False),
                                                         (32, 2, 7, True, False, True), (24, 1, 7, False, True, True)])
def test_direct_stem_kernel(orc, K, stride, qbit, relu, two, e4m3):
    """3x3 RGB stems through the CUDA-core direct kernel (conv_stem_direct.cu; MobileNetV1 3->32, ShuffleNetV2 3->24,
    VGG-16 3->64): float32 operands and accumulation, folded affine (+ ReLU), quantize-on-store in the fused code formats
    (post-ReLU codes, signed fast SFP<3,3>, exact signed codes; one or two consumers) against a float64 convolution of
    the decoded operands pushed through the reference quantizer."""
    import torch.nn as nn
    from cnns_slfp_quantization_b200 import engine, _native as nv
    from cnns_slfp_quantization_b200.utils import conv2d_func as cf
    torch.manual_seed(K + stride)
    dev = torch.device("cuda:0")
    N, H = 3, 21
    x = torch.randn(N, 3, H, H, device=dev) * 2
    ka = float(x.abs().max() / 15.5)
    conv = cf.conv2d_Q(qbit, 0.02, ka)(3, K, 3, stride=stride, padding=1).to(dev)
    conv.Kw = torch.tensor(float(conv.weight.detach().abs().max() / 15.5))
    bn = nn.BatchNorm2d(K).to(dev).eval()
    bn.weight.data.uniform_(0.5, 1.5); bn.bias.data.normal_(0, 0.5); bn.running_mean.normal_(0, 0.3); bn.running_var.uniform_(0.5, 1.5)
    P = engine.Plan(N, dev, qbit, e4m3=e4m3)
    xin = P.input_nchw(3, H, H)
    xin.copy_(x)
    xc = P.quantize_input(xin, engine._k32(conv.Ka))
    assert xc.cp == 4
    ks = [0.23, 0.31] if two else [0.23]
    out = P.conv(xc, conv, bn=bn, relu=relu, codes=ks, relu_codes=True, signed_fast=True)
    P.run()
    torch.cuda.synchronize()
    afmt, wfmt = orc.fmt_for(qbit, "act"), orc.fmt_for(qbit, "weight")
    _, xq = orc.quantize(x.cpu().numpy(), afmt, engine._k32(conv.Ka), want_codes=False)
    _, wq = orc.quantize(conv.weight.detach().cpu().numpy(), wfmt, engine._k32(conv.Kw), want_codes=False)
    wq = wq.astype(np.float16).astype(np.float64)                    # the kernel reads the float16 image of weight_q
    acc = torch.nn.functional.conv2d(torch.from_numpy(xq).double(), torch.from_numpy(wq), None, stride, 1).numpy()
    mul, add = P._affine(conv, bn, K)
    y = acc * mul.cpu().numpy()[None, :, None, None] + add.cpu().numpy()[None, :, None, None]
    if relu:
        y = np.maximum(y, 0)
    y = y.transpose(0, 2, 3, 1)
    for kd in ks:
        t = out["codes"][kd]
        if e4m3:
            assert t.fmt == nv.FMT_E4M3
        elif qbit == 7:
            assert t.fmt == (nv.FMT_SFP33_RELU if relu else nv.FMT_SFP33_SFAST)
        elif relu:
            assert t.fmt == nv.FMT_SLFP34_RELU
        codes = t.buf.cpu().numpy()
        assert (codes[..., K:] == 0).all()
        got = _h(_decode_any(orc, codes[..., :K], t.fmt))
        q = y / float(np.float32(kd))
        ok = np.zeros(q.shape, bool)
        for eps in (0.0, -3e-6, 3e-6):
            _, want = orc.quantize((q * (1.0 + eps)).astype(np.float32), afmt, want_codes=False)
            ok |= got == _h(want)
        assert ok.all(), (kd, int((~ok).sum()), ok.size)


@pytest.mark.parametrize("e4m3", [False, True])
@pytest.mark.parametrize("C,H,stride", [(58, 14, 1), (24, 20, 2), (116, 9, 1), (232, 6, 2), (116, 28, 1)])
def test_shufflenet_branch_fast_forms_equal_generic_forms(orc, C, H, stride, e4m3):
    """One ShuffleNetV2 residual branch (1x1 -> BN -> layerout -> ReLU -> dw 3x3 -> BN -> 1x1 -> BN -> layerout -> ReLU,
    58 channels: not a multiple of 16) through the fast forms - output channels padded to 64 for the vectorised
    epilogues, relu(quantize_layerout(.)) as three FMA-pipe operations, signed fast codes out of the depthwise conv -
    against the generic exact epilogues: the float16 branch output must agree (identical up to the code-boundary
    slivers of the two fast encoders)."""
    import torch.nn as nn
    from cnns_slfp_quantization_b200 import engine, _native as nv
    from cnns_slfp_quantization_b200.utils import conv2d_func as cf
    torch.manual_seed(11)
    dev = torch.device("cuda:0")
    N = 2
    Cs = engine.Plan._cp(C)

    def mk(kind):
        m = (cf.conv2d_Q(7, 0.05, 0.2)(C, C, 1) if kind == "pw" else cf.conv2d_Q(7, 0.05, 0.2)(C, C, 3, stride=stride, padding=1, groups=C)).to(dev)
        m.Kw = torch.tensor(float(m.weight.detach().abs().max() / 15.5))
        return m

    def mkbn():
        bn = nn.BatchNorm2d(C).to(dev).eval()
        bn.weight.data.uniform_(0.5, 1.5); bn.bias.data.normal_(0, 0.5); bn.running_mean.normal_(0, 0.3); bn.running_var.uniform_(0.5, 1.5)
        return bn
    c0, dw, c2 = mk("pw"), mk("dw"), mk("pw")
    b0, b1, b2 = mkbn(), mkbn(), mkbn()
    src = torch.rand(N, H, H, Cs, device=dev).half() * 3
    outs = []
    for fast in (True, False):
        P = engine.Plan(N, dev, 7, e4m3=e4m3)
        t = engine._T(src, N, H, H, Cs, Cs, "f16")
        t.c_logical = C
        xc = P.gather_quantize([(t, (j * 7) % C) for j in range(C)], engine._k32(c0.Ka))
        lo = 2 if fast else 1
        kd = engine._k32(0.2)
        a = P.conv(xc, c0, bn=b0, relu=True, layerout=lo, codes=[kd], pad_k=fast)["codes"][kd]
        b = P.conv(a, dw, bn=b1, relu=False, codes=[kd], relu_codes=False, signed_fast=fast)["codes"][kd]
        if e4m3:
            assert a.fmt == nv.FMT_E4M3 and b.fmt == nv.FMT_E4M3
        elif fast:
            assert a.fmt == nv.FMT_SFP33 and b.fmt == nv.FMT_SFP33_SFAST
        y = P.conv(b, c2, bn=b2, relu=True, layerout=lo, f16=True, pad_k=fast)["f16"]
        P.run()
        torch.cuda.synchronize()
        outs.append((_decode_any(orc, a.buf.cpu().numpy()[..., :C], a.fmt), _decode_any(orc, b.buf.cpu().numpy()[..., :C], b.fmt),
                     y.buf.float().cpu().numpy()[..., :C]))
    (af, bf, yf), (ag, bg, yg) = outs
    assert np.isfinite(yg).all() and np.isfinite(yf).all()
    assert (_h(af) == _h(ag)).mean() > 0.999                       # same grid values but for boundary slivers (reciprocal, ties)
    assert (_h(bf) == _h(bg)).mean() > 0.99
    # layer-out values live on the SFP<4,4> grid: float16 holds them exactly
    _, lq = orc.quantize(yf, 3, want_codes=False, bugcompat=False)
    assert (lq == yf).all()
    assert (yf == yg).mean() > 0.97 and np.abs(yf - yg).max() <= 0.15 * np.abs(yg).max()


@pytest.mark.parametrize("N,C,H,K,k,stride,pad,res", [
    (2, 64, 14, 64, 3, 1, 1, False),        # stage-1 3x3 (64-column tiles)
    (2, 128, 12, 128, 3, 2, 1, False),      # strided 3x3
    (1, 256, 9, 256, 3, 1, 1, False),       # 36 K blocks
    (2, 64, 11, 256, 1, 1, 0, True),        # block tail: residual in, float16 + two code tensors out (staged epilogue)
    (2, 512, 7, 2048, 1, 1, 0, True),       # 8 K blocks, 16 N tiles
    (1, 1024, 5, 256, 1, 1, 0, False),      # 1x1 reduce, 16 K blocks
    (2, 192, 6, 320, 3, 1, 1, False),       # ragged N (2.5 tiles of 128), 27 K blocks
    (3, 64, 7, 1024, 1, 1, 0, True),        # M tail (147 rows)
])
def test_f16_image_activations_equal_the_codes_path(orc, N, C, H, K, k, stride, pad, res):
    """SLFP_FMT_F16Q: a dense layer fed with the float16 image of its input codes (TMA -> swizzled shared memory -> MMA, no
    decode stage) produces bit-identical outputs to the same layer fed with the codes; SlfpEpilogue.store_f16 writes
    exactly float16(decode(code)) of the code the byte path stores."""
    from cnns_slfp_quantization_b200 import _native as nv
    lib, dev, st = nv.lib(), torch.device("cuda:0"), nv.stream()
    rng = np.random.default_rng(N * 1000 + C + K)
    rfmt = nv.FMT_SLFP34_RELU
    codes_in = rng.integers(0, 256, (N, H, H, C), dtype=np.uint8)
    codes_in[rng.random(codes_in.shape) < 0.3] = 0
    xv = orc.decode_relu(codes_in, False)
    xc = torch.from_numpy(codes_in).to(dev)
    xh = torch.from_numpy(_h(xv)).to(dev)
    w = torch.from_numpy((rng.standard_normal((K, C, k, k)) * 0.2).astype(np.float32)).to(dev)
    kw = float(w.abs().max() / 15.5)
    Ho = (H + 2 * pad - (k - 1) - 1) // stride + 1
    mul_t = torch.from_numpy((rng.uniform(0.5, 1.5, K) * kw * 0.05).astype(np.float32)).to(dev)
    add_t = torch.from_numpy((rng.standard_normal(K) * 0.3).astype(np.float32)).to(dev)
    res_t = torch.from_numpy(np.abs(rng.standard_normal((N, Ho, Ho, K))).astype(np.float16)).to(dev) if res else None

    def run(f16_in, store_f16):
        d = nv.SlfpConvDesc(N, H, H, C, C, K, k, k, stride, stride, pad, pad, 1, 1, 1, nv.FMT_F16Q if f16_in else rfmt, 0, 0)
        pitch = lib.slfp_conv_wpitch(ctypes.byref(d))
        wh = torch.empty((K * pitch,), dtype=torch.float16, device=dev)
        so, sc, sr, ss = w.stride()
        nv.check(lib.slfp_prepare_weights(ctypes.byref(d), w.data_ptr(), so, sc, sr, ss, float(np.float32(kw)), nv.FMT_SLFP34_WGT,
                                          wh.data_ptr(), None, None, st))
        e = nv.SlfpEpilogue()
        e.ch_mul, e.ch_add, e.relu = mul_t.data_ptr(), add_t.data_ptr(), 1
        outs = {}
        outs["c1"] = torch.full((N, Ho, Ho, K), 7, dtype=torch.float16 if store_f16 else torch.uint8, device=dev)
        e.y_codes, e.next_k_div, e.next_fmt, e.k_phys_out, e.store_f16 = outs["c1"].data_ptr(), 0.23, rfmt, K, 1 if store_f16 else 0
        if res:
            outs["c2"] = torch.full((N, Ho, Ho, K), 7, dtype=torch.uint8, device=dev)
            outs["y16"] = torch.full((N, Ho, Ho, K), 7, dtype=torch.float16, device=dev)
            e.y_codes2, e.next_k_div2, e.y_f16 = outs["c2"].data_ptr(), 0.37, outs["y16"].data_ptr()
            e.residual, e.residual_f16 = res_t.data_ptr(), 1
        nv.check(lib.slfp_conv2d_fwd(ctypes.byref(d), (xh if f16_in else xc).data_ptr(), wh.data_ptr(), ctypes.byref(e), st))
        torch.cuda.synchronize()
        return {k_: v.cpu().numpy() for k_, v in outs.items()}

    a, b = run(False, False), run(True, False)
    assert a["c1"].std() > 10, "degenerate case: the output codes barely vary"
    for key in a:
        assert (a[key].view(np.uint8) == b[key].view(np.uint8)).all(), (key, float((a[key] != b[key]).mean()))
    if not res:
        want = _h(orc.decode_relu(a["c1"], False))
        c_, d_ = run(False, True), run(True, True)
        assert (c_["c1"].view(np.uint16) == want.view(np.uint16)).all()
        assert (d_["c1"].view(np.uint16) == want.view(np.uint16)).all()


def _wprep_case(nv, lib, dev, w, Cp, fmt, out_kind, groups=1, row_scale=None, out_pitch=0, out_off=0):
    """One slfp_prepare_weights_jobs call; returns (float16 / e4m3-byte operand, codes) as numpy arrays."""
    K, Cg, R, S = w.shape
    C = Cg * groups
    flags = nv.CONV_E4M3_OPERANDS if out_kind == "e4m3" else 0
    d = nv.SlfpConvDesc(1, 8, 8, C, Cp, K, R, S, 1, 1, 0, 0, 1, 1, groups, nv.FMT_SLFP34_ACT, 0, 0, flags)
    pitch = lib.slfp_conv_wpitch(ctypes.byref(d))
    rowlen = out_pitch or pitch
    f16 = torch.full((K * rowlen,), 7.0, dtype=torch.float16, device=dev) if out_kind in ("f16", "both", "e4m3") else None
    codes = torch.full((K * rowlen,), 0x55, dtype=torch.uint8, device=dev) if out_kind in ("codes", "both") else None
    job = (nv.SlfpWeightJob * 1)()
    job[0].desc, job[0].w, job[0].kw = ctypes.pointer(d), w.data_ptr(), float(np.float32(0.07))
    job[0].w_stride[:] = w.stride()
    job[0].w_f16 = f16.data_ptr() if f16 is not None else None
    job[0].w_codes = codes.data_ptr() if codes is not None else None
    job[0].out_pitch, job[0].out_offset = out_pitch, out_off
    job[0].row_scale = row_scale.data_ptr() if row_scale is not None else None
    nv.check(lib.slfp_prepare_weights_jobs(1, job, fmt, nv.stream()))
    torch.cuda.synchronize()
    return (None if f16 is None else f16.view(torch.int16).cpu().numpy().copy(),
            None if codes is None else codes.cpu().numpy().copy(), pitch)


@pytest.mark.parametrize("shape", [
    # K, C/groups, R, S, c_phys, groups
    (64, 64, 1, 1, 64, 1),            # many rows per CTA
    (70, 64, 3, 3, 64, 1),            # 8 rows per CTA, ragged last group
    (24, 256, 3, 3, 256, 1),          # two rows per CTA
    (9, 512, 3, 3, 512, 1),           # one row per CTA (4 608 elements)
    (5, 1024, 3, 3, 1024, 1),         # a row split over channel pieces
    (6, 2048, 1, 1, 2048, 1),         # 1x1, two rows per CTA
    (3, 9216, 1, 1, 9216, 1),         # linear row longer than a CTA's capacity
    (16, 24, 3, 3, 32, 1),            # padding channels + K-padding taps
    (12, 20, 5, 5, 32, 1),            # 25 taps, padding channels
    (40, 1, 3, 3, 4, 40),             # depthwise: not eligible (one channel per group) -> gather form on both sides
    (32, 8, 3, 3, 8, 4),              # grouped, 8 channels per group
    (8, 3, 7, 7, 4, 1),               # the RGB stem: gather form
])
@pytest.mark.parametrize("fmt_name,out_kind", [("wgt", "f16"), ("wgt", "both"), ("wgt", "codes"), ("sfp33", "e4m3"), ("sfp33", "both"),
                                               ("act", "f16")])
def test_row_staged_weight_preparation_equals_gather_form(orc, shape, fmt_name, out_kind):
    """wprep_rows_kernel (coalesced source rows staged through shared memory) against the per-element gather kernels on
    the same jobs (SLFP_WPREP_GATHER=1), bit for bit, and the codes against the oracle's encoder; special values
    (zeros, NaN, infinities, tiny and huge magnitudes) are sprinkled into the weights."""
    from cnns_slfp_quantization_b200 import _native as nv
    lib = nv.lib()
    dev = torch.device("cuda:0")
    K, Cg, R, S, Cp, groups = shape
    fmt = {"wgt": nv.FMT_SLFP34_WGT, "sfp33": nv.FMT_SFP33, "act": nv.FMT_SLFP34_ACT}[fmt_name]
    if out_kind == "e4m3" and (groups > 1 or Cp % 16):
        pytest.skip("e4m3 operands are dense")
    rng = np.random.default_rng(K * 131 + Cg)
    w_np = (rng.standard_normal((K, Cg, R, S)) * 0.05).astype(np.float32)
    flat = w_np.reshape(-1)
    special = np.array([0.0, -0.0, np.nan, np.inf, -np.inf, 1e-30, -1e-30, 3e30, 1e-12, 0.07 * 15.32165, 0.07 * 0.0625, 0.07 * 0.125],
                       np.float32)
    pos = rng.choice(flat.size, size=min(flat.size // 2, 96), replace=False)
    flat[pos] = special[np.arange(pos.size) % special.size]
    w = torch.from_numpy(w_np).to(dev)
    row_scale = torch.from_numpy(rng.uniform(0.5, 2.0, K).astype(np.float32)).to(dev) if out_kind == "f16" and K % 2 == 0 else None
    pitch0 = R * S * (Cp if groups == 1 else Cg)
    pitch0 = pitch0 if groups > 1 else (pitch0 + 63) // 64 * 64
    out_pitch, out_off = (pitch0 + 128, 64) if (out_kind == "f16" and groups == 1) else (0, 0)
    try:
        os.environ.pop("SLFP_WPREP_GATHER", None)
        fast = _wprep_case(nv, lib, dev, w, Cp, fmt, out_kind, groups, row_scale, out_pitch, out_off)
        os.environ["SLFP_WPREP_GATHER"] = "1"
        gather = _wprep_case(nv, lib, dev, w, Cp, fmt, out_kind, groups, row_scale, out_pitch, out_off)
    finally:
        os.environ.pop("SLFP_WPREP_GATHER", None)
    for a, b in zip(fast[:2], gather[:2]):
        assert (a is None) == (b is None)
        if a is not None:
            if out_pitch:                                   # bytes outside [out_off, out_off + pitch) stay untouched on both sides
                assert (a.reshape(K, out_pitch)[:, :out_off] == 0x4700).all()
            assert np.array_equal(a, b), int((a != b).sum())
    if fast[1] is not None and groups == 1:
        codes = fast[1].reshape(K, fast[2])[:, :R * S * Cp].reshape(K, R * S, Cp)[..., :Cg]
        want, _ = orc.quantize(np.ascontiguousarray(w_np.reshape(K, Cg, R * S).transpose(0, 2, 1)),
                               {"wgt": 2, "sfp33": 0, "act": 1}[fmt_name], 0.07)
        assert np.array_equal(codes, want)


@pytest.mark.parametrize("fmt_name", ["act", "sfp33"])
def test_fused_avgpool_quantizer_equals_two_launches(orc, fmt_name):
    """slfp_avgpool_quantize_nhwc_f16 against slfp_avgpool_nhwc + slfp_quantize_nhwc_f32: same float32 means, same code
    bytes, and the codes equal the oracle's quantizer of the means."""
    from cnns_slfp_quantization_b200 import _native as nv
    lib = nv.lib()
    dev = torch.device("cuda:0")
    fmt = {"act": nv.FMT_SLFP34_ACT, "sfp33": nv.FMT_SFP33}[fmt_name]
    rng = np.random.default_rng(77)
    for (n, hw, c) in [(5, 49, 2048), (3, 16, 1024), (2, 1, 64), (4, 7, 24)]:
        x = torch.from_numpy((np.abs(rng.standard_normal((n, hw, c))) * 2).astype(np.float16)).to(dev)
        x[0, :, :8] = 0
        kdiv = float(np.float32(0.21))
        mean_a = torch.empty((n, c), dtype=torch.float32, device=dev)
        codes_a = torch.empty((n, c), dtype=torch.uint8, device=dev)
        nv.check(lib.slfp_avgpool_nhwc(x.data_ptr(), 1, n, hw, c, mean_a.data_ptr(), nv.stream()))
        nv.check(lib.slfp_quantize_nhwc_f32(mean_a.data_ptr(), n, c, c, kdiv, fmt, codes_a.data_ptr(), nv.stream()))
        mean_b = torch.full((n, c), -1.0, dtype=torch.float32, device=dev)
        codes_b = torch.full((n, c), 0x55, dtype=torch.uint8, device=dev)
        nv.check(lib.slfp_avgpool_quantize_nhwc_f16(x.data_ptr(), n, hw, c, mean_b.data_ptr(), kdiv, fmt, codes_b.data_ptr(), nv.stream()))
        codes_c = torch.full((n, c), 0x55, dtype=torch.uint8, device=dev)
        nv.check(lib.slfp_avgpool_quantize_nhwc_f16(x.data_ptr(), n, hw, c, None, kdiv, fmt, codes_c.data_ptr(), nv.stream()))
        torch.cuda.synchronize()
        assert torch.equal(mean_a.view(torch.int32), mean_b.view(torch.int32))
        assert torch.equal(codes_a, codes_b) and torch.equal(codes_a, codes_c)
        want, _ = orc.quantize(mean_a.cpu().numpy(), {"act": 1, "sfp33": 0}[fmt_name], kdiv)
        assert np.array_equal(codes_b.cpu().numpy(), want)


def test_row_staged_weight_encoder_every_mantissa(orc):
    """The one-look-up SLFP<3,4> weight encoder of wprep_rows_kernel against the oracle for EVERY float32 mantissa of the
    exponents around and inside the code range (2^-6 .. 2^4), both signs: codes bit-exact, float16 operand = float16 of
    the decoded value."""
    from cnns_slfp_quantization_b200 import _native as nv
    lib = nv.lib()
    dev = torch.device("cuda:0")
    K, C = 2048, 4096                                        # 2^23 weights per tensor = one binade
    d = nv.SlfpConvDesc(1, 1, 1, C, C, K, 1, 1, 1, 1, 0, 0, 1, 1, 1, nv.FMT_SLFP34_ACT)
    mant = np.arange(1 << 23, dtype=np.uint32)
    codes = torch.empty(K * C, dtype=torch.uint8, device=dev)
    f16 = torch.empty(K * C, dtype=torch.float16, device=dev)
    for e in range(120, 132):
        bits = mant | np.uint32(e << 23)
        bits[1::2] |= np.uint32(0x80000000)
        w = torch.from_numpy(bits.view(np.float32).reshape(K, C, 1, 1)).to(dev)
        nv.check(lib.slfp_prepare_weights(ctypes.byref(d), w.data_ptr(), *w.stride(), 1.0, nv.FMT_SLFP34_WGT, f16.data_ptr(), codes.data_ptr(),
                                          None, nv.stream()))
        torch.cuda.synchronize()
        want, fq = orc.quantize(bits.view(np.float32), 2, 1.0)
        got = codes.cpu().numpy()
        assert np.array_equal(got, want), (e, int((got != want).sum()))
        assert np.array_equal(f16.view(torch.int16).cpu().numpy(), fq.astype(np.float16).view(np.int16)), e


def test_host_pipeline_equals_the_single_graph():
    """Plan.submit_host (H2D straight into the input buffer on a copy stream, head / tail graphs, logits D2H) gives the logits
    of the single-graph path bit for bit, for a sequence of DIFFERENT batches submitted back to back (the copy of batch i+1
    overlaps the layers of batch i and must not disturb them)."""
    sys.path.insert(0, ROOT)
    import bench
    from cnns_slfp_quantization_b200 import engine, nets_common as nc
    dev = torch.device("cuda", 0)
    model = bench.build_model_gpu(224, dev)
    plan = engine.compile_resnet50(model, 32, 224, device=dev)
    batches = [nc.synth_images(32, 224, seed=100 + i) for i in range(5)]
    plan.capture()
    want = []
    for b in batches:
        plan(b.to(dev))
        torch.cuda.synchronize()
        want.append(plan.output.clone())
    assert not torch.equal(want[0], want[1])
    hosts = [b.pin_memory() for b in batches]
    outs = [torch.empty(tuple(plan.output.shape), dtype=torch.float32).pin_memory() for _ in batches]
    for rep in range(3):
        for h, o in zip(hosts, outs):
            plan.submit_host(h, o)
        torch.cuda.synchronize()
        for i, (o, w) in enumerate(zip(outs, want)):
            assert torch.equal(o, w.cpu()), (rep, i)
