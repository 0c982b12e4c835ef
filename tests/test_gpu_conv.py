"""GPU: convolution forward / backward through the C ABI against the oracle and the
reference-generated fixtures.

Tolerances (floating point, stated here as the north star requires):
  * against the oracle evaluated on the SAME float16-rounded operands the tensor cores see
    (double accumulation):  |y - y16| <= 2e-6 * L1 + 1e-6,  L1 = conv(|x_q|, |w_q|) * Ka * Kw
    -- only fp32 accumulation order separates the two;
  * against the reference's float32 result: |y - y_ref| <= 1.2e-3 * L1 (SLFP-8: the irrational
    grid 2^(j/16) is rounded to float16, rel. error <= 2^-11 per operand) and <= 4e-6 * L1 for
    SFP-7, whose grid is exact in float16.
"""
import numpy as np
import pytest
import torch

from conftest import same_bits

pytestmark = pytest.mark.gpu


def _case(g, name):
    return g[name + ".cfg"], {k[len(name) + 1:]: g[k] for k in g.files if k.startswith(name + ".")}


def _check_fwd(orc, x, w, b, ka, kw, qbit, st, pad, dil, groups, y_ref=None, tag=""):
    from gpu_util import conv_fwd_gpu
    xq, wq, y_or = orc.conv2d_Q_forward(x, w, b, ka, kw, qbit, st, pad, dil, groups)
    bq = None if b is None else ((np.asarray(b, np.float32) / np.float32(ka)) / np.float32(kw)).astype(np.float32)
    out = conv_fwd_gpu(x, w, bq, ka, kw, qbit, st, pad, dil, groups)
    y = out["y"]
    assert np.isfinite(y).all(), tag
    l1 = orc.conv2d_q(np.abs(xq), np.abs(wq), None, st, pad, dil, groups, ka, kw)
    if groups == 1:
        x16 = xq.astype(np.float16).astype(np.float32)
        w16 = wq.astype(np.float16).astype(np.float32)
        y16 = orc.conv2d_q(x16, w16, bq, st, pad, dil, groups, ka, kw)
        err16 = np.abs(y - y16)
        assert (err16 <= 2e-6 * l1 + 1e-6).all(), (tag, float((err16 / (l1 + 1e-9)).max()))
        tol = 4e-6 if qbit == 7 else 1.2e-3
    else:
        tol = 4e-6                      # the stencil path multiplies exact float32 operands
    err = np.abs(y - y_or)
    assert (err <= tol * l1 + 1e-6).all(), (tag, float((err / (l1 + 1e-9)).max()))
    if y_ref is not None:
        err = np.abs(y - y_ref)
        assert (err <= tol * l1 + 1e-5 * np.abs(y_ref).max()).all(), (tag, "vs reference fixture")
    return out, (xq, wq)


def test_golden_conv_cases_forward(orc, g_conv):
    for name in g_conv["names"]:
        name = str(name)
        if name.startswith("fc_") or name.endswith("fp32"):
            continue
        cfg, d = _case(g_conv, name)
        qbit, N, C, H, W, O, k, st, pad, dil, groups, has_bias = [int(v) for v in cfg]
        ka, kw = d["k"]
        out, (xq, wq) = _check_fwd(orc, d["x"], d["w"], d.get("b"), ka, kw, qbit, st, pad, dil, groups, d["y"], name)
        # the codes the kernels produced decode bit-exactly to the reference's input_q / weight_q
        Cp = out["x_codes"].shape[-1]
        xdec = orc.decode(out["x_codes"], 0 if qbit == 7 else 1)[..., :C].transpose(0, 3, 1, 2)
        assert same_bits(d["input_q"], xdec).all(), name
        cw = Cp if groups == 1 else C // groups
        wdec = orc.decode(out["w_codes"][:, :k * k * cw], 0 if qbit == 7 else 2).reshape(O, k, k, cw)[..., :C // groups]
        assert same_bits(d["weight_q"], wdec.transpose(0, 3, 1, 2)).all(), name


SHAPES = [
    # N, C, H, W, K, k, stride, pad, dil   (dense)
    (2, 64, 14, 14, 64, 1, 1, 0, 1),        # 1x1, single K block, BLOCK_N 64
    (1, 64, 12, 12, 256, 1, 1, 0, 1),       # BLOCK_N 256
    (3, 256, 7, 7, 64, 1, 1, 0, 1),         # 4 K blocks
    (2, 64, 10, 10, 64, 3, 1, 1, 1),        # 3x3, 9 K blocks
    (2, 128, 9, 9, 128, 3, 2, 1, 1),        # stride 2, BLOCK_N 128
    (1, 32, 17, 13, 48, 3, 1, 1, 1),        # Cin 32 (half a K block per tap), odd sizes, N tile 64 with K=48
    (2, 48, 8, 8, 24, 3, 1, 1, 1),          # BLOCK_N 32 with K=24
    (1, 512, 5, 5, 520, 1, 1, 0, 1),        # three N tiles of 256, last one ragged
    (2, 3, 33, 33, 64, 7, 2, 3, 1),         # the stem: Cin 3 -> c_phys 4, GRAN 4
    (1, 16, 20, 20, 32, 5, 1, 2, 1),        # 5x5
    (1, 16, 16, 16, 16, 3, 1, 2, 2),        # dilation 2
    (5, 24, 6, 6, 58, 1, 1, 0, 1),          # ShuffleNet-like odd channels (padded to 32)
    (300, 16, 3, 3, 16, 3, 1, 1, 1),        # many images, many tiles (> SM count on small GPUs)
]


# The distinct layer shapes of the BASELINE configs at batch 2: full channel counts and K depths
# (up to K = 4 608), spatial sizes of the 224x224 / 32x32 nets.  N, C, H, W, K, k, stride, pad, dil
CONFIG_SHAPES = [
    (2, 64, 56, 56, 64, 1, 1, 0, 1), (2, 64, 56, 56, 64, 3, 1, 1, 1), (2, 64, 56, 56, 256, 1, 1, 0, 1), (2, 256, 56, 56, 64, 1, 1, 0, 1),
    (2, 256, 56, 56, 128, 1, 1, 0, 1), (2, 128, 56, 56, 128, 3, 2, 1, 1), (2, 128, 28, 28, 512, 1, 1, 0, 1), (2, 256, 56, 56, 512, 1, 2, 0, 1),
    (2, 512, 28, 28, 128, 1, 1, 0, 1), (2, 128, 28, 28, 128, 3, 1, 1, 1), (2, 512, 28, 28, 256, 1, 1, 0, 1), (2, 256, 28, 28, 256, 3, 2, 1, 1),
    (2, 256, 14, 14, 1024, 1, 1, 0, 1), (2, 512, 28, 28, 1024, 1, 2, 0, 1), (2, 1024, 14, 14, 256, 1, 1, 0, 1), (2, 256, 14, 14, 256, 3, 1, 1, 1),
    (2, 1024, 14, 14, 512, 1, 1, 0, 1), (2, 512, 14, 14, 512, 3, 2, 1, 1), (2, 512, 7, 7, 2048, 1, 1, 0, 1), (2, 1024, 14, 14, 2048, 1, 2, 0, 1),
    (2, 2048, 7, 7, 512, 1, 1, 0, 1), (2, 512, 7, 7, 512, 3, 1, 1, 1), (4, 2048, 1, 1, 1000, 1, 1, 0, 1),          # ResNet-50
    (2, 3, 224, 224, 64, 7, 2, 3, 1),                                                                              # its 7x7 / 2 stem
    (4, 3, 32, 32, 64, 3, 1, 1, 1), (4, 64, 32, 32, 64, 3, 1, 1, 1), (4, 128, 16, 16, 256, 3, 1, 1, 1), (4, 256, 8, 8, 512, 3, 1, 1, 1),
    (8, 512, 2, 2, 512, 3, 1, 1, 1),                                                                               # VGG-16 CIFAR
    (2, 3, 224, 224, 32, 3, 2, 1, 1), (2, 32, 112, 112, 64, 1, 1, 0, 1), (2, 512, 14, 14, 512, 1, 1, 0, 1), (2, 1024, 7, 7, 1024, 1, 1, 0, 1),  # MobileNetV1
    (2, 3, 224, 224, 64, 11, 4, 2, 1), (2, 64, 27, 27, 192, 5, 1, 2, 1), (2, 192, 13, 13, 384, 3, 1, 1, 1), (4, 9216, 1, 1, 4096, 1, 1, 0, 1),  # AlexNet
    (2, 3, 224, 224, 96, 7, 2, 0, 1), (2, 96, 54, 54, 16, 1, 1, 0, 1), (2, 16, 54, 54, 64, 3, 1, 1, 1), (2, 512, 13, 13, 1000, 1, 1, 0, 1),     # SqueezeNet 1.0
    (2, 24, 56, 56, 58, 1, 1, 0, 1), (2, 116, 28, 28, 116, 1, 1, 0, 1), (2, 464, 7, 7, 1024, 1, 1, 0, 1),          # ShuffleNetV2 (odd channel counts)
]


@pytest.mark.parametrize("shape", CONFIG_SHAPES)
def test_config_layer_shapes_vs_oracle(orc, shape):
    """Every distinct layer shape of the BASELINE configs through the C ABI against
    the oracle's double-accumulating convolution of the reference's fake-quant operands: 1.2e-3 * L1 (SLFP-8), and
    2e-6 * L1 against the same convolution of the float16-rounded operands (accumulation order only)."""
    N, C, H, W, K, k, st, pad, dil = shape
    rng = np.random.default_rng(abs(hash(shape)) % (1 << 31))
    x = (rng.standard_normal((N, C, H, W)) * 2).astype(np.float32)
    w = (rng.standard_normal((K, C, k, k)) * 0.3).astype(np.float32)
    b = (rng.standard_normal(K) * 0.5).astype(np.float32) if (k in (5, 11) or C == 9216) else None
    ka, kw = float(np.abs(x).max() / 15.5), float(np.abs(w).max() / 15.5)
    _check_fwd(orc, x, w, b, ka, kw, 8, st, pad, dil, 1, None, str(shape))


@pytest.mark.parametrize("qbit", [8, 7])
@pytest.mark.parametrize("shape", SHAPES)
def test_dense_shapes_vs_oracle(orc, shape, qbit):
    N, C, H, W, K, k, st, pad, dil = shape
    rng = np.random.default_rng(abs(hash(shape)) % (1 << 31))
    x = (rng.standard_normal((N, C, H, W)) * 2).astype(np.float32)
    w = (rng.standard_normal((K, C, k, k)) * 0.3).astype(np.float32)
    ka, kw = float(np.abs(x).max() / 15.5), float(np.abs(w).max() / 15.5)
    _check_fwd(orc, x, w, None, ka, kw, qbit, st, pad, dil, 1, None, str(shape))


def test_many_tiles_persistent_schedule(orc):
    """More tiles than SMs and several K blocks: exercises the smem ring across tile boundaries and
    both TMEM accumulator buffers many times."""
    rng = np.random.default_rng(11)
    x = (rng.standard_normal((64, 64, 28, 28)) * 2).astype(np.float32)     # M = 50176 -> 392 tiles
    w = (rng.standard_normal((128, 64, 3, 3)) * 0.2).astype(np.float32)
    ka, kw = float(np.abs(x).max() / 15.5), float(np.abs(w).max() / 15.5)
    from gpu_util import conv_fwd_gpu
    out = conv_fwd_gpu(x, w, None, ka, kw, 8, 1, 1, 1, 1)
    # float32 torch reference on the exact fake-quant operands (CPU, fp32)
    _, xq = orc.quantize(x, 1, ka, want_codes=False)
    _, wq = orc.quantize(w, 2, kw, want_codes=False)
    x16 = torch.from_numpy(xq.astype(np.float16).astype(np.float32)).double()
    w16 = torch.from_numpy(wq.astype(np.float16).astype(np.float32)).double()
    y16 = (torch.nn.functional.conv2d(x16, w16, None, 1, 1) * float(np.float32(ka)) * float(np.float32(kw))).numpy()
    l1 = (torch.nn.functional.conv2d(x16.abs(), w16.abs(), None, 1, 1) * ka * kw).numpy()
    err = np.abs(out["y"] - y16)
    assert (err <= 2e-6 * l1 + 1e-6).all(), float((err / l1).max())


def test_fused_epilogue(orc):
    """bias + post-scale + folded-BN affine + residual + ReLU + fp16 + quantize-on-store (two scales)."""
    from gpu_util import conv_fwd_gpu
    rng = np.random.default_rng(21)
    N, C, H, W, K = 2, 64, 9, 9, 80
    x = (rng.standard_normal((N, C, H, W)) * 2).astype(np.float32)
    w = (rng.standard_normal((K, C, 3, 3)) * 0.2).astype(np.float32)
    ka, kw = float(np.abs(x).max() / 15.5), float(np.abs(w).max() / 15.5)
    bq = rng.standard_normal(K).astype(np.float32)
    sc = rng.uniform(0.5, 1.5, K).astype(np.float32)
    sh = rng.standard_normal(K).astype(np.float32) * 0.3
    res = rng.standard_normal((N, K, H, W)).astype(np.float32)
    for res_arr in (res, res.astype(np.float16)):
        out = conv_fwd_gpu(x, w, bq, ka, kw, 8, 1, 1, 1, 1,
                           dict(ch_scale=sc, ch_shift=sh, residual=res_arr, relu=True, want_f16=True, next_k=0.21, next_k2=0.4))
        y = out["y"]
        _, xq = orc.quantize(x, 1, ka, want_codes=False)
        _, wq = orc.quantize(w, 2, kw, want_codes=False)
        x16, w16 = xq.astype(np.float16).astype(np.float32), wq.astype(np.float16).astype(np.float32)
        base = orc.conv2d_q(x16, w16, bq, 1, 1, 1, 1, ka, kw)
        want = np.maximum(base * sc[None, :, None, None] + sh[None, :, None, None] + res_arr.astype(np.float32), 0)
        l1 = orc.conv2d_q(np.abs(x16), np.abs(w16), np.abs(bq), 1, 1, 1, 1, ka, kw) * 1.5 + 1
        assert (np.abs(y - want) <= 4e-6 * l1).all()
        # secondary outputs are exact functions of the float32 output
        assert (out["y_f16"] == y.astype(np.float16)).all()
        yn = y.transpose(0, 2, 3, 1)
        for key, kd in (("y_codes", 0.21), ("y_codes2", 0.4)):
            oc, _ = orc.quantize(yn, 1, kdiv=kd)
            assert (out[key][..., :K] == oc).all() and (out[key][..., K:] == 0).all()


@pytest.mark.parametrize("qbit", [8, 7])
def test_depthwise_and_grouped(orc, qbit):
    rng = np.random.default_rng(31)
    for (N, C, H, W, K, k, st, pad, groups) in [(2, 32, 12, 12, 32, 3, 1, 1, 32), (2, 24, 11, 11, 24, 3, 2, 1, 24),
                                                 (1, 116, 7, 7, 116, 3, 1, 1, 116), (2, 32, 6, 6, 64, 3, 1, 1, 4),
                                                 (1, 6, 5, 5, 6, 3, 1, 1, 6)]:
        x = (rng.standard_normal((N, C, H, W)) * 2).astype(np.float32)
        w = (rng.standard_normal((K, C // groups, k, k)) * 0.3).astype(np.float32)
        ka, kw = float(np.abs(x).max() / 15.5), float(np.abs(w).max() / 15.5)
        _check_fwd(orc, x, w, None, ka, kw, qbit, st, pad, 1, groups, None, f"g{groups} C{C}")


def test_golden_conv_cases_backward(orc, g_conv):
    import ctypes
    from cnns_slfp_quantization_b200 import _native as nv
    from gpu_util import conv_fwd_gpu
    for name in g_conv["names"]:
        name = str(name)
        if name.startswith("fc_") or name.endswith("fp32"):
            continue
        cfg, d = _case(g_conv, name)
        qbit, N, C, H, W, O, k, st, pad, dil, groups, has_bias = [int(v) for v in cfg]
        ka, kw = d["k"]
        out = conv_fwd_gpu(d["x"], d["w"], None, ka, kw, qbit, st, pad, dil, groups)
        gy = torch.from_numpy(d["gy"]).cuda().permute(0, 2, 3, 1).contiguous()
        dx = torch.empty((N, H, W, C), dtype=torch.float32, device="cuda")
        dw = torch.empty((O, C // groups, k, k), dtype=torch.float32, device="cuda")
        db = torch.empty((O,), dtype=torch.float32, device="cuda")
        so, sc, sr, ss = dw.stride()
        nv.check(nv.lib().slfp_conv2d_bwd(ctypes.byref(out["desc"]), gy.data_ptr(), out["dev"]["xc"].data_ptr(),
                                          out["dev"]["wc"].data_ptr(), nv.fmt_for(qbit, "weight"), float(np.float32(ka)),
                                          float(np.float32(kw)), dx.data_ptr(), dw.data_ptr(), so, sc, sr, ss,
                                          db.data_ptr(), nv.stream()))
        torch.cuda.synchronize()
        # tolerance: float32 accumulation vs the reference's float32 cuDNN-free CPU result
        for got, want, what in ((dx.permute(0, 3, 1, 2).cpu().numpy(), d["dx"], "dx"), (dw.cpu().numpy(), d["dw"], "dw")):
            np.testing.assert_allclose(got, want, rtol=1e-4, atol=2e-5 * np.abs(want).max(), err_msg=f"{name} {what}")
        if has_bias:
            np.testing.assert_allclose(db.cpu().numpy(), d["db"], rtol=1e-4, atol=1e-4, err_msg=name)


def _bwd_both(out, gy, N, H, W, C, O, k, groups, qbit, ka, kw, use_ws):
    """dx / dw through slfp_conv2d_bwd (direct CUDA-core kernels) or slfp_conv2d_bwd_ws (tcgen05 implicit GEMMs)."""
    import ctypes
    from cnns_slfp_quantization_b200 import _native as nv
    lib = nv.lib()
    dx = torch.full((N, H, W, C), float("nan"), dtype=torch.float32, device="cuda")
    dw = torch.full((O, C // groups, k, k), float("nan"), dtype=torch.float32, device="cuda")
    so, sc, sr, ss = dw.stride()
    args = (ctypes.byref(out["desc"]), gy.data_ptr(), out["dev"]["xc"].data_ptr(), out["dev"]["wc"].data_ptr(),
            nv.fmt_for(qbit, "weight"), float(np.float32(ka)), float(np.float32(kw)), dx.data_ptr(), dw.data_ptr(), so, sc, sr, ss, None)
    if use_ws:
        nbytes = lib.slfp_conv2d_bwd_workspace_size(ctypes.byref(out["desc"]), 1, 1)
        assert nbytes > 0, "shape not covered by the tensor-core backward"
        ws = torch.empty((nbytes,), dtype=torch.uint8, device="cuda")
        nv.check(lib.slfp_conv2d_bwd_ws(*args, ws.data_ptr(), nbytes, nv.stream()))
    else:
        nv.check(lib.slfp_conv2d_bwd(*args, nv.stream()))
    torch.cuda.synchronize()
    return dx.cpu().numpy(), dw.cpu().numpy()


# Tensor-core backward (csrc/conv_bwd_sm100.cu).  Tolerance, stated: float16 operands (gy scaled by a power of two,
# float16 images of the codes) -> relative rounding 2^-11 per operand, so |d - d_ref| <= 1.2e-3 * L1 where L1 is the same
# sum with every term replaced by its magnitude (computed by the direct float32 kernels on |gy| and sign-stripped codes).
@pytest.mark.parametrize("qbit", [8, 7])
@pytest.mark.parametrize("shape", [
    # N, C, H, W, K, k, stride, pad
    (2, 64, 14, 14, 64, 1, 1, 0),        # plain 1x1, single unit
    (3, 64, 9, 11, 128, 3, 1, 1),        # 3x3, 9 units (groups of 3), ragged pixel tail
    (2, 128, 12, 12, 64, 3, 2, 1),       # strided 3x3: four parity classes (1 / 2 / 2 / 4 taps)
    (2, 256, 10, 10, 320, 1, 2, 0),      # strided 1x1: empty parity classes, K not a multiple of 64, three k tiles
    (2, 192, 7, 7, 96, 3, 1, 1),         # C = 192: 27 units, K = 96 (padded to 128)
    (1, 64, 20, 20, 64, 5, 1, 2),        # 25 taps
    (2, 64, 13, 13, 72, 3, 2, 0),        # stride 2 without padding, odd sizes
    (4, 512, 7, 7, 512, 3, 1, 1),        # ResNet stage-4 shape (small batch)
    (2, 30, 9, 9, 40, 3, 1, 1),          # c_phys = 32: dgrad on tensor cores (scalar stores, C % 4 != 0), wgrad direct
    (3, 24, 8, 8, 64, 1, 2, 0),          # c_phys = 32, strided 1x1
    (5, 128, 1, 1, 1000, 1, 1, 0),       # classifier-like: one pixel per image, K = 1000
])
def test_tensor_core_backward_matches_direct(orc, qbit, shape):
    from gpu_util import conv_fwd_gpu
    N, C, H, W, K, k, st, pad = shape
    rng = np.random.default_rng(hash(shape) % (2 ** 31))
    x = (rng.standard_normal((N, C, H, W)) * 2).astype(np.float32)
    w = (rng.standard_normal((K, C, k, k)) * 0.3).astype(np.float32)
    ka, kw = float(np.abs(x).max() / 15.5), float(np.abs(w).max() / 15.5)
    out = conv_fwd_gpu(x, w, None, ka, kw, qbit, st, pad, 1, 1)
    Ho, Wo = out["y"].shape[2], out["y"].shape[3]
    gy = torch.from_numpy((rng.standard_normal((N, Ho, Wo, K)) * 3e-6 * np.exp(rng.standard_normal((N, Ho, Wo, K)))).astype(np.float32)).cuda()
    dx0, dw0 = _bwd_both(out, gy, N, H, W, C, K, k, 1, qbit, ka, kw, False)
    dx1, dw1 = _bwd_both(out, gy, N, H, W, C, K, k, 1, qbit, ka, kw, True)
    # magnitude sums: |gy| and codes with the sign bit cleared
    out["dev"]["xc"].bitwise_and_(0x7f)
    out["dev"]["wc"].bitwise_and_(0x7f)
    dxa, dwa = _bwd_both(out, gy.abs(), N, H, W, C, K, k, 1, qbit, ka, kw, False)
    assert np.isfinite(dx1).all() and np.isfinite(dw1).all()
    for got, want, l1, what in ((dx1, dx0, dxa, "dx"), (dw1, dw0, dwa, "dw")):
        err = np.abs(got - want)
        assert (err <= 1.2e-3 * l1 + 1e-30).all(), (what, shape, float((err / (l1 + 1e-30)).max()))


@pytest.mark.parametrize("shape", [(2, 64, 56, 56, 64, 3, 1, 1), (2, 256, 14, 14, 256, 3, 1, 1), (2, 128, 56, 56, 128, 3, 2, 1),
                                   (2, 64, 56, 56, 256, 1, 1, 0), (2, 512, 7, 7, 512, 3, 1, 1), (2, 64, 27, 27, 192, 5, 1, 2)])
def test_tensor_core_backward_config_shapes_vs_oracle(orc, shape):
    """dgrad / wgrad implicit GEMMs (tcgen05, float16 operands, STE = identity) at layer shapes of the BASELINE configs
    against the ORACLE's backward (double accumulation over the reference's fake-quant operands):
    |d - d_ref| <= 1.2e-3 * the same sums over magnitudes."""
    from gpu_util import conv_fwd_gpu
    N, C, H, W, K, k, st, pad = shape
    rng = np.random.default_rng(hash(shape) % (2 ** 31))
    x = (rng.standard_normal((N, C, H, W)) * 2).astype(np.float32)
    w = (rng.standard_normal((K, C, k, k)) * 0.3).astype(np.float32)
    ka, kw = float(np.abs(x).max() / 15.5), float(np.abs(w).max() / 15.5)
    out = conv_fwd_gpu(x, w, None, ka, kw, 8, st, pad, 1, 1)
    Ho, Wo = out["y"].shape[2], out["y"].shape[3]
    gy_np = (rng.standard_normal((N, K, Ho, Wo)) * 1e-3).astype(np.float32)
    gy = torch.from_numpy(np.ascontiguousarray(gy_np.transpose(0, 2, 3, 1))).cuda()
    dx, dw = _bwd_both(out, gy, N, H, W, C, K, k, 1, 8, ka, kw, True)
    _, xq = orc.quantize(x, 1, ka, want_codes=False)
    _, wq = orc.quantize(w, 2, kw, want_codes=False)
    dx_r, dw_r, _ = orc.conv2d_q_bwd(xq, wq, gy_np, st, pad, 1, 1, ka, kw)
    dx_a, dw_a, _ = orc.conv2d_q_bwd(np.abs(xq), np.abs(wq), np.abs(gy_np), st, pad, 1, 1, ka, kw)
    dx_r, dx_a = dx_r.transpose(0, 2, 3, 1), dx_a.transpose(0, 2, 3, 1)          # the C ABI's dx is NHWC
    for got, want, l1, what in ((dx, dx_r, dx_a, "dx"), (dw, dw_r, dw_a, "dw")):
        err = np.abs(got - want)
        assert (err <= 1.2e-3 * l1 + 1e-30).all(), (what, shape, float((err / (l1 + 1e-30)).max()))


def test_tensor_core_backward_zero_and_huge_gradients(orc):
    """The power-of-two scaling of gy: an all-zero gradient gives exact zeros, a 1e30-sized one stays finite."""
    from gpu_util import conv_fwd_gpu
    rng = np.random.default_rng(5)
    N, C, H, W, K, k = 2, 64, 8, 8, 64, 3
    x = (rng.standard_normal((N, C, H, W)) * 2).astype(np.float32)
    w = (rng.standard_normal((K, C, k, k)) * 0.3).astype(np.float32)
    ka, kw = float(np.abs(x).max() / 15.5), float(np.abs(w).max() / 15.5)
    out = conv_fwd_gpu(x, w, None, ka, kw, 8, 1, 1, 1, 1)
    gy = torch.zeros((N, H, W, K), dtype=torch.float32, device="cuda")
    dx, dw = _bwd_both(out, gy, N, H, W, C, K, k, 1, 8, ka, kw, True)
    assert (dx == 0).all() and (dw == 0).all()
    gy = torch.from_numpy((rng.standard_normal((N, H, W, K)) * 1e30).astype(np.float32)).cuda()
    dx0, dw0 = _bwd_both(out, gy, N, H, W, C, K, k, 1, 8, ka, kw, False)
    dx1, dw1 = _bwd_both(out, gy, N, H, W, C, K, k, 1, 8, ka, kw, True)
    assert np.isfinite(dx1).all() and np.isfinite(dw1).all()
    np.testing.assert_allclose(dx1, dx0, rtol=0, atol=2e-3 * np.abs(dx0).max())
    np.testing.assert_allclose(dw1, dw0, rtol=0, atol=2e-3 * np.abs(dw0).max())
