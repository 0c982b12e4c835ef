"""Drop-in for the reference's utils/conv2d_func.py: the class factories conv2d_Q (:8-26),
conv2d_Q_bias (:28-48) and linear_Q (:50-66) with the reference's constructor order
(in, out, kernel, Kw, Ka, stride, ...) and attributes (.q_bit .Kw .Ka .quantize_weight
.quantize_act, and after a forward .input_q .weight_q .output).

forward(input) computes   F.conv2d(qa(input/Ka), qw(weight/Kw), bias/Ka/Kw) * Ka * Kw
as three launches of hand-written sm_100a kernels through the C ABI (include/slfp_b200.h):
  1. fused pre-scale + quantize of the activations into 8-bit NHWC codes   (csrc/quantize.cu)
  2. fused pre-scale + quantize + KRSC re-layout of the weights            (csrc/quantize.cu)
  3. implicit-GEMM convolution on tcgen05 tensor cores with the post-scale and bias in the
     epilogue (csrc/conv_igemm_sm100.cu), or the depthwise / grouped stencil (csrc/conv_direct.cu)
Backward is the straight-through estimator of the reference (utils/sfp_quant.py:50-53): dgrad /
wgrad on the saved 8-bit codes with the constant scale factors fused (csrc/conv_direct.cu).

The returned tensors are channels-last in memory (NCHW logical shape), so BatchNorm / ReLU /
pooling between quantized layers keep running on the NHWC layout without transposes.
Star-import leaks torch / nn / F / np exactly like the reference module does.
"""
import ctypes as _ctypes

import torch
import torch.nn as nn
import torch.nn.functional as F
import numpy as np
from .sfp_quant import *

from .. import _native as _nv


# High-fidelity switch of the drop-in modules: True runs every dense layer in the split-operand tensor-core mode
# (include/slfp_b200.h SLFP_CONV_SPLIT_OPERANDS: float16 hi + lo operand pairs, three passes over K; conv outputs agree
# with the reference's float32 arithmetic to accumulation order) at about 3x the tensor work.  Forward only - the
# backward keeps single float16 operands.  Also settable with the environment variable SLFP_HIGH_FIDELITY=1.
import os as _os
HIGH_FIDELITY = bool(_os.environ.get("SLFP_HIGH_FIDELITY"))


def _ceil_to(v, m):
    return (v + m - 1) // m * m


def _k32(K):
    """float32 value the reference's arithmetic uses for a scale (K is a 0-dim float64 tensor)."""
    return float(np.float32(float(K)))


class _ConvCfg:
    __slots__ = ("q_bit", "ka", "kw", "stride", "padding", "dilation", "groups", "post_a", "post_b", "keep")

    def __init__(self, q_bit, ka, kw, stride, padding, dilation, groups, post_a, post_b, keep):
        self.q_bit, self.ka, self.kw = q_bit, ka, kw
        self.stride, self.padding, self.dilation, self.groups = stride, padding, dilation, groups
        self.post_a, self.post_b, self.keep = post_a, post_b, keep


class _QConvNHWC(torch.autograd.Function):
    """x: [N,H,W,C] float32 contiguous; weight: [K, C/groups, R, S] (any strides); bias_q: [K] or None.
    Returns y: [N,Ho,Wo,K] float32 contiguous."""

    @staticmethod
    def forward(ctx, x, weight, bias_q, cfg, x_codes=None, prepared=None):
        lib = _nv.lib()
        _nv.require_cuda(x, "Conv2d_Q / Linear_Q input")
        _nv.require_cuda(weight, "Conv2d_Q / Linear_Q weight")
        x = x.detach().contiguous()
        N, H, W, C = x.shape
        K, Cg, R, S = weight.shape
        groups = cfg.groups
        if Cg * groups != C:
            raise RuntimeError(f"Conv2d_Q: weight {tuple(weight.shape)} does not match input channels {C} (groups={groups})")
        dense = groups == 1
        hifi = HIGH_FIDELITY and dense
        Cp = ((4 if (C <= 4 and not hifi) else _ceil_to(C, 16))) if dense else _ceil_to(C, 4)
        afmt, wfmt = _nv.fmt_for(cfg.q_bit, "act"), _nv.fmt_for(cfg.q_bit, "weight")
        d = _nv.SlfpConvDesc(N, H, W, C, Cp, K, R, S, cfg.stride[0], cfg.stride[1], cfg.padding[0], cfg.padding[1],
                             cfg.dilation[0], cfg.dilation[1], groups, afmt, 0, 0, _nv.CONV_SPLIT_OPERANDS if hifi else 0)
        Ho = (H + 2 * cfg.padding[0] - cfg.dilation[0] * (R - 1) - 1) // cfg.stride[0] + 1
        Wo = (W + 2 * cfg.padding[1] - cfg.dilation[1] * (S - 1) - 1) // cfg.stride[1] + 1
        if Ho <= 0 or Wo <= 0:
            raise RuntimeError("Conv2d_Q: output size is too small")
        st = _nv.stream()
        # 1. activations: x / Ka -> 8-bit codes (NHWC, channel-padded)
        # (codes that the producer of x already wrote - utils/bn_act.py, same quantizer, same Ka - are used as they are)
        if x_codes is None or tuple(x_codes.shape) != (N, H, W, Cp) or x_codes.dtype != torch.uint8:
            x_codes = torch.empty((N, H, W, Cp), dtype=torch.uint8, device=x.device)
            _nv.check(lib.slfp_quantize_nhwc_f32(x.data_ptr(), N * H * W, C, Cp, cfg.ka, afmt, x_codes.data_ptr(), st))
        # 2. weights: w / Kw -> codes (+ the float16 tensor-core operand), KRSC
        pitch = lib.slfp_conv_wpitch(ctypes_byref(d))
        # (weights re-quantized for the whole net in one launch at the top of its forward: prepare_weights_batched)
        ready = prepared is not None and dense and not hifi and prepared[2] == Cp and prepared[3] == pitch and prepared[4] == cfg.kw
        if ready:
            w_f16, w_codes = prepared[0], prepared[1]
        else:
            w_codes = torch.empty((K * pitch,), dtype=torch.uint8, device=x.device)
            w_f16 = torch.empty((K * pitch * (2 if hifi else 1),), dtype=torch.float16, device=x.device) if dense else None
        so, sc, sr, ss = weight.stride()
        if ready:
            pass
        elif hifi:
            # split operands: rows of [hi | lo] halves - the jobs entry point carries the lo placement
            job = (_nv.SlfpWeightJob * 1)()
            job[0].desc, job[0].w, job[0].kw = _ctypes.pointer(d), weight.data_ptr(), cfg.kw
            job[0].w_stride[:] = (so, sc, sr, ss)
            job[0].w_f16, job[0].w_codes = w_f16.data_ptr(), None
            job[0].out_pitch, job[0].out_offset, job[0].lo_offset = 2 * pitch, 0, pitch
            _nv.check(lib.slfp_prepare_weights_jobs(1, job, wfmt, st))
            _nv.check(lib.slfp_prepare_weights(ctypes_byref(d), weight.data_ptr(), so, sc, sr, ss, cfg.kw, wfmt,
                                               None, w_codes.data_ptr(), None, st))          # codes (backward, taps): own pitch
        else:
            _nv.check(lib.slfp_prepare_weights(ctypes_byref(d), weight.data_ptr(), so, sc, sr, ss, cfg.kw, wfmt,
                                               _nv.ptr(w_f16), w_codes.data_ptr(), None, st))
        # 3. convolution with the reference's post-scale (and bias) in the epilogue
        y = torch.empty((N, Ho, Wo, K), dtype=torch.float32, device=x.device)
        epi = _nv.SlfpEpilogue()
        epi.bias_q = _nv.ptr(bias_q.detach().contiguous()) if bias_q is not None else None
        epi.post_a, epi.post_b = cfg.post_a, cfg.post_b
        epi.y_f32 = y.data_ptr()
        _nv.check(lib.slfp_conv2d_fwd(ctypes_byref(d), x_codes.data_ptr(), (w_f16 if dense else w_codes).data_ptr(),
                                      ctypes_byref(epi), st))
        ctx.save_for_backward(x_codes, w_codes)
        ctx.cfg, ctx.desc, ctx.wfmt = cfg, d, wfmt
        ctx.wshape, ctx.has_bias = tuple(weight.shape), bias_q is not None
        if cfg.keep is not None:                      # calibration taps read the codes lazily
            cfg.keep["x_codes"], cfg.keep["w_codes"] = x_codes, w_codes
            cfg.keep["meta"] = (N, H, W, C, Cp, K, Cg, R, S, pitch, afmt, wfmt, dense)
        return y

    @staticmethod
    def backward(ctx, gy):
        lib = _nv.lib()
        x_codes, w_codes = ctx.saved_tensors
        cfg, d = ctx.cfg, ctx.desc
        gy = gy.contiguous()
        K, Cg, R, S = ctx.wshape
        need_x, need_w = ctx.needs_input_grad[:2]
        dx = torch.empty((d.n, d.h, d.w, d.c), dtype=torch.float32, device=gy.device) if need_x else None
        dw = torch.empty(ctx.wshape, dtype=torch.float32, device=gy.device) if need_w else None
        db = torch.empty((K,), dtype=torch.float32, device=gy.device) if ctx.has_bias else None
        so, sc, sr, ss = (dw.stride() if dw is not None else (0, 0, 0, 0))
        # dgrad / wgrad as tcgen05 implicit GEMMs (csrc/conv_bwd_sm100.cu); the library never allocates, so the
        # float16 operand images and the split-K accumulator live in a scratch tensor from torch's caching allocator.
        # Shapes the tensor-core path does not cover (grouped, > 32 taps) report 0 bytes and run the direct kernels.
        # max |gy| tracked by the kernel that wrote gy (fused BatchNorm backward): the tensor-core paths skip their abs-max pass
        hint, _nv.pending_grad_absmax = _nv.pending_grad_absmax, None
        amax = hint[2] if (hint is not None and hint[0] == gy.data_ptr() and hint[1] == gy.numel()
                           and not _os.environ.get("SLFP_NO_ABSMAX_HINT")) else None
        if need_w and not need_x and db is None and _folded_stem_wgrad(lib, cfg, d, ctx.wfmt, gy, amax, x_codes, w_codes, dw):
            return None, dw, None, None, None, None
        ws_bytes = lib.slfp_conv2d_bwd_workspace_size(ctypes_byref(d), int(need_x), int(need_w))
        if ws_bytes:
            ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=gy.device)
            _nv.check(lib.slfp_conv2d_bwd_ws_absmax(ctypes_byref(d), gy.data_ptr(), _nv.ptr(amax), x_codes.data_ptr(), w_codes.data_ptr(),
                                                    ctx.wfmt, cfg.ka, cfg.kw, _nv.ptr(dx), _nv.ptr(dw), so, sc, sr, ss,
                                                    _nv.ptr(db), ws.data_ptr(), ws_bytes, _nv.stream()))
        else:
            _nv.check(lib.slfp_conv2d_bwd(ctypes_byref(d), gy.data_ptr(), x_codes.data_ptr(), w_codes.data_ptr(), ctx.wfmt,
                                          cfg.ka, cfg.kw, _nv.ptr(dx), _nv.ptr(dw), so, sc, sr, ss,
                                          _nv.ptr(db), _nv.stream()))
        if db is not None:                             # y = (acc + bias_q) * post_a * post_b
            db = (db * cfg.post_b) * cfg.post_a
        return dx, dw, db, None, None, None


def prepare_weights_batched(model):
    """Re-quantize the weights of every dense Conv2d_Q / Linear_Q of `model` in ONE launch (slfp_prepare_weights_jobs)
    instead of one launch per layer inside each forward (54 launches = 0.69 ms of the ResNet-50 QAT step).  Called by a
    net at the top of its forward; each layer's next forward consumes its entry exactly once (weights may change
    between forwards through raw pointers - the revised-SGD kernel - so nothing is cached across forwards).  Layers the
    job form does not cover (grouped, high-fidelity mode, q_bit 32) prepare their weights themselves as before."""
    if HIGH_FIDELITY or _os.environ.get("SLFP_NO_BATCHED_WPREP"):
        return
    lib = _nv.lib()
    cache = model.__dict__.get("_slfp_wprep")
    mods = cache[0] if cache else [m for m in model.modules() if isinstance(m, _QuantTapsMixin) and getattr(m, "q_bit", 32) in (7, 8)
                                   and getattr(m, "groups", 1) == 1]
    if not mods or not mods[0].weight.is_cuda:
        return
    sig = tuple((m.weight.data_ptr(), float(m.Kw)) for m in mods)
    if cache is None or cache[1] != sig:
        wfmt = _nv.fmt_for(mods[0].q_bit, "weight")
        if any(_nv.fmt_for(m.q_bit, "weight") != wfmt for m in mods):
            return
        jobs = (_nv.SlfpWeightJob * len(mods))()
        descs, metas, off = [], [], 0
        for m in mods:
            w = m.weight if m.weight.dim() == 4 else m.weight.view(m.weight.shape[0], m.weight.shape[1], 1, 1)
            K, C, R, S = w.shape
            Cp = 4 if C <= 4 else _ceil_to(C, 16)
            d = _nv.SlfpConvDesc(1, R, S, C, Cp, K, R, S, 1, 1, 0, 0, 1, 1, 1, _nv.fmt_for(m.q_bit, "act"), 0, 0, 0)
            pitch = lib.slfp_conv_wpitch(ctypes_byref(d))
            descs.append(d)
            metas.append((w, K, Cp, pitch, off))
            off += _ceil_to(K * pitch, 128)
        dev = mods[0].weight.device
        f16 = torch.empty((off,), dtype=torch.float16, device=dev)
        codes = torch.empty((off,), dtype=torch.uint8, device=dev)
        for j, d, m, (w, K, Cp, pitch, o) in zip(jobs, descs, mods, metas):
            j.desc, j.w, j.kw = _ctypes.pointer(d), w.data_ptr(), _k32(m.Kw)
            j.w_stride[:] = w.stride()
            j.w_f16, j.w_codes = f16[o:o + K * pitch].data_ptr(), codes[o:o + K * pitch].data_ptr()
        cache = (mods, sig, jobs, descs, metas, f16, codes, wfmt)
        model.__dict__["_slfp_wprep"] = cache
    mods, sig, jobs, descs, metas, f16, codes, wfmt = cache
    _nv.check(lib.slfp_prepare_weights_jobs(len(mods), jobs, wfmt, _nv.stream()))
    for m, (w, K, Cp, pitch, o) in zip(mods, metas):
        m.__dict__["_prepared_w"] = (f16[o:o + K * pitch], codes[o:o + K * pitch], Cp, pitch, _k32(m.Kw))


def _folded_stem_wgrad(lib, cfg, d, wfmt, gy, gy_absmax, x_codes, w_codes, dw):
    """wgrad of a network stem (<= 4 input channels, stride 2, more than 32 taps: ResNet's 7x7/2) on the tensor cores.

    The tensor-core wgrad covers <= 32 taps on c_phys % 64 == 0 channels, so the stem used to fall to a CUDA-core kernel
    (1.8 ms of the batch-128 QAT step).  A stride-2 RxR convolution on c channels is a stride-1 R2xR2 convolution
    (R2 = 4 for 7x7 / padding 3) on the 2x2 space-to-depth image with 4c channels - the fold engine.Plan.s2d_stem uses
    forward: input row 2*ho - P + r = 2*(ho - lo) + (r + off) with lo = ceil(P / 2), off = 2*lo - P, folded tap
    a = (r + off) // 2, parity dy = (r + off) % 2.  The folded code image is built from the saved codes (byte copies),
    zero-padded in space (so the folded layer needs no padding) and to 64 channels; the folded gradient is scattered
    back to [K, c, R, R].  Same function of (gy, codes) as the direct kernel.  Returns False when the shape does not
    qualify (the caller then takes the generic paths)."""
    if _os.environ.get("SLFP_NO_FOLDED_STEM_WGRAD"):
        return False
    N, H, W, C, Cp, K, R, S = d.n, d.h, d.w, d.c, d.c_phys, d.k, d.r, d.s
    if not (d.groups == 1 and Cp == 4 and R * S > 32 and R == S and cfg.stride == (2, 2) and cfg.dilation == (1, 1)
            and cfg.padding[0] == cfg.padding[1] and H % 2 == 0 and W % 2 == 0 and dw.is_contiguous()):
        return False
    P = cfg.padding[0]
    lo = (P + 1) // 2
    off = 2 * lo - P
    R2 = (R - 1 + off) // 2 + 1
    Ho, Wo = gy.shape[1], gy.shape[2]
    Hp, Wp = Ho + R2 - 1, Wo + R2 - 1
    if R2 * R2 > 32 or Hp - H // 2 - lo < 0 or Wp - W // 2 - lo < 0:
        return False
    d2 = _nv.SlfpConvDesc(N, Hp, Wp, 16, 64, K, R2, R2, 1, 1, 0, 0, 1, 1, 1, d.fmt, 0, 0, 0)
    ws_bytes = lib.slfp_conv2d_bwd_workspace_size(ctypes_byref(d2), 0, 1)
    if not ws_bytes:
        return False
    # [n, h/2, dy, w/2, dx, 4] -> [n, h/2, w/2, (dy, dx, c4)] inside the zero frame (code 0 decodes to 0)
    x2 = torch.zeros((N, Hp, Wp, 64), dtype=torch.uint8, device=gy.device)
    x2[:, lo:lo + H // 2, lo:lo + W // 2, :16] = x_codes.view(N, H // 2, 2, W // 2, 2, 4).permute(0, 1, 3, 2, 4, 5).reshape(N, H // 2, W // 2, 16)
    dw2 = torch.empty((K, 16, R2, R2), dtype=torch.float32, device=gy.device)
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=gy.device)
    _nv.check(lib.slfp_conv2d_bwd_ws_absmax(ctypes_byref(d2), gy.data_ptr(), _nv.ptr(gy_absmax), x2.data_ptr(), w_codes.data_ptr(), wfmt,
                                            cfg.ka, cfg.kw, None, dw2.data_ptr(), *dw2.stride(), None, ws.data_ptr(), ws_bytes,
                                            _nv.stream()))
    # dw2[k, (dy, dx, c4), a, b] = gradient of wp[k, c4, 2a + dy, 2b + dx];  w = wp[:, :c, off:off + R, off:off + R]
    wp = dw2.view(K, 2, 2, 4, R2, R2).permute(0, 3, 4, 1, 5, 2).reshape(K, 4, 2 * R2, 2 * R2)
    dw.copy_(wp[:, :C, off:off + R, off:off + S])
    return True


def ctypes_byref(obj):
    return _ctypes.byref(obj)


def _dequant_tap(keep, which):
    """Materialise the reference's `input_q` / `weight_q` fake-quant tensors from the kept codes."""
    if not keep or "meta" not in keep:
        raise AttributeError(f"'{which}' is only available after a forward pass")
    lib = _nv.lib()
    N, H, W, C, Cp, K, Cg, R, S, pitch, afmt, wfmt, dense = keep["meta"]
    lead = keep.get("lead")                     # Linear_Q: leading shape of the 2-D+ input
    if which == "input_q":
        codes = keep["x_codes"]
        out = torch.empty(codes.shape, dtype=torch.float32, device=codes.device)
        _nv.check(lib.slfp_dequantize(codes.data_ptr(), codes.numel(), afmt, out.data_ptr(), _nv.stream()))
        if lead is not None:
            return out[..., :C].reshape(*lead, C)
        return out[..., :C].permute(0, 3, 1, 2)
    codes = keep["w_codes"]
    out = torch.empty(codes.shape, dtype=torch.float32, device=codes.device)
    _nv.check(lib.slfp_dequantize(codes.data_ptr(), codes.numel(), wfmt, out.data_ptr(), _nv.stream()))
    cw = Cp if dense else Cg
    wq = out.view(K, pitch)[:, :R * S * cw].reshape(K, R, S, cw)[..., :Cg].permute(0, 3, 1, 2)
    return wq.reshape(K, Cg) if lead is not None else wq


class _QuantTapsMixin:
    """`.input_q` / `.weight_q`: the tensors the reference keeps alive after every forward (read by
    the calibration taps of the nets, e.g. nets_cifar/vgg16.py:136-182).  Here they are decoded on
    demand from the 8-bit codes the last forward produced."""

    @property
    def input_q(self):
        eager = self.__dict__.get("_eager_input_q")
        return eager if eager is not None else _dequant_tap(self.__dict__.get("_keep"), "input_q")

    @input_q.setter
    def input_q(self, v):
        self.__dict__["_eager_input_q"] = v

    @property
    def weight_q(self):
        eager = self.__dict__.get("_eager_weight_q")
        return eager if eager is not None else _dequant_tap(self.__dict__.get("_keep"), "weight_q")

    @weight_q.setter
    def weight_q(self, v):
        self.__dict__["_eager_weight_q"] = v


def _conv_forward(self, input, bias_q):
    if self.q_bit == 32:
        # identity quantizers: the reference's own float32 expression (conv2d_func.py:21-24)
        self.input_q = self.quantize_act(input / self.Ka)
        self.weight_q = self.quantize_weight(self.weight / self.Kw)
        return F.conv2d(self.input_q, self.weight_q, bias_q, self.stride, self.padding, self.dilation,
                        self.groups) * self.Ka * self.Kw
    if self.q_bit not in (7, 8):
        raise UnboundLocalError("cannot access local variable 'act_q' where it is not associated with a value")
    if isinstance(self.padding, str) or self.padding_mode != "zeros":
        raise NotImplementedError("Conv2d_Q: only numeric zero padding is supported")
    if input.dim() != 4:
        raise RuntimeError("Conv2d_Q: expected a 4-D NCHW input")
    self.__dict__["_eager_input_q"] = self.__dict__["_eager_weight_q"] = None
    keep = self.__dict__.setdefault("_keep", {})
    ka, kw = _k32(self.Ka), _k32(self.Kw)
    cfg = _ConvCfg(self.q_bit, ka, kw, tuple(self.stride), tuple(self.padding), tuple(self.dilation), self.groups,
                   ka, kw, keep)
    ready = getattr(input, "_slfp_codes", None)                 # written by the fused BatchNorm + ReLU that produced `input`
    x_codes = ready.get((ka, _nv.fmt_for(self.q_bit, "act"))) if (ready and self.groups == 1 and not HIGH_FIDELITY) else None
    y = _QConvNHWC.apply(input.permute(0, 2, 3, 1), self.weight, bias_q, cfg, x_codes, self.__dict__.pop("_prepared_w", None))
    return y.permute(0, 3, 1, 2)


def conv2d_Q(q_bit, Kw, Ka):
    """utils/conv2d_func.py:8-26."""
    class Conv2d_Q(_QuantTapsMixin, nn.Conv2d):
        _slfp_bias_scaled = False       # bias goes into the conv un-scaled (conv2d_func.py:23)

        def __init__(self, in_channels, out_channels, kernel_size, Kw=Kw, Ka=Ka,
                     stride=1, padding=0, dilation=1, groups=1, bias=False):
            super(Conv2d_Q, self).__init__(in_channels, out_channels, kernel_size, stride,
                                           padding, dilation, groups, bias)
            self.q_bit = q_bit
            self.quantize_weight = weight_quantize_func(q_bit=q_bit)
            self.quantize_act = act_quantize_func(q_bit=q_bit)
            self.Kw = torch.tensor(Kw)
            self.Ka = torch.tensor(Ka)

        def forward(self, input, order=None):
            # the reference passes self.bias straight into F.conv2d (no /Ka/Kw) in this variant (:23)
            self.output = _conv_forward(self, input, self.bias)
            return self.output
    return Conv2d_Q


def conv2d_Q_bias(q_bit, Kw, Ka):
    """utils/conv2d_func.py:28-48."""
    class Conv2d_Q(_QuantTapsMixin, nn.Conv2d):
        _slfp_bias_scaled = True        # bias / Ka / Kw (conv2d_func.py:44)

        def __init__(self, in_channels, out_channels, kernel_size, Kw=Kw, Ka=Ka, stride=1, padding=0, dilation=1,
                     groups=1, bias=True):
            super(Conv2d_Q, self).__init__(in_channels, out_channels, kernel_size, stride,
                                           padding, dilation, groups, bias)
            self.q_bit = q_bit
            self.quantize_weight = weight_quantize_func(q_bit=q_bit)
            self.quantize_act = act_quantize_func(q_bit=q_bit)
            self.Kw = torch.tensor(Kw)
            self.Ka = torch.tensor(Ka)

        def forward(self, input, order=None):
            self.bias_q = self.bias / self.Ka / self.Kw                    # (:44)
            self.output = _conv_forward(self, input, self.bias_q)
            return self.output
    return Conv2d_Q


def linear_Q(q_bit, Kw, Ka):
    """utils/conv2d_func.py:50-66: F.linear(qa(x/Ka), qw(W/Kw), bias/Kw/Ka) * Kw * Ka."""
    class Linear_Q(_QuantTapsMixin, nn.Linear):
        def __init__(self, in_features, out_features, Kw=Kw, Ka=Ka, bias=True):
            super(Linear_Q, self).__init__(in_features, out_features, bias)
            self.q_bit = q_bit
            self.quantize_weight = weight_quantize_func(q_bit=q_bit)
            self.quantize_act = act_quantize_func(q_bit=q_bit)
            self.Kw = torch.tensor(Kw)
            self.Ka = torch.tensor(Ka)

        def forward(self, input):
            self.bias_q = self.bias / self.Kw / self.Ka                    # (:63)
            if self.q_bit == 32:
                self.input_q = self.quantize_act(input / self.Ka)
                self.weight_q = self.quantize_weight(self.weight / self.Kw)
                return F.linear(self.input_q, self.weight_q, self.bias_q) * self.Kw * self.Ka
            if self.q_bit not in (7, 8):
                raise UnboundLocalError("cannot access local variable 'act_q' where it is not associated with a value")
            self.__dict__["_eager_input_q"] = self.__dict__["_eager_weight_q"] = None
            keep = self.__dict__.setdefault("_keep", {})
            ka, kw = _k32(self.Ka), _k32(self.Kw)
            cfg = _ConvCfg(self.q_bit, ka, kw, (1, 1), (0, 0), (1, 1), 1, kw, ka, keep)
            lead = tuple(input.shape[:-1])
            keep["lead"] = lead
            x = input.reshape(-1, 1, 1, self.in_features)
            y = _QConvNHWC.apply(x, self.weight.view(self.out_features, self.in_features, 1, 1), self.bias_q, cfg, None,
                                 self.__dict__.pop("_prepared_w", None))
            return y.reshape(*lead, self.out_features)
    return Linear_Q
