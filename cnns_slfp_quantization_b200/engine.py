"""Fused eval pipeline: whole-network inference where quantized layers exchange
8-bit codes and everything between two convolutions lives in the producing conv's epilogue.

The module-level drop-in (utils/conv2d_func.py) keeps float32 tensors between layers because
BatchNorm / ReLU / pooling / the residual add are stock PyTorch modules of the caller nets.  In eval
mode those are pure per-channel affine maps and elementwise ops, so a compiled plan can fold them:

    conv epilogue:  acc (+bias_q) *Ka *Kw -> *bn_scale + bn_shift -> (+ residual) -> ReLU
                    -> next layer's quantizer (x / Ka_next -> SLFP/SFP code)   [quantize-on-store]
                    -> float16 copy where a later residual add needs the un-quantized value

exactly the reference's arithmetic order (nets_imgnet/resnet50.py:71-90: bn -> relu, out += identity,
relu; utils/conv2d_func.py:21: the next layer quantizes its input / Ka), except that eval BatchNorm
is applied as one fused multiply-add (scale = gamma / sqrt(var + eps), shift = beta - mean * scale)
and the residual stream is carried in float16.  Max-pooling runs directly on codes (the quantizer is
monotone, so pooling commutes with it bit for bit).

A Plan is a flat list of pre-bound C-ABI calls on pre-allocated buffers; `capture()` records it into
a CUDA graph.  Weights are re-quantized on every forward like the reference does
(utils/conv2d_func.py:22) unless `static_weights=True`.
"""
import ctypes
import os

import numpy as np
import torch
import torch.nn as nn

from . import _native as nv


def _k32(K):
    return float(np.float32(float(K)))


class _T:
    """A device tensor of the plan: NHWC, kind in {'codes', 'q16', 'f16', 'f32'}.  'q16' (SLFP_FMT_F16Q) holds the float16
    image of the codes of format `qfmt` - the tensor-core operand itself - for a decode-bound dense consumer."""
    __slots__ = ("buf", "n", "h", "w", "c", "cp", "kind", "kdiv", "fmt", "c_logical", "qfmt", "pad", "im2col")

    def __init__(self, buf, n, h, w, c, cp, kind, kdiv=None, fmt=None):
        self.buf, self.n, self.h, self.w, self.c, self.cp, self.kind, self.kdiv = buf, n, h, w, c, cp, kind, kdiv
        self.fmt = fmt                  # code format of a 'codes' tensor (signed quantizer codes or post-ReLU codes)
        self.c_logical = c              # float16 outputs of a pad_k layer: c is the physical channel count
        self.qfmt = None                # 'q16': the code format whose values the halves are
        self.im2col = False             # buf is the 3x3 / pad 1 im2col matrix [n, h, w, 64] of a 3-channel input (q16)
        self.pad = None                 # (top, left, hp, wp): buf is physically zero-padded [n, hp, wp, cp]; h, w stay logical


def _ceil(v, m):
    return (v + m - 1) // m * m


def fold_bn(bn):
    """Eval BatchNorm as y = x * scale + shift (float64 arithmetic, rounded once to float32)."""
    g = bn.weight.detach().double() if bn.weight is not None else torch.ones_like(bn.running_mean, dtype=torch.float64)
    b = bn.bias.detach().double() if bn.bias is not None else torch.zeros_like(bn.running_mean, dtype=torch.float64)
    scale = g / torch.sqrt(bn.running_var.detach().double() + bn.eps)
    shift = b - bn.running_mean.detach().double() * scale
    return scale.float().contiguous(), shift.float().contiguous()


class Plan:
    def __init__(self, batch, device, q_bit, static_weights=False, high_fidelity=False, e4m3=None):
        self.lib = nv.lib()
        # high_fidelity: split-operand tensor-core mode (SLFP_CONV_SPLIT_OPERANDS: float16 hi + lo pairs, three passes over K)
        # for every dense layer with c_phys % 16 == 0, the exact signed code formats and float32 instead of float16 value
        # tensors - the plan then differs from the reference's float32 arithmetic by accumulation order only
        self.high_fidelity = high_fidelity
        # SFP-7 plans exchange e4m3 bytes (SLFP_FMT_E4M3): every SFP<3,3> value is an e4m3 number, so the codes ARE the
        # tensor-core operand (kind::f8f6f4, no decode) and the encoder is an exact round-half-even cvt
        self.e4m3 = (q_bit == 7 and not high_fidelity and not os.environ.get("SLFP_NO_E4M3")) if e4m3 is None else bool(e4m3)
        # SLFP-8 plans: a dense producer whose only consumer is a decode-bound dense layer (3x3, or a short-K block tail)
        # stores the float16 image of its codes (SLFP_FMT_F16Q, SlfpEpilogue.store_f16) and the consumer's TMA drops that
        # straight into the MMA's operand layout - no decode table, no decode warps, bit-identical results
        self.f16q = q_bit == 8 and not high_fidelity and not os.environ.get("SLFP_NO_F16Q")
        self.batch, self.dev, self.q_bit = batch, device, q_bit
        self.afmt, self.wfmt = nv.fmt_for(q_bit, "act"), nv.fmt_for(q_bit, "weight")
        self.static_weights = static_weights
        self.weight_table, self.ops, self.keep = [], [], []
        self.pre_weight_ops = []          # torch-side weight re-layouts run before the batched re-quantization
        self._wbatch = None
        self.input = None
        self.output = None
        self.graph = None
        self.flops = 0.0
        self.taps = []                # (module, input code tensor _T, op index) of every quantized layer (parity tests)
        self.conv_flops = []          # per conv launch, in launch order (bench.py's roofline pass)
        self.bytes_hbm = 0
        # what the plan froze at compile time: weight storage and scales are checked by verify(), the folded
        # bias / BatchNorm vectors can be recomputed in place by refresh()
        self._guards = []             # (module, weight view, its data_ptr, Ka, Kw)
        self._affines = []            # (module, bn, K, linear, mul32, add32) of single-branch layers

    # ---- buffers ----------------------------------------------------------------------------------------
    def _alloc(self, n, h, w, c, kind, kdiv=None, cp=None, fmt=None):
        if kind == "codes":
            cp = cp or _ceil(c, 16)
            buf = torch.zeros((n, h, w, cp), dtype=torch.uint8, device=self.dev)
        elif kind == "q16":
            cp = cp or c
            assert cp % 64 == 0
            buf = torch.zeros((n, h, w, cp), dtype=torch.float16, device=self.dev)
            self.bytes_hbm += buf.numel() * 2
            self.keep.append(buf)
            t = _T(buf, n, h, w, c, cp, kind, kdiv, nv.FMT_F16Q)
            t.qfmt = fmt
            return t
        else:
            cp = c
            buf = torch.empty((n, h, w, c), dtype=torch.float16 if kind == "f16" else torch.float32, device=self.dev)
        self.bytes_hbm += buf.numel() * buf.element_size()
        self.keep.append(buf)               # the pre-bound calls hold raw pointers: the plan owns the storage
        return _T(buf, n, h, w, c, cp, kind, kdiv, (fmt if fmt is not None else self.afmt) if kind == "codes" else None)

    def _weight_job(self, d, wview, kw, wbuf, dense, out_pitch=0, out_offset=0, row_scale=None, lo_offset=0):
        self.weight_table.append((d, wview.data_ptr(), tuple(wview.stride()), kw, wbuf.data_ptr() if dense else None,
                                  None if dense else wbuf.data_ptr(), out_pitch, out_offset,
                                  None if row_scale is None else row_scale.data_ptr(), lo_offset))
        self._wbatch = None

    def _call(self, fn, *args):
        def op(st):
            nv.check(fn(*args, st))
        return op

    # ---- ops ----------------------------------------------------------------------------------------------
    def input_nchw(self, c, h, w):
        self.input = torch.zeros((self.batch, c, h, w), dtype=torch.float32, device=self.dev)
        return self.input

    def quantize_input(self, x_nchw, kdiv):
        n, c, h, w = x_nchw.shape
        cp = 4 if c <= 4 else _ceil(c, 16)
        t = self._alloc(n, h, w, c, "codes", kdiv, cp=cp)
        self.ops.append(self._call(self.lib.slfp_quantize_nchw_f32, x_nchw.data_ptr(), n, c, h * w, cp, kdiv, self.afmt,
                                   t.buf.data_ptr()))
        self._head_ops = len(self.ops)          # Plan.submit_host: everything up to here reads self.input
        return t

    def quantize_input_s2d(self, x_nchw, kdiv):
        """Network input -> codes of the 2x2 space-to-depth image [n, h/2, w/2, round_up(4c, 16)]."""
        n, c, h, w = x_nchw.shape
        cp = _ceil(4 * c, 16)
        t = self._alloc(n, h // 2, w // 2, 4 * c, "codes", kdiv, cp=cp)
        self.ops.append(self._call(self.lib.slfp_quantize_nchw_s2d_f32, x_nchw.data_ptr(), n, c, h, w, cp, kdiv, self.afmt,
                                   t.buf.data_ptr()))
        self._head_ops = len(self.ops)          # Plan.submit_host: everything up to here reads self.input
        return t

    def quantize_input_im2col3x3(self, x_nchw, kdiv):
        """Network input -> the im2col matrix of a 3x3 / stride 1 / padding 1 RGB stem as SLFP_FMT_F16Q halves [n, h, w, 64]
        (entry (r * 3 + s) * 4 + c): the stem becomes a plain 1x1 layer of the no-decode dense kernel."""
        n, c, h, w = x_nchw.shape
        assert c == 3
        buf = torch.zeros((n, h, w, 64), dtype=torch.float16, device=self.dev)
        self.bytes_hbm += buf.numel() * 2
        self.keep.append(buf)
        t = _T(buf, n, h, w, c, 4, "q16", kdiv, nv.FMT_F16Q)
        t.qfmt, t.im2col = self.afmt, True
        self.ops.append(self._call(self.lib.slfp_quantize_nchw_im2col3x3_f16q, x_nchw.data_ptr(), n, h, w, kdiv, self.afmt, buf.data_ptr()))
        self._head_ops = len(self.ops)          # Plan.submit_host: everything up to here reads self.input
        return t

    def quantize_input_s2d_f16q(self, x_nchw, kdiv, pad_top, pad_left, pad_bottom, pad_right):
        """Network input -> SLFP_FMT_F16Q halves of the 2x2 space-to-depth image inside a zero-padded buffer
        [n, h/2 + pads, w/2 + pads, 16]: the input of the width-folded stem (SLFP_CONV_FOLD_W)."""
        n, c, h, w = x_nchw.shape
        assert c == 3 and w % 4 == 0 and h % 2 == 0
        hp, wp = h // 2 + pad_top + pad_bottom, w // 2 + pad_left + pad_right
        buf = torch.zeros((n, hp, wp, 16), dtype=torch.float16, device=self.dev)      # the border stays zero
        self.bytes_hbm += buf.numel() * 2
        self.keep.append(buf)
        t = _T(buf, n, h // 2, w // 2, 4 * c, 16, "q16", kdiv, nv.FMT_F16Q)
        t.qfmt, t.pad = self.afmt, (pad_top, pad_left, hp, wp)
        self.ops.append(self._call(self.lib.slfp_quantize_nchw_s2d_f16q, x_nchw.data_ptr(), n, h, w, kdiv, self.afmt, pad_top, pad_left,
                                   hp, wp, buf.data_ptr()))
        self._head_ops = len(self.ops)          # Plan.submit_host: everything up to here reads self.input
        return t

    def s2d_stem(self, conv):
        """A stride-2 RxR convolution (padding P) on c channels == a stride-1 convolution on the 2x2-folded image
        with 4c channels: row 2*ho - P + r = 2*(ho - ceil(P/2)) + (r + off), off = 2*ceil(P/2) - P, so the folded
        filter tap is a = (r + off) // 2 with parity dy = (r + off) % 2; taps outside the original filter are 0.
        Returns a module-like object whose .weight is refreshed from conv.weight before every re-quantization."""
        K, C, R, S = conv.weight.shape
        assert conv.stride == (2, 2) and conv.dilation == (1, 1) and conv.groups == 1 and conv.bias is None
        P = conv.padding
        lo = [(p + 1) // 2 for p in P]
        off = [2 * l - p for l, p in zip(lo, P)]
        R2, S2 = (R - 1 + off[0]) // 2 + 1, (S - 1 + off[1]) // 2 + 1
        w2 = torch.zeros((K, 4 * C, R2, S2), dtype=torch.float32, device=self.dev)

        def refresh():
            wp = torch.zeros((K, C, 2 * R2, 2 * S2), dtype=torch.float32, device=self.dev)
            wp[:, :, off[0]:off[0] + R, off[1]:off[1] + S] = conv.weight.detach()
            # [k, c, a, dy, b, dx] -> [k, (dy, dx, c), a, b]
            w2.copy_(wp.view(K, C, R2, 2, S2, 2).permute(0, 3, 5, 1, 2, 4).reshape(K, 4 * C, R2, S2))
        self.pre_weight_ops.append(refresh)

        shim = type("S2DStem", (), {})()
        shim.weight, shim.bias, shim.Ka, shim.Kw = w2, None, conv.Ka, conv.Kw
        shim.stride, shim.padding, shim.dilation, shim.groups = (1, 1), tuple(lo), (1, 1), 1
        shim.fold = (lo, off, R2, S2)
        shim.orig = conv                 # Plan.taps: the module whose input this layer reads (in the folded layout)
        return shim

    def quantize_flat(self, x_f32, c, kdiv):
        """[n, c] float32 features -> codes [n,1,1,cp] (classifier input)."""
        n = x_f32.shape[0]
        t = self._alloc(n, 1, 1, c, "codes", kdiv)
        self.ops.append(self._call(self.lib.slfp_quantize_nhwc_f32, x_f32.data_ptr(), n, c, t.cp, kdiv, self.afmt,
                                   t.buf.data_ptr()))
        return t

    @staticmethod
    def _cp(k):
        """Physical channel count of a code tensor with k logical channels: k itself when it is a multiple of 16, else
        padded to 16 (small) or to 64 (so that the consumer's TMA path moves whole 64-byte rows: ShuffleNetV2's 58 /
        116 / 232 channels)."""
        if k % 16 == 0:
            return k
        return _ceil(k, 16) if k < 48 else _ceil(k, 64)

    def conv(self, x, mod, bn=None, relu=False, residual=None, codes=(), f16=False, f32=False, linear=False,
             relu_codes=True, layerout=0, pad_k=False, signed_fast=False, q16=False):
        """One fused convolution / linear layer.  `codes`: divisors (Ka of the consumers) to quantize-on-store
        with (at most two distinct).  Returns {'codes': {kdiv: _T}, 'f16': _T | None, 'f32': _T | None}.
        relu_codes: the consumers are dense layers / max-pools (which read the unsigned post-ReLU code format,
        written by the 2-instruction encoder); pass False when a depthwise / grouped layer consumes the codes.
        q16: the single consumer is a dense layer that should read the float16 image of the codes (Plan.f16q); pass the
        consumer's tap count (1 for a 1x1 layer, 9 for a 3x3) - the size limit of the edge depends on it."""
        assert x.kind in ("codes", "q16")
        if linear:
            K, C = mod.weight.shape
            R = S = 1
            stride, pad, dil, groups = (1, 1), (0, 0), (1, 1), 1
            wview = mod.weight.view(K, C, 1, 1)
        else:
            K, Cg, R, S = mod.weight.shape
            stride, pad, dil, groups = mod.stride, mod.padding, mod.dilation, mod.groups
            C = Cg * groups
            wview = mod.weight
        assert C == x.c, (C, x.c)
        ka, kw = _k32(mod.Ka), _k32(mod.Kw)
        assert abs(ka - x.kdiv) == 0.0, "input codes were quantized with a different Ka"
        self.taps.append((mod, x, len(self.ops)))       # (module, input codes, index of this layer's op in self.ops)
        pex = getattr(mod, "pad_extra", (0, 0))          # bottom/right minus top/left padding (space-to-depth stem)
        dense = groups == 1
        # pad_k: present a dense layer whose output-channel count is not a multiple of 16 (ShuffleNetV2: 24 / 58 / 116 /
        # 232) to the kernel with K rounded up - weight rows and affine entries of the pad channels are zero, so they
        # produce exact zeros (code 0 / 0.0) - which makes it eligible for the vectorised fast epilogues.
        Kk = self._cp(K) if (pad_k and dense and K % 16 != 0) else K
        hifi = self.high_fidelity and dense and x.cp % 16 == 0
        nodec = dense and not hifi and x.fmt == nv.FMT_E4M3 and x.cp % 64 == 0      # the codes are the tensor-core operand
        flags = nv.CONV_SPLIT_OPERANDS if hifi else (nv.CONV_E4M3_OPERANDS if nodec else 0)
        d = nv.SlfpConvDesc(x.n, x.h, x.w, C, x.cp, Kk, R, S, stride[0], stride[1], pad[0], pad[1], dil[0], dil[1], groups,
                            x.fmt, pex[0], pex[1], flags)
        dw_ = d if Kk == K else nv.SlfpConvDesc(x.n, x.h, x.w, C, x.cp, K, R, S, stride[0], stride[1], pad[0], pad[1], dil[0],
                                                 dil[1], groups, x.fmt, pex[0], pex[1], flags)       # the weight job writes K rows
        Ho = (x.h + 2 * pad[0] + pex[0] - dil[0] * (R - 1) - 1) // stride[0] + 1
        Wo = (x.w + 2 * pad[1] + pex[1] - dil[1] * (S - 1) - 1) // stride[1] + 1
        pitch = self.lib.slfp_conv_wpitch(ctypes.byref(d))
        wrow = 2 * pitch if hifi else pitch                    # split operands: rows of [hi | lo]
        wbuf = torch.zeros((Kk * wrow,), dtype=torch.float16 if (dense and not nodec) else torch.uint8, device=self.dev)
        # weight re-quantization: one table entry; Plan.prepare_weights() runs the whole table in ONE launch
        if hifi:
            self._weight_job(dw_, wview, kw, wbuf, dense, out_pitch=wrow, out_offset=0, lo_offset=pitch)
        else:
            self._weight_job(dw_, wview, kw, wbuf, dense)
        if self.high_fidelity:
            relu_codes, signed_fast = False, False          # exact quantizer codes everywhere
        epi = nv.SlfpEpilogue()
        # Fold bias, post-scale and eval BatchNorm into one per-channel affine y = acc * mul + add
        # (float64, rounded once): mul = Ka*Kw*bn_scale, add = bias_q*Ka*Kw*bn_scale + bn_shift.
        pa, pb = ((kw, ka) if linear else (ka, kw))
        mul, add = self._affine(mod, bn, K, linear)
        mul32, add32 = torch.zeros(Kk, dtype=torch.float32, device=self.dev), torch.zeros(Kk, dtype=torch.float32, device=self.dev)
        mul32[:K], add32[:K] = mul.float(), add.float()
        self.keep += [mul32, add32]
        self._guards.append((mod, wview, wview.data_ptr(), float(mod.Ka), float(mod.Kw)))
        self._affines.append((mod, bn, K, linear, mul32, add32))
        epi.ch_mul, epi.ch_add = mul32.data_ptr(), add32.data_ptr()
        epi.post_a, epi.post_b = pa, pb
        if residual is not None:
            assert residual.kind in ("f16", "f32") and (residual.n, residual.h, residual.w, residual.c) == (x.n, Ho, Wo, Kk)
            epi.residual, epi.residual_f16 = residual.buf.data_ptr(), 1 if residual.kind == "f16" else 0
        epi.relu = 1 if relu else 0
        epi.layerout = layerout            # conv -> BN -> layerout_quantize_func -> ReLU (0 = none, 1 = NaN at 0, 2 = 0 stays 0)
        if layerout:
            # codes after quantize_layerout are written with the EXACT encoder in the signed formats: 5-bit layer-out values
            # over a scale that is max / 15.5 of such values sit exactly on rounding ties of the next grid, where the fast
            # post-ReLU formats (ties up) would differ from the reference (ties to even) on a visible share of elements
            relu_codes = False
        out = {"codes": {}, "f16": None, "f32": None}
        kds = []
        for kd in codes:
            if kd not in kds:
                kds.append(kd)
        assert len(kds) <= 2
        kp = self._cp(K)
        # post-ReLU codes: dense producer (the tcgen05 kernel with the TMA-im2col gather) ending in a ReLU
        depthwise = groups > 1 and groups == C == K
        fused_dw = depthwise and bn is not None and len(kds) == 1 and not f16 and not f32 and residual is None \
            and x.cp % 16 == 0 and mod.bias is None and not layerout
        fast_dw = fused_dw and relu
        # depthwise conv -> BN -> (no ReLU) -> next quantizer: sign bit + 7-bit magnitude code (SFP<3,3> only)
        sfast_dw = fused_dw and not relu and signed_fast and self.afmt == nv.FMT_SFP33 and kp == x.cp
        # 3x3 RGB stems (c_phys = 4) run the CUDA-core direct kernel, which writes the fused pipeline's fast code formats
        stem_direct = (dense and x.cp == 4 and not x.im2col and (R, S) == (3, 3) and tuple(dil) == (1, 1) and K in (24, 32) and bn is not None
                       and not f16 and not f32 and residual is None and not layerout and stride[0] == stride[1] and pad[0] == pad[1])
        if self.e4m3 and (x.cp % 16 == 0 or stem_direct):
            ofmt = nv.FMT_E4M3                 # every producer of an SFP-7 plan (dense, depthwise, direct stem) writes e4m3 bytes
        elif sfast_dw:
            ofmt = nv.FMT_SFP33_SFAST
        elif stem_direct and relu and relu_codes:
            ofmt = nv.relu_fmt(self.afmt)
        elif stem_direct and not relu and signed_fast and self.afmt == nv.FMT_SFP33:
            ofmt = nv.FMT_SFP33_SFAST
        else:
            ofmt = nv.relu_fmt(self.afmt) if (relu_codes and relu and (x.cp % 16 == 0 or x.im2col) and (groups == 1 or fast_dw)) else self.afmt
        # store_f16: the codes-only fast epilogue of the warp-specialised dense kernel (relu, one consumer, post-ReLU format)
        # Storing halves instead of bytes costs the PRODUCER ~0.5 us per 2^20 output elements (measured on the 1x1 reduce
        # layers, whose 8 epilogue warps are their critical path) against a flat ~12-18 us the 3x3 consumer saves by not
        # decoding: worth it below ~16 M elements (ResNet-50 stages 3-4 at batch 256)
        # (a producer that itself reads float16 images has no decode warps competing with its epilogue's table look-ups: 64 M)
        # (measured per edge kind, SLFP_F16Q_MAX_ELEMS A/B on one box: a 1x1 CONSUMER decodes each element once anyway and pays double
        # the bytes - ResNet-50 3.80 -> 3.74 ms with those edges on codes; a 3x3 consumer saves nine decodes - VGG-16 0.521 -> 0.504 ms)
        if int(q16) == 1 or (R * S == 1 and x.kind != "q16"):
            q16_max = int(os.environ.get("SLFP_F16Q_MAX_ELEMS_1X1", 4 << 20))
        else:
            q16_max = int(os.environ.get("SLFP_F16Q_MAX_ELEMS", (64 << 20) if x.kind == "q16" else (16 << 20)))
        as_q16 = (q16 and self.f16q and dense and (x.cp % 16 == 0 or x.im2col) and len(kds) == 1 and relu and bn is not None and K % 64 == 0
                  and Kk == K and ofmt == nv.relu_fmt(self.afmt) and not f16 and not f32 and residual is None and not layerout
                  and x.n * Ho * Wo * K <= q16_max)
        for i, kd in enumerate(kds):
            t = self._alloc(x.n, Ho, Wo, K, "q16" if as_q16 else "codes", kd, cp=kp, fmt=ofmt)
            out["codes"][kd] = t
            if i == 0:
                epi.y_codes, epi.next_k_div = t.buf.data_ptr(), kd
            else:
                epi.y_codes2, epi.next_k_div2 = t.buf.data_ptr(), kd
        epi.next_fmt, epi.k_phys_out, epi.store_f16 = ofmt, kp, 1 if as_q16 else 0
        assert Kk == K or not kds or kp == Kk
        if f16:
            out["f16"] = self._alloc(x.n, Ho, Wo, Kk, "f16")      # physical channels (pad channels hold 0.0)
            out["f16"].c_logical = K
            epi.y_f16 = out["f16"].buf.data_ptr()
        if f32:
            out["f32"] = self._alloc(x.n, Ho, Wo, Kk, "f32")
            epi.y_f32 = out["f32"].buf.data_ptr()
        d_launch = d
        if x.im2col:
            # 3x3 RGB stem on its explicit im2col matrix: a 1x1 layer with 64 input channels (36 used), same weight operand
            assert dense and x.cp == 4 and (R, S) == (3, 3) and tuple(stride) == (1, 1) and tuple(pad) == (1, 1) and tuple(dil) == (1, 1)
            assert pitch == 64
            d_launch = nv.SlfpConvDesc(x.n, x.h, x.w, 64, 64, Kk, 1, 1, 1, 1, 0, 0, 1, 1, 1, nv.FMT_F16Q, 0, 0, 0)
        if x.pad is not None:
            # width-folded stem: R x 4 taps on 16 channels presented as R x 1 on 64 over the physically padded input
            top, left, hp, wp = x.pad
            assert dense and x.fmt == nv.FMT_F16Q and x.cp == 16 and S == 4 and wp == Wo + 3 and hp == Ho + R - 1 and tuple(stride) == (1, 1)
            d_launch = nv.SlfpConvDesc(x.n, hp, Wo, 64, 64, Kk, R, 1, 1, 1, 0, 0, 1, 1, 1, nv.FMT_F16Q, 0, 0, nv.CONV_FOLD_W)
            assert self.lib.slfp_conv_wpitch(ctypes.byref(d_launch)) == pitch
        self.keep += [d, d_launch, epi, wbuf]
        self.ops.append(self._call(self.lib.slfp_conv2d_fwd, ctypes.byref(d_launch), x.buf.data_ptr(), wbuf.data_ptr(),
                                   ctypes.byref(epi)))
        fl = 2.0 * x.n * Ho * Wo * K * (C // groups) * R * S
        if hasattr(mod, "orig"):        # space-to-depth stem: count the ALGORITHMIC taps (7x7x3 = 147), not the folded 4x4x12 = 192
            fl = 2.0 * x.n * Ho * Wo * K * float(np.prod(mod.orig.weight.shape[1:]))
        self.flops += fl
        # algorithmic HBM bytes of the launch: codes in, everything written, the weight operand, the residual read
        by = x.buf.numel() * x.buf.element_size() + wbuf.numel() * wbuf.element_size() + (residual.buf.numel() * residual.buf.element_size() if residual is not None else 0)
        by += sum(t.buf.numel() * t.buf.element_size() for t in out["codes"].values()) + sum(out[k].buf.numel() * out[k].buf.element_size() for k in ("f16", "f32") if out[k] is not None)
        self.conv_flops.append((fl, groups == 1, f"{C}->{K} {R}x{S} s{stride[0]} @{x.h}", by))
        return out

    def _affine(self, mod, bn, K, linear=False):
        """(mul, add) float64 per output channel: y = acc * mul + add folds bias, the post-scale Ka*Kw and eval BN."""
        ka, kw = _k32(mod.Ka), _k32(mod.Kw)
        pa, pb = ((kw, ka) if linear else (ka, kw))
        post = torch.full((K,), float(np.float32(pa)) * float(np.float32(pb)), dtype=torch.float64, device=self.dev)
        mul, add = post.clone(), torch.zeros((K,), dtype=torch.float64, device=self.dev)
        if mod.bias is not None:
            b = mod.bias.detach().double().to(self.dev)
            # conv2d_Q_bias: bias/Ka/Kw (conv2d_func.py:44); linear_Q: bias/Kw/Ka (:63); plain conv2d_Q: raw bias (:23)
            if linear:
                bq = b / float(mod.Kw) / float(mod.Ka)
            elif getattr(mod, "_slfp_bias_scaled", True):
                bq = b / float(mod.Ka) / float(mod.Kw)
            else:
                bq = b
            add = bq * post
        if bn is not None:
            g = bn.weight.detach().double() if bn.weight is not None else torch.ones(K, dtype=torch.float64, device=self.dev)
            be = bn.bias.detach().double() if bn.bias is not None else torch.zeros(K, dtype=torch.float64, device=self.dev)
            sc_ = (g / torch.sqrt(bn.running_var.detach().double() + bn.eps)).to(self.dev)
            sh_ = be.to(self.dev) - bn.running_mean.detach().double().to(self.dev) * sc_
            mul, add = mul * sc_, add * sc_ + sh_
        return mul, add

    def conv_dual(self, x1, mod1, bn1, x2, mod2, bn2, relu=True, codes=(), f16=False):
        """Fused residual-block tail (nets_imgnet/resnet50.py:80-88 for a block with a downsample branch):
        relu(bn1(conv1x1(x1)) + bn2(conv1x1_stride(x2))) as ONE GEMM over the concatenated K dimension.  Per output
        channel the branch with the larger folded scale keeps its exact float16 weights and supplies the epilogue's
        multiplier; the other branch's weights carry the ratio (<= 1 in magnitude)."""
        assert x1.kind == "codes" and x2.kind == "codes" and x1.fmt == x2.fmt
        K, C1 = mod1.weight.shape[:2]
        K2, C2 = mod2.weight.shape[:2]
        assert K == K2 and mod1.weight.shape[2:] == (1, 1) and mod2.weight.shape[2:] == (1, 1)
        assert mod1.bias is None and mod2.bias is None and x1.cp % 64 == 0 and x2.cp % 64 == 0 and K % 16 == 0
        assert abs(_k32(mod1.Ka) - x1.kdiv) == 0.0 and abs(_k32(mod2.Ka) - x2.kdiv) == 0.0
        self.taps += [(mod1, x1, len(self.ops)), (mod2, x2, len(self.ops))]
        d1 = nv.SlfpConvDesc(x1.n, x1.h, x1.w, C1, x1.cp, K, 1, 1, 1, 1, 0, 0, 1, 1, 1, x1.fmt, 0, 0)
        s2 = mod2.stride
        d2 = nv.SlfpConvDesc(x2.n, x2.h, x2.w, C2, x2.cp, K, 1, 1, s2[0], s2[1], 0, 0, 1, 1, 1, x2.fmt, 0, 0)
        Ho, Wo = x1.h, x1.w
        assert ((x2.h - 1) // s2[0] + 1, (x2.w - 1) // s2[1] + 1) == (Ho, Wo) and x1.n == x2.n
        p1 = self.lib.slfp_conv_wpitch(ctypes.byref(d1))
        p2 = self.lib.slfp_conv_wpitch(ctypes.byref(d2))
        wbuf = torch.empty((K * (p1 + p2),), dtype=torch.float16, device=self.dev)
        mul1, add1 = self._affine(mod1, bn1, K)
        mul2, add2 = self._affine(mod2, bn2, K)
        ref = torch.where(mul1.abs() >= mul2.abs(), mul1, mul2)
        ref = torch.where(ref == 0, torch.ones_like(ref), ref)
        r1, r2 = (mul1 / ref).float().contiguous(), (mul2 / ref).float().contiguous()
        self._weight_job(d1, mod1.weight, _k32(mod1.Kw), wbuf, True, p1 + p2, 0, r1)
        self._weight_job(d2, mod2.weight, _k32(mod2.Kw), wbuf, True, p1 + p2, p1, r2)
        epi = nv.SlfpEpilogue()
        mul32, add32 = ref.float().contiguous(), (add1 + add2).float().contiguous()
        epi.ch_mul, epi.ch_add, epi.relu = mul32.data_ptr(), add32.data_ptr(), 1 if relu else 0
        out = {"codes": {}, "f16": None, "f32": None}
        kds = []
        for kd in codes:
            if kd not in kds:
                kds.append(kd)
        assert len(kds) <= 2
        ofmt = nv.relu_fmt(self.afmt) if relu else self.afmt
        for i, kd in enumerate(kds):
            t = self._alloc(x1.n, Ho, Wo, K, "codes", kd, cp=K, fmt=ofmt)
            out["codes"][kd] = t
            if i == 0:
                epi.y_codes, epi.next_k_div = t.buf.data_ptr(), kd
            else:
                epi.y_codes2, epi.next_k_div2 = t.buf.data_ptr(), kd
        epi.next_fmt, epi.k_phys_out = ofmt, K
        if f16:
            out["f16"] = self._alloc(x1.n, Ho, Wo, K, "f16")
            epi.y_f16 = out["f16"].buf.data_ptr()
        self.keep += [d1, d2, epi, wbuf, mul32, add32, r1, r2]
        self._guards += [(mod1, mod1.weight, mod1.weight.data_ptr(), float(mod1.Ka), float(mod1.Kw)),
                         (mod2, mod2.weight, mod2.weight.data_ptr(), float(mod2.Ka), float(mod2.Kw))]
        self.ops.append(self._call(self.lib.slfp_conv2d_fwd_dual, ctypes.byref(d1), x1.buf.data_ptr(), ctypes.byref(d2),
                                   x2.buf.data_ptr(), wbuf.data_ptr(), ctypes.byref(epi)))
        fl = 2.0 * x1.n * Ho * Wo * K * (C1 + C2)
        self.flops += fl
        by = x1.buf.numel() + x2.buf.numel() + wbuf.numel() * 2 + sum(t.buf.numel() for t in out["codes"].values()) + \
            (out["f16"].buf.numel() * 2 if out["f16"] is not None else 0)
        self.conv_flops.append((fl, True, f"{C1}+{C2}->{K} 1x1 dual @{x1.h}", by))
        return out

    def gather_quantize(self, chans, kdiv):
        """Codes of a LOGICAL tensor given as a channel list [(float16 _T, channel), ...]: the consumer-side form of
        torch.split / torch.cat / channel_shuffle (nets_cifar/shufflenet_v2.py:20-45, 100-115) - no data is moved for the
        shuffle itself, the activation quantizer reads each logical channel from wherever it lives."""
        t0 = chans[0][0]
        assert all(t.kind == "f16" and (t.n, t.h, t.w) == (t0.n, t0.h, t0.w) for t, _ in chans)
        c = len(chans)
        fmt = nv.FMT_E4M3 if self.e4m3 else self.afmt
        out = self._alloc(t0.n, t0.h, t0.w, c, "codes", kdiv, cp=self._cp(c) if not self.e4m3 else _ceil(c, 64), fmt=fmt)
        # runs: per source tensor, consecutive channels whose logical positions form an arithmetic sequence
        by_src = {}
        for j, (t, ch) in enumerate(chans):
            assert 0 <= ch < t.c_logical
            by_src.setdefault(id(t), (t, []))[1].append((ch, j))
        runs = []
        for t, lst in by_src.values():
            lst.sort()
            i = 0
            while i < len(lst):
                k, step = i + 1, None
                while k < len(lst) and lst[k][0] == lst[k - 1][0] + 1 and (step is None or lst[k][1] - lst[k - 1][1] == step) \
                        and lst[k][1] > lst[k - 1][1]:
                    step = lst[k][1] - lst[k - 1][1]
                    k += 1
                runs.append((t, lst[i][0], k - i, lst[i][1], step or 1))
                i = k
        tab = (nv.SlfpGatherRun * len(runs))()
        mg, sh = ctypes.c_uint(), ctypes.c_uint()
        stage_bpp = 0
        for e, (t, ch0, ln, d0, st_) in zip(tab, runs):
            assert t.c % 8 == 0, "source rows must be 16-byte multiples (8 float16)"
            nck = (((ch0 + ln + 7) & ~7) - (ch0 & ~7)) >> 3        # 16-byte chunks that cover the run: the kernel divides by this
            stage_bpp += 16 * nck
            self.lib.slfp_magic_u32(max(nck, 1), ctypes.byref(mg), ctypes.byref(sh))
            e.src, e.stride, e.ch0, e.len, e.dst_start, e.dst_step, e.magic, e.shift = t.buf.data_ptr(), t.c, ch0, ln, d0, st_, mg.value, sh.value
        dtab = torch.frombuffer(bytearray(bytes(tab)), dtype=torch.uint8).to(self.dev)
        self.keep.append(dtab)
        self.ops.append(self._call(self.lib.slfp_gather_quantize_runs_f16, dtab.data_ptr(), len(runs), t0.n * t0.h * t0.w, c, out.cp,
                                   stage_bpp, kdiv, fmt, out.buf.data_ptr()))
        return out

    def maxpool(self, x, k, stride, pad):
        # nn.MaxPool2d accepts ints or pairs; the kernel takes square windows
        def one(v):
            if isinstance(v, (tuple, list)):
                assert len(v) == 2 and v[0] == v[1], "square max-pool windows / strides / paddings only"
                return int(v[0])
            return int(v)
        k, stride, pad = one(k), one(stride if stride is not None else k), one(pad)
        assert x.kind == "codes" and x.fmt != nv.FMT_E4M3
        Ho = (x.h + 2 * pad - k) // stride + 1
        Wo = (x.w + 2 * pad - k) // stride + 1
        t = self._alloc(x.n, Ho, Wo, x.c, "codes", x.kdiv, cp=x.cp, fmt=x.fmt)
        self.ops.append(self._call(self.lib.slfp_maxpool_codes, x.buf.data_ptr(), x.n, x.h, x.w, x.cp, x.fmt, k, k, stride, pad,
                                   t.buf.data_ptr()))
        return t

    def avgpool(self, x):
        """Global average pool of an f16 / f32 NHWC tensor -> [n, c] float32."""
        out = torch.empty((x.n, x.c), dtype=torch.float32, device=self.dev)
        self.keep.append(out)
        self.ops.append(self._call(self.lib.slfp_avgpool_nhwc, x.buf.data_ptr(), 1 if x.kind == "f16" else 0, x.n,
                                   x.h * x.w, x.c, out.data_ptr()))
        return out

    def avgpool_quantize(self, x, kdiv):
        """Global average pool + the classifier's activation quantizer: [n, h, w, c] float16 -> codes [n, 1, 1, c].  One
        launch (slfp_avgpool_quantize_nhwc_f16) when the formats allow it, else avgpool() + quantize_flat() - same bits."""
        if (x.kind == "f16" and x.c % 16 == 0 and self.afmt in (nv.FMT_SFP33, nv.FMT_SLFP34_ACT)
                and not os.environ.get("SLFP_NO_FUSED_AVGPOOL")):
            t = self._alloc(x.n, 1, 1, x.c, "codes", kdiv)
            self.ops.append(self._call(self.lib.slfp_avgpool_quantize_nhwc_f16, x.buf.data_ptr(), x.n, x.h * x.w, x.c, None, kdiv,
                                       self.afmt, t.buf.data_ptr()))
            return t
        return self.quantize_flat(self.avgpool(x), x.c, kdiv)

    def torch_op(self, fn):
        """Escape hatch for plain (un-quantized) library layers of the caller net, e.g. MobileNetV1's nn.Linear."""
        self.ops.append(lambda st: fn())

    # ---- consistency with the model the plan was compiled from ---------------------------------------------------
    def verify(self):
        """Raise if the model changed in a way the pre-bound calls cannot follow: a weight tensor that moved (model.to(),
        .half(), a re-assigned Parameter) or a scale (Ka / Kw, e.g. after set_scales()) that differs from the compiled one -
        both are baked into descriptors and epilogues, so the plan has to be compiled again."""
        for mod, wview, ptr, ka, kw in self._guards:
            # (the space-to-depth stem shim owns a derived weight tensor that is refreshed before every re-quantization)
            if not hasattr(mod, "orig") and mod.weight.data_ptr() != ptr:
                raise RuntimeError("engine.Plan: a weight tensor was re-allocated after the plan was compiled; compile the plan again")
            if float(mod.Ka) != ka or float(mod.Kw) != kw:
                raise RuntimeError("engine.Plan: Ka / Kw changed after the plan was compiled (set_scales?); compile the plan again")

    @torch.no_grad()
    def refresh(self):
        """Recompute the folded per-channel vectors (bias, eval BatchNorm) IN PLACE from the modules' current parameters - the
        buffers keep their addresses, so a captured CUDA graph sees the new values.  Weights are re-quantized on every run
        anyway; scales need a new plan (verify())."""
        self.verify()
        for mod, bn, K, linear, mul32, add32 in self._affines:
            mul, add = self._affine(mod, bn, K, linear)
            mul32[:K].copy_(mul.float())
            add32[:K].copy_(add.float())

    # ---- execution ----------------------------------------------------------------------------------------
    def prepare_weights(self):
        """Re-quantize every layer's weights (utils/conv2d_func.py:22 does it on every forward): one launch."""
        n = len(self.weight_table)
        if n == 0:
            return
        for op in self.pre_weight_ops:
            op()
        if self._wbatch is None:
            jobs = (nv.SlfpWeightJob * n)()
            for j, (d, w, strides, kw, f16, codes, out_pitch, out_offset, row_scale, lo_offset) in zip(jobs, self.weight_table):
                j.desc, j.w, j.kw, j.w_f16, j.w_codes = ctypes.pointer(d), w, kw, f16, codes
                j.w_stride[:] = strides
                j.out_pitch, j.out_offset, j.row_scale, j.lo_offset = out_pitch, out_offset, row_scale, lo_offset
            self._wbatch = jobs
        nv.check(self.lib.slfp_prepare_weights_jobs(n, self._wbatch, self.wfmt, nv.stream()))

    def _run_head(self):
        """Weight re-quantization and the input quantizer do not depend on each other: the first runs on a side stream
        (a fork / join that a CUDA-graph capture records as two parallel branches) while the second - HBM-bound, where
        the weight kernel is latency-bound - runs on the caller's stream.  Returns the number of ops issued."""
        h = getattr(self, "_head_ops", 0)
        if self.static_weights or not self.weight_table:
            return 0
        if h == 0 or os.environ.get("SLFP_NO_WPREP_OVERLAP"):
            self.prepare_weights()
            return 0
        main = torch.cuda.current_stream(self.dev)
        if getattr(self, "_side_stream", None) is None:
            self._side_stream = torch.cuda.Stream(device=self.dev)
        side = self._side_stream
        side.wait_stream(main)
        with torch.cuda.stream(side):
            self.prepare_weights()
        st = nv.stream()
        for op in self.ops[:h]:
            op(st)
        main.wait_stream(side)
        return h

    @torch.no_grad()
    def run(self):
        done = self._run_head()
        st = nv.stream()
        for op in self.ops[done:]:
            op(st)
        return self.output

    @property
    def launches_per_step(self):
        return len(self.ops) + (0 if self.static_weights or not self.weight_table else 1)

    def capture(self):
        """Record the plan into a CUDA graph (after one eager warm-up so kernel attributes are set)."""
        self.verify()
        self.prepare_weights()
        self.run()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self.run()
        self.graph = g
        return g

    # ---- host-buffer pipeline -----------------------------------------------------------------------------------------
    def capture_host_pipeline(self):
        """Two CUDA graphs for submit_host(): HEAD = weight re-quantization + the input quantizer (the only reader of
        self.input), TAIL = every other layer."""
        h = getattr(self, "_head_ops", 0)
        if h == 0:
            raise RuntimeError("engine.Plan: the plan has no input quantizer")
        self.verify()
        self.prepare_weights()
        self.run()
        torch.cuda.synchronize()
        gh, gt = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        with torch.no_grad():
            with torch.cuda.graph(gh):
                done = self._run_head()
                for op in self.ops[done:h]:
                    op(nv.stream())
            with torch.cuda.graph(gt):
                for op in self.ops[h:]:
                    op(nv.stream())
        self._host_graphs = (gh, gt)
        self._copy_stream = torch.cuda.Stream(device=self.dev)
        self._ev_ready, self._ev_free = torch.cuda.Event(), torch.cuda.Event()
        self._ev_free_valid = False

    def submit_host(self, x_host, out_host=None):
        """One forward from a PINNED host batch [batch, c, h, w] float32, fully asynchronous: the host -> device copy goes
        straight into the plan's input buffer on a copy stream and only waits for the PREVIOUS submit's input quantizer
        (the buffer's only reader), so it overlaps that submit's remaining layers; then head graph, tail graph and - if
        out_host (pinned) is given - the device -> host copy of the logits on the caller's stream.  No staging copy on
        the device.  Call torch.cuda.synchronize() (or wait on your own event) before reading out_host."""
        if getattr(self, "_host_graphs", None) is None:
            self.capture_host_pipeline()
        gh, gt = self._host_graphs
        main = torch.cuda.current_stream(self.dev)
        with torch.cuda.stream(self._copy_stream):
            if self._ev_free_valid:
                self._copy_stream.wait_event(self._ev_free)
            self.input.copy_(x_host, non_blocking=True)
            self._ev_ready.record(self._copy_stream)
        main.wait_event(self._ev_ready)
        gh.replay()
        self._ev_free.record(main)
        self._ev_free_valid = True
        gt.replay()
        if out_host is not None:
            out_host.copy_(self.output, non_blocking=True)
        return self.output

    def __call__(self, x=None):
        if x is not None and x.data_ptr() != self.input.data_ptr():
            self.input.copy_(x, non_blocking=True)
        if self.graph is not None:
            self.graph.replay()
        else:
            self.run()
        return self.output


# ---- per-architecture compilers -------------------------------------------------------------------------------
def compile_resnet50(model, batch, size=224, device="cuda", residual="f16", static_weights=False, fuse_downsample=True,
                     high_fidelity=False):
    """nets_imgnet.ResNet50 (or the reference's own class: same attribute names) -> Plan.
    high_fidelity: split-operand tensor-core mode, exact signed codes, float32 residual stream, no fused block tails -
    about 3x the tensor work; the parity instrument (tests/test_gpu_parity.py), not the benchmark path."""
    assert not model.training, "the fused pipeline folds BatchNorm: call model.eval() first"
    if high_fidelity:
        residual, fuse_downsample = "f32", False
    P = Plan(batch, device, model.qbit if hasattr(model, "qbit") else model.conv1.q_bit, static_weights, high_fidelity)
    res_f16 = residual == "f16"
    q16_edges = P.f16q
    x = P.input_nchw(3, size, size)
    blocks = [b for li in range(1, 5) for b in getattr(model, f"layer{li}")]

    def consumers(block):
        ks = [_k32(block.conv1.Ka)]
        if block.downsample is not None:
            ks.append(_k32(block.downsample[0].Ka))
        return ks

    c1 = model.conv1
    if c1.stride == (2, 2) and size % 2 == 0 and c1.groups == 1 and c1.bias is None and c1.dilation == (1, 1):
        # 7x7/2 stem on 3 channels -> 4x4/1 on the space-to-depth image (16-byte channel vectors: TMA im2col path)
        shim = P.s2d_stem(c1)
        lo, off, R2, S2 = shim.fold
        ho = (size + 2 * c1.padding[0] - (c1.kernel_size[0] - 1) - 1) // 2 + 1
        wo = (size + 2 * c1.padding[1] - (c1.kernel_size[1] - 1) - 1) // 2 + 1
        shim.pad_extra = (ho - size // 2 + (R2 - 1) - 2 * lo[0], wo - size // 2 + (S2 - 1) - 2 * lo[1])
        if P.f16q and R2 == 4 and S2 == 4 and size % 4 == 0 and not os.environ.get("SLFP_NO_FOLD_STEM"):
            # float16 images in a zero-padded buffer: the stem runs without a decode stage, one 128-byte TMA row per
            # pixel and filter row (SLFP_CONV_FOLD_W)
            xc = P.quantize_input_s2d_f16q(x, _k32(c1.Ka), lo[0], lo[1], lo[0] + shim.pad_extra[0], lo[1] + shim.pad_extra[1])
        else:
            xc = P.quantize_input_s2d(x, _k32(c1.Ka))
        stem = P.conv(xc, shim, bn=model.bn1, relu=True, codes=consumers(blocks[0]))
    else:
        xc = P.quantize_input(x, _k32(c1.Ka))
        stem = P.conv(xc, c1, bn=model.bn1, relu=True, codes=consumers(blocks[0]))
    mp = model.maxpool
    k_, s_, p_ = (mp.kernel_size, mp.stride, mp.padding)
    cur_codes = {kd: P.maxpool(t, k_, s_, p_) for kd, t in stem["codes"].items()}
    cur_res = None
    for i, b in enumerate(blocks):
        nxt = blocks[i + 1] if i + 1 < len(blocks) else None
        ds0 = b.downsample[0] if b.downsample is not None else None
        fuse = (fuse_downsample and ds0 is not None and res_f16 and ds0.kernel_size == (1, 1) and ds0.padding == (0, 0)
                and ds0.bias is None and b.conv3.bias is None and cur_codes[_k32(ds0.Ka)].cp % 64 == 0
                and b.conv2.out_channels % 64 == 0 and b.conv3.out_channels % 16 == 0)
        # conv1 -> conv2 (3x3: nine taps x N tiles of table look-ups per input code) and conv2 -> conv3 (short K) exchange
        # float16 images; the dual-input tail of a down-sampling block reads codes on both inputs
        o1 = P.conv(cur_codes[_k32(b.conv1.Ka)], b.conv1, bn=b.bn1, relu=True, codes=[_k32(b.conv2.Ka)], q16=9 if q16_edges else 0)
        o2 = P.conv(o1["codes"][_k32(b.conv2.Ka)], b.conv2, bn=b.bn2, relu=True, codes=[_k32(b.conv3.Ka)], q16=1 if (q16_edges and not fuse) else 0)
        need_val = nxt is None or nxt.downsample is None        # someone adds / pools the un-quantized value
        if fuse:
            # block tail and downsample branch as ONE GEMM: no downsample launch, no float16 round trip of its output
            o3 = P.conv_dual(o2["codes"][_k32(b.conv3.Ka)], b.conv3, b.bn3, cur_codes[_k32(ds0.Ka)], ds0, b.downsample[1],
                             relu=True, codes=consumers(nxt) if nxt is not None else [], f16=need_val)
            cur_codes = o3["codes"]
            cur_res = o3["f16"]
            continue
        if b.downsample is not None:
            ds = P.conv(cur_codes[_k32(b.downsample[0].Ka)], b.downsample[0], bn=b.downsample[1], relu=False,
                        f16=res_f16, f32=not res_f16)
            res = ds["f16"] if res_f16 else ds["f32"]
        else:
            res = cur_res
        o3 = P.conv(o2["codes"][_k32(b.conv3.Ka)], b.conv3, bn=b.bn3, relu=True, residual=res,
                    codes=consumers(nxt) if nxt is not None else [], f16=need_val and res_f16, f32=need_val and not res_f16)
        cur_codes = o3["codes"]
        cur_res = o3["f16"] if res_f16 else o3["f32"]
    fcq = P.avgpool_quantize(cur_res, _k32(model.fc.Ka))
    out = P.conv(fcq, model.fc, f32=True, linear=True)
    P.output = out["f32"].buf.view(batch, -1)
    return P


def compile_vgg16(model, batch, size=32, device="cuda", static_weights=False):
    """nets_cifar.VGG16_Q -> Plan (conv+bias -> BN -> ReLU chains, 2x2 max-pools on codes, 3 quantized FCs)."""
    assert not model.training
    first = model.layer1[0]
    P = Plan(batch, device, first.q_bit, static_weights)
    x = P.input_nchw(3, size, size)
    seq = []
    for li in range(1, 6):
        seq += list(getattr(model, f"layer{li}"))
    fcs = [model.fc1[2], model.fc2[0], model.fc3]
    convs = [m for m in seq if isinstance(m, nn.Conv2d)]
    c0 = convs[0]
    if (P.f16q and c0.kernel_size == (3, 3) and c0.stride == (1, 1) and c0.padding == (1, 1) and c0.dilation == (1, 1) and c0.groups == 1
            and c0.in_channels == 3 and not os.environ.get("SLFP_NO_IM2COL_STEM")):
        cur = P.quantize_input_im2col3x3(x, _k32(c0.Ka))      # the 3 -> 64 stem as a 1x1 layer on its im2col matrix
    else:
        cur = P.quantize_input(x, _k32(c0.Ka))
    i = 0
    while i < len(seq):
        m = seq[i]
        if isinstance(m, nn.Conv2d):
            bn = seq[i + 1]
            assert isinstance(bn, nn.BatchNorm2d) and isinstance(seq[i + 2], nn.ReLU)
            ci = convs.index(m)
            nxt_k = _k32(convs[ci + 1].Ka) if ci + 1 < len(convs) else _k32(fcs[0].Ka)
            # conv -> conv edges (no pool in between) hand over float16 images: the 3x3 consumer skips its decode stage
            to_conv = i + 3 < len(seq) and isinstance(seq[i + 3], nn.Conv2d)
            cur = P.conv(cur, m, bn=bn, relu=True, codes=[nxt_k], q16=9 if to_conv else 0)["codes"][nxt_k]
            i += 3
        elif isinstance(m, nn.MaxPool2d):
            cur = P.maxpool(cur, m.kernel_size, m.stride, m.padding)
            i += 1
        else:
            raise NotImplementedError(type(m))
    assert cur.h == 1 and cur.w == 1, "VGG16_Q plan expects a 1x1 map before the classifier (32x32 input)"
    for j, fc in enumerate(fcs):
        last = j == len(fcs) - 1
        nk = None if last else _k32(fcs[j + 1].Ka)
        o = P.conv(cur, fc, relu=not last, codes=[] if last else [nk], f32=last, linear=True)
        cur = o["f32"] if last else o["codes"][nk]
    P.output = cur.buf.view(batch, -1)
    return P


def compile_mobilenetv1(model, batch, size, device="cuda", static_weights=False):
    """nets_imgnet / nets_cifar MobileNetV1_Q -> Plan (stem, 13 depthwise + pointwise pairs, pool, classifier)."""
    assert not model.training
    feats = list(model.model)
    pool = feats[-1]
    blocks = feats[:-1]
    stem = blocks[0]
    P = Plan(batch, device, stem[0].q_bit, static_weights)
    x = P.input_nchw(3, size, size)
    layers = []                                   # (conv, bn) in order
    for blk in blocks:
        mods = list(blk)
        for j in range(0, len(mods), 3):
            layers.append((mods[j], mods[j + 1]))
    quant_fc = hasattr(model.fc, "Ka")
    cur = P.quantize_input(x, _k32(layers[0][0].Ka))
    for j, (conv, bn) in enumerate(layers):
        last = j == len(layers) - 1
        if last:
            o = P.conv(cur, conv, bn=bn, relu=True, f16=True)
            cur = o["f16"]
        else:
            nk = _k32(layers[j + 1][0].Ka)
            cur = P.conv(cur, conv, bn=bn, relu=True, codes=[nk])["codes"][nk]
    if isinstance(pool, nn.AvgPool2d):
        assert cur.h == pool.kernel_size and cur.w == pool.kernel_size, "AvgPool2d(7) expects a 7x7 map (224x224 input)"
    if quant_fc:
        fcq = P.avgpool_quantize(cur, _k32(model.fc.Ka))
        P.output = P.conv(fcq, model.fc, f32=True, linear=True)["f32"].buf.view(batch, -1)
    else:
        feat = P.avgpool(cur)
        out = torch.empty((batch, model.fc.out_features), dtype=torch.float32, device=device)
        P.keep.append(out)
        P.torch_op(lambda: torch.addmm(model.fc.bias, feat, model.fc.weight.t(), out=out))
        P.output = out
    return P


def compile_shufflenetv2(model, batch, size, device="cuda", static_weights=False, nan_at_zero=False):
    """nets_cifar.ShuffleNetV2 -> Plan (BASELINE config 5's second net; reference nets_cifar/shufflenet_v2.py:47-252).

    A ShuffleUnit is  out = channel_shuffle(cat(shortcut(x1), residual(x2)), 2)  with (x1, x2) = split(x) in a basic
    unit and x1 = x2 = x in a down-sampling unit.  Here split / cat / shuffle never move data: a unit's output is kept as
    a LOGICAL channel list [(float16 tensor, channel)] - the pass-through half keeps pointing at the tensors that hold
    its values (float16 is an exact carrier for post-layerout, post-ReLU SFP<4,4> values), the branch output is one new
    float16 tensor - and the next consumer's activation quantizer gathers its input channels through that list
    (Plan.gather_quantize).  Inside a branch the layers exchange 8-bit codes:
        1x1 conv -> BN -> layerout -> ReLU -> [quantize]  ->  dw 3x3 -> BN -> [quantize]  ->  1x1 conv -> BN -> layerout -> ReLU
    with BN, quantize_layerout (SFP<4,4>), ReLU and the next layer's quantizer in the producing conv's epilogue.
    Output-channel counts that are not multiples of 16 (24 / 58 / 116 / 232) are padded for the kernels (zero weight
    rows), the depthwise layers - which have no ReLU - write the signed fast code format.
    nan_at_zero: the reference's quantize_layerout returns NaN for an exact 0 (its `2^(-8)` is an XOR,
    utils/sfp_quant.py:122-123); the fast epilogues map 0 to 0.  True selects the generic, bug-compatible epilogues
    (several times slower); the module-level drop-in is always bug-compatible."""
    assert not model.training
    from .utils import sfp_quant
    lo = 1 if (nan_at_zero and not getattr(sfp_quant, "LAYEROUT_ZERO_IS_ZERO", False)) else 2
    fast = lo == 2
    pre_conv, pre_bn = model.pre[0], model.pre[1]
    P = Plan(batch, device, pre_conv.q_bit, static_weights)
    x = P.input_nchw(3, size, size)
    units = [u for si in (2, 3, 4) for u in getattr(model, f"stage{si}")]

    def is_down(u):
        return len(u.shortcut) > 0

    def residual_branch(xc, mods):
        c0, b0, dw, b1, c2, b2 = mods[0], mods[1], mods[4], mods[5], mods[6], mods[7]
        k1, k2 = _k32(dw.Ka), _k32(c2.Ka)
        t = P.conv(xc, c0, bn=b0, relu=True, layerout=lo, codes=[k1], pad_k=fast)["codes"][k1]
        t = P.conv(t, dw, bn=b1, relu=False, codes=[k2], relu_codes=False, signed_fast=fast)["codes"][k2]
        return P.conv(t, c2, bn=b2, relu=True, layerout=lo, f16=True, pad_k=fast)["f16"]

    def shortcut_branch(xc, mods):
        dw, b0, c1, b1 = mods[0], mods[1], mods[2], mods[3]
        k1 = _k32(c1.Ka)
        t = P.conv(xc, dw, bn=b0, relu=False, codes=[k1], relu_codes=False, signed_fast=fast)["codes"][k1]
        return P.conv(t, c1, bn=b1, relu=True, layerout=lo, f16=True, pad_k=fast)["f16"]

    # stem: conv -> BN (no ReLU, no layerout); its output is consumed only by the first unit's two quantizers
    first = units[0]
    assert is_down(first)
    kr, ks = _k32(first.residual[0].Ka), _k32(first.shortcut[0].Ka)
    xc = P.quantize_input(x, _k32(pre_conv.Ka))
    stem = P.conv(xc, pre_conv, bn=pre_bn, relu=False, codes=[kr, ks], relu_codes=False, signed_fast=fast)
    direct = stem["codes"]                 # {Ka: code tensor} of the whole stem output
    chans = None
    for u in units:
        if is_down(u):
            kr, ks = _k32(u.residual[0].Ka), _k32(u.shortcut[0].Ka)
            if chans is None:
                xr, xs = direct[kr], direct[ks]
            else:
                xr = P.gather_quantize(chans, kr)
                xs = xr if ks == kr else P.gather_quantize(chans, ks)
            sc = shortcut_branch(xs, list(u.shortcut))
            rs = residual_branch(xr, list(u.residual))
            half = rs.c_logical
            assert sc.c_logical == half
            chans = [(sc, j // 2) if j % 2 == 0 else (rs, j // 2) for j in range(2 * half)]
        else:
            c = len(chans)
            x1, x2 = chans[:c // 2], chans[c // 2:]
            rs = residual_branch(P.gather_quantize(x2, _k32(u.residual[0].Ka)), list(u.residual))
            assert rs.c_logical == c // 2
            chans = [x1[j // 2] if j % 2 == 0 else (rs, j // 2) for j in range(c)]
    c5, b5 = model.conv5[0], model.conv5[1]
    feat_map = P.conv(P.gather_quantize(chans, _k32(c5.Ka)), c5, bn=b5, relu=True, layerout=lo, f16=True)["f16"]
    fcq = P.avgpool_quantize(feat_map, _k32(model.fc.Ka))
    P.output = P.conv(fcq, model.fc, f32=True, linear=True)["f32"].buf.view(batch, -1)
    return P
