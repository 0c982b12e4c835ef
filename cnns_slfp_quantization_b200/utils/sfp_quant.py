"""Drop-in for the reference's utils/sfp_quant.py: same callables, backed by the fused sm_100a
quantizer kernel (csrc/quantize.cu) through the C ABI.

Reference: utils/sfp_quant.py:7-54 quantize_weight, :56-103 quantize_act, :105-133
quantize_layerout, :135-175 the nn.Module wrappers.  k = 32 is the identity, k = 7 is SFP<3,3>,
k = 8 is SLFP<3,4>; the layer-out quantizer is SFP<4,4> for any k <= 8.  Backward is the identity
straight-through estimator (`grad_output.clone()`, :50-53).

Like the reference module, this one is meant to be star-imported and therefore leaks `torch`,
`nn`, `F` and `np` (the reference nets rely on that); no `__all__`.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F
import numpy as np

from .. import _native as _nv


def _quantize_fakeq(x, fmt, flags=0):
    """One fused pass: float32 tensor -> fake-quant float32 tensor of the same shape / layout."""
    _nv.require_cuda(x, "SLFP quantizer")
    xs = _nv.dense_flat(x.detach())
    out = torch.empty_like(xs)          # preserves the (dense) memory format
    _nv.check(_nv.lib().slfp_quantize_f32(xs.data_ptr(), xs.numel(), 1.0, fmt, flags, None, out.data_ptr(), None,
                                          _nv.stream()))
    return out


def _make_qfn(k, fmt_of_k, flags=0):
    class qfn(torch.autograd.Function):
        @staticmethod
        def forward(ctx, input):
            if k == 32:
                return input
            fmt = fmt_of_k(k)
            if fmt is None:
                # the reference falls through both branches and hits `return out` unbound (:48)
                raise UnboundLocalError("cannot access local variable 'out' where it is not associated with a value")
            return _quantize_fakeq(input, fmt, flags)

        @staticmethod
        def backward(ctx, grad_output):        # STE: identity
            return grad_output.clone()

    return qfn.apply


def quantize_weight(k):
    """utils/sfp_quant.py:7-54."""
    return _make_qfn(k, lambda kk: {7: _nv.FMT_SFP33, 8: _nv.FMT_SLFP34_WGT}.get(kk))


def quantize_act(k):
    """utils/sfp_quant.py:56-103."""
    return _make_qfn(k, lambda kk: {7: _nv.FMT_SFP33, 8: _nv.FMT_SLFP34_ACT}.get(kk))


# The reference's layer-out quantizer returns NaN for an exact 0 input (Python `^` is XOR on
# sfp_quant.py:122-123, so the low clamp is dead code and 0 * NaN survives).  That behaviour is
# reproduced by default; set LAYEROUT_ZERO_IS_ZERO = True for the evidently intended result (0).
LAYEROUT_ZERO_IS_ZERO = False


def quantize_layerout(k):
    """utils/sfp_quant.py:105-133 (SFP<4,4> for every k <= 8)."""
    class qfn(torch.autograd.Function):
        @staticmethod
        def forward(ctx, input):
            if k == 32:
                return input
            if k <= 8:
                flags = _nv.Q_LAYEROUT_ZERO_IS_ZERO if LAYEROUT_ZERO_IS_ZERO else 0
                return _quantize_fakeq(input, _nv.FMT_SFP44_OUT, flags)
            raise UnboundLocalError("cannot access local variable 'out' where it is not associated with a value")

        @staticmethod
        def backward(ctx, grad_output):
            return grad_output.clone()

    return qfn.apply


class weight_quantize_func(nn.Module):
    """utils/sfp_quant.py:135-147."""

    def __init__(self, q_bit):
        super(weight_quantize_func, self).__init__()
        assert q_bit <= 8 or q_bit == 32
        self.q_bit = q_bit
        self.quantize = quantize_weight(k=q_bit)

    def forward(self, x):
        if self.q_bit == 32:
            weight_q = x
        elif self.q_bit == 8 or self.q_bit == 7:
            weight_q = self.quantize(x)
        return weight_q          # any other q_bit: UnboundLocalError, like the reference (:147)


class act_quantize_func(nn.Module):
    """utils/sfp_quant.py:149-161."""

    def __init__(self, q_bit):
        super(act_quantize_func, self).__init__()
        assert q_bit <= 8 or q_bit == 32
        self.q_bit = q_bit
        self.quantize = quantize_act(k=q_bit)

    def forward(self, x):
        if self.q_bit == 32:
            act_q = x
        elif self.q_bit == 8 or self.q_bit == 7:
            act_q = self.quantize(x)
        return act_q


class layerout_quantize_func(nn.Module):
    """utils/sfp_quant.py:163-175."""

    def __init__(self, q_bit):
        super(layerout_quantize_func, self).__init__()
        assert q_bit <= 8 or q_bit == 32
        self.q_bit = q_bit
        self.quantize = quantize_layerout(k=q_bit)

    def forward(self, x):
        if self.q_bit == 32:
            out_q = x
        elif self.q_bit == 8 or self.q_bit == 7:
            out_q = self.quantize(x)
        return out_q
