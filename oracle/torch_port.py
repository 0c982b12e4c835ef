"""Torch-CPU port of the reference's hot path -- TEST INFRASTRUCTURE / CPU BASELINE ONLY.

The reference is pure Python on ATen CPU kernels and cannot travel to the GPU box, so its
CPU implementation is restated here with the SAME ATen operator sequence per quantizer (abs, log2,
floor, pow, div, round, masked writes, ... -- which is what makes its CPU cost what it is, about 25
elementwise passes per call), organised as one parametrised routine instead
of five copies.  bench.py times this as `cpu_baseline` / `--impl reference` (kind: "port"); the
tests check it bit-exactly against the reference-generated fixtures in tests/golden/.

Reference anchors: utils/sfp_quant.py:14-47, 63-96, 111-126 (quantizers), :50-53 (identity STE),
utils/conv2d_func.py:8-66 (module factories), utils/activation_func.py:30-36.
Only tests/, __graft_entry__.smoke() and bench.py may import this module.
"""
import types

import torch
import torch.nn as nn
import torch.nn.functional as F


def _grid_round(x, n_levels, log_domain, pre_round, lo, lo_to, hi, hi_to, hi_inclusive):
    """sign(x) * round-to-grid(|x|) with the reference's clamp writes, in the reference's op order."""
    sign = torch.sign(x)
    a = torch.abs(x)
    e = torch.floor(torch.log2(a))
    scale = pow(2, e)
    m = a / scale
    if log_domain:
        if pre_round:
            m = torch.round(m * n_levels) / n_levels            # linear pre-round (sfp_quant.py:88)
        lm = torch.round(torch.log2(m) * n_levels) / n_levels    # log converter (:40 / :89)
        out = pow(2, (e + lm))
    else:
        out = torch.mul(torch.round(m * n_levels) / n_levels, scale)
    if lo is not None:
        out[a < lo[0]] = lo_to[0]
        out[(a >= lo[0]) & (a < lo[1])] = lo_to[1]
    out[(a >= hi) if hi_inclusive else (a > hi)] = hi_to
    return torch.mul(sign, out)


def _quantizer(kind, k):
    class qfn(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x):
            if k == 32:
                return x
            if kind == "layerout":
                # SFP<4,4>; the reference's low clamp is dead code ('^' is XOR, :122-123) -> none here
                return _grid_round(x, 16, False, False, None, None, 248, 248, True)
            if k == 7:                                            # SFP<3,3>
                return _grid_round(x, 8, False, False, (0.0625, 0.125), (1e-10, 0.125), 15, 15, True)
            if k == 8:                                            # SLFP<3,4>
                return _grid_round(x, 16, True, kind == "act", (0.0625, 0.125), (1e-10, 0.125), 15.32165, 15.32165, False)
            raise UnboundLocalError("out")

        @staticmethod
        def backward(ctx, g):
            return g.clone()
    return qfn.apply


def quantize_weight(k):
    return _quantizer("weight", k)


def quantize_act(k):
    return _quantizer("act", k)


def quantize_layerout(k):
    return _quantizer("layerout", k)


class _QModule(nn.Module):
    _kind = "act"

    def __init__(self, q_bit):
        super().__init__()
        assert q_bit <= 8 or q_bit == 32
        self.q_bit = q_bit
        self.quantize = _quantizer(self._kind, q_bit)

    def forward(self, x):
        if self.q_bit == 32:
            return x
        if self.q_bit in (7, 8):
            return self.quantize(x)
        raise UnboundLocalError("q")


class weight_quantize_func(_QModule):
    _kind = "weight"


class act_quantize_func(_QModule):
    _kind = "act"


class layerout_quantize_func(_QModule):
    _kind = "layerout"


def _conv_factory(q_bit, Kw, Ka, bias_default, scale_bias):
    class Conv2d_Q(nn.Conv2d):
        def __init__(self, in_channels, out_channels, kernel_size, Kw=Kw, Ka=Ka, stride=1, padding=0, dilation=1,
                     groups=1, bias=bias_default):
            super().__init__(in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias)
            self.q_bit = q_bit
            self.quantize_weight = weight_quantize_func(q_bit)
            self.quantize_act = act_quantize_func(q_bit)
            self.Kw, self.Ka = torch.tensor(Kw), torch.tensor(Ka)

        def forward(self, input, order=None):
            self.input_q = self.quantize_act(input / self.Ka)
            self.weight_q = self.quantize_weight(self.weight / self.Kw)
            b = self.bias
            if scale_bias:
                b = self.bias_q = self.bias / self.Ka / self.Kw
            self.output = F.conv2d(self.input_q, self.weight_q, b, self.stride, self.padding, self.dilation,
                                   self.groups) * self.Ka * self.Kw
            return self.output
    return Conv2d_Q


def conv2d_Q(q_bit, Kw, Ka):
    return _conv_factory(q_bit, Kw, Ka, False, False)


def conv2d_Q_bias(q_bit, Kw, Ka):
    return _conv_factory(q_bit, Kw, Ka, True, True)


def linear_Q(q_bit, Kw, Ka):
    class Linear_Q(nn.Linear):
        def __init__(self, in_features, out_features, Kw=Kw, Ka=Ka, bias=True):
            super().__init__(in_features, out_features, bias)
            self.q_bit = q_bit
            self.quantize_weight = weight_quantize_func(q_bit)
            self.quantize_act = act_quantize_func(q_bit)
            self.Kw, self.Ka = torch.tensor(Kw), torch.tensor(Ka)

        def forward(self, input):
            self.input_q = self.quantize_act(input / self.Ka)
            self.weight_q = self.quantize_weight(self.weight / self.Kw)
            self.bias_q = self.bias / self.Kw / self.Ka
            return F.linear(self.input_q, self.weight_q, self.bias_q) * self.Kw * self.Ka
    return Linear_Q


class Swish(nn.Module):
    def forward(self, x):
        return x * torch.sigmoid(x)


def ops():
    """The factory namespace the caller nets accept (cnns_slfp_quantization_b200.nets_common)."""
    return types.SimpleNamespace(conv2d_Q=conv2d_Q, conv2d_Q_bias=conv2d_Q_bias, linear_Q=linear_Q,
                                 layerout_quantize_func=layerout_quantize_func, Swish=Swish)
