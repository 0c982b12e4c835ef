"""Drop-in for the reference's utils/activation_func.py (STLFunction / STL :6-28, Swish :30-32,
Sigmoid :34-36) on the fused elementwise kernels of csrc/elementwise.cu."""
import torch
import torch.nn as nn
import torch.nn.functional as F
import numpy as np

from .. import _native as _nv


def _act_fwd(x, kind):
    _nv.require_cuda(x, "SLFP activation")
    xs = _nv.dense_flat(x.detach())
    y = torch.empty_like(xs)
    _nv.check(_nv.lib().slfp_act_fwd(xs.data_ptr(), xs.numel(), kind, y.data_ptr(), _nv.stream()))
    return xs, y


def _act_bwd(xs, gy, kind):
    gy = gy.contiguous(memory_format=torch.channels_last) if (xs.dim() == 4 and not xs.is_contiguous()) else gy.contiguous()
    gx = torch.empty_like(xs)
    _nv.check(_nv.lib().slfp_act_bwd(_nv.ptr(xs), gy.data_ptr(), gy.numel(), kind, gx.data_ptr(), _nv.stream()))
    return gx


def STLFunction():
    """activation_func.py:6-19: y = x if |x| <= 1 else sign(x)(ln|x| + 1); backward clips the
    incoming gradient to [-1, 1] by its own magnitude (:16)."""
    class stl(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x):
            xs, y = _act_fwd(x, _nv.ACT_STL)
            ctx.save_for_backward(xs)
            return y

        @staticmethod
        def backward(ctx, grad_output):
            (xs,) = ctx.saved_tensors
            return _act_bwd(xs, grad_output, _nv.ACT_STL)

    return stl.apply


class _PointwiseAct(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, kind):
        xs, y = _act_fwd(x, kind)
        ctx.save_for_backward(xs)
        ctx.kind = kind
        return y

    @staticmethod
    def backward(ctx, grad_output):
        (xs,) = ctx.saved_tensors
        return _act_bwd(xs, grad_output, ctx.kind), None


class STL(nn.Module):
    def __init__(self):
        super(STL, self).__init__()
        self.stl = STLFunction()

    def forward(self, x):
        stlout = self.stl(x)
        return stlout


class Swish(nn.Module):
    def forward(self, x):
        return _PointwiseAct.apply(x, _nv.ACT_SWISH)


class Sigmoid(nn.Module):
    def forward(self, x):
        return _PointwiseAct.apply(x, _nv.ACT_SIGMOID)
