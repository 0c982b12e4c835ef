"""Fused training-mode BatchNorm2d (+ residual add) (+ ReLU) around the quantized convolutions of a QAT step.

The reference's nets run `relu(bn(conv(x)))` and `relu(bn3(conv3(out)) + identity)` as separate PyTorch modules
(nets_imgnet/resnet50.py:71-88); in the SLFP-8 QAT step (cifar100_train_eval.py:170-179) those stock kernels were half
of the step time.  `bn_act(x, bn, relu, residual)` computes the same function of the same nn.BatchNorm2d module - its
weight / bias / running statistics / num_batches_tracked are used and updated in place, so state_dict keys and the
optimizer's parameter list do not change - with two HBM passes forward and two backward (csrc/bn_act.cu) on the
channels-last float32 tensors Conv2d_Q returns.  Anything the kernels do not cover (eval mode, CPU tensors, no affine /
no running statistics, momentum=None, channel counts that are not multiples of 4, SLFP_NO_FUSED_BN=1) takes the stock
modules, so the function is always safe to call.
"""
import ctypes
import os

import torch
import torch.nn.functional as F

from .. import _native as _nv

_workspaces = {}


def _workspace(dev, c):
    """One zero-initialised workspace per (device, stream); the kernels leave it zeroed."""
    key = (dev.index, torch.cuda.current_stream(dev).cuda_stream)
    need = _nv.lib().slfp_bn_act_workspace_floats(c)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < need:
        ws = torch.zeros(max(need, 4 + 4 * 4096), dtype=torch.float32, device=dev)
        _workspaces[key] = ws
    return ws


class _BnAct(torch.autograd.Function):
    """x, residual: [N, H, W, C] float32 contiguous (NHWC).  Returns y of the same shape."""

    @staticmethod
    def forward(ctx, x, weight, bias, residual, running_mean, running_var, eps, momentum, relu, quant=None):
        """quant: None or (fmt, [k_div, ...]) - the kernel also writes the codes of y / k_div for each scale (returned
        through quant[2], a list: code tensors are not autograd outputs)."""
        lib = _nv.lib()
        x = x.contiguous()
        res = residual.contiguous() if residual is not None else None
        c = x.shape[-1]
        m = x.numel() // c
        y = torch.empty_like(x)
        mean = torch.empty(c, dtype=torch.float32, device=x.device)
        invstd = torch.empty(c, dtype=torch.float32, device=x.device)
        ws = _workspace(x.device, c)
        coef = torch.empty(4 * c, dtype=torch.float32, device=x.device)
        if quant is not None and relu:
            fmt, kdivs, out = quant
            codes = [torch.empty(x.shape, dtype=torch.uint8, device=x.device) for _ in kdivs]
            ks = (ctypes.c_float * len(kdivs))(*kdivs)
            ps = (ctypes.c_void_p * len(kdivs))(*[t.data_ptr() for t in codes])
            _nv.check(lib.slfp_bn_act_fwd_train_quant(x.data_ptr(), m, c, weight.data_ptr(), bias.data_ptr(), _nv.ptr(res), eps, momentum,
                                                      _nv.ptr(running_mean), _nv.ptr(running_var), y.data_ptr(), mean.data_ptr(),
                                                      invstd.data_ptr(), ws.data_ptr(), coef.data_ptr(), fmt, len(kdivs), ks, ps,
                                                      _nv.stream()))
            out.extend(codes)
        else:
            _nv.check(lib.slfp_bn_act_fwd_train(x.data_ptr(), m, c, weight.data_ptr(), bias.data_ptr(), _nv.ptr(res), int(relu), eps,
                                                momentum, _nv.ptr(running_mean), _nv.ptr(running_var), y.data_ptr(), mean.data_ptr(),
                                                invstd.data_ptr(), ws.data_ptr(), coef.data_ptr(), _nv.stream()))
        # the ReLU mask of a layer without residual is recomputed from x in the backward: y is not kept alive
        ctx.save_for_backward(x, y if (relu and res is not None) else None, weight, bias, mean, invstd)
        ctx.relu, ctx.has_res = bool(relu), res is not None
        return y

    @staticmethod
    def backward(ctx, gy):
        lib = _nv.lib()
        x, y, weight, bias, mean, invstd = ctx.saved_tensors
        gy = gy.contiguous()
        c = x.shape[-1]
        m = x.numel() // c
        dx = torch.empty_like(x)
        dres = torch.empty_like(x) if (ctx.has_res and ctx.needs_input_grad[3]) else None
        dgamma = torch.empty(c, dtype=torch.float32, device=x.device)
        dbeta = torch.empty(c, dtype=torch.float32, device=x.device)
        ws = _workspace(x.device, c)
        coef = torch.empty(4 * c, dtype=torch.float32, device=x.device)
        amax = torch.empty(1, dtype=torch.float32, device=x.device)           # zeroed by the reduce kernel
        _nv.check(lib.slfp_bn_act_bwd(gy.data_ptr(), x.data_ptr(), _nv.ptr(y), m, c, weight.data_ptr(), bias.data_ptr(), mean.data_ptr(),
                                      invstd.data_ptr(), int(ctx.relu), dx.data_ptr(), _nv.ptr(dres), dgamma.data_ptr(),
                                      dbeta.data_ptr(), ws.data_ptr(), coef.data_ptr(), amax.data_ptr(), _nv.stream()))
        _nv.pending_grad_absmax = (dx.data_ptr(), dx.numel(), amax)
        return dx, dgamma, dbeta, dres, None, None, None, None, None, None


_pending_counters = {}


def flush_batch_counters():
    """nn.BatchNorm2d increments num_batches_tracked once per training forward - 53 one-element launches per ResNet-50
    step when done layer by layer.  bn_act() queues the counters; the net calls this once at the end of its forward (any
    reader of the counters - state_dict(), a checkpoint - after a forward that did not flush should call it first)."""
    global _pending_counters
    if _pending_counters:
        todo, _pending_counters = list(_pending_counters.values()), {}
        torch._foreach_add_([t for t, _ in todo], [k for _, k in todo])


def fused_ok(bn, x):
    return (bn.training and x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 and bn.affine and bn.track_running_stats
            and bn.momentum is not None and bn.num_features % 4 == 0 and bn.weight.dtype == torch.float32
            and not os.environ.get("SLFP_NO_FUSED_BN"))


def _consumer_key(conv):
    """(k_div as float32, activation format) of a Conv2d_Q that can take ready-made codes, else None."""
    import numpy as np
    if getattr(conv, "q_bit", None) not in (7, 8) or getattr(conv, "groups", 1) != 1 or conv.in_channels % 16:
        return None
    return float(np.float32(float(conv.Ka))), _nv.fmt_for(conv.q_bit, "act")


def bn_act(x, bn, relu=True, residual=None, consumers=()):
    """relu?(bn(x) + residual?) for an nn.BatchNorm2d module `bn` and NCHW-shaped tensors (channels-last memory is used
    as is; other layouts are converted).

    consumers: the Conv2d_Q modules that will read the result.  With relu=True the fused kernel also writes their
    activation codes (quantize_act(y / Ka), one code tensor per distinct Ka, at most two) and attaches them to the
    returned tensor (`._slfp_codes`); Conv2d_Q.forward picks them up instead of quantizing y again."""
    if not fused_ok(bn, x):
        out = bn(x)
        if residual is not None:
            out = out + residual
        return F.relu(out) if relu else out
    t = bn.num_batches_tracked                               # += 1, flushed in one launch (flush_batch_counters)
    ent = _pending_counters.get(id(t))
    if ent is None:
        _pending_counters[id(t)] = [t, 1]
    else:
        ent[1] += 1
    if len(_pending_counters) >= 512:
        flush_batch_counters()
    xn = x.permute(0, 2, 3, 1)
    rn = residual.permute(0, 2, 3, 1) if residual is not None else None
    quant, keys = None, []
    if relu and consumers and not os.environ.get("SLFP_NO_FUSED_ACT_QUANT"):
        for cv in consumers:
            k = _consumer_key(cv)
            if k is not None and k not in keys:
                keys.append(k)
        if keys and len(keys) <= 2 and len({f for _, f in keys}) == 1 and x.shape[1] % 16 == 0:
            quant = (keys[0][1], [k for k, _ in keys], [])
    y = _BnAct.apply(xn, bn.weight, bn.bias, rn, bn.running_mean, bn.running_var, float(bn.eps), float(bn.momentum), bool(relu), quant)
    out = y.permute(0, 3, 1, 2)
    if quant is not None and quant[2]:
        out._slfp_codes = {key: t for key, t in zip(keys, quant[2])}
    return out


class _MaxPool3x3s2(torch.autograd.Function):
    """x: [N, H, W, C] float32 contiguous -> [N, Ho, Wo, C]; the window positions of the maxima (one byte each) are kept
    for the gather backward (csrc/bn_act.cu)."""

    @staticmethod
    def forward(ctx, x):
        lib = _nv.lib()
        x = x.contiguous()
        n, h, w, c = x.shape
        ho, wo = (h - 1) // 2 + 1, (w - 1) // 2 + 1
        y = torch.empty((n, ho, wo, c), dtype=torch.float32, device=x.device)
        idx = torch.empty((n, ho, wo, c), dtype=torch.uint8, device=x.device)
        _nv.check(lib.slfp_maxpool3x3s2_fwd_f32(x.data_ptr(), n, h, w, c, y.data_ptr(), idx.data_ptr(), _nv.stream()))
        ctx.save_for_backward(idx)
        ctx.shape = (n, h, w, c)
        return y

    @staticmethod
    def backward(ctx, gy):
        lib = _nv.lib()
        (idx,) = ctx.saved_tensors
        n, h, w, c = ctx.shape
        gy = gy.contiguous()
        dx = torch.empty((n, h, w, c), dtype=torch.float32, device=gy.device)
        _nv.check(lib.slfp_maxpool3x3s2_bwd_f32(gy.data_ptr(), idx.data_ptr(), n, h, w, c, dx.data_ptr(), _nv.stream()))
        _nv.pending_grad_absmax = None
        return dx


def maxpool_train(x, pool):
    """`pool(x)` for an nn.MaxPool2d; a 3x3 / stride 2 / padding 1 pool of a float32 CUDA tensor that needs a gradient runs
    through the pair of kernels above (NCHW shape, channels-last memory), everything else through the stock module."""
    def _pair(v):
        return tuple(v) if isinstance(v, (tuple, list)) else (v, v)
    if (x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 and x.requires_grad and torch.is_grad_enabled() and x.shape[1] % 4 == 0
            and _pair(pool.kernel_size) == (3, 3) and _pair(pool.stride) == (2, 2) and _pair(pool.padding) == (1, 1)
            and _pair(pool.dilation) == (1, 1) and not pool.ceil_mode and not pool.return_indices
            and not os.environ.get("SLFP_NO_FUSED_BN")):
        return _MaxPool3x3s2.apply(x.permute(0, 2, 3, 1)).permute(0, 3, 1, 2)
    return pool(x)
