"""Drop-in for the reference's utils/optimizer.py: DSGD (:9-73, the "revised SGD"), SSGD (:75-132)
and NormalSGD (:134-190), each step being ONE fused multi-tensor launch per parameter group
(csrc/elementwise.cu: sgd_step_kernel) instead of ~60 elementwise kernels per parameter.

Semantics kept from the reference, quirks included:
  * weight decay is added into p.grad in place (:46);
  * momentum buffer starts as a copy of the (decayed) gradient (:50);
  * DSGD quantizes the RAW parameter (no division by Kw) before and after the SGD step and gives
    elements whose quantized value moved by less than 1e-4 two extra steps (:58-64);
  * SSGD's extra step is scaled by |p| + 1 (:130-131);
  * every parameter with a gradient is touched, BatchNorm / bias included.
"""
import ctypes as _ctypes

import torch
import torch.optim as optim
from torch.optim.optimizer import required
from torch.optim import Optimizer
from .sfp_quant import *

from .. import _native as _nv


def _check_hparams(lr, momentum, dampening, weight_decay, nesterov):
    if lr is not required and lr < 0.0:
        raise ValueError("Invalid learning rate: {}".format(lr))
    if momentum < 0.0:
        raise ValueError("Invalid momentum value: {}".format(momentum))
    if weight_decay < 0.0:
        raise ValueError("Invalid weight_decay value: {}".format(weight_decay))
    if nesterov and (momentum <= 0 or dampening != 0):
        raise ValueError("Nesterov momentum requires a momentum and zero dampening")


class _FusedSGDBase(Optimizer):
    _mode = _nv.SGD_NORMAL

    def __init__(self, params, qbit, lr, momentum, dampening, weight_decay, nesterov):
        _check_hparams(lr, momentum, dampening, weight_decay, nesterov)
        defaults = dict(lr=lr, momentum=momentum, dampening=dampening,
                        weight_decay=weight_decay, nesterov=nesterov)
        super().__init__(params, defaults)
        self._qbit = qbit
        if qbit is not None:
            self.quantize_fn = weight_quantize_func(q_bit=qbit)

    def __setstate__(self, state):
        super().__setstate__(state)
        for group in self.param_groups:
            group.setdefault('nesterov', False)

    def _qfmt(self):
        if self._qbit is None or self._qbit == 32:
            return -1
        if self._qbit in (7, 8):
            return _nv.fmt_for(self._qbit, "weight")
        # the reference's weight_quantize_func.forward raises for any other q_bit (sfp_quant.py:147)
        raise UnboundLocalError("cannot access local variable 'weight_q' where it is not associated with a value")

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _nv.lib()
        for group in self.param_groups:
            momentum = group['momentum']
            # tensors are bucketed by "has a momentum buffer yet": the first step copies the gradient
            fresh, warm = [], []
            for p in group['params']:
                if p.grad is None:
                    continue
                _nv.require_cuda(p, type(self).__name__)
                if _nv.dense_flat(p) is not p:
                    raise RuntimeError("fused SGD needs dense (contiguous or channels-last) parameters")
                g = p.grad
                if g.stride() != p.stride():
                    g = p.grad = _restride(p, g)
                buf = None
                if momentum != 0:
                    st = self.state[p]
                    if 'momentum_buffer' not in st:
                        buf = st['momentum_buffer'] = torch.empty_like(p, memory_format=torch.preserve_format)
                        fresh.append((p, g, buf))
                        continue
                    buf = st['momentum_buffer']
                warm.append((p, g, buf))
            for first, items in ((1, fresh), (0, warm)):
                if not items:
                    continue
                n = len(items)
                P = (_ctypes.c_void_p * n)(*[p.data_ptr() for p, _, _ in items])
                G = (_ctypes.c_void_p * n)(*[g.data_ptr() for _, g, _ in items])
                B = (_ctypes.c_void_p * n)(*[(b.data_ptr() if b is not None else None) for _, _, b in items])
                S = (_ctypes.c_size_t * n)(*[p.numel() for p, _, _ in items])
                _nv.check(lib.slfp_sgd_step(n, P, G, B if momentum != 0 else None, S, self._mode, self._qfmt(),
                                            group['lr'], momentum, group['dampening'], group['weight_decay'],
                                            1 if group['nesterov'] else 0, first, _nv.stream()))
        return loss


def _restride(p, g):
    out = torch.empty_like(p, memory_format=torch.preserve_format)
    out.copy_(g)
    return out


class DSGD(_FusedSGDBase):
    """utils/optimizer.py:9-73."""
    _mode = _nv.SGD_DSGD

    def __init__(self, params, qbit, lr=required, momentum=0, dampening=0, weight_decay=0, nesterov=False):
        super().__init__(params, qbit, lr, momentum, dampening, weight_decay, nesterov)


class SSGD(_FusedSGDBase):
    """utils/optimizer.py:75-132."""
    _mode = _nv.SGD_SSGD

    def __init__(self, params, qbit, lr=required, momentum=0, dampening=0, weight_decay=0, nesterov=False):
        super().__init__(params, qbit, lr, momentum, dampening, weight_decay, nesterov)


class NormalSGD(_FusedSGDBase):
    """utils/optimizer.py:134-190."""
    _mode = _nv.SGD_NORMAL

    def __init__(self, params, lr=required, momentum=0, dampening=0, weight_decay=0, nesterov=False):
        super().__init__(params, None, lr, momentum, dampening, weight_decay, nesterov)
