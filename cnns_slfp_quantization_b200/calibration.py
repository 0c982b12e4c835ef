"""Max-scaling calibration: per-layer Ka = max|input| / 15.5, Kw = max|weight| / 15.5.

Replaces the reference's `get_scale_factor` (cifar100_train_eval.py:213-277, imgnet_train_eval.py:
220-287), which copies every layer input to the host, concatenates them and takes torch.max on the CPU,
with the fused abs-max kernel (csrc/quantize.cu: warp shuffles -> block reduction -> atomicMax) folding
into one device vector, and -- when the batch is sharded over ranks -- ONE allreduce(MAX) over that
vector (max is exact and order independent, so every rank ends with bit-identical scales).

The reference obtains the raw tensors by running the net at Qbits=32 with all K = 1 so that `input_q`
is the un-scaled layer input; here a forward pre-hook reads the layer input directly, which is the same
tensor.
"""
import numpy as np
import torch

from . import _native as nv
from .nets_common import quantized_layers


class ScaleCalibrator:
    def __init__(self, model, divisor=15.5):
        self.model, self.divisor = model, divisor
        self.layers = quantized_layers(model)
        dev = next(model.parameters()).device
        self.act_max = torch.zeros(len(self.layers), dtype=torch.float32, device=dev)
        self.wgt_max = torch.zeros(len(self.layers), dtype=torch.float32, device=dev)
        self._hooks = []

    def _hook(self, idx):
        lib = nv.lib()

        def pre(mod, inp):
            x = inp[0].detach()
            nv.require_cuda(x, "calibration")
            x = nv.dense_flat(x)
            nv.check(lib.slfp_absmax_f32(x.data_ptr(), x.numel(), self.act_max[idx:idx + 1].data_ptr(), 0, nv.stream()))
        return pre

    def __enter__(self):
        self._hooks = [l.register_forward_pre_hook(self._hook(i)) for i, l in enumerate(self.layers)]
        return self

    def __exit__(self, *exc):
        for h in self._hooks:
            h.remove()
        self._hooks = []

    def observe_weights(self):
        lib = nv.lib()
        for i, l in enumerate(self.layers):
            w = nv.dense_flat(l.weight.detach())
            nv.check(lib.slfp_absmax_f32(w.data_ptr(), w.numel(), self.wgt_max[i:i + 1].data_ptr(), 0, nv.stream()))

    def scales(self, group=None):
        """(Ka, Kw) as float64 numpy arrays; with torch.distributed initialised, maxima are all-reduced."""
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            both = torch.cat([self.act_max, self.wgt_max])
            dist.all_reduce(both, op=dist.ReduceOp.MAX, group=group)
            n = len(self.layers)
            self.act_max, self.wgt_max = both[:n].clone(), both[n:].clone()
        ka = self.act_max.double().cpu().numpy() / self.divisor
        kw = self.wgt_max.double().cpu().numpy() / self.divisor
        return ka, kw


def calibrate_scales(model_fp32, batches, divisor=15.5, group=None):
    """Run `model_fp32` (a q_bit = 32 instance on a CUDA device) over the batches and return (Ka, Kw)."""
    cal = ScaleCalibrator(model_fp32, divisor)
    cal.observe_weights()
    # the Qbits = 32 pass runs stock cuDNN / cuBLAS: keep it in true float32 (PyTorch's default would be TF32 for
    # convolutions), so the observed maxima are the reference's CPU float32 maxima up to summation order
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        with cal, torch.no_grad():
            for x in batches:
                model_fp32(x)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    return cal.scales(group)


def quantize_dynamic(x, q_bit, kind="act", divisor=15.5, group=None, want="codes", absmax=None):
    """Dynamic max-scaling quantizer (north_star: abs-max reduction -> max-scaling -> round to nearest), all on the
    device: slfp_absmax_f32 -> [allreduce(MAX) when torch.distributed is initialised, so that every rank of a sharded
    batch uses the bit-identical K] -> slfp_quantize_dyn_f32, which reads K = float32(max / divisor) from device memory.
    Three asynchronous launches on the current stream, no host synchronisation, CUDA-graph capturable.
    Returns (codes | fake-quant float32, K as a 0-dim device tensor).  Reference recipe: K = max|x| / 15.5
    (cifar100_train_eval.py:261-271, nets_cifar/mobilenetv1.py:15)."""
    import torch.distributed as dist
    lib = nv.lib()
    nv.require_cuda(x, "quantize_dynamic")
    xf = nv.dense_flat(x.detach())
    if absmax is None:
        absmax = torch.empty((), dtype=torch.float32, device=x.device)
        nv.check(lib.slfp_absmax_f32(xf.data_ptr(), xf.numel(), absmax.data_ptr(), 1, nv.stream()))
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(absmax, op=dist.ReduceOp.MAX, group=group)
    fmt = nv.fmt_for(q_bit, kind)
    k = torch.empty((), dtype=torch.float32, device=x.device)
    if want == "codes":
        out = torch.empty(xf.shape, dtype=torch.uint8, device=x.device)
        nv.check(lib.slfp_quantize_dyn_f32(xf.data_ptr(), xf.numel(), absmax.data_ptr(), float(divisor), fmt, 0, out.data_ptr(), None, None,
                                           k.data_ptr(), nv.stream()))
    else:
        out = torch.empty_like(xf)
        nv.check(lib.slfp_quantize_dyn_f32(xf.data_ptr(), xf.numel(), absmax.data_ptr(), float(divisor), fmt, 0, None, out.data_ptr(), None,
                                           k.data_ptr(), nv.stream()))
    return out, k
