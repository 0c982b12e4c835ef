#!/usr/bin/env python
"""Verify every fused conv op of a compiled Plan in isolation: recompute its output with torch (float64)
from the op's OWN input buffers and compare with what the kernel wrote."""
import os, sys
import numpy as np, torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cnns_slfp_quantization_b200 import nets_common as nc, engine, _native as nv
from tools.netcheck import prepare

name = sys.argv[1] if len(sys.argv) > 1 else "vgg16"
m, comp, batch, size = prepare(name)
x = nc.synth_images(batch, size).cuda()
recs = []
orig_conv = engine.Plan.conv
def conv_spy(self, xt, mod, bn=None, relu=False, residual=None, codes=(), f16=False, f32=False, linear=False, relu_codes=True):
    out = orig_conv(self, xt, mod, bn=bn, relu=relu, residual=residual, codes=codes, f16=f16, f32=f32, linear=linear,
                    relu_codes=relu_codes)
    recs.append((xt, mod, bn, relu, residual, out, linear))
    return out
engine.Plan.conv = conv_spy
plan = comp(m, batch, size)
ye = plan(x)
torch.cuda.synchronize()
lib = nv.lib()

def deq(t, fmt):
    out = torch.empty(t.buf.shape, dtype=torch.float32, device="cuda")
    nv.check(lib.slfp_dequantize(t.buf.data_ptr(), t.buf.numel(), fmt, out.data_ptr(), nv.stream()))
    return out[..., :t.c]

from oracle import slfp_oracle as orc
for i, (xt, mod, bn, relu, residual, out, linear) in enumerate(recs):
    xq = deq(xt, xt.fmt).double().permute(0, 3, 1, 2)
    ka, kw = engine._k32(mod.Ka), engine._k32(mod.Kw)
    w = mod.weight.detach().float().cpu().numpy()
    _, wq = orc.quantize(w, 2 if plan.q_bit == 8 else 0, kdiv=kw, want_codes=False)
    wq = torch.from_numpy(wq).cuda().double()
    if linear:
        wq = wq.view(wq.shape[0], wq.shape[1], 1, 1)
        bq = (mod.bias.detach() / mod.Kw / mod.Ka).double()
        y = F.conv2d(xq, wq, bq) * kw * ka
    else:
        bq = None
        if mod.bias is not None:
            bq = (mod.bias.detach() / mod.Ka / mod.Kw).double()
        y = F.conv2d(xq, wq, bq, mod.stride, mod.padding, mod.dilation, mod.groups) * ka * kw
    if bn is not None:
        sc, sh = engine.fold_bn(bn)
        y = y * sc.double().view(1, -1, 1, 1) + sh.double().view(1, -1, 1, 1)
    if residual is not None:
        y = y + residual.buf.double().permute(0, 3, 1, 2)
    if relu:
        y = y.clamp_min(0)
    y = y.permute(0, 2, 3, 1)
    msgs = []
    scale = float(y.abs().max()) + 1e-12
    for kind in ("f16", "f32"):
        if out[kind] is not None:
            d = (out[kind].buf.double() - y).abs().max().item()
            msgs.append(f"{kind} max|d|/max|y| {d / scale:.2e}")
    for kd, t in out["codes"].items():
        got = deq(t, t.fmt)
        _, want = orc.quantize(y.float().cpu().numpy(), 1 if plan.q_bit == 8 else 0, kdiv=kd, want_codes=False)
        want = torch.from_numpy(want).cuda()
        mism = (got.half() != want.half()).float().mean().item()      # compared as tensor-core operands (float16)
        big = ((got - want).abs() > 0.1 * want.abs() + 1e-6).float().mean().item()
        msgs.append(f"codes(k={kd:.4f}) mismatch {mism:.4f} gross {big:.5f}")
    print(f"op {i:2d} {'lin' if linear else 'conv'} in{tuple(xt.buf.shape)} K={mod.weight.shape[0]} g={getattr(mod,'groups',1)} res={residual is not None}: " + "; ".join(msgs))
print("final logits", ye[0, :5].tolist())
