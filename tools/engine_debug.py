#!/usr/bin/env python
"""Layer-by-layer comparison of the fused engine's code tensors with the module path's input_q taps."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cnns_slfp_quantization_b200 import nets_common as nc, engine, _native as nv
from tools.netcheck import prepare

name = sys.argv[1] if len(sys.argv) > 1 else "vgg16"
m, comp, batch, size = prepare(name)
x = nc.synth_images(batch, size).cuda()
with torch.no_grad():
    ym = m(x)
layers = nc.quantized_layers(m)
taps = [l.input_q.detach() for l in layers]
plan = comp(m, batch, size)
# instrument: record conv op inputs
recs = []
orig_conv = engine.Plan.conv
plan2 = None
def conv_spy(self, xt, mod, *a, **k):
    recs.append((mod, xt))
    return orig_conv(self, xt, mod, *a, **k)
engine.Plan.conv = conv_spy
plan = comp(m, batch, size)
ye = plan(x)
torch.cuda.synchronize()
lib = nv.lib()
for mod, xt in recs:
    i = layers.index(mod)
    out = torch.empty(xt.buf.shape, dtype=torch.float32, device="cuda")
    nv.check(lib.slfp_dequantize(xt.buf.data_ptr(), xt.buf.numel(), plan.afmt, out.data_ptr(), nv.stream()))
    got = out[..., :xt.c]
    want = taps[i]
    if want.dim() == 4:
        want = want.permute(0, 2, 3, 1)
    want = want.reshape(got.shape)
    diff = (got != want)
    nan = torch.isnan(got).sum().item()
    print(f"layer {i:2d} {type(mod).__name__} in {tuple(got.shape)} mismatching codes {diff.float().mean().item():.4f} nan {nan} "
          f"max|d| {(got - want).abs().max().item():.4g}")
print("logits max diff", (ye - ym).abs().max().item())
