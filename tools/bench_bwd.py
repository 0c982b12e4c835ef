#!/usr/bin/env python
"""Per-layer timing of the tensor-core backward (slfp_conv2d_bwd_ws: dgrad + wgrad + operand preparation) on the
ResNet-50 layer shapes, CUDA events, batch 256 by default.
    python tools/bench_bwd.py [batch] [only-substring]"""
import ctypes, json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from cnns_slfp_quantization_b200 import _native as nv
import tools.microbench as mb

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 256
only = sys.argv[2] if len(sys.argv) > 2 else None
lib = nv.lib()
st = nv.stream()
for name, C, K, k, stride, pad, H in mb.RESNET50_LAYERS:
    if C < 16 or (only and only not in name):
        continue
    Cp = (C + 15) // 16 * 16
    Ho = (H + 2 * pad - k) // stride + 1
    afmt, wfmt = nv.fmt_for(8, "act"), nv.fmt_for(8, "weight")
    d = nv.SlfpConvDesc(batch, H, H, C, Cp, K, k, k, stride, stride, pad, pad, 1, 1, 1, afmt)
    x = torch.randn(batch, H, H, C, device="cuda") * 4
    xc = torch.empty((batch, H, H, Cp), dtype=torch.uint8, device="cuda")
    nv.check(lib.slfp_quantize_nhwc_f32(x.data_ptr(), batch * H * H, C, Cp, 1.0, afmt, xc.data_ptr(), st))
    w = torch.randn(K, C, k, k, device="cuda")
    pitch = lib.slfp_conv_wpitch(ctypes.byref(d))
    wc = torch.empty((K * pitch,), dtype=torch.uint8, device="cuda")
    nv.check(lib.slfp_prepare_weights(ctypes.byref(d), w.data_ptr(), *w.stride(), 0.25, wfmt, None, wc.data_ptr(), None, st))
    gy = torch.randn(batch, Ho, Ho, K, device="cuda") * 1e-4
    dx = torch.empty((batch, H, H, C), dtype=torch.float32, device="cuda")
    dw = torch.empty((K, C, k, k), dtype=torch.float32, device="cuda")
    nbytes = lib.slfp_conv2d_bwd_workspace_size(ctypes.byref(d), 1, 1)
    ws = torch.empty((max(nbytes, 1),), dtype=torch.uint8, device="cuda")
    so, sc, sr, ss = dw.stride()
    fn = lambda: nv.check(lib.slfp_conv2d_bwd_ws(ctypes.byref(d), gy.data_ptr(), xc.data_ptr(), wc.data_ptr(), wfmt, 1.0, 0.25,
                                                 dx.data_ptr(), dw.data_ptr(), so, sc, sr, ss, None, ws.data_ptr(), nbytes, st))
    med, best = mb.timeit(fn, iters=10)
    fl = 2 * 2.0 * batch * Ho * Ho * K * C * k * k
    print(json.dumps({"layer": name, "batch": batch, "ms_dgrad_wgrad_prep": round(med, 4), "TFLOPs": round(fl / med / 1e9, 1)}), flush=True)
