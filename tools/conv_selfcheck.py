#!/usr/bin/env python
"""Consistency experiments for the conv kernel's output modes on the GPU (no oracle needed):
codes-only vs codes+f32 runs must agree bitwise, and repeated runs must be deterministic."""
import os, sys, ctypes
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from gpu_util import conv_fwd_gpu

rng = np.random.default_rng(0)
SH = [(8, 512, 2, 2, 512, 3, 1, 1), (8, 512, 4, 4, 512, 3, 1, 1), (8, 256, 8, 8, 256, 3, 1, 1), (8, 64, 16, 16, 256, 1, 1, 0),
      (8, 512, 1, 1, 512, 1, 1, 0), (8, 2048, 1, 1, 1000, 1, 1, 0), (8, 256, 16, 16, 64, 1, 1, 0)]
for (N, C, H, W, K, k, st, pad) in SH:
    x = (rng.standard_normal((N, C, H, W)) * 2).astype(np.float32)
    w = (rng.standard_normal((K, C, k, k)) * 0.1).astype(np.float32)
    ka, kw = float(np.abs(x).max() / 15.5), float(np.abs(w).max() / 15.5)
    sc = rng.uniform(0.5, 1.5, K).astype(np.float32); sh = (rng.standard_normal(K) * 0.3).astype(np.float32)
    a = conv_fwd_gpu(x, w, None, ka, kw, 8, st, pad, 1, 1, dict(ch_scale=sc, ch_shift=sh, relu=True, next_k=0.2))
    b = conv_fwd_gpu(x, w, None, ka, kw, 8, st, pad, 1, 1, dict(ch_scale=sc, ch_shift=sh, relu=True, next_k=0.2))
    # reference: torch fp64 conv on fp16-rounded operands is not needed here; check self-consistency + f32->codes
    import oracle.slfp_oracle as orc
    oc, _ = orc.quantize(a["y"].transpose(0, 2, 3, 1), 1, kdiv=0.2)
    print((N, C, H, W, K, k), "deterministic", bool((a["y_codes"] == b["y_codes"]).all() and (a["y"] == b["y"]).all()),
          "codes==quant(f32)", bool((a["y_codes"][..., :K] == oc).all()), "finite", bool(np.isfinite(a["y"]).all()),
          "absmax", float(np.abs(a["y"]).max()))
