#!/usr/bin/env python
"""Kernel micro-benchmarks on one B200 (CUDA events, warm-up, inputs larger than L2 or L2 flushed).

    python tools/microbench.py quant          # fused quantizer GB/s at the BASELINE sizes
    python tools/microbench.py conv           # per-layer implicit-GEMM TFLOP/s on ResNet-50 shapes (batch 256)
"""
import ctypes
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cnns_slfp_quantization_b200 import _native as nv   # noqa: E402

PEAKS = {"hbm_gbs": 6536.7, "bf16_tflops": 1616.7}
try:
    PEAKS.update(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))))
except Exception:
    pass


def timeit(fn, iters=20, warm=3, flush=None):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def bench_quant(kdiv=1.0):
    """x = 4 randn(n) with K = 1 (about 1 % flushed, 0.01 % clamped, every octave hit)."""
    lib = nv.lib()
    for n in (1 << 24, 205520896):
        x = torch.randn(n, device="cuda") * 4
        codes = torch.empty(n, dtype=torch.uint8, device="cuda")
        fq = torch.empty(n, dtype=torch.float32, device="cuda")
        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda") if n < (1 << 26) else None
        for label, c, f, bpe in (("codes", codes, None, 5), ("fakeq", None, fq, 8)):
            for fmt, fname in ((1, "slfp34_act"), (0, "sfp33"), (2, "slfp34_wgt")):
                fn = lambda: nv.check(lib.slfp_quantize_f32(x.data_ptr(), n, kdiv, fmt, 0, nv.ptr(c), nv.ptr(f), None, nv.stream()))
                med, best = timeit(fn, flush=flush)
                gbs = n * bpe / med / 1e6
                print(json.dumps({"kernel": "quantize", "fmt": fname, "mode": label, "n": n, "ms": round(med, 4),
                                  "GBps": round(gbs, 1), "frac_of_measured_hbm": round(gbs / PEAKS["hbm_gbs"], 3)}), flush=True)
        out = torch.zeros(1, device="cuda")
        med, _ = timeit(lambda: nv.check(lib.slfp_absmax_f32(x.data_ptr(), n, out.data_ptr(), 1, nv.stream())), flush=flush)
        print(json.dumps({"kernel": "absmax", "n": n, "ms": round(med, 4), "GBps": round(n * 4 / med / 1e6, 1)}), flush=True)


RESNET50_LAYERS = [
    # name, Cin, Cout, k, stride, pad, Hin
    ("stem7x7", 3, 64, 7, 2, 3, 224), ("s2dstem 16-64 4x4", 16, 64, 4, 1, 2, 112), ("l1.c1 64-64 1x1", 64, 64, 1, 1, 0, 56), ("l1.c2 64-64 3x3", 64, 64, 3, 1, 1, 56),
    ("l1.c3 64-256 1x1", 64, 256, 1, 1, 0, 56), ("l1.c1b 256-64 1x1", 256, 64, 1, 1, 0, 56),
    ("l2.c1 256-128 1x1", 256, 128, 1, 1, 0, 56), ("l2.c2 128-128 3x3 s2", 128, 128, 3, 2, 1, 56),
    ("l2.c3 128-512 1x1", 128, 512, 1, 1, 0, 28), ("l2.ds 256-512 1x1 s2", 256, 512, 1, 2, 0, 56),
    ("l2.c1b 512-128 1x1", 512, 128, 1, 1, 0, 28), ("l2.c2b 128-128 3x3", 128, 128, 3, 1, 1, 28),
    ("l3.c2b 256-256 3x3", 256, 256, 3, 1, 1, 14), ("l3.c3 256-1024 1x1", 256, 1024, 1, 1, 0, 14),
    ("l3.c1b 1024-256 1x1", 1024, 256, 1, 1, 0, 14), ("l4.c2b 512-512 3x3", 512, 512, 3, 1, 1, 7),
    ("l4.c3 512-2048 1x1", 512, 2048, 1, 1, 0, 7), ("l4.c1b 2048-512 1x1", 2048, 512, 1, 1, 0, 7),
]


def bench_conv(batch=256, mode="f32", only=None, iters=10):
    """mode: f32 | codes | rcodes | res; a trailing "16" (rcodes16, res16) feeds the layer with SLFP_FMT_F16Q float16 images;
    "rq16" = rcodes with store_f16 (the producer side of that format)."""
    lib = nv.lib()
    a16 = mode.endswith("16") and mode != "rq16"
    store16 = mode == "rq16"
    mode = "rcodes" if store16 else (mode[:-2] if a16 else mode)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for name, C, K, k, st, pad, H in RESNET50_LAYERS:
        if only and only not in name:
            continue
        Cp = 4 if C <= 4 else (C + 15) // 16 * 16
        if (a16 and Cp % 64) or (store16 and (Cp % 16 or K % 64)):
            continue
        d = nv.SlfpConvDesc(batch, H, H, C, Cp, K, k, k, st, st, pad, pad, 1, 1, 1, nv.FMT_F16Q if a16 else nv.FMT_SLFP34_ACT)
        Ho = (H + 2 * pad - (k - 1) - 1) // st + 1
        xc = torch.randint(16, 128, (batch, H, H, Cp), dtype=torch.uint8, device="cuda")
        if a16:
            xc = (xc.to(torch.float16) / 64).contiguous()
        pitch = lib.slfp_conv_wpitch(ctypes.byref(d))
        w = torch.randn(K, C, k, k, device="cuda") * 0.1
        wh = torch.empty(K * pitch, dtype=torch.float16, device="cuda")
        so, sc, sr, ss = w.stride()
        nv.check(lib.slfp_prepare_weights(ctypes.byref(d), w.data_ptr(), so, sc, sr, ss, 0.02, nv.FMT_SLFP34_WGT,
                                          wh.data_ptr(), None, None, nv.stream()))
        epi = nv.SlfpEpilogue()
        epi.post_a, epi.post_b = 0.3, 0.02
        if mode == "f32":
            y = torch.empty((batch, Ho, Ho, K), dtype=torch.float32, device="cuda")
            epi.y_f32 = y.data_ptr()
            out_b = 4
        elif mode == "res":                 # bottleneck tail: + float16 residual, ReLU, float16 out, post-ReLU codes
            y = torch.empty((batch, Ho, Ho, K), dtype=torch.uint8, device="cuda")
            y16 = torch.empty((batch, Ho, Ho, K), dtype=torch.float16, device="cuda")
            res = torch.randn((batch, Ho, Ho, K), dtype=torch.float16, device="cuda")
            mul = torch.full((K,), 0.006, device="cuda"); add = torch.zeros(K, device="cuda")
            epi.ch_mul, epi.ch_add = mul.data_ptr(), add.data_ptr()
            epi.residual, epi.residual_f16, epi.y_f16 = res.data_ptr(), 1, y16.data_ptr()
            epi.y_codes, epi.next_k_div, epi.next_fmt, epi.k_phys_out, epi.relu = y.data_ptr(), 0.2, 4, K, 1
            out_b = 5
        else:
            y = torch.empty((batch, Ho, Ho, K), dtype=torch.float16 if store16 else torch.uint8, device="cuda")
            epi.store_f16 = 1 if store16 else 0
            nfmt = 4 if (mode == "rcodes" and Cp >= 16) else nv.FMT_SLFP34_ACT
            epi.y_codes, epi.next_k_div, epi.next_fmt, epi.k_phys_out, epi.relu = y.data_ptr(), 0.2, nfmt, K, 1
            if mode == "rcodes":            # the fused pipeline's form: folded per-channel affine, post-ReLU codes
                mul = torch.full((K,), 0.006, device="cuda"); add = torch.zeros(K, device="cuda")
                epi.ch_mul, epi.ch_add = mul.data_ptr(), add.data_ptr()
            out_b = 2 if store16 else 1
        fn = lambda: nv.check(lib.slfp_conv2d_fwd(ctypes.byref(d), xc.data_ptr(), wh.data_ptr(), ctypes.byref(epi), nv.stream()))
        med, best = timeit(fn, iters=iters, flush=flush)
        flops = 2.0 * batch * Ho * Ho * K * C * k * k
        byts = xc.numel() * xc.element_size() + y.numel() * out_b + wh.numel() * 2
        print(json.dumps({"kernel": "conv_igemm", "layer": name, "out": mode + ("16" if a16 else "") + ("+store_f16" if store16 else ""), "ms": round(med, 4),
                          "TFLOPs": round(flops / med / 1e9, 1), "frac_tensor": round(flops / med / 1e9 / PEAKS["bf16_tflops"], 3),
                          "GBps": round(byts / med / 1e6, 1), "frac_hbm": round(byts / med / 1e6 / PEAKS["hbm_gbs"], 3)}), flush=True)


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "quant"
    if what == "quant":
        bench_quant()
    elif what == "conv":
        bench_conv(mode=sys.argv[2] if len(sys.argv) > 2 else "f32", only=sys.argv[3] if len(sys.argv) > 3 else None,
                   iters=int(sys.argv[4]) if len(sys.argv) > 4 else 10)
    elif what == "quant1":
        lib = nv.lib()
        n = 205520896
        x = torch.randn(n, device="cuda") * 4
        codes = torch.empty(n, dtype=torch.uint8, device="cuda")
        for _ in range(3):
            nv.check(lib.slfp_quantize_f32(x.data_ptr(), n, 0.7, 1, 0, codes.data_ptr(), None, None, nv.stream()))
        torch.cuda.synchronize()
