#!/usr/bin/env python
"""Timing of the other BASELINE.json configurations through the fused plans (CUDA graph, CUDA events):
VGG-16 CIFAR SLFP-8 batch 512, MobileNetV1 CIFAR SLFP-8 batch 128, MobileNetV1 ImageNet SFP-7 batch 256."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cnns_slfp_quantization_b200 import engine, nets_common as nc, calibration
from cnns_slfp_quantization_b200.nets_imgnet import MobileNetV1_Q as MobileNetImg
from cnns_slfp_quantization_b200.nets_cifar import VGG16_Q, MobileNetV1_Q as MobileNetCifar

dev = torch.device("cuda", 0)
CASES = [("vgg16_cifar_slfp8", lambda q: VGG16_Q(q), 8, 512, 32, engine.compile_vgg16),
         ("mobilenetv1_cifar_slfp8", lambda q: MobileNetCifar(3, q), 8, 128, 32, engine.compile_mobilenetv1),
         ("mobilenetv1_imgnet_sfp7", lambda q: MobileNetImg(3, q), 7, 256, 224, engine.compile_mobilenetv1)]
for name, ctor, qbit, batch, size, comp in CASES:
    m32 = ctor(32).eval()
    sd = nc.synth_state_dict(m32)
    m32.load_state_dict(sd)
    n_layers = len(nc.quantized_layers(m32))
    nc.set_scales(m32, np.ones(n_layers), np.ones(n_layers))
    m32 = m32.to(dev)
    ka, kw = calibration.calibrate_scales(m32, [nc.synth_images(8, size).to(dev)])
    m = ctor(qbit).eval()
    m.load_state_dict(sd)
    nc.set_scales(m, ka, kw)
    m = m.to(dev)
    plan = comp(m, batch, size)
    x = nc.synth_images(batch, size).to(dev)
    plan.input.copy_(x)
    plan.capture()
    for _ in range(3):
        plan()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        plan()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 20
    if os.environ.get("SLFP_BENCH_LAYERS"):
        from cnns_slfp_quantization_b200 import _native as nv
        plan.run(); torch.cuda.synchronize()
        nv.profile = {}
        plan.run(); torch.cuda.synchronize()
        prof, nv.profile = nv.profile, None
        conv = sorted(prof.get("slfp_conv2d_fwd", []) + prof.get("slfp_conv2d_fwd_dual", []), key=lambda t: t[2])
        for (ea, eb, _), (fl, is_dense, desc, _by) in zip(conv, plan.conv_flops):
            print(f"   {desc:34s} {'dense' if is_dense else 'dw   '} {ea.elapsed_time(eb) * 1e3:9.1f} us")
        print("  ", {k: round(sum(x.elapsed_time(y) for x, y, _ in v), 4) for k, v in prof.items()})
    print(json.dumps({"net": name, "batch": batch, "size": size, "ms_per_step": round(ms, 4), "images_per_s": round(batch / ms * 1e3, 1),
                      "launches": plan.launches_per_step, "tflops": round(plan.flops / ms / 1e9, 1)}), flush=True)

# ShuffleNetV2 x1 (BASELINE config 5's second net): no fused plan yet - the module-level drop-in (float32 between layers,
# stock BatchNorm / ReLU / channel shuffle), SFP-7, calibrated scales, eager launches timed with CUDA events.
from cnns_slfp_quantization_b200.nets_cifar import ShuffleNetV2
for batch, size in ((512, 32), (64, 224)):
    m32 = ShuffleNetV2(32).eval()
    sd = nc.synth_state_dict(m32)
    m32.load_state_dict(sd)
    nc.set_scales(m32, np.ones(57), np.ones(57))
    m32 = m32.to(dev)
    ka, kw = calibration.calibrate_scales(m32, [nc.synth_images(8, size).to(dev)])
    m = ShuffleNetV2(7).eval()
    m.load_state_dict(sd)
    nc.set_scales(m, ka, kw)
    m = m.to(dev)
    x = nc.synth_images(batch, size).to(dev).contiguous(memory_format=torch.channels_last)
    with torch.no_grad():
        for _ in range(3):
            m(x)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            m(x)
        b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    print(json.dumps({"net": "shufflenetv2_sfp7_modules", "batch": batch, "size": size, "ms_per_step": round(ms, 4),
                      "images_per_s": round(batch / ms * 1e3, 1)}), flush=True)
