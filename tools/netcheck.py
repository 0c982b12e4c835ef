#!/usr/bin/env python
"""Whole-net parity diagnostic on the GPU: module-level drop-in path and fused engine against the
reference-generated fixture tests/golden/net_cases.npz (logits computed by the REFERENCE nets on CPU)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cnns_slfp_quantization_b200 import nets_common as nc, engine   # noqa: E402
from cnns_slfp_quantization_b200.nets_imgnet import ResNet50, MobileNetV1_Q as MobileNetImg   # noqa: E402
from cnns_slfp_quantization_b200.nets_cifar import VGG16_Q, MobileNetV1_Q as MobileNetCifar, ShuffleNetV2     # noqa: E402

G = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "net_cases.npz"))


def build(name, qbit):
    if name == "resnet50":
        return ResNet50(qbit), lambda m, b, s: engine.compile_resnet50(m, b, s)
    if name == "vgg16":
        return VGG16_Q(qbit), lambda m, b, s: engine.compile_vgg16(m, b, s)
    if name == "mobilenetv1_cifar":
        return MobileNetCifar(3, qbit), lambda m, b, s: engine.compile_mobilenetv1(m, b, s)
    if name == "shufflenetv2":
        return ShuffleNetV2(qbit), None            # module-level drop-in only (no fused plan yet)
    return MobileNetImg(3, qbit), lambda m, b, s: engine.compile_mobilenetv1(m, b, s)


def prepare(name):
    qbit, batch, size = [int(v) for v in G[f"{name}.cfg"]]
    m, comp = build(name, qbit)
    m.load_state_dict(nc.synth_state_dict(m))
    nc.apply_classifier(m, float(G[f"{name}.fc_scale"]), G[f"{name}.fc_bias"])
    nc.set_scales(m, G[f"{name}.ka"], G[f"{name}.kw"])
    return m.cuda().eval(), comp, batch, size


if __name__ == "__main__":
    for name in ("resnet50", "vgg16", "mobilenetv1_cifar", "mobilenetv1_imgnet"):
        m, comp, batch, size = prepare(name)
        x = nc.synth_images(batch, size).cuda()
        ref = G[f"{name}.logits"]
        with torch.no_grad():
            ym = m(x.contiguous(memory_format=torch.channels_last)).float().cpu().numpy()
        plan = comp(m, batch, size)
        ye = plan(x).float().cpu().numpy().copy()
        plan.capture()
        yg = plan(x).float().cpu().numpy()
        srt = np.sort(ref, 1)
        print(name, "ref top1", ref.argmax(1).tolist())
        print("   modules top1", ym.argmax(1).tolist(), "max|d|", float(np.abs(ym - ref).max()), "rms", float(np.sqrt(((ym - ref) ** 2).mean())))
        print("   engine  top1", ye.argmax(1).tolist(), "max|d|", float(np.abs(ye - ref).max()), "rms", float(np.sqrt(((ye - ref) ** 2).mean())),
              "graph==eager", bool((ye == yg).all()))
        print("   margins", (srt[:, -1] - srt[:, -2]).round(3).tolist(), "launches/step", plan.launches_per_step)
