#!/usr/bin/env python
"""Whole-net golden fixtures from the REFERENCE nets (build container only; needs /root/reference).

For each net: the reference class (nets_imgnet/resnet50.py, nets_cifar/vgg16.py, ... on the
reference's own utils) is built, loaded with nets_common.synth_state_dict (parameters keyed by
NAME, so this repo's table-driven re-statements load identical values), calibrated the way the
reference does (Qbits=32, K=1 forward; scale = max|.|/15.5 -- cifar100_train_eval.py:213-277), and
run at Qbits 8 / 7 on nets_common.synth_images.  Stored: the calibrated scales, logits and top-1.
"""
import os
import sys
import types

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
sys.modules.setdefault("torchsummary", types.SimpleNamespace(summary=lambda *a, **k: None))

import numpy as np
import torch

from cnns_slfp_quantization_b200 import nets_common as nc

torch.set_num_threads(8)


def ref_net(name, qbit):
    if name == "resnet50":
        from nets_imgnet.resnet50 import ResNet50
        return ResNet50(qbit=qbit)
    if name == "vgg16":
        from nets_cifar.vgg16 import VGG16_Q
        return VGG16_Q(qbit)
    if name == "mobilenetv1_cifar":
        from nets_cifar.mobilenetv1 import MobileNetV1_Q
        return MobileNetV1_Q(3, qbit)
    if name == "mobilenetv1_imgnet":
        from nets_imgnet.mobilenetv1 import MobileNetV1_Q
        return MobileNetV1_Q(3, qbit)
    if name == "alexnet":
        from nets_imgnet.alexnet import AlexNet
        return AlexNet(qbit)
    if name == "squeezenet":
        from nets_imgnet.squeezenet1_0 import SqueezeNet
        return SqueezeNet(qbit)
    if name == "shufflenetv2":
        from nets_cifar.shufflenet_v2 import ShuffleNetV2
        m = ShuffleNetV2(qbit)
        m.reset_layer_inputs_outputs()            # the reference's forward needs these dicts (shufflenet_v2.py:175-183, :197)
        m.reset_layer_weights()
        return m
    raise KeyError(name)


def calibrate(name, sd, x):
    """The reference's calibration workflow: Qbits=32 with K=1 so input_q / weight_q are the raw tensors."""
    m = ref_net(name, 32).eval()
    m.load_state_dict(sd, strict=False)
    nc.set_scales(m, np.ones(64), np.ones(64))
    with torch.no_grad():
        m(x)
    layers = nc.quantized_layers(m)
    ka = np.array([float(l.input_q.abs().max()) for l in layers]) / 15.5
    kw = np.array([float(l.weight_q.abs().max()) for l in layers]) / 15.5
    return ka, kw


CASES = [("resnet50", 8, 8, 64), ("vgg16", 8, 8, 32), ("mobilenetv1_cifar", 8, 8, 32), ("mobilenetv1_imgnet", 7, 4, 224),
         ("shufflenetv2", 7, 8, 32)]

if __name__ == "__main__":
    # `make_golden_nets.py NAME ...` regenerates only the named cases and keeps the others of the existing fixture
    only = sys.argv[1:]
    path = os.path.join(HERE, "net_cases.npz")
    out = dict(np.load(path)) if only and os.path.exists(path) else {}
    for name, qbit, batch, size in CASES:
        if only and name not in only:
            continue
        x = nc.synth_images(batch, size)
        m0 = ref_net(name, 32).eval()
        sd = nc.synth_state_dict(m0)
        m0.load_state_dict(sd, strict=False)
        nc.set_scales(m0, np.ones(64), np.ones(64))
        fc_scale, fcb = nc.recenter_classifier(m0, x)
        fc_name = [n for n, mod in m0.named_modules() if mod is nc.classifier_module(m0)][0]
        sd[fc_name + ".weight"] = nc.classifier_weight(nc.classifier_module(m0), fc_scale)
        sd[fc_name + ".bias"] = fcb
        out[f"{name}.fc_scale"], out[f"{name}.fc_bias"] = np.float64(fc_scale), fcb.numpy()
        ka, kw = calibrate(name, sd, x)
        m = ref_net(name, qbit).eval()
        m.load_state_dict(sd, strict=False)
        nc.set_scales(m, ka, kw)
        with torch.no_grad():
            y = m(x)
            m32 = ref_net(name, 32).eval()
            m32.load_state_dict(sd, strict=False)
            y32 = m32(x)
        out[f"{name}.cfg"] = np.array([qbit, batch, size])
        out[f"{name}.ka"], out[f"{name}.kw"] = ka, kw
        out[f"{name}.logits"] = y.numpy()
        out[f"{name}.logits_fp32"] = y32.numpy()
        top = y.argmax(1).numpy()
        srt = np.sort(y.numpy(), 1)
        print(name, "q", qbit, "top1", top.tolist(), "fp32 top1", y32.argmax(1).tolist(),
              "margin", (srt[:, -1] - srt[:, -2]).round(4).tolist(), "logit std", float(y.std()))
    np.savez_compressed(os.path.join(HERE, "net_cases.npz"), **out)
