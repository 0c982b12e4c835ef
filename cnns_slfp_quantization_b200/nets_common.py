"""Shared plumbing of the caller nets (nets_imgnet/, nets_cifar/).

The nets are the CALLERS of the hot path: small table-driven re-statements of
the reference's model graphs with the reference's parameter names (so its state_dict files load
unchanged), built from an `ops` namespace that supplies the quantized-module factories:
  * default: this package's drop-in modules (sm_100a kernels);
  * tests / bench.py's CPU baseline pass the oracle's torch port instead (oracle/torch_port.py).
"""
import json
import os
import types

import numpy as np
import torch
import torch.nn as nn

_HERE = os.path.dirname(os.path.abspath(__file__))


def product_ops():
    from .utils import conv2d_func, sfp_quant, activation_func
    return types.SimpleNamespace(conv2d_Q=conv2d_func.conv2d_Q, conv2d_Q_bias=conv2d_func.conv2d_Q_bias,
                                 linear_Q=conv2d_func.linear_Q, layerout_quantize_func=sfp_quant.layerout_quantize_func,
                                 Swish=activation_func.Swish)


def reference_scales(name):
    """Per-layer Ka / Kw = max|.| / 15.5 the reference hard-codes in its nets (calibrated on its
    private checkpoints); data extracted to ref_scales.json, source lines recorded there."""
    with open(os.path.join(_HERE, "ref_scales.json")) as f:
        t = json.load(f)[name]
    return np.array(t["ka_max"]) / t["divisor"], np.array(t["kw_max"]) / t["divisor"]


def quantized_layers(model):
    """Quantized conv / linear modules in forward-definition order."""
    return [m for m in model.modules() if hasattr(m, "Ka") and hasattr(m, "Kw") and hasattr(m, "q_bit")]


def set_scales(model, ka=None, kw=None):
    """Re-assign the per-layer scales (plain attributes in the reference: not buffers, not in the
    state_dict)."""
    for i, m in enumerate(quantized_layers(model)):
        if ka is not None:
            m.Ka = torch.tensor(float(ka[i]), dtype=torch.float64)
        if kw is not None:
            m.Kw = torch.tensor(float(kw[i]), dtype=torch.float64)


def synth_state_dict(model, seed=0):
    """Deterministic synthetic parameters keyed by parameter NAME (independent of construction order,
    so the reference's own model classes can load the very same values for golden generation).
    Conv / linear weights ~ N(0, sqrt(2/fan_in)); BatchNorm weight ~ U(0.5, 1.5), bias ~ N(0, 0.2),
    running_mean ~ N(0, 0.2), running_var ~ U(0.5, 1.5)  (SURVEY.md section 8c caveat 3)."""
    import zlib
    out = {}
    for name, t in model.state_dict().items():
        rng = np.random.default_rng([seed, zlib.crc32(name.encode())])
        shape = tuple(t.shape)
        if name.endswith("num_batches_tracked"):
            v = np.zeros(shape, np.int64)
        elif name.endswith("running_var"):
            v = rng.uniform(0.5, 1.5, shape)
        elif name.endswith("running_mean"):
            v = rng.normal(0.0, 0.2, shape)
        elif t.dim() == 1 and name.endswith("bn3.weight"):
            v = rng.uniform(0.1, 0.3, shape)          # last BN of a residual branch: keep the stream bounded
        elif t.dim() == 1 and name.endswith("weight"):
            v = rng.uniform(0.5, 1.5, shape)
        elif t.dim() == 1:
            v = rng.normal(0.0, 0.2, shape)
        else:
            fan_in = int(np.prod(shape[1:]))
            v = rng.normal(0.0, np.sqrt(2.0 / fan_in), shape)
        out[name] = torch.from_numpy(np.asarray(v)).to(t.dtype)
    return out


def classifier_module(model):
    """The final classifier layer: the last nn.Linear (quantized or plain), or - SqueezeNet, whose classifier is a 1x1
    convolution followed by ReLU and a global average pool (nets_imgnet/squeezenet1_0.py:87-93) - the last 1x1 conv."""
    last = None
    for m in model.modules():
        if isinstance(m, nn.Linear):
            last = m
    if last is None:
        for m in model.modules():
            if isinstance(m, nn.Conv2d) and tuple(m.kernel_size) == (1, 1):
                last = m
    return last


def classifier_dims(fc):
    return (fc.out_features, fc.in_features) if isinstance(fc, nn.Linear) else (fc.out_channels, fc.in_channels)


def classifier_weight(fc, scale, seed=7):
    """Seeded N(0,1) classifier weight times `scale` (the fixture stores only the scalar and the bias)."""
    g = torch.Generator().manual_seed(seed)
    w = torch.randn(fc.out_features, fc.in_features, generator=g, dtype=torch.float64)
    return (w * float(scale)).float()


def recenter_classifier(model, x, seed=7):
    """Random-init nets predict one class for every image (SURVEY.md section 8c caveat 1).  Make top-1
    image dependent: weight = scale * N(0,1) with unit-variance logits on the probe batch x, and
    bias = -W . mean(feature).  Call on a float32 (q_bit 32) instance; returns (scale, bias)."""
    fc = classifier_module(model)
    feats = []
    h = fc.register_forward_hook(lambda mod, inp, out: feats.append(inp[0].detach()))
    with torch.no_grad():
        model(x)
    h.remove()
    f = feats[0].reshape(-1, fc.in_features).double().cpu()
    w = classifier_weight(fc, 1.0, seed).double()
    f_c = f - f.mean(0, keepdim=True)
    scale = 1.0 / float((f_c @ w.T).std().clamp_min(1e-12))
    w = classifier_weight(fc, scale, seed)
    b = -(w.double() @ f.mean(0)).float()
    return scale, b


def apply_classifier(model, scale, bias, seed=7):
    fc = classifier_module(model)
    with torch.no_grad():
        fc.weight.copy_(classifier_weight(fc, scale, seed).to(fc.weight.device))
        fc.bias.copy_(torch.as_tensor(bias, dtype=torch.float32).to(fc.bias.device))


def synth_images(batch, size, seed=1234, channels=3):
    """Synthetic input batch: N(0,1) noise plus a low-frequency per-image pattern (gives the
    classifier something image-dependent)."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(batch, channels, size, size, generator=g)
    yy, xx = torch.meshgrid(torch.linspace(-1, 1, size), torch.linspace(-1, 1, size), indexing="ij")
    ph = torch.rand(batch, channels, 3, generator=g) * 6.28318
    pat = torch.sin(3.0 * xx[None, None] + ph[..., 0:1, None]) * torch.cos(2.0 * yy[None, None] + ph[..., 1:2, None])
    return x + 1.5 * pat * (0.5 + ph[..., 2:3, None] / 6.28318)


def prototype_classifier(feats, out_features, target=8.0, seed=11, offset=0.0):
    """Nearest-prototype classifier for decisive whole-net fixtures: class i (i < batch) is image i.

    Row i = target * d_i / |d_i|^2 with d_i = f_i - mean(f) (so image i scores `target` on its own class and
    target * <d_i, d_j> / |d_i|^2 on the others), rows >= batch are seeded random directions of comparable norm,
    bias = -W . mean(f) + offset.  `feats` are the REFERENCE net's classifier inputs on the fixture batch (averaged over
    the pixels when the classifier is SqueezeNet's 1x1 convolution; `offset` then lifts every pre-activation above its
    ReLU so that the pooled logits stay linear in the features); the fixture stores the prototype rows, the scalar for
    the random rows and the bias (tests/golden/make_golden_net224.py).
    Returns (protos [batch, in] float32, rest_scale float, bias [out] float32)."""
    f = torch.as_tensor(feats).double()
    fbar = f.mean(0, keepdim=True)
    d = f - fbar
    n2 = (d * d).sum(1, keepdim=True).clamp_min(1e-30)
    protos = (target * d / n2).float()
    rest_scale = float(0.5 * target / n2.sqrt().median())
    w = prototype_weight(protos, out_features, rest_scale, seed)
    bias = (-(w.double() @ fbar[0]) + offset).float()
    return protos, rest_scale, bias


def prototype_weight(protos, out_features, rest_scale, seed=11):
    protos = torch.as_tensor(protos, dtype=torch.float32)
    b, cin = protos.shape
    assert b <= out_features
    g = torch.Generator().manual_seed(seed)
    u = torch.randn(out_features, cin, generator=g, dtype=torch.float64)
    u = u / u.norm(dim=1, keepdim=True)
    w = (u * float(rest_scale)).float()
    w[:b] = protos
    return w


def apply_prototype_classifier(model, protos, rest_scale, bias, seed=11):
    fc = classifier_module(model)
    out_f, _ = classifier_dims(fc)
    with torch.no_grad():
        fc.weight.copy_(prototype_weight(protos, out_f, rest_scale, seed).to(fc.weight.device).view_as(fc.weight))
        fc.bias.copy_(torch.as_tensor(bias, dtype=torch.float32).to(fc.bias.device))
