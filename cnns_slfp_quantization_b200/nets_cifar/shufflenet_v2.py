"""ShuffleNetV2 (x0.5 / x1 / x1.5 / x2) for CIFAR-100 built from the SLFP quantized modules - the caller that exercises
the depthwise / grouped path with channel counts that are not multiples of 16 (24, 58, 116, 232; BASELINE config 5,
reference: nets_cifar/shufflenet_v2.py:47-252).

Same topology, parameter names (pre.{0,1}, stage{2,3,4}.{i}.residual.{0..9} / shortcut.{0..5}, conv5.{0,1}, fc) and
per-layer scale indexing as the reference, so its state_dicts load unchanged:
    pre 0;  stage offsets 0 / 14 / 40;  a stage's down-sampling unit uses off+1..off+5 (residual 1x1, dw 3x3, 1x1,
    shortcut dw 3x3, 1x1), its i-th basic unit off+6+3i..off+8+3i;  conv5 55;  fc 56   (:159-170, :243-251).
Every unit ends in concat + channel_shuffle(2); BatchNorm outputs that feed a ReLU go through layerout_quantize_func
(SFP<4,4>, NaN at exact 0 like the reference - sfp_quant.LAYEROUT_ZERO_IS_ZERO selects the intended value).

Deviation, on purpose: the reference's forward raises AttributeError unless reset_layer_inputs_outputs() and
reset_layer_weights() were called first (:175-183, :197 - its __init__ never creates the dicts).  Here the calibration
taps are recorded only after those calls and forward() works without them.  Written table-driven, not copied.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from ..nets_common import product_ops, quantized_layers, reference_scales

_WIDTHS = {0.5: (48, 96, 192, 1024), 1: (116, 232, 464, 1024), 1.5: (176, 352, 704, 1024), 2: (244, 488, 976, 2048)}
_STAGES = ((3, 0), (7, 14), (3, 40))            # (basic units after the down-sampling unit, scale offset)


def channel_shuffle(x, groups):
    n, c, h, w = x.shape
    return x.view(n, groups, c // groups, h, w).transpose(1, 2).contiguous().view(n, c, h, w)


class ShuffleUnit(nn.Module):
    def __init__(self, ops, qbit, cin, cout, stride, Kw, Ka, idx):
        """idx: the five consecutive scale indices of this unit (a basic unit uses the first three)."""
        super().__init__()
        self.stride, self.in_channels, self.out_channels = stride, cin, cout
        mk = lambda j: ops.conv2d_Q(q_bit=qbit, Kw=Kw[idx + j], Ka=Ka[idx + j])
        lq = lambda: ops.layerout_quantize_func(q_bit=qbit)
        down = stride != 1 or cin != cout
        c, half = (cin, cout // 2) if down else (cin // 2, cin // 2)
        branch = [mk(0)(c, c, 1), nn.BatchNorm2d(c), lq(), nn.ReLU(),
                  mk(1)(c, c, 3, stride=stride, padding=1, groups=c), nn.BatchNorm2d(c),
                  mk(2)(c, half, 1), nn.BatchNorm2d(half), lq(), nn.ReLU()]
        if down:
            self.residual = nn.Sequential(*branch)
            self.shortcut = nn.Sequential(mk(3)(c, c, 3, stride=stride, padding=1, groups=c), nn.BatchNorm2d(c),
                                          mk(4)(c, half, 1), nn.BatchNorm2d(half), lq(), nn.ReLU())
        else:
            self.shortcut = nn.Sequential()
            self.residual = nn.Sequential(*branch)

    def forward(self, x):
        if self.stride == 1 and self.out_channels == self.in_channels:
            shortcut, residual = torch.split(x, self.in_channels // 2, dim=1)
        else:
            shortcut = residual = x
        return channel_shuffle(torch.cat([self.shortcut(shortcut), self.residual(residual)], dim=1), 2)


class ShuffleNetV2(nn.Module):
    def __init__(self, qbit, ratio=1, class_num=100, ops=None, scales=None):
        super().__init__()
        ops = ops or product_ops()
        Ka, Kw = scales if scales is not None else reference_scales("shufflenetv2_cifar")
        widths = _WIDTHS[ratio]
        self.pre = nn.Sequential(ops.conv2d_Q(q_bit=qbit, Kw=Kw[0], Ka=Ka[0])(3, 24, 3, padding=1), nn.BatchNorm2d(24))
        cin = 24
        for si, ((repeat, off), cout) in enumerate(zip(_STAGES, widths[:3]), start=2):
            units = [ShuffleUnit(ops, qbit, cin, cout, 2, Kw, Ka, off + 1)]
            units += [ShuffleUnit(ops, qbit, cout, cout, 1, Kw, Ka, off + 6 + 3 * i) for i in range(repeat)]
            setattr(self, f"stage{si}", nn.Sequential(*units))
            cin = cout
        self.conv5 = nn.Sequential(ops.conv2d_Q(q_bit=qbit, Kw=Kw[55], Ka=Ka[55])(cin, widths[3], 1), nn.BatchNorm2d(widths[3]),
                                   ops.layerout_quantize_func(q_bit=qbit), nn.ReLU())
        self.fc = ops.linear_Q(q_bit=qbit, Kw=Kw[56], Ka=Ka[56])(widths[3], class_num)

    # calibration taps (cifar100_train_eval.py:213-277 reads them after a q_bit = 32 forward)
    def get_layer_inputs(self):
        return self.layer_inputs

    def get_layer_outputs(self):
        return self.layer_outputs

    def reset_layer_inputs_outputs(self):
        self.layer_inputs, self.layer_outputs = {}, {}

    def get_layer_weights(self):
        return self.layer_weights

    def reset_layer_weights(self):
        self.layer_weights = {}

    def forward(self, x):
        x = self.conv5(self.stage4(self.stage3(self.stage2(self.pre(x)))))
        x = F.adaptive_avg_pool2d(x, 1)
        x = self.fc(x.view(x.size(0), -1))
        if hasattr(self, "layer_inputs") and hasattr(self, "layer_weights"):
            for i, layer in enumerate(quantized_layers(self)):        # module order == the reference's index order
                self.layer_inputs[i] = layer.input_q
                self.layer_weights[i] = layer.weight_q
            self.layer_outputs[55] = x
        return x
