"""MobileNetV1 (ImageNet) from the SLFP quantized modules: depthwise 3x3 + pointwise 1x1 pairs
(reference: nets_imgnet/mobilenetv1.py:10-83; the classifier there is a plain nn.Linear, :61).
Parameter names follow the reference's nn.Sequential nesting: model.{i}.{0,1,3,4}.*, fc.*."""
import torch.nn as nn

from ..nets_common import product_ops, reference_scales

# (in, out, stride) of the 13 depthwise-separable pairs
_PAIRS = ((32, 64, 1), (64, 128, 2), (128, 128, 1), (128, 256, 2), (256, 256, 1), (256, 512, 2),
          (512, 512, 1), (512, 512, 1), (512, 512, 1), (512, 512, 1), (512, 512, 1), (512, 1024, 2), (1024, 1024, 1))


def _features(ops, qbit, Ka, Kw, ch_in):
    conv = lambda i: ops.conv2d_Q(q_bit=qbit, Kw=Kw[i], Ka=Ka[i])
    blocks = [nn.Sequential(conv(0)(ch_in, 32, 3, stride=2, padding=1, bias=False), nn.BatchNorm2d(32), nn.ReLU(inplace=True))]
    for j, (inp, oup, stride) in enumerate(_PAIRS):
        i = 1 + 2 * j
        blocks.append(nn.Sequential(
            conv(i)(inp, inp, 3, stride=stride, padding=1, groups=inp, bias=False), nn.BatchNorm2d(inp), nn.ReLU(inplace=True),
            conv(i + 1)(inp, oup, 1, stride=1, padding=0, bias=False), nn.BatchNorm2d(oup), nn.ReLU(inplace=True)))
    return blocks


class MobileNetV1_Q(nn.Module):
    def __init__(self, ch_in, qbit, ops=None, scales=None, num_classes=1000):
        super().__init__()
        ops = ops or product_ops()
        Ka, Kw = scales if scales is not None else reference_scales("mobilenetv1_imgnet")
        self.model = nn.Sequential(*_features(ops, qbit, Ka, Kw, ch_in), nn.AvgPool2d(7))
        self.fc = nn.Linear(1024, num_classes)

    def forward(self, x):
        x = self.model(x)
        return self.fc(x.view(-1, 1024))
