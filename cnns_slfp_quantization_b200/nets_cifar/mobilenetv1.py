"""MobileNetV1 for CIFAR-100 (BASELINE config 1; reference: nets_cifar/mobilenetv1.py:10-82): the
ImageNet feature stack at 32x32 with a global average pool and a quantized classifier (scale 27)."""
import torch.nn as nn

from ..nets_common import product_ops, reference_scales
from ..nets_imgnet.mobilenetv1 import _features


class MobileNetV1_Q(nn.Module):
    def __init__(self, ch_in, qbit, ops=None, scales=None, num_classes=100):
        super().__init__()
        ops = ops or product_ops()
        Ka, Kw = scales if scales is not None else reference_scales("mobilenetv1_cifar")
        self.model = nn.Sequential(*_features(ops, qbit, Ka, Kw, ch_in), nn.AdaptiveAvgPool2d(1))
        self.fc = ops.linear_Q(q_bit=qbit, Kw=Kw[27], Ka=Ka[27])(1024, num_classes)

    def forward(self, x):
        x = self.model(x)
        return self.fc(x.view(-1, 1024))
