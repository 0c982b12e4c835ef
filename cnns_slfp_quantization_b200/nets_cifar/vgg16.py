"""VGG-16 for CIFAR-100 from the SLFP quantized modules (reference: nets_cifar/vgg16.py:13-132):
13 biased 3x3 convolutions + BN + ReLU in five stages with 2x2 max-pools, then three quantized
linear layers that all use scale index 13 (nets_cifar/vgg16.py:99,104,108 -- kept).
Parameter names: layer{1..5}.{j}.*, fc1.2.*, fc2.0.*, fc3.*."""
import torch.nn as nn

from ..nets_common import product_ops, reference_scales

_CFG = ((64, 64), (128, 128), (256, 256, 256), (512, 512, 512), (512, 512, 512))


class VGG16_Q(nn.Module):
    def __init__(self, qbit, ops=None, scales=None, num_classes=100):
        super().__init__()
        ops = ops or product_ops()
        Ka, Kw = scales if scales is not None else reference_scales("vgg16_cifar")
        Conv2d = ops.conv2d_Q_bias(q_bit=qbit, Kw=Kw, Ka=Ka)
        Linear = ops.linear_Q(q_bit=qbit, Kw=Kw, Ka=Ka)
        cin, i = 3, 0
        for li, widths in enumerate(_CFG, start=1):
            mods = []
            for cout in widths:
                mods += [Conv2d(cin, cout, 3, Kw[i], Ka[i], 1, 1), nn.BatchNorm2d(cout), nn.ReLU()]
                cin, i = cout, i + 1
            mods.append(nn.MaxPool2d(2, 2))
            setattr(self, f"layer{li}", nn.Sequential(*mods))
        self.fc1 = nn.Sequential(nn.AdaptiveAvgPool2d(1), nn.Flatten(), Linear(512, 512, Kw[13], Ka[13]), nn.ReLU(), nn.Dropout())
        self.fc2 = nn.Sequential(Linear(512, 256, Kw[13], Ka[13]), nn.ReLU(), nn.Dropout())
        self.fc3 = Linear(256, num_classes, Kw[13], Ka[13])

    def forward(self, x):
        for li in range(1, 6):
            x = getattr(self, f"layer{li}")(x)
        return self.fc3(self.fc2(self.fc1(x)))
