"""SqueezeNet 1.0 from the SLFP quantized modules (reference: nets_imgnet/squeezenet1_0.py:20-112) - 26 biased
quantized convolutions: a 7x7 / stride-2 stem without padding, eight Fire modules (1x1 squeeze, then 1x1 and 3x3
expand branches concatenated along the channels), ceil-mode 3x3 / 2 max-pools, and a 1x1 classifier convolution followed
by ReLU and a global average pool.  Same parameter names (features.0, features.{3,4,5,7,8,9,10,12}.{squeeze,
expand1x1, expand3x3}, classifier.1) and scale indexing (stem 0, Fire j uses 3j+1 .. 3j+3, classifier 25) as the
reference.  Table-driven, not copied."""
import torch
import torch.nn as nn
import torch.nn.init as init

from ..nets_common import product_ops, reference_scales

# (in, squeeze, expand1x1, expand3x3) of the eight Fire modules; "P" = ceil-mode max-pool
_LAYOUT = ((96, 16, 64, 64), (128, 16, 64, 64), (128, 32, 128, 128), "P", (256, 32, 128, 128), (256, 48, 192, 192),
           (384, 48, 192, 192), (384, 64, 256, 256), "P", (512, 64, 256, 256))


class Fire(nn.Module):
    def __init__(self, ops, qbit, inplanes, squeeze, e1, e3, Kw, Ka):
        super().__init__()
        mk = lambda j: ops.conv2d_Q_bias(q_bit=qbit, Kw=Kw[j], Ka=Ka[j])
        self.inplanes = inplanes
        self.squeeze = mk(0)(inplanes, squeeze, kernel_size=1)
        self.squeeze_activation = nn.ReLU(inplace=True)
        self.expand1x1 = mk(1)(squeeze, e1, kernel_size=1)
        self.expand1x1_activation = nn.ReLU(inplace=True)
        self.expand3x3 = mk(2)(squeeze, e3, kernel_size=3, padding=1)
        self.expand3x3_activation = nn.ReLU(inplace=True)

    def forward(self, x):
        x = self.squeeze_activation(self.squeeze(x))
        return torch.cat([self.expand1x1_activation(self.expand1x1(x)), self.expand3x3_activation(self.expand3x3(x))], 1)


class SqueezeNet(nn.Module):
    def __init__(self, qbit, version=1.0, num_classes=1000, ops=None, scales=None):
        super().__init__()
        if version != 1.0:
            raise ValueError(f"Unsupported SqueezeNet version {version}: the reference builds 1.0 only")
        ops = ops or product_ops()
        Ka, Kw = scales if scales is not None else reference_scales("squeezenet1_0_imgnet")
        self.num_classes = num_classes
        mods = [ops.conv2d_Q_bias(q_bit=qbit, Kw=Kw[0], Ka=Ka[0])(3, 96, kernel_size=7, stride=2), nn.ReLU(inplace=True),
                nn.MaxPool2d(kernel_size=3, stride=2, ceil_mode=True)]
        j = 1
        for item in _LAYOUT:
            if item == "P":
                mods.append(nn.MaxPool2d(kernel_size=3, stride=2, ceil_mode=True))
            else:
                mods.append(Fire(ops, qbit, *item, Kw=Kw[j:], Ka=Ka[j:]))
                j += 3
        self.features = nn.Sequential(*mods)
        final_conv = ops.conv2d_Q_bias(q_bit=qbit, Kw=Kw[25], Ka=Ka[25])(512, num_classes, kernel_size=1)
        self.classifier = nn.Sequential(nn.Dropout(p=0.5), final_conv, nn.ReLU(inplace=True), nn.AdaptiveAvgPool2d((1, 1)))
        for m in self.modules():                       # nets_imgnet/squeezenet1_0.py:96-103
            if isinstance(m, nn.Conv2d):
                if m is final_conv:
                    init.normal_(m.weight, mean=0.0, std=0.01)
                else:
                    init.kaiming_uniform_(m.weight)
                if m.bias is not None:
                    init.constant_(m.bias, 0)

    def forward(self, x):
        return self.classifier(self.features(x)).view(x.size(0), self.num_classes)
