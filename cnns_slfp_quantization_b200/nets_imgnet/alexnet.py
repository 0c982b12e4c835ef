"""AlexNet from the SLFP quantized modules (reference: nets_imgnet/alexnet.py:18-69) - the caller with the 11x11 /
stride-4 stem, a 5x5 layer and the 4096-wide quantized linear layers (58.6 M of its 714 M MAC per image), bias in every
layer, no BatchNorm.  Same parameter names (features.{0,3,6,8,10}, classifier.{1,4,6}) and scale indexing (layer i uses
Kw[i], Ka[i]) as the reference, so its state_dicts load unchanged.  Table-driven, not copied."""
import torch.nn as nn

from ..nets_common import product_ops, reference_scales

# (out channels, kernel, stride, padding, max-pool after the ReLU)
_FEATURES = ((64, 11, 4, 2, True), (192, 5, 1, 2, True), (384, 3, 1, 1, False), (256, 3, 1, 1, False), (256, 3, 1, 1, True))


class AlexNet(nn.Module):
    def __init__(self, qbit, num_classes=1000, ops=None, scales=None):
        super().__init__()
        ops = ops or product_ops()
        Ka, Kw = scales if scales is not None else reference_scales("alexnet_imgnet")
        Conv2d = ops.conv2d_Q_bias(q_bit=qbit, Kw=Kw, Ka=Ka)
        Linear = ops.linear_Q(q_bit=qbit, Kw=Kw, Ka=Ka)
        mods, cin = [], 3
        for i, (cout, k, s, p, pool) in enumerate(_FEATURES):
            mods += [Conv2d(cin, cout, k, Kw[i], Ka[i], stride=s, padding=p), nn.ReLU(inplace=True)]
            if pool:
                mods.append(nn.MaxPool2d(kernel_size=3, stride=2))
            cin = cout
        self.features = nn.Sequential(*mods)
        self.classifier = nn.Sequential(nn.Dropout(), Linear(256 * 6 * 6, 4096, Kw[5], Ka[5]), nn.ReLU(inplace=True),
                                        nn.Dropout(), Linear(4096, 4096, Kw[6], Ka[6]), nn.ReLU(inplace=True),
                                        Linear(4096, num_classes, Kw[7], Ka[7]))

    def forward(self, x):
        x = self.features(x)
        return self.classifier(x.reshape(x.size(0), 256 * 6 * 6))
