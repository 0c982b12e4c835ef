"""ResNet-50 v1.5 built from the SLFP quantized modules -- the caller of the hot path that
BASELINE.json's headline metric is quoted on (reference: nets_imgnet/resnet50.py:24-245).

Same topology, parameter names (conv1, bn1, layer{1..4}.{i}.conv{1,2,3} / bn{1,2,3} /
downsample.{0,1}, fc) and per-layer scale indexing as the reference: stem 0; stage offsets
1 / 11 / 24 / 43 with downsample = off, block b convs = off + 3b + {1,2,3}; fc 53
(nets_imgnet/resnet50.py:113-145, 180-211).  Written table-driven, not copied.
"""
import torch
import torch.nn as nn

from ..nets_common import product_ops, reference_scales
from ..utils.bn_act import bn_act, flush_batch_counters, maxpool_train
from ..utils.conv2d_func import prepare_weights_batched

_STAGES = ((64, 3, 1, 1), (128, 4, 2, 11), (256, 6, 2, 24), (512, 3, 2, 43))   # planes, blocks, stride, scale offset


class Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, ops, qbit, Ka, Kw, idx, inplanes, planes, stride, downsample_idx=None):
        super().__init__()
        out = planes * self.expansion
        mk = lambda i: ops.conv2d_Q(q_bit=qbit, Kw=Kw[i], Ka=Ka[i])
        self.conv1 = mk(idx + 1)(inplanes, planes, 1)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = mk(idx + 2)(planes, planes, 3, stride=stride, padding=1)       # v1.5: stride on the 3x3
        self.bn2 = nn.BatchNorm2d(planes)
        self.conv3 = mk(idx + 3)(planes, out, 1)
        self.bn3 = nn.BatchNorm2d(out)
        self.relu = nn.ReLU()
        self.downsample = None
        if downsample_idx is not None:
            self.downsample = nn.Sequential(mk(downsample_idx)(inplanes, out, 1, stride=stride), nn.BatchNorm2d(out))
        self.stride = stride

    def forward(self, x):
        # relu(bn(conv)) / relu(bn3(conv3) + identity) as in the reference (nets_imgnet/resnet50.py:71-88); in training mode on
        # the GPU each BatchNorm (+ add) (+ ReLU) group is one fused op (utils/bn_act.py), otherwise the stock modules run
        # (`consumers`: the fused op also writes the activation codes its readers would otherwise compute in a pass of their own)
        identity = x
        out = bn_act(self.conv1(x), self.bn1, consumers=(self.conv2,))
        out = bn_act(self.conv2(out), self.bn2, consumers=(self.conv3,))
        if self.downsample is not None:
            identity = bn_act(self.downsample[0](x), self.downsample[1], relu=False)
        return bn_act(self.conv3(out), self.bn3, relu=True, residual=identity, consumers=self.__dict__.get("_readers", ()))


class ResNet50(nn.Module):
    def __init__(self, qbit, num_classes=1000, ops=None, scales=None):
        super().__init__()
        ops = ops or product_ops()
        Ka, Kw = scales if scales is not None else reference_scales("resnet50_imgnet")
        self.qbit, self.Ka, self.Kw = qbit, Ka, Kw
        self.conv1 = ops.conv2d_Q(q_bit=qbit, Kw=Kw[0], Ka=Ka[0])(3, 64, 7, stride=2, padding=3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(kernel_size=3, stride=2, padding=1)
        inplanes = 64
        for li, (planes, blocks, stride, off) in enumerate(_STAGES, start=1):
            layer = []
            for b in range(blocks):
                layer.append(Bottleneck(ops, qbit, Ka, Kw, off + 3 * b, inplanes, planes, stride if b == 0 else 1,
                                        downsample_idx=off if b == 0 else None))
                inplanes = planes * Bottleneck.expansion
            setattr(self, f"layer{li}", nn.Sequential(*layer))
        blocks = [b for li in range(1, 5) for b in getattr(self, f"layer{li}")]
        for b, nxt in zip(blocks, blocks[1:]):                     # readers of a block's output (kept out of the module tree)
            b.__dict__["_readers"] = (nxt.conv1,) + ((nxt.downsample[0],) if nxt.downsample is not None else ())
        self.avgpool = nn.AdaptiveAvgPool2d((1, 1))
        self.fc = ops.linear_Q(q_bit=qbit, Kw=Kw[53], Ka=Ka[53])(inplanes, num_classes)
        for m in self.modules():                                   # nets_imgnet/resnet50.py:149-154
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)

    def forward(self, x):
        if x.is_cuda and self.qbit in (7, 8):
            prepare_weights_batched(self)                          # all 54 layers' weight re-quantization in one launch
        x = maxpool_train(bn_act(self.conv1(x), self.bn1), self.maxpool)
        x = self.layer4(self.layer3(self.layer2(self.layer1(x))))
        x = torch.flatten(self.avgpool(x), 1)
        flush_batch_counters()                                     # the BatchNorm layers' num_batches_tracked, one launch
        return self.fc(x)
