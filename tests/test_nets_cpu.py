"""CPU: the table-driven caller nets (nets_cifar/*) built on the ORACLE's torch port of the quantized modules
reproduce the logits the REFERENCE nets produced for tests/golden/net_cases.npz (same synthetic parameters keyed by
name, same calibrated scales).  This pins the nets' topology, parameter names and per-layer scale indexing - the part
of the drop-in that is host logic - without a GPU; the GPU tests then swap the oracle's modules for the product's."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _build(name, qbit, ops):
    from cnns_slfp_quantization_b200.nets_cifar import VGG16_Q, MobileNetV1_Q, ShuffleNetV2
    if name == "vgg16":
        return VGG16_Q(qbit, ops=ops)
    if name == "mobilenetv1_cifar":
        return MobileNetV1_Q(3, qbit, ops=ops)
    return ShuffleNetV2(qbit, ops=ops)


@pytest.mark.parametrize("name", ["shufflenetv2", "mobilenetv1_cifar", "vgg16"])
def test_caller_net_on_the_oracle_port_matches_the_reference_logits(name):
    from cnns_slfp_quantization_b200 import nets_common as nc
    from oracle import torch_port
    g = np.load(os.path.join(ROOT, "tests", "golden", "net_cases.npz"))
    qbit, batch, size = [int(v) for v in g[f"{name}.cfg"]]
    torch.set_num_threads(min(8, os.cpu_count() or 1))
    m = _build(name, qbit, torch_port.ops()).eval()
    m.load_state_dict(nc.synth_state_dict(m))
    nc.apply_classifier(m, float(g[f"{name}.fc_scale"]), g[f"{name}.fc_bias"])
    nc.set_scales(m, g[f"{name}.ka"], g[f"{name}.kw"])
    assert len(nc.quantized_layers(m)) == len(g[f"{name}.ka"])
    with torch.no_grad():
        y = m(nc.synth_images(batch, size)).numpy()
    ref = g[f"{name}.logits"]
    # same ATen CPU kernels as the reference run: equal up to float32 summation order inside oneDNN
    np.testing.assert_allclose(y, ref, rtol=0, atol=2e-4 * float(np.abs(ref).max()))
    assert (y.argmax(1) == ref.argmax(1)).all()


def test_resnet50_training_forward_on_cpu_takes_the_stock_modules():
    """Off the GPU the QAT glue (utils/bn_act.bn_act, maxpool_train, prepare_weights_batched) must be transparent: the
    ResNet-50 training forward / backward on CPU tensors equals the plain module sequence of the reference
    (nets_imgnet/resnet50.py:69-90, 231-245) - same outputs, same gradients, same BatchNorm statistics - with the
    oracle's port of the quantized modules in both."""
    import copy
    import torch.nn as nn
    from cnns_slfp_quantization_b200.nets_imgnet import ResNet50
    from oracle import torch_port
    torch.set_num_threads(min(8, os.cpu_count() or 1))
    torch.manual_seed(0)
    scales = (np.full(54, 0.25), np.full(54, 0.02))
    m = ResNet50(8, ops=torch_port.ops(), scales=scales).train()
    ref = copy.deepcopy(m)

    def plain_block(b, x):                                      # the reference's Bottleneck.forward
        identity = x
        out = b.relu(b.bn1(b.conv1(x)))
        out = b.relu(b.bn2(b.conv2(out)))
        out = b.bn3(b.conv3(out))
        if b.downsample is not None:
            identity = b.downsample(x)
        out = out + identity
        return b.relu(out)

    def plain_forward(net, x):
        x = net.maxpool(net.relu(net.bn1(net.conv1(x))))
        for li in range(1, 5):
            for b in getattr(net, f"layer{li}"):
                x = plain_block(b, x)
        return net.fc(torch.flatten(net.avgpool(x), 1))

    x = torch.randn(2, 3, 32, 32)
    y = m(x)
    y.square().mean().backward()
    yr = plain_forward(ref, x)
    yr.square().mean().backward()
    assert torch.equal(y, yr)
    for (n1, p1), (n2, p2) in zip(m.named_parameters(), ref.named_parameters()):
        assert n1 == n2 and torch.equal(p1.grad, p2.grad), n1
    for (n1, b1), (n2, b2) in zip(m.named_buffers(), ref.named_buffers()):
        assert n1 == n2 and torch.equal(b1, b2), n1
    assert int(m.bn1.num_batches_tracked) == 1
