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
