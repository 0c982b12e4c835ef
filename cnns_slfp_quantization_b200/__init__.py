"""cnns_slfp_quantization_b200 -- B200-native (sm_100a) SLFP-quantized convolution hot path.

A from-scratch implementation of the hot path of happyxtt/CNNs_SLFP_quantization behind the
reference's own Python interface:

    utils.sfp_quant        quantize_weight / quantize_act / quantize_layerout and the nn.Module wrappers
    utils.conv2d_func      conv2d_Q / conv2d_Q_bias / linear_Q class factories
    utils.activation_func  STLFunction / STL / Swish / Sigmoid
    utils.optimizer        DSGD / SSGD / NormalSGD

All arithmetic runs in hand-written CUDA kernels (csrc/) reached through the C ABI of
include/slfp_b200.h (libslfp_b200.so, bound with ctypes).  There is no CPU fallback.

`install_as_utils()` registers this package's `utils` under the top-level name `utils`, which is
how the reference's nets (`from utils.sfp_quant import *`) pick it up unchanged.
"""
import sys

__version__ = "0.1.0"


def install_as_utils():
    """Make `import utils.sfp_quant` (etc.) resolve to this package's drop-in modules."""
    from . import utils as _u
    from .utils import sfp_quant, conv2d_func, activation_func, optimizer
    sys.modules["utils"] = _u
    sys.modules["utils.sfp_quant"] = sfp_quant
    sys.modules["utils.conv2d_func"] = conv2d_func
    sys.modules["utils.activation_func"] = activation_func
    sys.modules["utils.optimizer"] = optimizer
    return _u
