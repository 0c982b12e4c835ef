from .vgg16 import VGG16_Q                             # noqa: F401
from .mobilenetv1 import MobileNetV1_Q                 # noqa: F401
from .shufflenet_v2 import ShuffleNetV2                # noqa: F401
