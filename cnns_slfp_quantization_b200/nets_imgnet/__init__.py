from .resnet50 import ResNet50, Bottleneck            # noqa: F401
from .mobilenetv1 import MobileNetV1_Q                # noqa: F401
from .alexnet import AlexNet                          # noqa: F401
from .squeezenet1_0 import SqueezeNet, Fire           # noqa: F401
