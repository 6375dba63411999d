"""paddle.vision.models. Parity: python/paddle/vision/models/__init__.py (51 names)."""
from .resnet import (BasicBlock, BottleneckBlock, ResNet, resnet18, resnet34, resnet50, resnet101, resnet152, resnext50_32x4d,  # noqa: F401
                     resnext50_64x4d, resnext101_32x4d, resnext101_64x4d, resnext152_32x4d, resnext152_64x4d, wide_resnet50_2, wide_resnet101_2)
from .small_nets import (VGG, AlexNet, DenseNet, GoogLeNet, InceptionV3, LeNet, MobileNetV1, MobileNetV2, MobileNetV3Large,  # noqa: F401
                         MobileNetV3Small, ShuffleNetV2, SqueezeNet, alexnet, densenet121, densenet161, densenet169, densenet201,
                         densenet264, googlenet, inception_v3, mobilenet_v1, mobilenet_v2, mobilenet_v3_large, mobilenet_v3_small,
                         shufflenet_v2_swish, shufflenet_v2_x0_25, shufflenet_v2_x0_33, shufflenet_v2_x0_5, shufflenet_v2_x1_0,
                         shufflenet_v2_x1_5, shufflenet_v2_x2_0, squeezenet1_0, squeezenet1_1, vgg11, vgg13, vgg16, vgg19)
