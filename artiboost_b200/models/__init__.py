"""B200 counterparts of `anakin.models` for the clasbased path: same registries, class names, cfg keys, forward
contracts and state_dict names (anakin/utils/builder.py:5-11, anakin/models/{resnet,simplebaseline,hybridbaseline,
mlp,arch}.py); the arithmetic runs on the tcgen05 GEMM + NHWC kernels of the C-ABI."""
from .registry import BACKBONE, HEAD, MODEL, Registry, build_arch_model_list, build_backbone, build_from_cfg, build_head, build_model  # noqa: F401
from .resnet import ResNet, ResNet18, ResNet34, ResNet50, ResNet101, ResNet152, FrozenBatchNorm2d  # noqa: F401
from .simplebaseline import IntegralDeconvHead  # noqa: F401
from .mlp import MLP_O  # noqa: F401
from .hybridbaseline import HybridBaseline  # noqa: F401
from .arch import Arch  # noqa: F401
