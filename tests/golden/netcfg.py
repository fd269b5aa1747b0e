"""Shared by make_golden_network.py (reference side, build container) and tests/test_gpu_network.py (our side, GPU box):
the clasbased ARCH / DATA_PRESET of config/ho3dv2_clasbased_jlol_artiboost2.yaml:98-170 at 256x256, the seeded
BatchNorm randomisation and the seeded inputs."""
import copy

import torch

SEED = 1  # TRAIN.MANUAL_SEED, yaml:139

ARCH = {
    "TYPE": "HybridBaseline", "PRETRAINED": "",
    "BACKBONE": {"TYPE": "ResNet34", "PRETRAINED": True, "FREEZE_BATCHNORM": False},
    "HYBRID_HEAD": {"TYPE": "IntegralDeconvHead", "NCLASSES": 22, "DECONV_WITH_BIAS": False, "NORM_TYPE": "softmax",
                    "INPUT_CHANNEL": 512, "DEPTH_RESOLUTION": 28, "NUM_DECONV_LAYERS": 2, "NUM_DECONV_FILTERS": [256, 256],
                    "NUM_DECONV_KERNELS": [4, 4], "FINAL_CONV_KERNEL": 1},
    "BOX_HEAD": {"TYPE": "MLP_O", "LAYERS_N": [512, 256, 128], "OUT_CHANNEL": 6},
    "PREVIOUS": [],
}
DATA_PRESET = {"PRESET_TYPE": "", "USE_CACHE": True, "FILTER_NO_CONTACT": False, "FILTER_THRESH": 0.0,
               "BBOX_EXPAND_RATIO": 1.2, "FULL_IMAGE": False, "IMAGE_SIZE": [224, 224], "HEATMAP_SIZE": [28, 28],
               "HEATMAP_SIGMA": 2.0, "CENTER_IDX": 0, "CROP_MODEL": "root_obj"}


def arch_cfg(backbone="ResNet34", size=256):
    arch, preset = copy.deepcopy(ARCH), copy.deepcopy(DATA_PRESET)
    arch["BACKBONE"]["PRETRAINED"] = False  # no network for the ImageNet weights (resnet.py:194-197)
    arch["BACKBONE"]["TYPE"] = backbone
    if backbone == "ResNet50":
        arch["HYBRID_HEAD"]["INPUT_CHANNEL"] = 2048
        arch["BOX_HEAD"]["LAYERS_N"] = [2048, 256, 128]
    preset["IMAGE_SIZE"] = [size, size]          # a 256x256 input needs HEATMAP_SIZE 32
    preset["HEATMAP_SIZE"] = [size // 8, size // 8]
    return arch, preset


def randomise_bn(model, seed=SEED + 1):
    """Non-trivial eval-mode BatchNorm statistics / affine, identical on both sides (CPU generator)."""
    g = torch.Generator().manual_seed(seed)
    for m in model.modules():
        if isinstance(m, torch.nn.BatchNorm2d) or type(m).__name__ == "FrozenBatchNorm2d":
            n = m.weight.shape[0]
            with torch.no_grad():
                m.weight.copy_(1.0 + 0.2 * torch.randn(n, generator=g))
                m.bias.copy_(0.1 * torch.randn(n, generator=g))
                m.running_mean.copy_(0.1 * torch.randn(n, generator=g))
                m.running_var.copy_(1.0 + 0.3 * torch.rand(n, generator=g))


def make_inputs(B, seed=SEED + 2, size=256):
    g = torch.Generator().manual_seed(seed)
    image = torch.rand((B, 3, size, size), generator=g) - 0.5
    root = torch.tensor([0.0, 0.0, 0.5]) + 0.05 * torch.randn((B, 3), generator=g)
    f = 217.5 * size / 256
    K = torch.tensor([[f, 0, size / 2.0], [0, f, size / 2.0], [0, 0, 1]]).repeat(B, 1, 1)
    corners = 0.2 * torch.rand((B, 8, 3), generator=g) - 0.1
    return {"image": image, "root_joint": root, "cam_intr": K, "corners_can": corners}
