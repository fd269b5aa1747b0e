"""ResNet backbones (anakin/models/resnet.py:44-275): same classes, cfg keys, forward contract and state_dict names;
forward runs NHWC bf16 on the tensor-core GEMM with BN / ReLU / residual fused into the epilogues."""
from collections import OrderedDict
from typing import Dict

import torch
import torch.nn as nn

from . import nhwc, train_ops
from .registry import BACKBONE, enable_lower_param


class FrozenBatchNorm2d(nn.Module):
    """BatchNorm2d with fixed statistics and affine parameters (resnet.py:24-69)."""

    def __init__(self, n):
        super().__init__()
        self.register_buffer("weight", torch.ones(n))
        self.register_buffer("bias", torch.zeros(n))
        self.register_buffer("running_mean", torch.zeros(n))
        self.register_buffer("running_var", torch.ones(n))

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
        num_batches_tracked_key = prefix + "num_batches_tracked"
        if num_batches_tracked_key in state_dict:
            del state_dict[num_batches_tracked_key]
        super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs)


def _prims():
    """Inference primitives under torch.no_grad(), differentiable ones (forward + backward kernels) otherwise."""
    return train_ops if torch.is_grad_enabled() else nhwc


def conv3x3(in_planes, out_planes, stride=1):
    return nn.Conv2d(in_planes, out_planes, kernel_size=3, stride=stride, padding=1, bias=False)


class BasicBlock(nn.Module):
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None, bn_layer=nn.BatchNorm2d):
        super().__init__()
        self.conv1 = conv3x3(inplanes, planes, stride)
        self.bn1 = bn_layer(planes)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = conv3x3(planes, planes)
        self.bn2 = bn_layer(planes)
        self.downsample = downsample
        self.stride = stride

    def forward_act(self, x: nhwc.Act) -> nhwc.Act:
        tr, ops_ = self.training, _prims()
        residual = x if self.downsample is None else ops_.conv_bn_act(x, self.downsample[0], self.downsample[1], training=tr)
        out = ops_.conv_bn_act(x, self.conv1, self.bn1, relu=True, training=tr)
        return ops_.conv_bn_act(out, self.conv2, self.bn2, relu=True, residual=residual, training=tr)


class Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, downsample=None, bn_layer=nn.BatchNorm2d):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, kernel_size=1, bias=False)
        self.bn1 = bn_layer(planes)
        self.conv2 = nn.Conv2d(planes, planes, kernel_size=3, stride=stride, padding=1, bias=False)
        self.bn2 = bn_layer(planes)
        self.conv3 = nn.Conv2d(planes, planes * self.expansion, kernel_size=1, bias=False)
        self.bn3 = bn_layer(planes * self.expansion)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample
        self.stride = stride

    def forward_act(self, x: nhwc.Act) -> nhwc.Act:
        tr, ops_ = self.training, _prims()
        residual = x if self.downsample is None else ops_.conv_bn_act(x, self.downsample[0], self.downsample[1], training=tr)
        out = ops_.conv_bn_act(x, self.conv1, self.bn1, relu=True, training=tr)
        out = ops_.conv_bn_act(out, self.conv2, self.bn2, relu=True, training=tr)
        return ops_.conv_bn_act(out, self.conv3, self.bn3, relu=True, residual=residual, training=tr)


class ResNet(nn.Module):

    def __init__(self, block, layers, num_classes=1000, interm_feat=True, **kwargs):
        super().__init__()
        self.bn_layer = FrozenBatchNorm2d if kwargs["FREEZE_BATCHNORM"] else nn.BatchNorm2d
        self.inplanes = 64
        self.interm_feat = kwargs["INTERM_FEAT"] if "INTERM_FEAT" in kwargs else interm_feat
        self.conv1 = nn.Conv2d(3, 64, kernel_size=7, stride=2, padding=3, bias=False)
        self.bn1 = self.bn_layer(64)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(kernel_size=3, stride=2, padding=1)
        self.layer1 = self._make_layer(block, 64, layers[0])
        self.layer2 = self._make_layer(block, 128, layers[1], stride=2)
        self.layer3 = self._make_layer(block, 256, layers[2], stride=2)
        self.layer4 = self._make_layer(block, 512, layers[3], stride=2)
        self.avgpool = nn.AdaptiveAvgPool2d((1, 1))
        self.fc = nn.Linear(512 * block.expansion, num_classes)  # unused ImageNet head, kept for checkpoint names
        self.features = 512 * block.expansion
        self.output_channel = self.inplanes
        self.init_weights()

    def init_weights(self):
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            elif isinstance(m, self.bn_layer):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)

    def _make_layer(self, block, planes, blocks, stride=1):
        downsample = None
        if stride != 1 or self.inplanes != planes * block.expansion:
            downsample = nn.Sequential(
                nn.Conv2d(self.inplanes, planes * block.expansion, kernel_size=1, stride=stride, bias=False),
                self.bn_layer(planes * block.expansion),
            )
        layers = [block(self.inplanes, planes, stride, downsample, self.bn_layer)]
        self.inplanes = planes * block.expansion
        for _ in range(1, blocks):
            layers.append(block(self.inplanes, planes))
        return nn.Sequential(*layers)

    # file names of the torchvision ImageNet checkpoints the reference fetches (anakin/models/resnet.py:14-20)
    PRETRAINED_FILES = {"ResNet18": "resnet18-5c106cde.pth", "ResNet34": "resnet34-333f7ec4.pth",
                        "ResNet50": "resnet50-19c8e357.pth", "ResNet101": "resnet101-5d3b4d8f.pth",
                        "ResNet152": "resnet152-b121ed2d.pth"}

    def load_pretrained(self, source=True):
        """PRETRAINED: true / "<path>" (anakin/models/resnet.py:194-197).  The reference downloads the torchvision
        checkpoint through model_zoo; this build has no network, so `true` looks where model_zoo would have cached the
        file ($ARTIBOOST_PRETRAINED_DIR, then $TORCH_HOME/hub/checkpoints) and a string is taken as the path of a
        state_dict file.  Missing files raise: silently training from scratch would not be the configured run."""
        import os
        if isinstance(source, (str, os.PathLike)):
            path = os.fspath(source)
        else:
            name = self.PRETRAINED_FILES[type(self).__name__]
            dirs = [os.environ.get("ARTIBOOST_PRETRAINED_DIR"), os.path.join(torch.hub.get_dir(), "checkpoints")]
            path = next((os.path.join(d, name) for d in dirs if d and os.path.exists(os.path.join(d, name))), None)
            if path is None:
                raise FileNotFoundError(f"PRETRAINED: true needs {name} in $ARTIBOOST_PRETRAINED_DIR or "
                                        f"{dirs[1]} (no network to download it); or set PRETRAINED to a state_dict path")
        state = torch.load(path, map_location="cpu")
        state = state.get("state_dict", state) if isinstance(state, dict) else state
        self.load_state_dict(state)
        nhwc.bump_params()  # packed bf16 filter copies are stale

    def forward_acts(self, image: torch.Tensor) -> Dict[str, object]:
        """NHWC bf16 feature maps (what the head consumes without a layout round trip)."""
        ops_ = _prims()
        x = nhwc.image_to_act(image)
        if self.training and ops_ is not nhwc and ops_.stem_fusable(self.conv1, self.bn1):
            x = ops_.stem_conv_bn_relu_maxpool(x, self.conv1, self.bn1)   # bn1 + ReLU evaluated inside the pooling kernel
        else:
            x = ops_.conv_bn_act(x, self.conv1, self.bn1, relu=True, training=self.training)
            x = ops_.maxpool3x3s2(x)
        feats = OrderedDict()
        for name in ("layer1", "layer2", "layer3", "layer4"):
            for blk in getattr(self, name):
                x = blk.forward_act(x)
            feats["res_" + name] = x
        if ops_ is nhwc:
            feats["res_layer4_mean"], feats["res_layer4_mean_bf16"] = nhwc.avgpool(x)
        else:
            feats["res_layer4_mean"] = train_ops.avgpool(x)
            feats["res_layer4_mean_bf16"] = feats["res_layer4_mean"]  # LinearFn casts (and routes the gradient) itself
        return feats

    def forward(self, **kwargs) -> Dict:
        """-> OrderedDict{res_layer1..4 fp32 NCHW, res_layer4_mean fp32 [B, C]} (resnet.py:199-224)."""
        feats = self.forward_acts(kwargs["image"])
        features = OrderedDict()
        for k in ("res_layer1", "res_layer2", "res_layer3", "res_layer4"):
            features[k] = feats[k].nchw()
        features["res_layer4_mean"] = feats["res_layer4_mean"]
        if self.interm_feat:
            return features
        out = {"res_output": nhwc.linear(feats["res_layer4_mean_bf16"], self.fc, out_fp32=True)}
        out.update(features)
        return out


@BACKBONE.register_module
class ResNet18(ResNet):

    @enable_lower_param
    def __init__(self, **cfg):
        super().__init__(BasicBlock, [2, 2, 2, 2], **cfg)
        if cfg["PRETRAINED"]:
            self.load_pretrained(cfg["PRETRAINED"])


@BACKBONE.register_module
class ResNet34(ResNet):

    @enable_lower_param
    def __init__(self, **cfg):
        super().__init__(BasicBlock, [3, 4, 6, 3], **cfg)
        if cfg["PRETRAINED"]:
            self.load_pretrained(cfg["PRETRAINED"])


@BACKBONE.register_module
class ResNet50(ResNet):

    @enable_lower_param
    def __init__(self, **cfg):
        super().__init__(Bottleneck, [3, 4, 6, 3], **cfg)
        if cfg["PRETRAINED"]:
            self.load_pretrained(cfg["PRETRAINED"])


@BACKBONE.register_module
class ResNet101(ResNet):

    @enable_lower_param
    def __init__(self, **cfg):
        super().__init__(Bottleneck, [3, 4, 23, 3], **cfg)
        if cfg["PRETRAINED"]:
            self.load_pretrained(cfg["PRETRAINED"])


@BACKBONE.register_module
class ResNet152(ResNet):

    @enable_lower_param
    def __init__(self, **cfg):
        super().__init__(Bottleneck, [3, 8, 36, 3], **cfg)
        if cfg["PRETRAINED"]:
            self.load_pretrained(cfg["PRETRAINED"])
