"""IntegralDeconvHead (anakin/models/simplebaseline.py:74-190): same cfg keys, parameter names and outputs.
deconv x N -> 1x1 conv -> softmax over D*H*W per class -> confidence, re-normalise, soft-argmax; the last four steps
are ONE kernel over the fp32 logits (ab_head_decode)."""
from typing import Dict

import torch
import torch.nn as nn

from . import nhwc, train_ops
from .registry import HEAD, enable_lower_param


@HEAD.register_module
class IntegralDeconvHead(nn.Module):

    @enable_lower_param
    def __init__(self, **cfg):
        super().__init__()
        self.inplanes = cfg["INPUT_CHANNEL"]
        self.depth_res = cfg["DEPTH_RESOLUTION"]
        self.height_res = cfg["HEATMAP_SIZE"][1]
        self.width_res = cfg["HEATMAP_SIZE"][0]
        self.deconv_with_bias = cfg["DECONV_WITH_BIAS"]
        self.nclasses = cfg["NCLASSES"]
        self.norm_type = cfg["NORM_TYPE"]
        if self.norm_type != "softmax":
            raise NotImplementedError("NORM_TYPE 'softmax' is the only one the shipped configs use")
        self.deconv_layers = self._make_deconv_layer(cfg["NUM_DECONV_LAYERS"], cfg["NUM_DECONV_FILTERS"],
                                                     cfg["NUM_DECONV_KERNELS"])
        self.final_layer = nn.Conv2d(in_channels=cfg["NUM_DECONV_FILTERS"][-1], out_channels=cfg["NCLASSES"] * self.depth_res,
                                     kernel_size=cfg["FINAL_CONV_KERNEL"], stride=1,
                                     padding=1 if cfg["FINAL_CONV_KERNEL"] == 3 else 0)
        self.init_weights()

    def init_weights(self):
        for m in self.deconv_layers.modules():
            if isinstance(m, nn.ConvTranspose2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
                if self.deconv_with_bias:
                    nn.init.constant_(m.bias, 0)
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)
        for m in self.final_layer.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
                nn.init.constant_(m.bias, 0)

    def _get_deconv_cfg(self, deconv_kernel, index):
        if deconv_kernel == 4:
            return 4, 1, 0
        if deconv_kernel == 3:
            return 3, 1, 1
        if deconv_kernel == 2:
            return 2, 0, 0
        raise ValueError()

    def _make_deconv_layer(self, num_layers, num_filters, num_kernels):
        assert num_layers == len(num_filters), "ERROR: num_deconv_layers is different len(num_deconv_filters)"
        assert num_layers == len(num_kernels), "ERROR: num_deconv_layers is different len(num_deconv_filters)"
        layers = []
        for i in range(num_layers):
            kernel, padding, output_padding = self._get_deconv_cfg(num_kernels[i], i)
            planes = num_filters[i]
            layers.append(nn.ConvTranspose2d(in_channels=self.inplanes, out_channels=planes, kernel_size=kernel, stride=2,
                                             padding=padding, output_padding=output_padding, bias=self.deconv_with_bias))
            layers.append(nn.BatchNorm2d(planes))
            layers.append(nn.ReLU(inplace=True))
            self.inplanes = planes
        return nn.Sequential(*layers)

    def forward_act(self, x: nhwc.Act) -> Dict[str, torch.Tensor]:
        ops_ = train_ops if torch.is_grad_enabled() else nhwc
        mods = list(self.deconv_layers)
        for i in range(0, len(mods), 3):
            x = ops_.deconv4x4s2_bn_relu(x, mods[i], mods[i + 1], training=self.training)
        if (x.H, x.W) != (self.height_res, self.width_res):
            # the reference's view_to_bcdhw would raise on this mismatch too (simplebaseline.py:120-135)
            raise RuntimeError(f"heatmap is {x.H}x{x.W} but HEATMAP_SIZE is {self.height_res}x{self.width_res}")
        fl = self.final_layer
        if ops_ is train_ops and fl.kernel_size == (1, 1) and fl.stride == (1, 1) and fl.padding == (0, 0):
            # one autograd node: the bf16 logit gradient goes from the decode's backward straight into the conv's
            kp3d, confd = train_ops.conv_head_decode(x, fl, self.nclasses, self.depth_res)
            return {"kp3d": kp3d, "kp3d_confd": confd}
        logits = ops_.conv_bn_act(x, fl, None, relu=False, out_fp32=True)  # [B*H*W, ncls*D], conv bias fused
        kp3d, confd = ops_.head_decode(logits, x.B, self.nclasses, self.depth_res, x.H, x.W)
        return {"kp3d": kp3d, "kp3d_confd": confd}

    def forward(self, **kwargs) -> Dict[str, torch.Tensor]:
        x = kwargs["feature"]
        if not isinstance(x, nhwc.Act):  # fp32 NCHW feature map, as the reference passes it
            B, C, H, W = x.shape
            x = nhwc.Act(x.permute(0, 2, 3, 1).reshape(B * H * W, C).to(torch.bfloat16).contiguous(), B, H, W, C)
        return self.forward_act(x)
