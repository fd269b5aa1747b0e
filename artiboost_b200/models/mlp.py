"""MLP_O box-rotation head (anakin/models/mlp.py:10-25) on the tensor-core GEMM (bias + ReLU fused)."""
import torch
import torch.nn as nn

from . import nhwc, train_ops
from .registry import MODEL


@MODEL.register_module
class MLP_O(nn.Module):

    def __init__(self, **cfg):
        super().__init__()
        layers_n = cfg["LAYERS_N"]
        out_channel = cfg["OUT_CHANNEL"]
        layers = nn.ModuleList()
        for (in_n, out_n) in zip(layers_n[:-1], layers_n[1:]):
            layers.append(nn.Linear(in_n, out_n))
            layers.append(nn.ReLU())
        layers.append(nn.Linear(layers_n[-1], out_channel))
        self.layers = nn.Sequential(*layers)

    def forward(self, x):
        fcs = [m for m in self.layers if isinstance(m, nn.Linear)]
        if torch.is_grad_enabled():
            h = x
            for fc in fcs[:-1]:
                h = train_ops.linear(h, fc, relu=True)
            return train_ops.linear(h, fcs[-1], relu=False, out_fp32=True)
        h = x.to(torch.bfloat16).contiguous()
        for fc in fcs[:-1]:
            h = nhwc.linear(h, fc, relu=True).contiguous()
        return nhwc.linear(h, fcs[-1], relu=False, out_fp32=True)
