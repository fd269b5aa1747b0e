#!/usr/bin/env python
"""Device time of the 3x3 / stride 1 convolutions of ResNet-34 at batch 128 (forward, eval-mode epilogue).
usage (GPU box): AB_CONV_HALO=0|1 python tools/time_conv.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from artiboost_b200.models import nhwc  # noqa: E402

dev = torch.device("cuda", 0)
B = int(os.environ.get("B", 128))
print("AB_CONV_HALO =", os.environ.get("AB_CONV_HALO", "1"))
for C, cout, hw in ((64, 64, 64), (128, 128, 32), (256, 256, 16), (512, 512, 8), (64, 128, 64), (256, 64, 32)):
    conv = torch.nn.Conv2d(C, cout, 3, 1, 1, bias=False).to(dev)
    x = nhwc.Act(torch.randn((B * hw * hw, C), device=dev).to(torch.bfloat16), B, hw, hw, C)
    with torch.no_grad():
        for _ in range(5):
            nhwc.conv_bn_act(x, conv, None, relu=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(30):
            nhwc.conv_bn_act(x, conv, None, relu=True)
        e1.record()
        torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 30 * 1e3
    fl = 2.0 * B * hw * hw * cout * 9 * C
    print(f"  C={C:4d} Cout={cout:4d} {hw:3d}x{hw:<3d}: {us:7.1f} us  {fl / us / 1e6:7.1f} TFLOP/s")
