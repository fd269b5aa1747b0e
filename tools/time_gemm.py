#!/usr/bin/env python
"""Device time of plain / implicit GEMM shapes of the clasbased networks at batch 128 (eval-mode epilogue, bf16 out).
usage (GPU box): AB_GEMM_PERSISTENT=0|1 python tools/time_gemm.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from artiboost_b200.models import nhwc  # noqa: E402

dev = torch.device("cuda", 0)
B = int(os.environ.get("B", 128))
print("AB_GEMM_PERSISTENT =", os.environ.get("AB_GEMM_PERSISTENT", "1"))
tot = 0.0
# (Cin, Cout, k, stride, H): ResNet-50 1x1 / 3x3 layers, ResNet-34 3x3 layers past layer1, the final 1x1 of the head
for C, cout, k, s, hw in ((64, 64, 1, 1, 64), (64, 256, 1, 1, 64), (256, 64, 1, 1, 64), (256, 128, 1, 1, 64), (128, 512, 1, 1, 32),
                          (512, 128, 1, 1, 32), (512, 256, 1, 1, 32), (256, 1024, 1, 1, 16), (1024, 256, 1, 1, 16), (1024, 512, 1, 1, 16),
                          (512, 2048, 1, 1, 8), (2048, 512, 1, 1, 8), (128, 128, 3, 1, 32), (128, 128, 3, 2, 64), (256, 256, 3, 1, 16),
                          (512, 512, 3, 1, 8), (256, 616, 1, 1, 32)):
    conv = torch.nn.Conv2d(C, cout, k, s, k // 2, bias=False).to(dev)
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
    ho = hw // s
    fl = 2.0 * B * ho * ho * cout * k * k * C
    by = 2.0 * B * (hw * hw * C + ho * ho * cout)
    tot += us
    print(f"  {k}x{k}/{s} C={C:4d} Cout={cout:4d} {hw:3d}x{hw:<3d}: {us:7.1f} us  {fl / us / 1e6:7.1f} TFLOP/s  {by / us / 1e6:5.2f} TB/s in+out")
print(f"  sum {tot:.1f} us")
