#!/usr/bin/env python
"""Device time of the 3x3 weight gradients of ResNet-34 at batch 128 (ab_conv_wgrad_bf16_nhwc: split kernel + reduce).
usage (GPU box): python tools/time_wgrad.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from artiboost_b200 import lib  # noqa: E402

dev = torch.device("cuda", 0)
L = lib.load()
B = 128
tot = 0.0
for C, cout, hw, n_layers in ((64, 64, 64, 6), (128, 128, 32, 7), (256, 256, 16, 11), (512, 512, 8, 5)):
    P = B * hw * hw
    x = [torch.randn(P, C, device=dev).bfloat16() for _ in range(2)]
    dy = [torch.randn(P, cout, device=dev).bfloat16() for _ in range(2)]
    dw = torch.zeros(cout, C, 3, 3, device=dev)
    ws = torch.empty(int(L.ab_wgrad_workspace_bytes(P, cout, 9 * C)) // 4, device=dev)

    def run(i):
        lib.check(L.ab_conv_wgrad_bf16_nhwc(x[i % 2].data_ptr(), B, hw, hw, C, dy[i % 2].data_ptr(), cout, 3, 3, 1, 1, dw.data_ptr(), 1,
                                            ws.data_ptr(), lib.stream_ptr(dev)), "wgrad")
    for i in range(4):
        run(i)
    torch.cuda.synchronize()
    lib.profile_enable(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(20):
        run(i)
    e1.record()
    torch.cuda.synchronize()
    lib.profile_enable(False)
    st = lib.profile_collect()
    us = e0.elapsed_time(e1) / 20 * 1e3
    fl = 2.0 * P * cout * 9 * C
    tot += us * n_layers
    print(f"  C={C:4d} Cout={cout:4d} {hw:3d}x{hw:<3d}: {us:7.1f} us (split + reduce)  {fl / us / 1e6:7.1f} TFLOP/s   x{n_layers} layers; workspace {ws.numel() * 4 / 1e6:.1f} MB",
          {k: round(v[0] / v[1] * 1e3, 1) for k, v in st.items()})
print(f"  ResNet-34 3x3 weight gradients: {tot:.0f} us per step")
