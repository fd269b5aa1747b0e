#!/usr/bin/env python
"""Stem contraction ([B*128*128, 200] im2col rows x [64, 200] filters) and the head's final 1x1 with fp32 logits.
usage (GPU box): AB_GEMM_PERSISTENT=0|1 python tools/time_stem_gemm.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from artiboost_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)
print("AB_GEMM_PERSISTENT =", os.environ.get("AB_GEMM_PERSISTENT", "1"))
for name, M, N, K, f32 in (("stem", 128 * 128 * 128, 64, 200, False), ("final 1x1", 128 * 32 * 32, 616, 256, True),
                           ("deconv1", 128 * 8 * 8, 4096, 512, False), ("deconv2", 128 * 16 * 16, 4096, 256, False)):
    a = [torch.randn(M, K, device=dev).bfloat16() for _ in range(2)]
    b = torch.randn(N, K, device=dev).bfloat16()
    for i in range(4):
        ops.gemm_bf16(a[i % 2], b, out_fp32=f32)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(20):
        ops.gemm_bf16(a[i % 2], b, out_fp32=f32)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    by = M * K * 2 + M * N * (4 if f32 else 2)
    print(f"  {name:10s} M={M} N={N} K={K}: {us:7.1f} us  {2.0 * M * N * K / us / 1e6:6.1f} TFLOP/s  {by / us / 1e6:5.2f} TB/s")
