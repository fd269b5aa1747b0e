#!/usr/bin/env python
"""Device time of the captured training step alone (fixed batch, CUDA-graph replay): A/B of step-level switches on one box.
usage (GPU box): [AB_BN_MASK_FROM_RAW=0] [AB_FUSED_TAIL=0] python tools/time_train_step.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import netcfg  # noqa: E402

import artiboost_b200.models as M  # noqa: E402
from artiboost_b200 import lib  # noqa: E402
from artiboost_b200.train import TrainStep, real_shaped_batch  # noqa: E402

dev = torch.device("cuda", 0)
B = int(os.environ.get("B", 128))
arch, preset = netcfg.arch_cfg(os.environ.get("BACKBONE", "ResNet34"))
torch.manual_seed(1)
model = M.Arch({"ARCH": arch}, M.build_arch_model_list(arch, preset_cfg=preset)).to(dev)
gen = torch.Generator(device=dev).manual_seed(3)
step = TrainStep(model, generator=gen, use_graph=True, graph_warmup=2)
batch = real_shaped_batch(B, dev, gen)
for _ in range(6):
    step(batch)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = int(os.environ.get("N", 30))
e0.record()
for _ in range(n):
    step(batch)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print(f"train step (graph, B={B}): {ms:.3f} ms  {B / ms * 1e3:.0f} img/s   switches: " +
      " ".join(f"{k}={os.environ.get(k, '1')}" for k in ("AB_BN_MASK_FROM_RAW", "AB_FUSED_TAIL")))
if os.environ.get("STAGES"):
    lib.profile_enable(True)
    for _ in range(3):
        step._eager(batch)
    torch.cuda.synchronize()
    lib.profile_enable(False)
    for k, v in lib.profile_collect().items():
        print(f"  {k}: {v[0] / 3:.3f} ms per step ({v[1] // 3} launches)")
step.close()
