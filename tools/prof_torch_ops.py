#!/usr/bin/env python
"""Which torch-native (non-library) kernels run inside one eager training step, grouped by the aten op that launched them
and by the Python source line.  usage (GPU box): python tools/prof_torch_ops.py > gpurun_out/torch_ops.txt"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import netcfg  # noqa: E402

import artiboost_b200.models as M  # noqa: E402
from artiboost_b200.synth import SynthPipeline  # noqa: E402
from artiboost_b200.train import ArtiBoostLoop  # noqa: E402

dev = torch.device("cuda", 0)
arch, preset = netcfg.arch_cfg("ResNet34")
torch.manual_seed(1)
model = M.Arch({"ARCH": arch}, M.build_arch_model_list(arch, preset_cfg=preset)).to(dev)
pipe = SynthPipeline(device=dev, seed=11)
loop = ArtiBoostLoop(model, pipe, batch_size=128, generator=torch.Generator(device=dev).manual_seed(100), use_graph=False)
fixed = loop.make_batch()
for _ in range(3):
    loop.step(fixed)
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile  # noqa: E402

with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True, record_shapes=True) as prof:
    loop.step(fixed)
    torch.cuda.synchronize()
print(prof.key_averages(group_by_stack_n=6).table(sort_by="cuda_time_total", row_limit=70, max_name_column_width=60,
                                                   max_src_column_width=110))
