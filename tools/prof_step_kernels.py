#!/usr/bin/env python
"""Every GPU kernel of one eager training step (fixed batch), grouped by name: count, total and mean time; then the aten ops
that launched non-library kernels with the Python line they came from.  usage (GPU box): python tools/prof_step_kernels.py"""
import collections
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
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
agg = collections.defaultdict(lambda: [0, 0.0])
for e in ev:
    a = agg[e.name[:90]]
    a[0] += 1
    a[1] += e.device_time
ours = sum(v[0] for k, v in agg.items() if "ab::" in k)
print(f"kernels {len(ev)} (ours {ours}), total {sum(v[1] for v in agg.values()) / 1e3:.2f} ms")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:60]:
    print(f"{v[1]:9.1f} us {v[0]:4d} x {v[1] / v[0]:7.1f}  {k}")
print("---- who launches the torch-native copy / add kernels (op <- parents, input shapes)")
seen = collections.Counter()
for e in prof.events():
    if e.device_type != torch.autograd.DeviceType.CPU or not e.kernels:
        continue
    if not any(("copy" in k.name or "CUDAFunctor_add<c10::BFloat" in k.name) for k in e.kernels):
        continue
    chain, p_ = [], e.cpu_parent
    while p_ is not None and len(chain) < 4:
        chain.append(p_.name[:40])
        p_ = p_.cpu_parent
    us = sum(k.duration for k in e.kernels)
    seen[(e.name, " <- ".join(chain), str(e.input_shapes)[:80])] += us
for (name, chain, shapes), us in seen.most_common(25):
    print(f"{us:8.1f} us  {name} <- {chain}  {shapes}")
print("---- aten ops by source line (non-library kernels)")
print(prof.key_averages(group_by_stack_n=4).table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=40,
                                                   max_src_column_width=90))
