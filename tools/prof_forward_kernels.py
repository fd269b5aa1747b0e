#!/usr/bin/env python
"""Every GPU kernel of one eval-mode forward pass (ResNet34 clasbased network, batch 128), grouped by name."""
import collections
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import netcfg  # noqa: E402

import artiboost_b200.models as M  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

dev = torch.device("cuda", 0)
arch, preset = netcfg.arch_cfg(os.environ.get("BACKBONE", "ResNet34"))
torch.manual_seed(1)
model = M.Arch({"ARCH": arch}, M.build_arch_model_list(arch, preset_cfg=preset)).to(dev).eval()
B = 128
from artiboost_b200.train import real_shaped_batch  # noqa: E402
batch = real_shaped_batch(B, dev, torch.Generator(device=dev).manual_seed(3))
with torch.no_grad():
    for _ in range(3):
        model(batch)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        model(batch)
        torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
agg = collections.defaultdict(lambda: [0, 0.0])
for e in ev:
    a = agg[e.name[:100]]
    a[0] += 1
    a[1] += e.device_time
ours = sum(v[0] for k, v in agg.items() if "ab::" in k)
print(f"kernels {len(ev)} (ours {ours}, torch-native {len(ev) - ours}), total {sum(v[1] for v in agg.values()) / 1e3:.2f} ms")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{v[1]:9.1f} us {v[0]:4d} x {v[1] / v[0]:7.1f}  {k}")
