#!/usr/bin/env python
"""GPU kernels of one ArtiBoostLoop.make_batch() (48 synthetic + 80 real-shaped samples) and of the loop's per-step work
outside the captured graph (CCV feedback), grouped by name.  usage (GPU box): python tools/prof_make_batch.py"""
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
from torch.profiler import ProfilerActivity, profile  # noqa: E402

dev = torch.device("cuda", 0)
arch, preset = netcfg.arch_cfg("ResNet34")
torch.manual_seed(1)
model = M.Arch({"ARCH": arch}, M.build_arch_model_list(arch, preset_cfg=preset)).to(dev)
pipe = SynthPipeline(device=dev, seed=11)
loop = ArtiBoostLoop(model, pipe, batch_size=128, generator=torch.Generator(device=dev).manual_seed(100), use_graph=True)
for _ in range(6):
    loop.step()
torch.cuda.synchronize()


def show(title, fn):
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        fn()
        torch.cuda.synchronize()
    ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for e in ev:
        a = agg[e.name[:100]]
        a[0] += 1
        a[1] += e.device_time
    print(f"== {title}: {len(ev)} kernels, {sum(v[1] for v in agg.values()):.1f} us of device time")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(os.environ.get("TOP", 25))]:
        print(f"{v[1]:9.1f} us {v[0]:4d} x {v[1] / v[0]:7.1f}  {k}")


show("make_batch", loop.make_batch)
b = loop.make_batch()
torch.cuda.synchronize()
preds = None


def feed():
    targ = b["corners_3d"] + b["root_joint"].unsqueeze(1)
    loop.feedback.feed(targ.clone(), targ, b["obj_id"], b["persp_id"], b["grasp_id"], b["is_synth"])


show("targets + CCVFeedback.feed", feed)
loop.close()
