"""Profiling driver: a few full train steps (synth -> batch -> fwd -> bwd -> all-reduce -> Adam) for ncu launch lists.
  ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/x.csv python tools/prof_train.py ResNet34 128 2
"""
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

backbone = sys.argv[1] if len(sys.argv) > 1 else "ResNet34"
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 128
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
dev = torch.device("cuda", 0)
arch, preset = netcfg.arch_cfg(backbone)
torch.manual_seed(1)
model = M.Arch({"ARCH": arch}, M.build_arch_model_list(arch, preset_cfg=preset)).to(dev)
pipe = SynthPipeline(device=dev, seed=11)
loop = ArtiBoostLoop(model, pipe, batch_size=batch, generator=torch.Generator(device=dev).manual_seed(100))
for _ in range(steps):
    loss = loop.step()
torch.cuda.synchronize()
print("loss", float(loss))
