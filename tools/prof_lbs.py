"""Runs the fused MANO LBS launch (ab_mano_forward through the pose generator) at batch 512 a few times: the target of
`ncu -k regex:mano_lbs` for profiles/r2_mano_lbs.*"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from artiboost_b200 import lib  # noqa: E402
from artiboost_b200.synth import SynthPipeline  # noqa: E402

pipe = SynthPipeline(device="cuda:0", seed=1, n_hand_tex=2, n_bg=2)
for _ in range(int(os.environ.get("N", 10))):
    pipe.sample_poses(512)
torch.cuda.synchronize()
lib.profile_enable(True)
for _ in range(20):
    pipe.sample_poses(512)
torch.cuda.synchronize()
lib.profile_enable(False)
for k, v in lib.profile_collect().items():
    print(f"{k}: {v[0] / v[1] * 1e3:.1f} us per launch ({v[1]} launches)")
