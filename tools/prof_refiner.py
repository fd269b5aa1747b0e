#!/usr/bin/env python
"""The pose generator with the hand_obj refiner at the bench's size (batch 512, 3 iterations), for ncu captures of
chamfer_nn_grouped_kernel / linear_f32_kernel on real hand-vertex geometry."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from artiboost_b200.synth import DEFAULT_CFG, SynthPipeline  # noqa: E402

cfg = dict(DEFAULT_CFG, SCRAMBLER=dict(DEFAULT_CFG["SCRAMBLER"], TYPE="random_2"),
           REFINER={"TYPE": "hand_obj", "PRETRAINED": None, "ITERS": 3})
pipe = SynthPipeline(device="cuda:0", seed=5, cfg=cfg, n_hand_tex=2, n_bg=2)
for _ in range(2):
    pipe.sample_poses(512)
torch.cuda.synchronize()
