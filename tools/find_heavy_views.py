#!/usr/bin/env python
"""Which views make a batch slow?  Times each resident batch of the bench (same construction), then looks at the slowest:
covered pixels and nearest depth per view, and the batch's time with its heaviest views replaced by an ordinary one."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from artiboost_b200.synth import SynthPipeline  # noqa: E402

dev = torch.device("cuda", 0)
B = 512
pipe = SynthPipeline(device=dev, seed=1, sample_seed=1 + int(os.environ.get("AB_BENCH_SEED_OFFSET", "3")), chunk=B)
res = [(pipe.sample_poses(B), pipe.draw_render_randoms(B)) for _ in range(10)]
out = {"rgba": torch.empty((B, 256, 256, 4), dtype=torch.uint8, device=dev),
       "depth": torch.empty((B, 256, 256), dtype=torch.float32, device=dev),
       "seg": torch.empty((B, 256, 256), dtype=torch.uint8, device=dev)}


def timed(p, r, n=10):
    for _ in range(2):
        pipe.render(p, r, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        pipe.render(p, r, out=out)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


ms = [timed(p, r) for p, r in res]
print("ms per batch:", [round(m, 3) for m in ms])
k = max(range(10), key=lambda i: ms[i])
p, r = res[k]
pipe.render(p, r, out=out)
torch.cuda.synchronize()
cov = (out["seg"] > 0).flatten(1).sum(1)
hand = (out["seg"] == 1).flatten(1).sum(1)
d = out["depth"].flatten(1)
dmin = torch.where(d > 0, d, torch.full_like(d, 9.0)).min(1).values
order = torch.argsort(cov, descending=True)
print("slowest batch", k, "covered px: median", int(cov.median()), "top:", [(int(i), int(cov[i]), int(hand[i]), round(float(dmin[i]), 3)) for i in order[:8]])
hv = p["final_hand_verts"]
print("hand z min over batch:", float(hv[..., 2].min()), " object z:", float(p["final_obj_pose"][:, 2, 3].min()), float(p["final_obj_pose"][:, 2, 3].max()))
for n_rep in (1, 2, 4, 8):
    q = {kk: (v.clone() if torch.is_tensor(v) else v) for kk, v in p.items()}
    rr = {kk: v.clone() for kk, v in r.items()}
    normal = int(order[B // 2])
    for i in order[:n_rep].tolist():
        for kk in ("final_hand_verts", "final_obj_pose", "obj_id"):
            q[kk][i] = q[kk][normal]
    print(f"  heaviest {n_rep} views replaced by a median view: {timed(q, rr):.3f} ms")
