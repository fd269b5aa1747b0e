"""Quick device timing of ab_render_batch on the bench workload (batch 512, 256x256): eager and CUDA-graph replay."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from artiboost_b200 import lib  # noqa: E402
from artiboost_b200.synth import SynthPipeline  # noqa: E402

B = int(os.environ.get("B", 512))
dev = torch.device("cuda", 0)
pipe = SynthPipeline(device=dev, seed=1, chunk=B)
poses = pipe.sample_poses(B)
rand = pipe.draw_render_randoms(B)
out = {"rgba": torch.empty((B, 256, 256, 4), dtype=torch.uint8, device=dev),
       "depth": torch.empty((B, 256, 256), dtype=torch.float32, device=dev),
       "seg": torch.empty((B, 256, 256), dtype=torch.uint8, device=dev)}


def step():
    pipe.render(poses, rand, out=out)


def timed(fn, n):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, (t1 - t0) / n * 1e3


N = int(os.environ.get("N", 200))
ms, host = timed(step, N)
print(f"eager: {ms:.4f} ms/step ({B / ms * 1e3:.0f} views/s), host submit {host:.4f} ms/step")
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    step()
    torch.cuda.synchronize()
    with torch.cuda.graph(g, stream=s):
        step()
ms, host = timed(g.replay, N)
print(f"graph: {ms:.4f} ms/step ({B / ms * 1e3:.0f} views/s), host submit {host:.4f} ms/step")
lib.profile_enable(True)
for _ in range(20):
    step()
torch.cuda.synchronize()
lib.profile_enable(False)
for k, v in lib.profile_collect().items():
    print(f"  {k}: {v[0] / v[1] * 1e3:.1f} us per launch ({v[1]} launches)")
seg = out["seg"]
print("covered fraction", float((seg > 0).float().mean()), "hand", float((seg == 1).float().mean()), "obj", float((seg == 2).float().mean()))
