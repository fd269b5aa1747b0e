"""Timeline of synthesise(prefetch=True): when do the pose kernels of batch k+1 run relative to the rasteriser of batch k?
Events are recorded on both streams by wrapping sample_poses / render; times are relative to the first event."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from artiboost_b200.synth import SynthPipeline  # noqa: E402

B = int(os.environ.get("B", 512))
dev = torch.device("cuda", 0)
pipe = SynthPipeline(device=dev, seed=1, chunk=B)
out = {"rgba": torch.empty((B, 256, 256, 4), dtype=torch.uint8, device=dev),
       "depth": torch.empty((B, 256, 256), dtype=torch.float32, device=dev),
       "seg": torch.empty((B, 256, 256), dtype=torch.uint8, device=dev)}
marks = []
orig_sp, orig_r = pipe.sample_poses, pipe.render


def ev():
    e = torch.cuda.Event(enable_timing=True)
    e.record(torch.cuda.current_stream(dev))
    return e


def sp(n):
    a = ev(); r = orig_sp(n); b = ev()
    marks.append(("poses", a, b, time.perf_counter()))
    return r


def rd(*a, **k):
    x = ev(); r = orig_r(*a, **k); y = ev()
    marks.append(("raster", x, y, time.perf_counter()))
    return r


for pf in (False, True):
    pipe.drop_prefetch()
    pipe.sample_poses, pipe.render = orig_sp, orig_r
    for _ in range(5):
        pipe.synthesise(B, out=out, prefetch=pf)
    torch.cuda.synchronize()
    pipe.sample_poses, pipe.render = sp, rd
    marks.clear()
    t0 = time.perf_counter()
    base = ev()
    for _ in range(6):
        pipe.synthesise(B, out=out, prefetch=pf)
    th = time.perf_counter() - t0
    torch.cuda.synchronize()
    print(f"prefetch={pf}: host {th / 6 * 1e3:.3f} ms per call")
    for name, a, b, tm in marks:
        print(f"  {name:7s} {base.elapsed_time(a) * 1e3:8.1f} -> {base.elapsed_time(b) * 1e3:8.1f} us   (host issued at {(tm - t0) * 1e6:8.1f} us)")

# the same loops without any event inside (what bench.py times)
pipe.sample_poses, pipe.render = orig_sp, orig_r
for pf in (False, True, False, True):
    pipe.drop_prefetch()
    for _ in range(5):
        pipe.synthesise(B, out=out, prefetch=pf)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(40):
        pipe.synthesise(B, out=out, prefetch=pf)
    e1.record()
    torch.cuda.synchronize()
    print(f"no marks, prefetch={pf}: {e0.elapsed_time(e1) / 40:.4f} ms per batch")
