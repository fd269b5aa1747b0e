"""Per-CTA timeline of raster_tile_kernel on the bench workload.  Needs a trace build:
`AB_EXTRA_NVCC_FLAGS=-DAB_RASTER_TRACE python -m artiboost_b200.build --force` (debug only; rebuild without it afterwards)."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from artiboost_b200 import lib  # noqa: E402
from artiboost_b200.synth import SynthPipeline  # noqa: E402

B = int(os.environ.get("B", 512))
dev = torch.device("cuda", 0)
pipe = SynthPipeline(device=dev, seed=1, sample_seed=int(os.environ.get("SAMPLE_SEED", "1")), chunk=B)
poses = pipe.sample_poses(B)
rand = pipe.draw_render_randoms(B)
out = {"rgba": torch.empty((B, 256, 256, 4), dtype=torch.uint8, device=dev),
       "depth": torch.empty((B, 256, 256), dtype=torch.float32, device=dev),
       "seg": torch.empty((B, 256, 256), dtype=torch.uint8, device=dev)}
for _ in range(10):
    pipe.render(poses, rand, out=out)
torch.cuda.synchronize()
n = B * 16
buf = np.zeros((n, 6), dtype=np.uint64)
L = lib.load()
L.ab_debug_raster_trace.restype = C.c_int
rc = L.ab_debug_raster_trace(buf.ctypes.data_as(C.c_void_p), n)
assert rc == 0, rc
t0 = buf[:, 0].min()
start = (buf[:, 0] - t0).astype(np.float64) / 1e3
end = (buf[:, 1] - t0).astype(np.float64) / 1e3
patch_end = (buf[:, 5] - t0).astype(np.float64) / 1e3
sm = (buf[:, 2] & np.uint64(0xffffffff)).astype(int)
nbig = (buf[:, 2] >> np.uint64(32)).astype(int)   # triangles that took the >= 64 px (int64) path in this tile
cnt = buf[:, 3].astype(int)
hit = buf[:, 4].astype(int)
dur = end - start
print("kernel span us", end.max(), "CTAs", n, "busy", int((cnt > 0).sum()))
busy = cnt > 0
print("empty CTA dur us: mean %.2f p50 %.2f p99 %.2f" % (dur[~busy].mean(), np.median(dur[~busy]), np.percentile(dur[~busy], 99)))
print("busy CTA dur us: mean %.1f p50 %.1f p90 %.1f p99 %.1f max %.1f" % (dur[busy].mean(), np.median(dur[busy]), np.percentile(dur[busy], 90), np.percentile(dur[busy], 99), dur[busy].max()))
print("busy patch-loop part: mean %.1f; shade part mean %.1f" % ((patch_end - start)[busy].mean(), (end - patch_end)[busy].mean()))
print("count: mean %.1f p50 %d p90 %d max %d ; hits mean %.0f max %d" % (cnt[busy].mean(), np.median(cnt[busy]), np.percentile(cnt[busy], 90), cnt.max(), hit[busy].mean(), hit.max()))
# regression of duration on count and hits
A = np.stack([np.ones(busy.sum()), cnt[busy], hit[busy]], 1)
coef, *_ = np.linalg.lstsq(A, dur[busy], rcond=None)
res = dur[busy] - A @ coef
print("dur ~ %.1f + %.3f*count + %.4f*hits us; residual std %.1f; corr(count,dur) %.3f" % (coef[0], coef[1], coef[2], res.std(), np.corrcoef(cnt[busy], dur[busy])[0, 1]))
# per SM: last end, sum of busy time
last = np.zeros(148)
for s in range(148):
    m = sm == s
    if m.any():
        last[s] = end[m].max()
print("per-SM last end us: min %.1f mean %.1f max %.1f" % (last.min(), last.mean(), last.max()))
# occupancy over time: number of busy CTAs resident, sampled
ts = np.linspace(0, end.max(), 41)
for t in ts[::4]:
    r = (start <= t) & (end > t)
    print("t=%6.1f resident busy %4d empty %4d" % (t, int((r & busy).sum()), int((r & ~busy).sum())))
# the stragglers
order = np.argsort(-end)[:12]
for i in order:
    print("late CTA blk %5d sm %3d start %.1f end %.1f dur %.1f count %d hits %d big-path triangles %d" % (i, sm[i], start[i], end[i], dur[i], cnt[i], hit[i], nbig[i]))
print("tiles with big-path triangles:", int((nbig > 0).sum()), "triangles", int(nbig.sum()), "max per tile", int(nbig.max()))
np.save(os.path.join("gpurun_out", os.environ.get("TAG", "trace") + ".npy"), buf)
