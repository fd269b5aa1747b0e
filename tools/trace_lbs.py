"""Phase timeline of mano_lbs_kernel at batch 512.  Needs a trace build of mano.cu (-DAB_LBS_TRACE, debug only):
nvcc ... -DAB_LBS_TRACE -c artiboost_b200/csrc/mano.cu, linked into artiboost_b200/build/variants/lbstrace.so."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from artiboost_b200 import lib  # noqa: E402
from artiboost_b200.synth import SynthPipeline  # noqa: E402

pipe = SynthPipeline(device="cuda:0", seed=1, n_hand_tex=2, n_bg=2)
for _ in range(10):
    pipe.sample_poses(512)
torch.cuda.synchronize()
n = 32 * 19
buf = np.zeros((n, 8), dtype=np.uint64)
L = lib.load()
L.ab_debug_lbs_trace.restype = C.c_int
assert L.ab_debug_lbs_trace(buf.ctypes.data_as(C.c_void_p), n) == 0
t = (buf[:, :7] - buf[:, 0].min()).astype(np.float64) / 1e3
print("kernel span us %.1f; CTA start spread %.1f" % (t[:, 6].max(), t[:, 0].max()))
names = ["prelude (rodrigues, coef)", "rest joints + weight compaction", "chain (thread 0's finger)", "blend shapes", "skinning", "centre/joints/store"]
for i, nm in enumerate(names):
    d = t[:, i + 1] - t[:, i]
    print("%-34s mean %6.2f us  p90 %6.2f  max %6.2f" % (nm, d.mean(), np.percentile(d, 90), d.max()))
print("CTA total: mean %.2f max %.2f" % ((t[:, 6] - t[:, 0]).mean(), (t[:, 6] - t[:, 0]).max()))
