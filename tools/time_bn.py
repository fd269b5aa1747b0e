"""Device time of the BatchNorm backward passes and the stem im2col at the ResNet34 / batch-128 shapes, cold-ish (each
launch works on its own set of buffers, rotated so that consecutive launches do not find their inputs in L2).
Environment switches are read by the library at first use: run once per setting (tools/gpu_u3.sh)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from artiboost_b200 import lib  # noqa: E402

dev = torch.device("cuda", 0)
L = lib.load()
st = lib.stream_ptr(dev)
P = lambda t: None if t is None else t.data_ptr()  # noqa: E731
SHAPES = [("stem 128x128x64", 128 * 128 * 128, 64), ("layer1 64x64x64", 128 * 64 * 64, 64), ("layer2 32x32x128", 128 * 32 * 32, 128),
          ("layer3 16x16x256", 128 * 16 * 16, 256), ("layer4 8x8x512", 128 * 8 * 8, 512)]


def timed(fn, n_sets, reps=6):
    for i in range(n_sets):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for r in range(reps):
        for i in range(n_sets):
            fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (reps * n_sets) * 1e3


print("settings:", {k: v for k, v in os.environ.items() if k.startswith("AB_")})
tot_r = tot_a = 0.0
for name, M, C in SHAPES:
    bytes_t = M * C * 2
    n_sets = max(2, min(8, int(400e6 // (3 * bytes_t)) + 1))   # > 126 MB of L2 between two uses of a buffer
    mk = lambda: [torch.randn(M, C, device=dev).bfloat16() for _ in range(n_sets)]  # noqa: E731
    dy, raw, y, dx = mk(), mk(), mk(), mk()
    gamma, mean = torch.rand(C, device=dev) + 0.5, torch.randn(C, device=dev) * 0.1
    invstd, scale, shift = torch.rand(C, device=dev) + 0.5, torch.rand(C, device=dev) + 0.5, torch.randn(C, device=dev) * 0.1
    dg, db, coef = torch.zeros(C, device=dev), torch.zeros(C, device=dev), torch.zeros(3 * C, device=dev)
    ws = torch.empty(2 * lib.STAT_PARTS * C, device=dev)
    for mode in ("from_raw", "with_y"):
        yy = (lambda i: None) if mode == "from_raw" else (lambda i: P(y[i]))

        def red(i):
            lib.check(L.ab_bn_bwd_reduce(P(dy[i]), yy(i), P(raw[i]), M, C, P(gamma), P(mean), P(invstd), 1, P(dg), P(db), 0, P(coef),
                                         P(ws), P(scale), P(shift), st), "reduce")

        def app(i):
            lib.check(L.ab_bn_bwd_apply(P(dy[i]), yy(i), P(raw[i]), M, C, P(coef), 1, P(dx[i]), None, P(scale), P(shift), st), "apply")

        n_in = 2 if mode == "from_raw" else 3
        tr, ta = timed(red, n_sets), timed(app, n_sets)
        print(f"{name:18s} {mode:8s} reduce {tr:7.1f} us ({n_in * bytes_t / tr / 1e6:5.2f} TB/s)   apply {ta:7.1f} us ({(n_in + 1) * bytes_t / ta / 1e6:5.2f} TB/s)")
        if mode == "from_raw":
            tot_r += tr; tot_a += ta
    del dy, raw, y, dx
print(f"sum over the five shapes (from_raw): reduce {tot_r:.1f} us, apply {tot_a:.1f} us")
# stem im2col: 128 x 256 x 256 x 4 -> [128*128*128, 200]
B, H, W, Kp = 128, 256, 256, 200
x = [torch.randn(B, H, W, 4, device=dev).bfloat16() for _ in range(2)]
a = [torch.empty(B * 128 * 128, Kp, dtype=torch.bfloat16, device=dev) for _ in range(2)]
t = timed(lambda i: lib.check(L.ab_im2col_nhwc(P(x[i]), B, H, W, 4, 7, 7, 2, 3, Kp, P(a[i]), st), "im2col"), 2)
print(f"stem im2col: {t:.1f} us ({a[0].numel() * 2 / t / 1e6:.2f} TB/s written)")
