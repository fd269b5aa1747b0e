"""Device time of the clasbased tail (uvd->xyz, 6D->R, corners, projections) + Criterion forward and backward on [B,22,3] / [B,6]
inputs, as torch ops (eager and CUDA-graph replay): what a fused kernel would replace."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from artiboost_b200.criterions import DEFAULT_CRITERION_CFG, Criterion  # noqa: E402
from artiboost_b200.models.transform import batch_uvd2xyz, compute_rotation_matrix_from_ortho6d  # noqa: E402

B = int(os.environ.get("B", 128))
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(0)
kp = torch.rand((B, 22, 3), device=dev, generator=g).requires_grad_(True)
r6 = torch.randn((B, 6), device=dev, generator=g).requires_grad_(True)
batch = {"root_joint": torch.randn((B, 3), device=dev) * 0.05 + torch.tensor([0, 0, 0.5], device=dev),
         "cam_intr": torch.tensor([[600.0, 0, 128], [0, 600, 128], [0, 0, 1]], device=dev).repeat(B, 1, 1),
         "corners_can": torch.randn((B, 8, 3), device=dev) * 0.05, "joints_3d": torch.randn((B, 21, 3), device=dev) * 0.05,
         "corners_3d": torch.randn((B, 8, 3), device=dev) * 0.05, "joints_vis": torch.ones((B, 21), device=dev),
         "corners_vis": torch.ones((B, 8), device=dev)}
crit = Criterion(DEFAULT_CRITERION_CFG, generator=g)


def step():
    kp.grad = None
    r6.grad = None
    p = batch_uvd2xyz(kp, batch["root_joint"], batch["cam_intr"], [256, 256])
    j, br = p[:, :21], p[:, 21:22]
    R = compute_rotation_matrix_from_ortho6d(r6)
    c = torch.matmul(R, batch["corners_can"].permute(0, 2, 1)).permute(0, 2, 1) + br
    loss, _ = crit.compute_losses({"joints_3d_abs": j, "corners_3d_abs": c}, batch)
    loss.backward()
    return loss


for _ in range(3):
    step()
torch.cuda.synchronize()
gr = torch.cuda.CUDAGraph()
gr.register_generator_state(g)
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    step()
    torch.cuda.synchronize()
    with torch.cuda.graph(gr, stream=s):
        step()
for name, fn in (("eager", step), ("graph", gr.replay)):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1) / 50:.3f} ms per tail + losses fwd + bwd at B={B}")
from torch.profiler import ProfilerActivity, profile  # noqa: E402
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
print("kernels per step:", len(ev), "sum of kernel time us:", sum(e.device_time for e in ev))
