"""View engine (anakin/artiboost/view_engine.py:6-86): perspective id -> rotation aligning +z with a jittered
sphere-bin direction, free in-plane roll, camera z offset.  `get_view_batch` is the on-device batched form
(ab_view_from_id); `get_view` keeps the reference's one-id host signature on top of it."""
import ctypes as C

import numpy as np
import torch

from .. import lib


class ViewEngine:

    def __init__(self, cfg):
        self.persp_u_bins = cfg["PERSP_U_BINS"]
        self.persp_theta_bins = cfg["PERSP_THETA_BINS"]
        self.camera_z_range = cfg["CAMERA_Z_RANGE"]
        self.n_persp_center = self.persp_u_bins * self.persp_theta_bins

    @torch.no_grad()
    def get_view_batch(self, persp_id: torch.Tensor, rand4: torch.Tensor = None, generator=None):
        """persp_id int[n] on a CUDA device; rand4 f32[n,4] U[0,1) = (u jitter, theta jitter, roll, z) or None to
        draw them here.  -> persp_rotmat [n,3,3], camera_free_transf [n,4,4], z_offset [n,3] (fp32, device)."""
        lib.require_cuda(persp_id, "persp_id")
        dev = persp_id.device
        n = int(persp_id.shape[0])
        if rand4 is None:
            rand4 = torch.rand((n, 4), device=dev, generator=generator)
        pid = persp_id.to(torch.int32).contiguous()
        rand4 = rand4.float().contiguous()
        rot = torch.empty((n, 3, 3), device=dev, dtype=torch.float32)
        free = torch.empty((n, 4, 4), device=dev, dtype=torch.float32)
        zoff = torch.empty((n, 3), device=dev, dtype=torch.float32)
        zmin, zmax = self.camera_z_range
        with torch.cuda.device(dev):
            rc = lib.load().ab_view_from_id(lib.ptr(pid), n, self.persp_u_bins, self.persp_theta_bins, float(zmin),
                                            float(zmax), lib.ptr(rand4), lib.ptr(rot), lib.ptr(free), lib.ptr(zoff),
                                            lib.stream_ptr(dev))
        lib.check(rc, "ab_view_from_id")
        return rot, free, zoff

    def get_view(self, persp_id, device="cuda"):
        rot, free, zoff = self.get_view_batch(torch.as_tensor([int(persp_id)], device=device))
        return (rot[0].cpu().numpy().astype(np.float64), free[0].cpu().numpy().astype(np.float64),
                zoff[0].cpu().numpy().astype(np.float64))
