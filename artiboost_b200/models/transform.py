"""Keypoint-space helpers of the clasbased tail: tiny [B,22,3] / [B,6] tensors (torch plumbing, not a hot loop)."""
from typing import List, Optional

import torch


def batch_uvd2xyz(uvd: torch.Tensor, root_joint: torch.Tensor, intr: torch.Tensor, inp_res: Optional[List[int]] = None,
                  depth_range: float = 0.4, ref_bone_len: Optional[torch.Tensor] = None) -> torch.Tensor:
    """anakin/utils/transform.py:512-546."""
    if inp_res is None:
        inp_res = [256, 256]
    # scalar multiplies (no host->device tensor): the step must stay capturable into a CUDA graph
    uv = torch.stack((uvd[:, :, 0] * float(inp_res[0]), uvd[:, :, 1] * float(inp_res[1])), dim=-1)
    d = (uvd[:, :, 2] - 0.5) * depth_range
    if ref_bone_len is None:
        ref_bone_len = torch.ones((uvd.shape[0], 1), dtype=uvd.dtype, device=uvd.device)
    z = d * ref_bone_len + root_joint[:, -1].unsqueeze(-1).to(uvd.device)
    intr = intr.to(uvd.device)
    f = torch.stack((intr[:, 0, 0], intr[:, 1, 1]), dim=1).unsqueeze(1)
    c = torch.stack((intr[:, 0, 2], intr[:, 1, 2]), dim=1).unsqueeze(1)
    xy = (uv - c) / f * z.unsqueeze(-1)
    return torch.cat((xy, z.unsqueeze(-1)), -1)


def _normalize(v):
    mag = torch.clamp(torch.sqrt(v.pow(2).sum(1)), min=1e-8)
    return v / mag.unsqueeze(1)


def compute_rotation_matrix_from_ortho6d(poses):
    """anakin/utils/transform.py:578-598: x = norm(a), z = norm(x cross b), y = z cross x, columns [x y z]."""
    x = _normalize(poses[:, 0:3])
    z = _normalize(torch.cross(x, poses[:, 3:6], dim=1))
    y = torch.cross(z, x, dim=1)
    return torch.stack((x, y, z), dim=2)
