"""Pose scramblers (anakin/artiboost/scrambler.py).  On the fused path the scrambler only DRAWS the noise
(`sample_noise`); ab_pose_generate applies it inside the pose-generator prelude (csrc/posegen.cu) with the
arithmetic of RandomScrambler.forward (scrambler.py:65-81).  `forward` keeps the reference's dict-in/dict-out
contract for callers that use a scrambler on its own."""
from typing import Callable, Mapping, Optional, Tuple

import torch
from torch import nn


def register(reg, key):

    def fn(cls):
        reg[key] = cls
        return cls

    return fn


class Scrambler(nn.Module):
    build_mapping: Mapping[str, Callable] = {}

    @staticmethod
    def build(type, *args, **kwargs) -> "Scrambler":
        return Scrambler.build_mapping[type](*args, **kwargs)  # KeyError on unknown types, like the reference

    def sample_noise(self, batch_size: int, device, generator=None) -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor]]:
        """-> (noise_tsl [B,3], noise_angle [B,16]) already scaled by sigma, or (None, None)."""
        return None, None

    def forward(self, inp, **kwargs):
        pose, tsl = inp["hand_pose"], inp["hand_tsl"]
        n_tsl, n_ang = self.sample_noise(pose.shape[0], pose.device)
        if n_tsl is None:
            return {"hand_pose": pose, "hand_tsl": tsl}
        hp = pose.reshape(-1, 16, 3)
        nrm = torch.norm(hp, dim=-1, keepdim=True)
        hp = hp / torch.clip(nrm, min=1e-7) * (nrm + n_ang.unsqueeze(-1))
        return {"hand_pose": hp.reshape(-1, 48), "hand_tsl": tsl + n_tsl}


@register(reg=Scrambler.build_mapping, key="null")
class NullScrambler(Scrambler):

    def __init__(self, cfg=None) -> None:
        super().__init__()


@register(reg=Scrambler.build_mapping, key="naive")
class NaiveScrambler(Scrambler):
    """Translation noise only (scrambler.py:38-56)."""

    def __init__(self, cfg) -> None:
        super().__init__()
        self.tsl_sigma = float(cfg["HAND_TSL_SIGMA"])

    def sample_noise(self, batch_size, device, generator=None):
        n_tsl = torch.randn((batch_size, 3), device=device, generator=generator) * self.tsl_sigma
        return n_tsl, torch.zeros((batch_size, 16), device=device)


@register(reg=Scrambler.build_mapping, key="random")
class RandomScrambler(Scrambler):
    """Per-joint angle-magnitude noise + translation noise (scrambler.py:59-81); the shipped config's default
    (config/ho3dv2_clasbased_jlol_artiboost2.yaml:42-45)."""

    def __init__(self, cfg) -> None:
        super().__init__()
        self.tsl_sigma = float(cfg["HAND_TSL_SIGMA"])
        self.pose_sigma = float(cfg["HAND_POSE_SIGMA"])

    def sample_noise(self, batch_size, device, generator=None):
        n_tsl = torch.randn((batch_size, 3), device=device, generator=generator) * self.tsl_sigma
        n_ang = torch.randn((batch_size, 16), device=device, generator=generator) * self.pose_sigma
        return n_tsl, n_ang
