"""Pose scramblers (anakin/artiboost/scrambler.py).  On the fused path the scrambler only DRAWS the noise
(`sample_noise`); ab_pose_generate applies it inside the pose-generator prelude (csrc/posegen.cu) with the
arithmetic of RandomScrambler.forward (scrambler.py:65-81).  `forward` keeps the reference's dict-in/dict-out
contract for callers that use a scrambler on its own.  The anatomical scramblers `random_2` / `random_3`
(scrambler.py:84-260) run through ab_scramble_anatomical (csrc/refine.cu), which also holds manotorch's AxisLayer."""
from typing import Callable, Mapping, Optional, Tuple

import torch
from torch import nn

from .. import lib


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


def scramble_anatomical(pose, joints, transf, splay, bend, thumb, return_axes=False):
    """ab_scramble_anatomical: pose [B,48], joints [B,21,3], transf [B,16,4,4] (MANO chain order), splay [B,4],
    bend [B,14], thumb [B,2] -> pose' [B,48] (and the AxisLayer axes [B,15,3(b,u,l),3] when asked)."""
    lib.require_cuda(pose, "hand_pose")
    dev, B = pose.device, pose.shape[0]
    c = lambda t: t.to(dev).contiguous().float()  # noqa: E731
    pose, joints, transf, splay, bend, thumb = c(pose), c(joints), c(transf), c(splay), c(bend), c(thumb)
    assert joints.shape[1:] == (21, 3) and transf.shape[1:] == (16, 4, 4)
    assert splay.shape == (B, 4) and bend.shape == (B, 14) and thumb.shape == (B, 2)
    out = torch.empty((B, 48), device=dev, dtype=torch.float32)
    axes = torch.empty((B, 15, 3, 3), device=dev, dtype=torch.float32) if return_axes else None
    with torch.cuda.device(dev):
        rc = lib.load().ab_scramble_anatomical(B, lib.ptr(pose), lib.ptr(joints), lib.ptr(transf), lib.ptr(splay),
                                               lib.ptr(bend), lib.ptr(thumb), lib.ptr(out), lib.ptr(axes),
                                               lib.stream_ptr(dev))
    lib.check(rc, "ab_scramble_anatomical")
    return (out, axes) if return_axes else out


class AxisLayer(nn.Module):
    """manotorch.axislayer.AxisLayer as the scramblers call it (scrambler.py:116,219): back / up / left axes of the 15
    articulated joints, in MANO chain order, each in its joint's local frame."""

    def forward(self, hand_joints, transf):
        B = hand_joints.shape[0]
        z = torch.zeros
        dev = hand_joints.device
        _, axes = scramble_anatomical(z((B, 48), device=dev), hand_joints, transf, z((B, 4), device=dev),
                                      z((B, 14), device=dev), z((B, 2), device=dev), return_axes=True)
        return axes[:, :, 0], axes[:, :, 1], axes[:, :, 2]


class _AnatomicalScrambler(Scrambler):
    needs_hand_transf = True  # PreProcessorPoseGenerator runs the MANO forward that yields joints / hand_transf

    def __init__(self, cfg) -> None:
        super().__init__()
        self.tsl_sigma = float(cfg["HAND_TSL_SIGMA"])
        self.pose_sigma = float(cfg["HAND_POSE_SIGMA"])
        self.axis_layer = AxisLayer()

    def sample_anatomical_noise(self, batch_size, device, generator=None):
        """-> (n_tsl [B,3], splay [B,4], bend [B,14], thumb [B,2]) in the reference's draw order."""
        raise NotImplementedError

    def forward(self, inp, noise=None, **kwargs):
        pose, tsl = inp["hand_pose"], inp["hand_tsl"]
        assert inp["hand_transf"] is not None
        n_tsl, splay, bend, thumb = noise if noise is not None else self.sample_anatomical_noise(
            pose.shape[0], pose.device, kwargs.get("generator"))
        new_pose = scramble_anatomical(pose, inp["joints"], inp["hand_transf"], splay, bend, thumb)
        return {"hand_pose": new_pose, "hand_tsl": tsl + n_tsl.to(tsl.device)}


@register(reg=Scrambler.build_mapping, key="random_2")
class RandomScrambler2(_AnatomicalScrambler):
    """Knuckle splay + one bend draw per finger spread over its three joints with the interlink coefficients
    1 / 1.1 / 0.9 (thumb: 1 / 0.9) + thumb-base bend and splay (scrambler.py:84-188)."""

    def __init__(self, cfg) -> None:
        super().__init__(cfg)
        self.coef_1, self.coef_2 = 1.1, 0.9

    def expand_bend(self, bend5):
        """[B,5] per-finger draws (index, middle, ring, little, thumb: scrambler.py:139-170) -> [B,14] per-joint bend
        angles of chain joints (1..12, 14, 15) = (index, middle, little, ring, thumb)."""
        link = torch.tensor([1.0, self.coef_1, self.coef_2], device=bend5.device)
        per = lambda k: bend5[:, k:k + 1] * link  # noqa: E731
        return torch.cat([per(0), per(1), per(3), per(2), bend5[:, 4:5] * link[[0, 2]]], 1)

    def sample_anatomical_noise(self, batch_size, device, generator=None):
        rn = lambda *s: torch.randn(s, device=device, generator=generator)  # noqa: E731
        n_tsl = rn(batch_size, 3) * self.tsl_sigma
        splay = rn(batch_size, 4) * self.pose_sigma
        bend = self.expand_bend(rn(batch_size, 5) * self.pose_sigma)
        thumb = rn(batch_size, 2) * self.pose_sigma
        return n_tsl, splay, bend, thumb


@register(reg=Scrambler.build_mapping, key="random_3")
class RandomScrambler3(_AnatomicalScrambler):
    """Knuckle splay + an independent bend draw for each of the 14 bending joints + thumb base (scrambler.py:191-260)."""

    def sample_anatomical_noise(self, batch_size, device, generator=None):
        rn = lambda *s: torch.randn(s, device=device, generator=generator)  # noqa: E731
        n_tsl = rn(batch_size, 3) * self.tsl_sigma
        splay = rn(batch_size, 4) * self.pose_sigma
        bend = rn(batch_size, 14) * self.pose_sigma
        thumb = rn(batch_size, 2) * self.pose_sigma
        return n_tsl, splay, bend, thumb
