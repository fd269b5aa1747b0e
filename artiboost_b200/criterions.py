"""Losses of the clasbased training config (config/ho3dv2_clasbased_jlol_artiboost2.yaml:172-178), restated from
anakin/criterions/{criterion,jointloss,ordinal}.py on device tensors.  They act on [B,21,3] / [B,8,3] keypoints (a few
KB per batch): torch glue around the network kernels, with all random draws taken on the device so the step never syncs.

Same dict contract as the reference: preds["joints_3d_abs"], preds["corners_3d_abs"]; targs["joints_3d"], ["corners_3d"],
["root_joint"], ["joints_vis"], ["corners_vis"]; `compute_losses` returns (weighted sum, dict of parts)."""
from itertools import combinations, product
from typing import Dict, List, Tuple

import numpy as np
import torch
import torch.nn.functional as F

JOINTS_IDX_PARENTS = [0, 0, 1, 2, 3, 0, 5, 6, 7, 0, 9, 10, 11, 0, 13, 14, 15, 0, 17, 18, 19]  # CONST, anakin/utils/misc.py:71


def _targets(targs, device):
    root = targs["root_joint"].to(device)
    return targs["joints_3d"].to(device) + root.unsqueeze(1), targs["corners_3d"].to(device) + root.unsqueeze(1)


_dev_cache = {}
_CAMERA_AXIS = np.array([[0.0, 0.0, 1.0]], np.float32)


def _on(arr, device):
    """Constant index table on `device` (uploaded once, outside any graph capture)."""
    key = (id(arr), str(device))
    t = _dev_cache.get(key)
    if t is None:
        t = _dev_cache[key] = (arr, torch.as_tensor(arr, device=device))
    return t[1]


def sample_view_vectors(n_virtual_views, device, generator=None):
    """ordinal.py:59-72: the camera axis + n random directions on the upper hemisphere."""
    theta = torch.rand(n_virtual_views, device=device, generator=generator) * 2.0 * np.pi
    u = torch.rand(n_virtual_views, device=device, generator=generator)
    s = torch.sqrt(1.0 - u ** 2)
    nv = torch.stack([s * torch.cos(theta), s * torch.sin(theta), u], dim=1)
    return torch.cat([_on(_CAMERA_AXIS, device), nv], dim=0)


def _subsample(n, device, generator):
    """random.shuffle(idx)[: n // 3] (ordinal.py:167-170), drawn on the device."""
    return torch.randperm(n, device=device, generator=generator)[: n // 3]


class JointsLoss:
    """Masked MSE on absolute joints / corners (jointloss.py:14-67)."""

    def __init__(self, **cfg):
        self.lambda_joints_3d = cfg.get("LAMBDA_JOINTS_3D", 0.0)
        self.lambda_corners_3d = cfg.get("LAMBDA_CORNERS_3D", 0.0)

    def __call__(self, preds: Dict, targs: Dict, **kw) -> Tuple[torch.Tensor, Dict]:
        dev = preds["joints_3d_abs"].device
        tj, tc = _targets(targs, dev)
        final, losses = torch.zeros((), device=dev), {}
        if self.lambda_joints_3d:
            m = targs["joints_vis"].to(dev).unsqueeze(-1)
            losses["joints_3d_loss"] = F.mse_loss(preds["joints_3d_abs"] * m, tj * m)
            final = final + self.lambda_joints_3d * losses["joints_3d_loss"]
        if self.lambda_corners_3d:
            m = targs["corners_vis"].to(dev).unsqueeze(-1)
            losses["corners_3d_loss"] = F.mse_loss(preds["corners_3d_abs"] * m, tc * m)
            final = final + self.lambda_corners_3d * losses["corners_3d_loss"]
        return final, losses


def _joint_ord(pairs_a, pairs_b, view_vecs):
    return torch.einsum("bpk,vk->bpv", pairs_a - pairs_b, view_vecs)


class HandOrdLoss:
    """Joint-level and part-level ordinal relations of the hand under random virtual views (ordinal.py:75-227)."""

    def __init__(self, **cfg):
        self.lambda_part_lev = float(cfg.get("LAMBDA_PART_LEVEL", 1.0))
        self.lambda_joint_lev = float(cfg.get("LAMBDA_JOINTS_LEVEL", 1.0))
        self.n_virtual_views = int(cfg.get("N_VIRTUAL_VIEWS", 20))
        self.jp = np.array(list(combinations(range(21), 2)))
        self.pp = np.array(list(combinations(range(20), 2)))
        self.parents = np.array(JOINTS_IDX_PARENTS)
        self.generator = None

    def __call__(self, preds, targs, **kw):
        pj = preds["joints_3d_abs"]
        dev = pj.device
        tj, _ = _targets(targs, dev)
        m = targs["joints_vis"].to(dev).unsqueeze(-1)
        pj, tj = pj * m, tj * m
        vv = sample_view_vectors(self.n_virtual_views, dev, self.generator)
        jp = _on(self.jp, dev)[_subsample(len(self.jp), dev, self.generator)]
        sign = torch.sign(_joint_ord(tj[:, jp[:, 0]], tj[:, jp[:, 1]], vv))
        joint_ord_loss = torch.log(1.0 + F.relu(-sign * _joint_ord(pj[:, jp[:, 0]], pj[:, jp[:, 1]], vv))).mean()
        par = _on(self.parents, dev)
        pparts, tparts = (pj - pj[:, par])[:, 1:], (tj - tj[:, par])[:, 1:]
        pp = _on(self.pp, dev)[_subsample(len(self.pp), dev, self.generator)]
        t_ord = torch.einsum("bpk,vk->bpv", torch.cross(tparts[:, pp[:, 0]], tparts[:, pp[:, 1]], dim=-1), vv)
        p_ord = torch.einsum("bpk,vk->bpv", torch.cross(pparts[:, pp[:, 0]], pparts[:, pp[:, 1]], dim=-1), vv)
        part_ord_loss = F.relu(-torch.sign(t_ord) * p_ord).mean()
        final = self.lambda_joint_lev * joint_ord_loss + self.lambda_part_lev * part_ord_loss
        return final, {"joint_ord_loss": joint_ord_loss, "part_ord_loss": part_ord_loss}


class SceneOrdLoss:
    """Hand-joint / object-corner ordinal relations (ordinal.py:231-306)."""

    def __init__(self, **cfg):
        self.lambda_scene_lev = float(cfg.get("LAMBDA_SCENE_LEVEL", 1.0))
        self.n_virtual_views = int(cfg.get("N_VIRTUAL_VIEWS", 40))
        self.hp = np.array(list(product(range(21), range(8))))
        self.generator = None

    def __call__(self, preds, targs, **kw):
        pj, pc = preds["joints_3d_abs"], preds["corners_3d_abs"]
        dev = pj.device
        tj, tc = _targets(targs, dev)
        mj, mc = targs["joints_vis"].to(dev).unsqueeze(-1), targs["corners_vis"].to(dev).unsqueeze(-1)
        pj, tj, pc, tc = pj * mj, tj * mj, pc * mc, tc * mc
        vv = sample_view_vectors(self.n_virtual_views, dev, self.generator)
        hp = _on(self.hp, dev)[_subsample(len(self.hp), dev, self.generator)]
        sign = torch.sign(_joint_ord(tj[:, hp[:, 0]], tc[:, hp[:, 1]], vv))
        loss = torch.log(1.0 + F.relu(-sign * _joint_ord(pj[:, hp[:, 0]], pc[:, hp[:, 1]], vv))).mean()
        return self.lambda_scene_lev * loss, {"scene_ord_loss": loss}


LOSSES = {"JointsLoss": JointsLoss, "HandOrdLoss": HandOrdLoss, "SceneOrdLoss": SceneOrdLoss}


class Criterion:
    """criterion.py:30-67: weighted sum of the listed losses."""

    def __init__(self, cfg: Dict, loss_list: List = None, generator=None):
        if loss_list is None:
            loss_list = [LOSSES[c["TYPE"]](**{k: v for k, v in c.items() if k != "TYPE"}) for c in cfg["CRITERION"]]
        self._loss_list = loss_list
        self._loss_lambdas = {type(l).__name__: lam for l, lam in zip(loss_list, cfg["LAMBDAS"])}
        for l in loss_list:
            if hasattr(l, "generator"):
                l.generator = generator

    @property
    def loss_list(self):
        return self._loss_list

    @property
    def loss_lambdas(self):
        return self._loss_lambdas

    def compute_losses(self, preds, targs, **kw):
        total, parts = 0.0, {}
        for loss in self._loss_list:
            final, losses = loss(preds, targs, **kw)
            total = total + self._loss_lambdas[type(loss).__name__] * final
            parts.update(losses)
        assert "final_loss" not in parts, "unexpected premature final loss encountered"
        parts["final_loss"] = total
        return total, parts


DEFAULT_CRITERION_CFG = {"LAMBDAS": [0.5, 0.2, 0.1],
                         "CRITERION": [{"TYPE": "JointsLoss", "LAMBDA_JOINTS_3D": 1.0, "LAMBDA_CORNERS_3D": 0.2},
                                       {"TYPE": "HandOrdLoss"}, {"TYPE": "SceneOrdLoss"}]}
