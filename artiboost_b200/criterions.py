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


def get_symmetry_transformations(model_info: Dict, max_sym_disc_step: float):
    """anakin/utils/bop_toolkit/bop_misc.py:18-65: the identity + discrete symmetries, each combined with the discretised
    continuous ones (rotation about `axis` through `offset` in steps of 2 pi / ceil(pi / max_sym_disc_step))."""
    disc = [(np.eye(3), np.zeros((3, 1)))]
    for sym in model_info.get("symmetries_discrete", []):
        m = np.reshape(sym, (4, 4))
        disc.append((m[:3, :3], m[:3, 3].reshape(3, 1)))
    cont = []
    for sym in model_info.get("symmetries_continuous", []):
        axis = np.asarray(sym["axis"], np.float64)
        offset = np.asarray(sym["offset"], np.float64).reshape(3, 1)
        steps = int(np.ceil(np.pi / max_sym_disc_step))
        d = axis / np.linalg.norm(axis)
        for i in range(1, steps):
            ang = i * 2.0 * np.pi / steps
            sa, ca = np.sin(ang), np.cos(ang)  # bop_toolkit/transform.py:302-343 rotation_matrix
            R = np.diag([ca, ca, ca]) + np.outer(d, d) * (1.0 - ca) + sa * np.array([[0.0, -d[2], d[1]], [d[2], 0.0, -d[0]], [-d[1], d[0], 0.0]])
            cont.append((R, -R.dot(offset) + offset))
    if not cont:
        return disc
    return [(Rc.dot(Rd), Rc.dot(td) + tc) for Rd, td in disc for Rc, tc in cont]


class SymCornerLoss:
    """Symmetry-aware corner loss of the DexYCB config (symcornerloss.py:18-108, eval_dexycb_clasbased_sym_artiboost.yaml:
    89-91): the MSE to the closest of the object's symmetric corner sets.  `MODEL_INFO` (dict) or `MODEL_INFO_PATH` (json,
    BOP models_info layout keyed "1".."N"); translations in millimetres like the reference."""

    def __init__(self, **cfg):
        import json
        self.lambda_sym_corners_3d = cfg.get("LAMBDA_SYM_CORNERS_3D", 0.0)
        info = cfg["MODEL_INFO"] if "MODEL_INFO" in cfg else json.load(open(cfg["MODEL_INFO_PATH"], "r"))
        self.max_sym_disc_step = cfg.get("MAX_SYM_DISC_STEP", 0.01)
        self.use_ho3d_ycb = cfg.get("USE_HO3D_YCB", False)
        syms = [get_symmetry_transformations(info[str(i)], self.max_sym_disc_step) for i in range(1, len(info) + 1)]
        k = max(len(s_) for s_ in syms)
        R = np.stack([np.stack([s_[j][0] if j < len(s_) else np.eye(3) for j in range(k)]) for s_ in syms])
        t = np.stack([np.stack([s_[j][1] if j < len(s_) else np.zeros((3, 1)) for j in range(k)]) for s_ in syms]) / 1000.0
        self.R, self.t = R.astype(np.float32), t.astype(np.float32)  # [N, K, 3, 3], [N, K, 3, 1]

    def __call__(self, preds, targs, **kw):
        pc = preds["corners_3d_abs"]
        dev = pc.device
        if not self.lambda_sym_corners_3d:
            return torch.zeros((), device=dev), {"sym_corners_3d_loss": None}
        idx = targs["obj_idx"].to(dev).long() - 1  # device gather (the reference goes through .tolist())
        sym_R, sym_t = _on(self.R, dev)[idx], _on(self.t, dev)[idx]
        cc = targs["corners_can"].to(dev)
        transf = targs["obj_transf"].to(dev)
        if not self.use_ho3d_ycb:
            sym_can = (torch.einsum("bkmn,bcn->bkmc", sym_R, cc) + sym_t).transpose(-2, -1)
        else:
            ext = _on(_CAM_EXTR, dev)
            sym_can = (ext @ (torch.einsum("bkmn,bnc->bkmc", sym_R, ext @ cc.transpose(-2, -1)) + sym_t)).transpose(-2, -1)
        sym_abs = (torch.einsum("bij,bklj->bkil", transf[:, :3, :3], sym_can) + transf[:, None, :3, 3:]).transpose(-2, -1)
        vis = targs["corners_vis"].to(dev)
        pcm = pc * vis.unsqueeze(-1)
        sym_abs = sym_abs * vis[:, None, :, None]
        loss = ((sym_abs - pcm[:, None]) ** 2).mean(-1).mean(-1).min(dim=-1)[0].mean()
        return self.lambda_sym_corners_3d * loss, {"sym_corners_3d_loss": loss}


_CAM_EXTR = np.array([[1.0, 0.0, 0.0], [0.0, -1.0, 0.0], [0.0, 0.0, -1.0]], np.float32)
LOSSES = {"JointsLoss": JointsLoss, "HandOrdLoss": HandOrdLoss, "SceneOrdLoss": SceneOrdLoss, "SymCornerLoss": SymCornerLoss}


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
