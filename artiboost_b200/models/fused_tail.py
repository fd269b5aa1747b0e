"""Training-step fusion of the clasbased tail and its criterion: one `ab_tail_losses` launch instead of ~320 torch launches
(HybridBaseline tail, anakin/models/hybridbaseline.py:41-96, + Criterion.compute_losses, anakin/criterions/criterion.py:57-67,
+ their autograd backward).  The unfused modules (`hybridbaseline.py`, `criterions.py`) stay the definition of the
arithmetic -- they are pinned by fixtures recorded from the reference -- and `tests/test_gpu_losses.py` holds this path to
them, values and gradients.

`FusedTailCriterion.plan(criterion, ...)` returns None for a criterion it cannot express (an unknown loss type, a loss
listed twice); the caller then keeps the unfused path."""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import torch

from .. import criterions as crit
from .. import lib


class _TailLossFn(torch.autograd.Function):
    """(kp3d, rot6d) -> (total loss, parts[8], seven prediction tensors); the kernel returns the gradient with the value."""

    @staticmethod
    def forward(ctx, kp3d, rot6d, plan, inputs):
        dev = kp3d.device
        B = kp3d.shape[0]
        f = lambda t: t.to(device=dev, dtype=torch.float32).contiguous()  # noqa: E731
        kp, r6 = f(kp3d), f(rot6d)
        root, intr, can = f(inputs["root_joint"]), f(inputs["cam_intr"]), f(inputs["corners_can"])
        tj, tc, vj, vc = f(inputs["joints_3d"]), f(inputs["corners_3d"]), f(inputs["joints_vis"]), f(inputs["corners_vis"])
        draws = plan.draw(dev)
        cfg = lib.TailCfgStruct()
        cfg.batch, cfg.center_idx = B, plan.center_idx
        cfg.inp_w, cfg.inp_h = float(plan.inp_res[0]), float(plan.inp_res[1])
        cfg.img_w, cfg.img_h = float(inputs["image"].shape[3]), float(inputs["image"].shape[2])
        cfg.depth_range = 0.4
        w = plan.weights
        cfg.w_joints, cfg.w_corners, cfg.w_joint_ord, cfg.w_part_ord = w["joints"], w["corners"], w["joint_ord"], w["part_ord"]
        cfg.w_scene_ord, cfg.w_sym = w["scene_ord"], w["sym"]
        vv_h, jp, pp = draws.get("vv_hand"), draws.get("jp"), draws.get("pp")
        vv_s, hp = draws.get("vv_scene"), draws.get("hp")
        cfg.n_views_hand = 0 if vv_h is None else vv_h.shape[0]
        cfg.n_pairs_joint = 0 if jp is None else jp.shape[0]
        cfg.n_pairs_part = 0 if pp is None else pp.shape[0]
        cfg.n_views_scene = 0 if vv_s is None else vv_s.shape[0]
        cfg.n_pairs_scene = 0 if hp is None else hp.shape[0]
        sym_R = sym_t = obj_idx = obj_transf = None
        cfg.n_sym, cfg.sym_ho3d = 0, 0
        if plan.sym is not None and w["sym"] != 0.0:
            sym_R, sym_t = crit._on(plan.sym.R, dev), crit._on(plan.sym_t3, dev)
            obj_idx = inputs["obj_idx"].to(device=dev, dtype=torch.int32).contiguous()
            obj_transf = f(inputs["obj_transf"])
            cfg.n_sym, cfg.sym_ho3d = sym_R.shape[1], int(plan.sym.use_ho3d_ycb)
        new = lambda *s: torch.empty(s, dtype=torch.float32, device=dev)  # noqa: E731
        out = {"joints_3d_abs": new(B, 21, 3), "corners_3d_abs": new(B, 8, 3), "joints_3d": new(B, 21, 3), "corners_3d": new(B, 8, 3),
               "2d_uvd": new(B, 30, 3), "boxroot_3d_abs": new(B, 1, 3), "box_rot_rotmat": new(B, 3, 3)}
        parts, d_kp, d_r6 = new(8), new(B, 22, 3), new(B, 6)
        L = lib.load()
        ws = torch.empty(max(int(L.ab_tail_losses_workspace_bytes(B)), 4), dtype=torch.uint8, device=dev)
        P = lambda t: None if t is None else t.data_ptr()  # noqa: E731
        with torch.cuda.device(dev):
            lib.check(L.ab_tail_losses(C.byref(cfg), P(kp), P(r6), P(root), P(intr), P(can), P(tj), P(tc), P(vj), P(vc), P(vv_h), P(jp),
                                       P(pp), P(vv_s), P(hp), P(sym_R), P(sym_t), P(obj_idx), P(obj_transf), P(out["joints_3d_abs"]),
                                       P(out["corners_3d_abs"]), P(out["joints_3d"]), P(out["corners_3d"]), P(out["2d_uvd"]),
                                       P(out["boxroot_3d_abs"]), P(out["box_rot_rotmat"]), P(parts), P(d_kp), P(d_r6), P(ws),
                                       lib.stream_ptr(dev)), "ab_tail_losses")
        ctx.save_for_backward(d_kp, d_r6)
        ctx.in_dtypes = (kp3d.dtype, rot6d.dtype)
        outs = tuple(out[k] for k in plan.OUT_KEYS)
        ctx.mark_non_differentiable(parts, *outs)
        return (parts[7].clone(), parts) + outs

    @staticmethod
    def backward(ctx, g_loss, _g_parts, *_g_outs):
        d_kp, d_r6 = ctx.saved_tensors
        return (d_kp * g_loss).to(ctx.in_dtypes[0]), (d_r6 * g_loss).to(ctx.in_dtypes[1]), None, None


@torch.no_grad()
def tail_forward(kp3d: torch.Tensor, rot6d: torch.Tensor, inputs: Dict[str, torch.Tensor], center_idx: int, inp_res) -> Dict[str, torch.Tensor]:
    """The clasbased tail alone (hybridbaseline.py:41-96: uvd -> xyz, 6D -> R, corners, projection, the seven outputs) for the
    inference path: the same `ab_tail_losses` launch with every loss weight zero (no targets, no draws) instead of ~40 torch
    launches.  The torch composition in HybridBaseline.forward stays the definition; tests hold this call to it."""
    dev = kp3d.device
    B = kp3d.shape[0]
    f = lambda t: t.to(device=dev, dtype=torch.float32).contiguous()  # noqa: E731
    kp, r6 = f(kp3d), f(rot6d)
    root, intr, can = f(inputs["root_joint"]), f(inputs["cam_intr"]), f(inputs["corners_can"])
    cfg = lib.TailCfgStruct()
    cfg.batch, cfg.center_idx = B, int(center_idx)
    cfg.inp_w, cfg.inp_h = float(inp_res[0]), float(inp_res[1])
    cfg.img_w, cfg.img_h = float(inputs["image"].shape[3]), float(inputs["image"].shape[2])
    cfg.depth_range = 0.4
    cfg.w_joints = cfg.w_corners = cfg.w_joint_ord = cfg.w_part_ord = cfg.w_scene_ord = cfg.w_sym = 0.0
    cfg.n_views_hand = cfg.n_pairs_joint = cfg.n_pairs_part = cfg.n_views_scene = cfg.n_pairs_scene = cfg.n_sym = cfg.sym_ho3d = 0
    new = lambda *s_: torch.empty(s_, dtype=torch.float32, device=dev)  # noqa: E731
    out = {"joints_3d_abs": new(B, 21, 3), "corners_3d_abs": new(B, 8, 3), "joints_3d": new(B, 21, 3), "corners_3d": new(B, 8, 3),
           "2d_uvd": new(B, 30, 3), "boxroot_3d_abs": new(B, 1, 3), "box_rot_rotmat": new(B, 3, 3)}
    parts, d_kp, d_r6 = new(8), new(B, 22, 3), new(B, 6)
    zero = torch.zeros(B * (63 + 24 + 21 + 8), dtype=torch.float32, device=dev)   # the (unweighted) targets the entry point insists on
    tj, tc, vj, vc = zero[:63 * B], zero[63 * B:87 * B], zero[87 * B:108 * B], zero[108 * B:]
    L = lib.load()
    ws = torch.empty(max(int(L.ab_tail_losses_workspace_bytes(B)), 4), dtype=torch.uint8, device=dev)
    P = lambda t: None if t is None else t.data_ptr()  # noqa: E731
    with torch.cuda.device(dev):
        lib.check(L.ab_tail_losses(C.byref(cfg), P(kp), P(r6), P(root), P(intr), P(can), P(tj), P(tc), P(vj), P(vc), None, None, None, None,
                                   None, None, None, None, None, P(out["joints_3d_abs"]), P(out["corners_3d_abs"]), P(out["joints_3d"]),
                                   P(out["corners_3d"]), P(out["2d_uvd"]), P(out["boxroot_3d_abs"]), P(out["box_rot_rotmat"]), P(parts),
                                   P(d_kp), P(d_r6), P(ws), lib.stream_ptr(dev)), "ab_tail_losses")
    return out


class FusedTailCriterion:
    OUT_KEYS = ("joints_3d_abs", "corners_3d_abs", "joints_3d", "corners_3d", "2d_uvd", "boxroot_3d_abs", "box_rot_rotmat")
    TARGET_KEYS = ("root_joint", "cam_intr", "corners_can", "joints_3d", "corners_3d", "joints_vis", "corners_vis", "image")

    def __init__(self, criterion: crit.Criterion, center_idx: int, inp_res):
        self.criterion, self.center_idx, self.inp_res = criterion, int(center_idx), list(inp_res)
        by = {type(l).__name__: l for l in criterion.loss_list}
        lam = criterion.loss_lambdas
        j, h, s, y = by.get("JointsLoss"), by.get("HandOrdLoss"), by.get("SceneOrdLoss"), by.get("SymCornerLoss")
        self.hand, self.scene, self.sym = h, s, y
        self.order = [type(l).__name__ for l in criterion.loss_list]
        self.weights = {
            "joints": float(lam["JointsLoss"] * j.lambda_joints_3d) if j else 0.0,
            "corners": float(lam["JointsLoss"] * j.lambda_corners_3d) if j else 0.0,
            "joint_ord": float(lam["HandOrdLoss"] * h.lambda_joint_lev) if h else 0.0,
            "part_ord": float(lam["HandOrdLoss"] * h.lambda_part_lev) if h else 0.0,
            "scene_ord": float(lam["SceneOrdLoss"] * s.lambda_scene_lev) if s else 0.0,
            "sym": float(lam["SymCornerLoss"] * y.lambda_sym_corners_3d) if y else 0.0,
        }
        self.sym_t3 = None if y is None else y.t.reshape(y.t.shape[0], y.t.shape[1], 3).copy()

    @classmethod
    def plan(cls, criterion: crit.Criterion, center_idx: int, inp_res) -> Optional["FusedTailCriterion"]:
        names = [type(l).__name__ for l in criterion.loss_list]
        known = {"JointsLoss": crit.JointsLoss, "HandOrdLoss": crit.HandOrdLoss, "SceneOrdLoss": crit.SceneOrdLoss,
                 "SymCornerLoss": crit.SymCornerLoss}
        if len(set(names)) != len(names) or any(known.get(n) is not type(l) for n, l in zip(names, criterion.loss_list)):
            return None
        return cls(criterion, center_idx, inp_res)

    def draw(self, dev) -> Dict[str, torch.Tensor]:
        """The random draws of the ordinal losses, in the criterion's own order and from its generator
        (criterions.py HandOrdLoss.__call__ / SceneOrdLoss.__call__): both paths consume the same stream."""
        d = {}
        for name in self.order:
            if name == "HandOrdLoss":
                h = self.hand
                d["vv_hand"] = crit.sample_view_vectors(h.n_virtual_views, dev, h.generator).contiguous()
                d["jp"] = crit._on(h.jp, dev)[crit._subsample(len(h.jp), dev, h.generator)].to(torch.int32).contiguous()
                d["pp"] = crit._on(h.pp, dev)[crit._subsample(len(h.pp), dev, h.generator)].to(torch.int32).contiguous()
            elif name == "SceneOrdLoss":
                s = self.scene
                d["vv_scene"] = crit.sample_view_vectors(s.n_virtual_views, dev, s.generator).contiguous()
                d["hp"] = crit._on(s.hp, dev)[crit._subsample(len(s.hp), dev, s.generator)].to(torch.int32).contiguous()
        return d

    def usable(self, inputs: Dict) -> bool:
        need = self.TARGET_KEYS + (("obj_idx", "obj_transf") if self.sym is not None and self.weights["sym"] != 0.0 else ())
        return all(k in inputs for k in need)

    def __call__(self, kp3d: torch.Tensor, rot6d: torch.Tensor, inputs: Dict):
        """-> (preds dict with the reference's seven keys, total loss, parts dict with the criterion's keys)."""
        res = _TailLossFn.apply(kp3d, rot6d, self, inputs)
        loss, parts = res[0], res[1]
        preds = dict(zip(self.OUT_KEYS, res[2:]))
        w, named = self.weights, {}
        if "JointsLoss" in self.order:
            j = next(l for l in self.criterion.loss_list if isinstance(l, crit.JointsLoss))
            if j.lambda_joints_3d:
                named["joints_3d_loss"] = parts[0]
            if j.lambda_corners_3d:
                named["corners_3d_loss"] = parts[1]
        if self.hand is not None:
            named["joint_ord_loss"], named["part_ord_loss"] = parts[2], parts[3]
        if self.scene is not None:
            named["scene_ord_loss"] = parts[4]
        if self.sym is not None:
            named["sym_corners_3d_loss"] = parts[5] if w["sym"] != 0.0 else None
        named["final_loss"] = loss
        return preds, loss, named
