"""Refiners (anakin/artiboost/refiner.py): `NullRefine` (:118-147) and the shipped config's `hand_obj` refiner,
`HORefiner` + `_RefineNet` (:150-285; config/ho3dv2_clasbased_jlol_artiboost2.yaml:47-50).

Same classes, constructor arguments, parameter names (a GrabNet `refinenet.pt` loads with `load_state_dict(...,
strict=False)` exactly as in the reference) and dict-in / dict-out contract.  The arithmetic runs through the C-ABI
(include/artiboost_b200.h): ab_chamfer_nn (point2point_signed with the object rotation and the eval-mode
BatchNorm1d(778) folded in), ab_linear_f32 (ResBlock MLP, BatchNorm folded into the weights), ab_refine_encode /
ab_refine_decode (rotation 6D <-> axis-angle) and ab_mano_forward.  Inference only, eval-mode semantics (the
reference puts the net in eval() at construction, refiner.py:159, and calls it under no_grad,
artiboost_loader.py:369): Dropout is the identity and BatchNorm uses its running statistics."""
from copy import deepcopy
from typing import Callable, Dict, List, Mapping, Optional

import numpy as np
import torch
from torch import nn

from .. import lib
from ..manolayer import ManoLayer
from .scrambler import register

IN_SIZE = 778 + 16 * 6 + 3  # h2o distances + 6D pose + translation (refiner.py:229)
H_SIZE = 512


class Refiner:
    build_mapping: Mapping[str, Callable] = {}

    @staticmethod
    def build(type, *args, **kwargs):
        return Refiner.build_mapping[type](*args, **kwargs)


# --------------------------------------------------------------------------------------------- thin kernel wrappers
def chamfer_nn(x: torch.Tensor, y_points: torch.Tensor, obj_id: Optional[torch.Tensor] = None,
               rot: Optional[torch.Tensor] = None, scale: Optional[torch.Tensor] = None,
               shift: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None, return_idx: bool = True):
    """x [B,P1,3]; y_points [n_obj or B, P2, 3]; obj_id int32 [B] or None (cloud b); rot [B,3,3] / [B,4,4] or None.
    -> (dist [B,P1] (or written into `out`, a [B,P1] view with unit column stride), idx int32 [B,P1] or None)."""
    lib.require_cuda(x, "x")
    B, P1 = x.shape[0], x.shape[1]
    x = x.contiguous().float()
    y_points = y_points.contiguous().float()
    if obj_id is None and y_points.shape[0] != B:
        raise ValueError("y does not have the correct shape.")  # refiner.py:50-51
    if out is None:
        out = torch.empty((B, P1), device=x.device, dtype=torch.float32)
    assert out.stride(1) == 1 and out.dtype == torch.float32
    idx = torch.empty((B, P1), device=x.device, dtype=torch.int32) if return_idx else None
    rs = 0
    if rot is not None:
        rot = rot.contiguous().float()
        rs = 16 if rot.shape[-1] == 4 else 9
    if obj_id is not None:
        obj_id = obj_id.to(torch.int32).contiguous()
    L = lib.load()
    ws = torch.empty(max(int(L.ab_chamfer_nn_workspace_bytes(B, P1)), 8) // 8, dtype=torch.int64, device=x.device)
    with torch.cuda.device(x.device):
        rc = L.ab_chamfer_nn(B, P1, lib.ptr(x), y_points.shape[1], lib.ptr(y_points), lib.ptr(obj_id),
                             lib.ptr(rot), rs, lib.ptr(scale), lib.ptr(shift), out.data_ptr(),
                             out.stride(0) if B > 0 else P1, lib.ptr(idx), lib.ptr(ws), lib.stream_ptr(x.device))
    lib.check(rc, "ab_chamfer_nn")
    return out, idx


def build_nn_groups(points: np.ndarray, group: int = 32):
    """points [n_obj, P, 3] (static clouds) -> (sorted_points [n_obj, Pp, 3] f32, perm [n_obj, Pp] i32, boxes
    [n_obj, 6, Pp / 32] f32) for ab_chamfer_nn_grouped: each cloud sorted along a 30-bit Morton curve over its bounding
    box (stable, so equal codes keep their original order), padded to a multiple of 32 by repeating the last point, with
    the axis-aligned box of every run of 32 points."""
    points = np.asarray(points, np.float32)
    n_obj, P, _ = points.shape
    Pp = (P + group - 1) // group * group
    sp = np.empty((n_obj, Pp, 3), np.float32)
    pm = np.empty((n_obj, Pp), np.int32)
    bx = np.empty((n_obj, 6, Pp // group), np.float32)

    def spread(v):  # 10 bits -> every third bit
        v = v.astype(np.uint64)
        v = (v | (v << 16)) & 0x030000FF
        v = (v | (v << 8)) & 0x0300F00F
        v = (v | (v << 4)) & 0x030C30C3
        v = (v | (v << 2)) & 0x09249249
        return v

    for o in range(n_obj):
        p = points[o]
        lo, hi = p.min(0), p.max(0)
        q = np.clip(((p - lo) / np.maximum(hi - lo, 1e-12) * 1023.0).astype(np.int64), 0, 1023)
        code = spread(q[:, 0]) | (spread(q[:, 1]) << 1) | (spread(q[:, 2]) << 2)
        order = np.argsort(code, kind="stable")
        order = np.concatenate([order, np.full(Pp - P, order[-1])])
        sp[o], pm[o] = p[order], order
        g = sp[o].reshape(Pp // group, group, 3)
        bx[o, :3], bx[o, 3:] = g.min(1).T, g.max(1).T
    return sp, pm, bx


def chamfer_nn_grouped(x, groups, obj_id=None, rot=None, scale=None, shift=None, out=None, return_idx=True):
    """chamfer_nn over clouds prepared by build_nn_groups (`groups` = its three arrays as device tensors): same
    arguments, same bits, a fraction of the pairs evaluated."""
    lib.require_cuda(x, "x")
    sp, pm, bx = groups
    B, P1 = x.shape[0], x.shape[1]
    x = x.contiguous().float()
    if out is None:
        out = torch.empty((B, P1), device=x.device, dtype=torch.float32)
    assert out.stride(1) == 1 and out.dtype == torch.float32
    idx = torch.empty((B, P1), device=x.device, dtype=torch.int32) if return_idx else None
    rs = 0
    if rot is not None:
        rot = rot.contiguous().float()
        rs = 16 if rot.shape[-1] == 4 else 9
    if obj_id is not None:
        obj_id = obj_id.to(torch.int32).contiguous()
    elif sp.shape[0] != B:
        raise ValueError("y does not have the correct shape.")
    with torch.cuda.device(x.device):
        rc = lib.load().ab_chamfer_nn_grouped(B, P1, lib.ptr(x), bx.shape[2], lib.ptr(sp), lib.ptr(pm), lib.ptr(bx),
                                              lib.ptr(obj_id), lib.ptr(rot), rs, lib.ptr(scale), lib.ptr(shift),
                                              out.data_ptr(), out.stride(0) if B > 0 else P1, lib.ptr(idx),
                                              lib.stream_ptr(x.device))
    lib.check(rc, "ab_chamfer_nn_grouped")
    return out, idx


def point2point_signed(x, y, x_normals=None, y_normals=None):
    """refiner.py:21-85 as the hot path calls it (no normals): distance from every x point to its nearest y point."""
    if x_normals is not None or y_normals is not None:
        raise NotImplementedError("the ArtiBoost refiner calls point2point_signed without normals")
    if y.shape[0] != x.shape[0] or y.shape[2] != x.shape[2]:
        raise ValueError("y does not have the correct shape.")
    return chamfer_nn(x, y, return_idx=False)[0]


def linear_f32(x, W, bias, out, residual=None, leaky: Optional[float] = None):
    """out[M,N] = act(x[M,K] W[N,K]^T + bias (+ residual)); x / out / residual may be column slices of wider buffers."""
    M, K = x.shape
    N = W.shape[0]
    assert x.stride(1) == 1 and out.stride(1) == 1 and W.is_contiguous() and W.shape[1] == K and out.shape == (M, N)
    with torch.cuda.device(x.device):
        rc = lib.load().ab_linear_f32(M, N, K, x.data_ptr(), x.stride(0), W.data_ptr(), K, lib.ptr(bias),
                                      None if residual is None else residual.data_ptr(),
                                      0 if residual is None else residual.stride(0), int(leaky is not None),
                                      float(leaky or 0.0), out.data_ptr(), out.stride(0), lib.stream_ptr(x.device))
    lib.check(rc, "ab_linear_f32")
    return out


def CRot2rotmat(pose):
    """refiner.py:88-98 (host tensor ops; the hot path decodes inside ab_refine_decode)."""
    a = pose.view(-1, 3, 2)
    b1 = nn.functional.normalize(a[:, :, 0], dim=1)
    dot = torch.sum(b1 * a[:, :, 1], dim=1, keepdim=True)
    b2 = nn.functional.normalize(a[:, :, 1] - dot * b1, dim=-1)
    b3 = torch.cross(b1, b2, dim=1)
    return torch.stack([b1, b2, b3], dim=-1)


def parms_decode(pose_crot, trans):
    """refiner.py:101-107 through ab_refine_decode."""
    B = trans.shape[0]
    feat = torch.cat([pose_crot.reshape(B, 96), trans], 1).contiguous().float()
    pose = torch.empty((B, 48), device=feat.device, dtype=torch.float32)
    with torch.cuda.device(feat.device):
        rc = lib.load().ab_refine_decode(B, lib.ptr(feat), 99, None, 0, None, lib.ptr(pose), None, None,
                                         lib.stream_ptr(feat.device))
    lib.check(rc, "ab_refine_decode")
    return {"th_pose_coeffs": pose, "th_tsl": trans}


# ----------------------------------------------------------------------------------------------------- the network
class ResBlock(nn.Module):
    """refiner.py:288-319 (parameters only; the arithmetic is _RefineNet._resblock)."""

    def __init__(self, Fin, Fout, n_neurons=256):
        super().__init__()
        self.Fin, self.Fout = Fin, Fout
        self.fc1 = nn.Linear(Fin, n_neurons)
        self.bn1 = nn.BatchNorm1d(n_neurons)
        self.fc2 = nn.Linear(n_neurons, Fout)
        self.bn2 = nn.BatchNorm1d(Fout)
        if Fin != Fout:
            self.fc3 = nn.Linear(Fin, Fout)
        self.ll = nn.LeakyReLU(negative_slope=0.2)


def _bn_affine(bn: nn.BatchNorm1d):
    s = bn.weight.detach().float() / torch.sqrt(bn.running_var.float() + bn.eps)
    return s, bn.bias.detach().float() - bn.running_mean.float() * s


class _RefineNet(nn.Module):
    """GrabNet RefineNet (refiner.py:227-285) + the refiner's MANO layer at `refine_net.mano_layer`, where
    ArtiBoostLoader picks it up (artiboost_loader.py:172)."""

    def __init__(self, in_size=IN_SIZE, h_size=H_SIZE, n_iters=3, mano_model=None, mano_assets_root="assets/mano_v1_2"):
        super().__init__()
        self.n_iters = n_iters
        self.in_size, self.h_size = in_size, h_size
        if n_iters > 0:  # NullRefine builds the holder with n_iters=0 and never runs the MLP
            self.bn1 = nn.BatchNorm1d(778)
            self.rb1 = ResBlock(in_size, h_size)
            self.rb2 = ResBlock(in_size + h_size, h_size)
            self.rb3 = ResBlock(in_size + h_size, h_size)
            self.out_p = nn.Linear(h_size, 16 * 6)
            self.out_t = nn.Linear(h_size, 3)
            self.dout = nn.Dropout(0.3)
            self.actvf = nn.LeakyReLU(0.2, inplace=True)
            self.tanh = nn.Tanh()
        self.mano_layer = ManoLayer(rot_mode="axisang", side="right", center_idx=None, use_pca=False,
                                    flat_hand_mean=True, mano_assets_root=mano_assets_root, mano_model=mano_model)
        self._folded = None

    # -- eval-mode BatchNorm folded into the linear layers; rebuilt whenever a parameter / buffer changes
    def _fold(self):
        key = tuple((t.data_ptr(), t._version) for t in list(self.parameters()) + list(self.buffers()))
        if self._folded is not None and self._folded[0] == key:
            return self._folded[1]
        f = {}
        with torch.no_grad():
            f["bn_s"], f["bn_t"] = (t.contiguous() for t in _bn_affine(self.bn1))
            for name in ("rb1", "rb2", "rb3"):
                rb = getattr(self, name)
                s1, t1 = _bn_affine(rb.bn1)
                s2, t2 = _bn_affine(rb.bn2)
                # fc1 (+ bn1) and fc3 read the same input and both end in the LeakyReLU: one GEMM with N = 256 + Fout
                f[name] = dict(W13=torch.cat([rb.fc1.weight.float() * s1[:, None], rb.fc3.weight.float()], 0).contiguous(),
                               b13=torch.cat([rb.fc1.bias.float() * s1 + t1, rb.fc3.bias.float()], 0).contiguous(),
                               W2=(rb.fc2.weight.float() * s2[:, None]).contiguous(), b2=(rb.fc2.bias.float() * s2 + t2).contiguous(),
                               n1=rb.fc1.weight.shape[0])
            f["Wo"] = torch.cat([self.out_p.weight, self.out_t.weight], 0).detach().float().contiguous()
            f["bo"] = torch.cat([self.out_p.bias, self.out_t.bias], 0).detach().float().contiguous()
        self._folded = (key, f)
        return f

    def _resblock(self, w, x, out, hx):
        """ResBlock.forward (refiner.py:306-319): out = ll(ll(fc3 x) + bn2(fc2(ll(bn1(fc1 x))))).  hx [B, 256 + Fout]
        holds ll(bn1(fc1 x)) | ll(fc3 x)."""
        n1 = w["n1"]
        linear_f32(x, w["W13"], w["b13"], hx, leaky=0.2)
        linear_f32(hx[:, :n1], w["W2"], w["b2"], out, residual=hx[:, n1:], leaky=0.2)

    @torch.no_grad()
    def _iterate(self, feat, hand_pose, hand_tsl, cloud, h2o_first=None, rigid=None, offset=None):
        """The refinement loop of refiner.py:259-285 on the feature buffer `feat` [B, 512 + 877]:
        columns 0:512 = ResBlock output X, 512:1290 = normalised h2o distances, 1290:1386 = 6D pose, 1386:1389 =
        translation (so X0 = feat[:, 512:] and cat([X, X0]) = feat, no concatenation copies).
        cloud = kwargs of chamfer_nn (y_points, obj_id, rot).  hand_pose / hand_tsl feed the first MANO forward unless
        `h2o_first` [B,778] is given.  -> (pose [B,48], tsl [B,3], verts, joints) of the final MANO forward, with the rigid
        map x' = rigid (x + tsl + offset) fused into its store when `rigid` is given."""
        f = self._fold()
        B, dev = feat.shape[0], feat.device
        L = lib.load()
        hs, ld = self.h_size, feat.shape[1]
        X, X0, h2o, f6d = feat[:, :hs], feat[:, hs:], feat[:, hs:hs + 778], feat[:, hs + 778:]
        verts = torch.empty((B, 778, 3), device=dev, dtype=torch.float32)
        joints = torch.empty((B, 21, 3), device=dev, dtype=torch.float32)
        pose = torch.empty((B, 48), device=dev, dtype=torch.float32)
        tsl = torch.empty((B, 3), device=dev, dtype=torch.float32)
        post = torch.empty((B, 12), device=dev, dtype=torch.float32)
        hx = torch.empty((B, 256 + hs), device=dev, dtype=torch.float32)
        st = lib.stream_ptr(dev)

        def decode(rg=None, off=None):
            with torch.cuda.device(dev):
                rc = L.ab_refine_decode(B, f6d.data_ptr(), ld, lib.ptr(rg), 0 if rg is None else (16 if rg.shape[-1] == 4 else 9),
                                        lib.ptr(off), lib.ptr(pose), lib.ptr(tsl), lib.ptr(post), st)
            lib.check(rc, "ab_refine_decode")

        for i in range(self.n_iters):
            if i != 0 or h2o_first is None:
                if i == 0:
                    post[:, :9] = torch.eye(3, device=dev).reshape(1, 9)
                    post[:, 9:] = hand_tsl
                    self.mano_layer.lbs_into(hand_pose, None, post, verts, joints)
                else:
                    decode()
                    self.mano_layer.lbs_into(pose, None, post, verts, joints)
                if "groups" in cloud:  # static clouds grouped at setup: same bits, ~20x fewer pairs
                    chamfer_nn_grouped(verts, cloud["groups"], obj_id=cloud.get("obj_id"), rot=cloud.get("rot"),
                                       scale=f["bn_s"], shift=f["bn_t"], out=h2o, return_idx=False)
                else:
                    chamfer_nn(verts, scale=f["bn_s"], shift=f["bn_t"], out=h2o, return_idx=False, **cloud)
            else:
                h2o.copy_(h2o_first * f["bn_s"] + f["bn_t"])
            self._resblock(f["rb1"], X0, X, hx)
            self._resblock(f["rb2"], feat, X, hx)
            self._resblock(f["rb3"], feat, X, hx)
            linear_f32(X, f["Wo"], f["bo"], f6d, residual=f6d)   # init_pose += out_p(X); init_trans += out_t(X)
        decode(rigid, offset)
        self.mano_layer.lbs_into(pose, None, post, verts, joints)
        return pose, tsl, verts, joints

    def new_feat(self, B, device):
        return torch.empty((B, self.h_size + self.in_size), device=device, dtype=torch.float32)

    @torch.no_grad()
    def forward(self, h2o_dist, fpose_rhand_rotmat_f, trans_rhand_f, global_orient_rhand_rotmat_f, verts_object, **kwargs):
        """Reference signature (refiner.py:250-285): -> {"th_pose_coeffs" [B,48], "th_tsl" [B,3]}."""
        lib.require_cuda(h2o_dist, "h2o_dist")
        B = h2o_dist.shape[0]
        feat = self.new_feat(B, h2o_dist.device)
        hs = self.h_size
        feat[:, hs + 778:hs + 784] = global_orient_rhand_rotmat_f[..., :2].reshape(B, -1)
        feat[:, hs + 784:hs + 874] = fpose_rhand_rotmat_f[..., :2].reshape(B, -1)
        feat[:, hs + 874:] = trans_rhand_f
        pose, tsl, _, _ = self._iterate(feat, None, None, dict(y_points=verts_object), h2o_first=h2o_dist.float())
        return {"th_pose_coeffs": pose, "th_tsl": tsl}


@register(reg=Refiner.build_mapping, key="null")
class NullRefine(nn.Module):

    def __init__(self, cfg=None, mano_model=None):
        super().__init__()
        self.cfg = cfg
        self.resampled_objs = []
        self.obj_idx = {}
        self.refine_net = _RefineNet(n_iters=0, mano_model=mano_model)

    def setup(self, obj_meshes: Dict[str, object]):
        pass

    def forward(self, inp, obj_name: List[str] = None):
        hand_pose, hand_tsl = inp["hand_pose"], inp["hand_tsl"]
        mano_out = self.refine_net.mano_layer(hand_pose)  # betas=None (refiner.py:138)
        return {
            "hand_verts": mano_out.verts + hand_tsl.unsqueeze(1),
            "joints": mano_out.joints + hand_tsl.unsqueeze(1),
            "hand_pose": hand_pose,
            "hand_tsl": hand_tsl,
        }


def subdivide_mesh(vertices: np.ndarray, faces: np.ndarray):
    """One level of midpoint subdivision (what trimesh's `Trimesh.subdivide` does, refiner.py:175-176): every triangle
    becomes four, one new vertex per unique edge, appended after the original vertices."""
    vertices, faces = np.asarray(vertices, np.float64), np.asarray(faces, np.int64)
    edges = np.sort(np.concatenate([faces[:, [0, 1]], faces[:, [1, 2]], faces[:, [2, 0]]], 0), axis=1)
    uniq, inv = np.unique(edges, axis=0, return_inverse=True)
    mid = vertices[uniq].mean(axis=1)
    inv = inv.reshape(3, -1).T + len(vertices)  # [F,3] midpoint ids of edges (01, 12, 20)
    f = np.concatenate([np.stack([faces[:, 0], inv[:, 0], inv[:, 2]], 1), np.stack([inv[:, 0], faces[:, 1], inv[:, 1]], 1),
                        np.stack([inv[:, 2], inv[:, 1], faces[:, 2]], 1), np.stack([inv[:, 0], inv[:, 1], inv[:, 2]], 1)], 0)
    return np.concatenate([vertices, mid], 0), f


@register(reg=Refiner.build_mapping, key="hand_obj")
class HORefiner(nn.Module):
    """refiner.py:150-224.  cfg: {"PRETRAINED": path of the GrabNet refinenet.pt or None, "ITERS": n}.  `state_dict`
    may carry the weights directly (tests / synthetic runs: the licensed checkpoint is not redistributable)."""

    def __init__(self, cfg, mano_model=None, state_dict=None):
        super().__init__()
        self.refine_net = _RefineNet(n_iters=cfg["ITERS"], mano_model=mano_model)
        if state_dict is None and cfg.get("PRETRAINED"):
            state_dict = torch.load(cfg["PRETRAINED"], map_location=torch.device("cpu"))
        if state_dict is not None:
            self.refine_net.load_state_dict(state_dict, strict=False)
        self.refine_net.eval()
        self.resampled_objs = []
        self.obj_idx = {}
        # grouped search over the static clouds (ab_chamfer_nn_grouped): bit-identical to the scan as long as `obj_rot` is a
        # rotation, which is what the pose generator passes (preprocessor.py:79).  False: scan the whole cloud (ab_chamfer_nn).
        self.use_groups = True

    def setup(self, obj_meshes: Dict[str, object]):
        for name, m in obj_meshes.items():
            self.obj_idx[name] = len(self.obj_idx)
            self.resampled_objs.append(torch.Tensor(self.resample_obj(m)).float())
        self.resampled_objs = torch.stack(self.resampled_objs)
        self.register_buffer("resampled_objs_buffer", self.resampled_objs)
        # the clouds are static: group them once for ab_chamfer_nn_grouped (<= 480 groups of 32 points)
        if self.use_groups and self.resampled_objs.shape[1] <= 480 * 32:
            sp, pm, bx = build_nn_groups(self.resampled_objs.numpy())
            self.register_buffer("nn_sorted", torch.from_numpy(sp), persistent=False)
            self.register_buffer("nn_perm", torch.from_numpy(pm), persistent=False)
            self.register_buffer("nn_boxes", torch.from_numpy(bx), persistent=False)

    @staticmethod
    def resample_obj(obj_mesh, n_sample_verts: int = 10000):
        """refiner.py:171-182: subdivide until there are enough vertices, then draw n without replacement from
        np.random (seed it for reproducibility, like the reference)."""
        if hasattr(obj_mesh, "subdivide"):  # a real trimesh.Trimesh
            mesh = deepcopy(obj_mesh)
            while mesh.vertices.shape[0] < n_sample_verts:
                mesh = mesh.subdivide()
            verts_obj = np.asarray(mesh.vertices)
        else:
            verts_obj, faces = np.asarray(obj_mesh.vertices), np.asarray(obj_mesh.faces)[:, :3]
            while verts_obj.shape[0] < n_sample_verts:
                verts_obj, faces = subdivide_mesh(verts_obj, faces)
        ids = np.random.choice(verts_obj.shape[0], n_sample_verts, replace=False)
        return verts_obj[ids]

    @torch.no_grad()
    def refine(self, hand_pose, hand_tsl, obj_rot, obj_id, rigid=None, offset=None):
        """Device path: obj_id int [B] indexes `resampled_objs_buffer`; rigid / offset: see _RefineNet._iterate."""
        lib.require_cuda(hand_pose, "hand_pose")
        net = self.refine_net
        B, dev = hand_pose.shape[0], hand_pose.device
        hand_pose, hand_tsl = hand_pose.contiguous().float(), hand_tsl.contiguous().float()
        feat = net.new_feat(B, dev)
        f6d = feat[:, net.h_size + 778:]
        with torch.cuda.device(dev):
            rc = lib.load().ab_refine_encode(B, lib.ptr(hand_pose), lib.ptr(hand_tsl), f6d.data_ptr(), feat.shape[1],
                                             lib.stream_ptr(dev))
        lib.check(rc, "ab_refine_encode")
        cloud = dict(y_points=self.resampled_objs_buffer, obj_id=obj_id, rot=obj_rot)
        if self.use_groups and hasattr(self, "nn_sorted"):
            cloud["groups"] = (self.nn_sorted, self.nn_perm, self.nn_boxes)
        return net._iterate(feat, hand_pose, hand_tsl, cloud, rigid=rigid, offset=offset)

    def forward(self, inp, obj_name: List[str]):
        obj_rot = inp["obj_rot"]
        assert (len(obj_name) == obj_rot.shape[0]), \
            f"object name and rotation matrix do not match, got {len(obj_name)} and {obj_rot.shape[0]}"
        obj_id = torch.tensor([self.obj_idx[name] for name in obj_name], dtype=torch.int32, device=obj_rot.device)
        pose, tsl, verts, joints = self.refine(inp["hand_pose"], inp["hand_tsl"], obj_rot, obj_id)
        return {"hand_verts": verts, "joints": joints, "hand_pose": pose, "hand_tsl": tsl}
