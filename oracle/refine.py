"""ORACLE (test infrastructure, never imported by the product path).

CPU/NumPy restatement of the hand-object refiner and the anatomical scramblers:

  point2point_signed / chamfer NN   anakin/artiboost/refiner.py:21-85 over the third-party `chamfer_distance`
                                    (git+https://github.com/KailinLi/chamfer_distance.git, UNPINNED, requirements.txt:179;
                                    absent here).  Its published kernel (NmDistanceKernel) scans the target cloud in
                                    index order, keeps the squared distance (x1-x2)^2+(y1-y2)^2+(z1-z2)^2 and replaces the
                                    best on strict `<`, so the FIRST minimum wins; the reference then gathers the winner
                                    and takes the norm of the difference (:59-81).
  CRot2rotmat / parms_decode        refiner.py:88-107
  _RefineNet.forward, ResBlock      refiner.py:250-319 (eval mode: Dropout = identity, BatchNorm = running statistics)
  HORefiner.forward / resample_obj  refiner.py:171-224
  AxisLayer                         third-party manotorch (UNPINNED, absent): [recalled] -- back axis = joint - child
                                    keypoint rotated into the joint frame by transforms_abs[1:]^T, left = back x up_base
                                    (up_base = (0,1,0) for the 12 finger joints, (1,1,1) for the thumb's), up = left x back,
                                    all normalised; joints listed in MANO chain order.  parity unpinned for this piece.
  axis_angle_op, RandomScrambler2/3 anakin/artiboost/scrambler.py:19-28,84-260

Pinned by tests/golden/{refiner,scrambler23,preprocessor_staged}.npz: outputs of the reference's own refiner.py /
scrambler.py / preprocessor.py run in the build container (tests/golden/make_golden_refine.py) with the absent
third-party pieces (chamfer_distance, AxisLayer, manotorch, pytorch3d) replaced by the restatements in this package,
so the pins cover the reference's composition, iteration structure, indexing and BatchNorm / residual wiring.

Arithmetic notes.  The nearest-neighbour search is fp32 with the exact operation order of the CUDA kernel
(d = fma(dz,dz, fma(dy,dy, dx*dx)); fma emulated through float64, which is exact for the product and differs from a
true fused rounding only in double-rounding corner cases), so indices and distances compare bit for bit.
"""
import numpy as np

from . import rotations as rot
from .mano_lbs import ManoLayer

f32 = np.float32
AXIS_JOINTS = [5, 6, 7, 9, 10, 11, 17, 18, 19, 13, 14, 15, 1, 2, 3]


def _fma(a, b, c):
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(f32)


def rotate_cloud(R, pts):
    """y = R o in fp32, ((R0 ox + R1 oy) + R2 oz) with every operation rounded (refiner.py:190-191's bmm)."""
    R, pts = np.asarray(R, f32), np.asarray(pts, f32)
    ox, oy, oz = pts[:, 0], pts[:, 1], pts[:, 2]
    return np.stack([(R[i, 0] * ox + R[i, 1] * oy) + R[i, 2] * oz for i in range(3)], 1).astype(f32)


def chamfer_nn(x, y, chunk=128):
    """x [P1,3], y [P2,3] fp32 -> (dist [P1] = |x - y_nn|, idx [P1]); first minimum wins."""
    x, y = np.asarray(x, f32), np.asarray(y, f32)
    dist, idx = np.empty(len(x), f32), np.empty(len(x), np.int64)
    for s in range(0, len(x), chunk):
        q = x[s:s + chunk]
        dx, dy, dz = (q[:, None, k] - y[None, :, k] for k in range(3))
        d = _fma(dz, dz, _fma(dy, dy, dx * dx))
        i = d.argmin(axis=1)
        idx[s:s + chunk] = i
        dist[s:s + chunk] = np.sqrt(d[np.arange(len(q)), i])
    return dist, idx


def point2point_signed(x, y):
    """Batched refiner.py:21-85 without normals: [N,P1,3], [N,P2,3] -> [N,P1]."""
    return np.stack([chamfer_nn(a, b)[0] for a, b in zip(x, y)])


def crot2rotmat(pose6d):
    a = np.asarray(pose6d, f32).reshape(-1, 3, 2)
    n1 = np.maximum(np.linalg.norm(a[:, :, 0], axis=1, keepdims=True), f32(1e-12))
    b1 = a[:, :, 0] / n1
    dot = np.sum(b1 * a[:, :, 1], axis=1, keepdims=True)
    u = a[:, :, 1] - dot * b1
    b2 = u / np.maximum(np.linalg.norm(u, axis=1, keepdims=True), f32(1e-12))
    b3 = np.cross(b1, b2)
    return np.stack([b1, b2, b3], axis=-1).astype(f32)


def parms_decode(pose6d, trans):
    bs = trans.shape[0]
    return rot.rotmat_to_aa(crot2rotmat(pose6d)).reshape(bs, -1).astype(f32), trans


def _leaky(x, slope=0.2):
    return np.where(x > 0, x, x * f32(slope)).astype(f32)


class RefineNet:
    """_RefineNet in eval mode over a state dict of numpy arrays (reference parameter names)."""

    def __init__(self, state, mano_model, n_iters=3):
        self.s = {k: np.asarray(v, f32) for k, v in state.items()}
        self.n_iters = n_iters
        self.mano = ManoLayer(mano_model, center_idx=None, dtype=f32)

    def _bn(self, name, x, eps=1e-5):
        s = self.s
        return ((x - s[name + ".running_mean"]) / np.sqrt(s[name + ".running_var"] + f32(eps)) * s[name + ".weight"]
                + s[name + ".bias"]).astype(f32)

    def _lin(self, name, x):
        return (x @ self.s[name + ".weight"].T + self.s[name + ".bias"]).astype(f32)

    def _rb(self, name, x):
        xin = _leaky(self._lin(name + ".fc3", x))
        h = _leaky(self._bn(name + ".bn1", self._lin(name + ".fc1", x)))
        out = self._bn(name + ".bn2", self._lin(name + ".fc2", h))
        return _leaky(xin + out)

    def forward(self, h2o, rel_rotmat, trans, glob_rotmat, verts_object):
        bs = h2o.shape[0]
        init_pose = np.concatenate([glob_rotmat[..., :2].reshape(bs, -1), rel_rotmat[..., :2].reshape(bs, -1)], 1).astype(f32)
        init_trans = np.asarray(trans, f32)
        for i in range(self.n_iters):
            if i != 0:
                pose, tsl = parms_decode(init_pose, init_trans)
                verts = self.mano(pose).verts + tsl[:, None]
                h2o = point2point_signed(verts, verts_object)
            h = self._bn("bn1", h2o)
            X0 = np.concatenate([h, init_pose, init_trans], 1)
            X = self._rb("rb1", X0)
            X = self._rb("rb2", np.concatenate([X, X0], 1))
            X = self._rb("rb3", np.concatenate([X, X0], 1))
            init_trans = init_trans + self._lin("out_t", X)
            init_pose = init_pose + self._lin("out_p", X)
        return parms_decode(init_pose, init_trans)


def ho_refiner(net: RefineNet, resampled_objs, obj_ids, hand_pose, hand_tsl, obj_rot):
    """HORefiner.forward (refiner.py:184-224) -> dict(hand_verts, joints, hand_pose, hand_tsl, h2o)."""
    hand_pose, hand_tsl = np.asarray(hand_pose, f32), np.asarray(hand_tsl, f32)
    bs = hand_pose.shape[0]
    R = rot.aa_to_rotmat(hand_pose.reshape(bs, 16, 3)).astype(f32)
    verts = net.mano(hand_pose).verts + hand_tsl[:, None]
    verts_object = np.stack([rotate_cloud(obj_rot[b], resampled_objs[int(obj_ids[b])]) for b in range(bs)])
    h2o = np.abs(point2point_signed(verts, verts_object))
    pose, tsl = net.forward(h2o, R[:, 1:], hand_tsl, R[:, 0], verts_object)
    out = net.mano(pose)
    return {"hand_verts": (out.verts + tsl[:, None]).astype(f32), "joints": (out.joints + tsl[:, None]).astype(f32),
            "hand_pose": pose, "hand_tsl": tsl, "h2o": h2o}


def subdivide(vertices, faces):
    """Midpoint subdivision as trimesh's Trimesh.subdivide does it [recalled]: one new vertex per unique edge, appended."""
    vertices, faces = np.asarray(vertices, np.float64), np.asarray(faces, np.int64)
    edges = np.sort(np.concatenate([faces[:, [0, 1]], faces[:, [1, 2]], faces[:, [2, 0]]], 0), axis=1)
    uniq = np.unique(edges, axis=0)
    return np.concatenate([vertices, vertices[uniq].mean(axis=1)], 0)


# --------------------------------------------------------------------------------------------- anatomical scramblers
def axis_layer(joints, transf):
    """-> b_axis, u_axis, l_axis, each [B,15,3]."""
    joints, transf = np.asarray(joints, f32), np.asarray(transf, f32)
    b = joints[:, AXIS_JOINTS] - joints[:, [i + 1 for i in AXIS_JOINTS]]
    b = np.einsum("bkji,bkj->bki", transf[:, 1:, :3, :3], b).astype(f32)  # R^T d
    up = np.concatenate([np.tile([[0, 1, 0]], (12, 1)), np.tile([[1, 1, 1]], (3, 1))]).astype(f32)[None]
    l = np.cross(b, np.broadcast_to(up, b.shape)).astype(f32)
    u = np.cross(l, b).astype(f32)
    n = lambda v: (v / np.linalg.norm(v, axis=2, keepdims=True)).astype(f32)  # noqa: E731
    return n(b), n(u), n(l)


def axis_angle_op(aa1, aa2):
    """scrambler.py:19-28: aa of R(aa1) R(aa2)."""
    return rot.rotmat_to_aa(rot.aa_to_rotmat(np.asarray(aa1, f32)) @ rot.aa_to_rotmat(np.asarray(aa2, f32))).astype(f32)


def _thumb_base(hp, l_axis, u_axis, thumb):
    ob = l_axis[:, (12,)] * thumb[:, (0,), None]
    osp = u_axis[:, (12,)] * thumb[:, (1,), None]
    hp[:, (13,)] = axis_angle_op(osp, axis_angle_op(ob, hp[:, (13,)].copy()))


def random_scrambler_2(hand_pose, joints, transf, splay, bend5, thumb, coef=(1.0, 1.1, 0.9)):
    """RandomScrambler2.forward (scrambler.py:96-188) with the draws passed in: splay [B,4], bend5 [B,5], thumb [B,2]."""
    _, u_axis, l_axis = axis_layer(joints, transf)
    hp = np.array(hand_pose, f32).reshape(-1, 16, 3)
    splay, bend5, thumb = (np.asarray(a, f32) for a in (splay, bend5, thumb))
    sj = (1, 4, 7, 10)
    hp[:, sj] = axis_angle_op(hp[:, sj].copy(), u_axis[:, (0, 3, 6, 9)] * splay[:, :, None])
    link = np.asarray(coef, f32)[None]
    for k, ax, jt in ((0, (0, 1, 2), (1, 2, 3)), (1, (3, 4, 5), (4, 5, 6)), (2, (9, 10, 11), (10, 11, 12)),
                      (3, (6, 7, 8), (7, 8, 9))):
        ang = bend5[:, k:k + 1] * link
        hp[:, jt] = axis_angle_op(l_axis[:, ax] * ang[:, :, None], hp[:, jt].copy())
    ang = bend5[:, 4:5] * link[:, (0, 2)]
    hp[:, (14, 15)] = axis_angle_op(l_axis[:, (13, 14)] * ang[:, :, None], hp[:, (14, 15)].copy())
    _thumb_base(hp, l_axis, u_axis, thumb)
    return hp.reshape(-1, 48)


def random_scrambler_3(hand_pose, joints, transf, splay, bend14, thumb):
    """RandomScrambler3.forward (scrambler.py:201-260)."""
    _, u_axis, l_axis = axis_layer(joints, transf)
    hp = np.array(hand_pose, f32).reshape(-1, 16, 3)
    splay, bend14, thumb = (np.asarray(a, f32) for a in (splay, bend14, thumb))
    sj = (1, 4, 7, 10)
    hp[:, sj] = axis_angle_op(hp[:, sj].copy(), u_axis[:, (0, 3, 6, 9)] * splay[:, :, None])
    ax = (0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 14)
    jt = (1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 14, 15)
    hp[:, jt] = axis_angle_op(l_axis[:, ax] * bend14[:, :, None], hp[:, jt].copy())
    _thumb_base(hp, l_axis, u_axis, thumb)
    return hp.reshape(-1, 48)
