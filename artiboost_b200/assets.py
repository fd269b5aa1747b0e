"""Deterministic synthetic stand-ins for the licensed assets the reference loads from disk.

The reference needs MANO_RIGHT.pkl (docs/Installation.md:175-176), YCB meshes
(anakin/artiboost/object_engine.py:44-63), per-object grasp pickles
(anakin/artiboost/grasp_engine.py:17,28-32), 51 HTML hand textures
(anakin/artiboost/hand_texture.py:7-10) and background photos
(anakin/utils/renderer.py:139-149).  None of these exist on the build or GPU box, so the
benchmarks and parity tests run on procedurally generated assets with the SAME shapes:

* hand model: V=778, F=1538 (disk topology, 16-edge wrist boundary like MANO), 16 joints with
  the MANO kinematic tree, shapedirs 778x3x10, posedirs 778x3x135, J_regressor 16x778,
  skinning weights 778x16.
* objects: 21 closed meshes named after CONST.YCB_IDX2CLASSES (anakin/utils/misc.py:97-119),
  V=8192, F=16380 each, bbox-centred like object_engine.py:50-54, plus 8 bbox corners.
* grasp tables: per object a list of (hand_pose[48], hand_shape[10] | None, hand_tsl[3]).

`load_mano_pkl` reads a real MANO pickle into the same dict when a user has one.
"""
from __future__ import annotations

import pickle
from typing import Dict, List, Optional, Tuple

import numpy as np

MANO_PARENTS = [-1, 0, 1, 2, 0, 4, 5, 0, 7, 8, 0, 10, 11, 0, 13, 14]
MANO_TIP_VERTS = [745, 317, 444, 556, 673]
MANO_JOINT_REORDER = [0, 13, 14, 15, 16, 1, 2, 3, 17, 4, 5, 6, 18, 10, 11, 12, 19, 7, 8, 9, 20]

YCB_NAMES = [
    "002_master_chef_can", "003_cracker_box", "004_sugar_box", "005_tomato_soup_can", "006_mustard_bottle",
    "007_tuna_fish_can", "008_pudding_box", "009_gelatin_box", "010_potted_meat_can", "011_banana",
    "019_pitcher_base", "021_bleach_cleanser", "024_bowl", "025_mug", "035_power_drill", "036_wood_block",
    "037_scissors", "040_large_marker", "051_large_clamp", "052_extra_large_clamp", "061_foam_brick",
]
HO3D_TRAIN_OBJS = ["010_potted_meat_can", "021_bleach_cleanser", "006_mustard_bottle", "019_pitcher_base"]


def _strip(ring_a: List[int], ang_a: np.ndarray, ring_b: List[int], ang_b: np.ndarray) -> List[Tuple[int, int, int]]:
    """Triangulate the band between two closed rings (possibly different sizes), a -> outer, b -> inner."""
    na, nb = len(ring_a), len(ring_b)
    faces = []
    ia = ib = 0
    while ia < na or ib < nb:
        a0, b0 = ring_a[ia % na], ring_b[ib % nb]
        next_a = ang_a[(ia + 1) % na] + (2 * np.pi if ia + 1 >= na else 0.0)
        next_b = ang_b[(ib + 1) % nb] + (2 * np.pi if ib + 1 >= nb else 0.0)
        if ib >= nb or (ia < na and next_a <= next_b):
            faces.append((a0, ring_a[(ia + 1) % na], b0))
            ia += 1
        else:
            faces.append((a0, ring_b[(ib + 1) % nb], b0))
            ib += 1
    return faces


def make_synthetic_mano(seed: int = 0) -> Dict[str, np.ndarray]:
    """MANO-shaped right-hand model (mitten surface) with real MANO dimensions.  float64 arrays."""
    rng = np.random.RandomState(seed)
    n_around, n_rings, n_inner = 16, 48, 9
    # ---- template surface: flattened tube along +x, wrist at x=-0.01, tip at x~0.18
    verts = []
    xs = np.linspace(-0.01, 0.165, n_rings)
    for i, x in enumerate(xs):
        t = i / (n_rings - 1)
        half_w = 0.042 * (1.0 - 0.55 * t ** 2.5) * (0.75 + 0.25 * min(1.0, t * 6))
        half_h = 0.014 * (1.0 - 0.45 * t ** 2)
        for k in range(n_around):
            a = 2 * np.pi * k / n_around
            verts.append([x, half_h * np.sin(a), half_w * np.cos(a)])
    ang_outer = 2 * np.pi * np.arange(n_around) / n_around
    ang_inner = 2 * np.pi * (np.arange(n_inner) + 0.5) / n_inner
    for a in ang_inner:
        verts.append([0.172, 0.006 * np.sin(a), 0.009 * np.cos(a)])
    verts.append([0.178, 0.0, 0.0])
    verts = np.asarray(verts, dtype=np.float64)
    assert verts.shape == (778, 3)
    faces = []
    for i in range(n_rings - 1):
        for k in range(n_around):
            a, b = i * n_around + k, i * n_around + (k + 1) % n_around
            c, d = (i + 1) * n_around + k, (i + 1) * n_around + (k + 1) % n_around
            faces += [(a, b, c), (b, d, c)]
    last = [(n_rings - 1) * n_around + k for k in range(n_around)]
    inner = [n_rings * n_around + k for k in range(n_inner)]
    faces += _strip(last, ang_outer, inner, ang_inner)
    apex = 777
    for k in range(n_inner):
        faces.append((inner[k], inner[(k + 1) % n_inner], apex))
    faces = np.asarray(faces, dtype=np.int64)
    assert faces.shape == (1538, 3), faces.shape
    # make every face CCW seen from outside (normal away from the x axis / towards +x at the cap)
    p0, p1, p2 = verts[faces[:, 0]], verts[faces[:, 1]], verts[faces[:, 2]]
    nrm = np.cross(p1 - p0, p2 - p0)
    cen = (p0 + p1 + p2) / 3.0
    out_dir = cen * np.array([0.0, 1.0, 1.0]) + np.array([1e-3, 0, 0]) * (cen[:, :1] > 0.166)
    flip = (nrm * out_dir).sum(1) < 0
    faces[flip] = faces[flip][:, [0, 2, 1]]
    # shuffle vertex order so the MANO fingertip vertex ids land on scattered surface points
    perm = rng.permutation(778)
    inv = np.argsort(perm)
    verts = verts[perm]
    faces = inv[faces]

    # ---- 16-joint skeleton in MANO order (wrist; index; middle; pinky; ring; thumb)
    def finger(base, direction, lens):
        direction = np.asarray(direction, dtype=np.float64)
        direction /= np.linalg.norm(direction)
        pts, p = [], np.asarray(base, dtype=np.float64)
        for ln in [0.0] + list(lens[:-1]):
            p = p + direction * ln
            pts.append(p.copy())
        return pts

    joints_t = [np.zeros(3)]
    joints_t += finger([0.088, 0.0, 0.028], [1, 0, 0.10], [0.036, 0.024, 0.02])   # index
    joints_t += finger([0.092, 0.0, 0.008], [1, 0, 0.02], [0.040, 0.027, 0.02])   # middle
    joints_t += finger([0.078, 0.0, -0.030], [1, 0, -0.15], [0.028, 0.018, 0.02])  # pinky
    joints_t += finger([0.086, 0.0, -0.012], [1, 0, -0.06], [0.036, 0.024, 0.02])  # ring
    joints_t += finger([0.025, 0.004, 0.030], [0.7, 0, 0.7], [0.034, 0.028, 0.02])  # thumb
    joints_t = np.asarray(joints_t)
    assert joints_t.shape == (16, 3)

    # ---- J_regressor: each joint = convex combination of its 12 nearest template vertices
    J_reg = np.zeros((16, 778))
    for j in range(16):
        d = np.linalg.norm(verts - joints_t[j], axis=1)
        idx = np.argsort(d)[:12]
        w = np.exp(-d[idx] / 0.01)
        J_reg[j, idx] = w / w.sum()

    # ---- skinning weights: softmax of -distance to bone segment, top-4, renormalised
    j_real = J_reg @ verts
    child_of = {p: [] for p in range(16)}
    for c, p in enumerate(MANO_PARENTS):
        if p >= 0:
            child_of[p].append(c)
    dist = np.zeros((778, 16))
    for j in range(16):
        a = j_real[j]
        ends = [j_real[c] for c in child_of[j]] or [a + (a - j_real[MANO_PARENTS[j]]) * 0.8]
        dj = np.full(778, np.inf)
        for b in ends:
            ab = b - a
            t = np.clip(((verts - a) @ ab) / max(ab @ ab, 1e-12), 0, 1)
            dj = np.minimum(dj, np.linalg.norm(verts - (a + t[:, None] * ab), axis=1))
        dist[:, j] = dj
    w = np.exp(-dist / 0.008)
    kth = np.sort(w, axis=1)[:, -4][:, None]
    w = np.where(w >= kth, w, 0.0)
    weights = w / w.sum(1, keepdims=True)

    # ---- blend shapes: smooth low-amplitude fields
    def smooth_field(n_out, amp):
        centres = rng.uniform([-0.01, -0.02, -0.05], [0.18, 0.02, 0.05], size=(6, 3))
        coef = rng.normal(0, 1, size=(6, 3, n_out))
        rbf = np.exp(-((verts[:, None, :] - centres[None]) ** 2).sum(-1) / (2 * 0.04 ** 2))  # 778x6
        return amp * np.einsum("vc,cdk->vdk", rbf, coef)

    shapedirs = smooth_field(10, 0.003)
    posedirs = smooth_field(135, 0.0008)
    return {
        "v_template": verts,
        "f": faces.astype(np.int64),
        "shapedirs": shapedirs,
        "posedirs": posedirs,
        "J_regressor": J_reg,
        "weights": weights,
        "kintree_table": np.stack([np.array([2 ** 32 - 1] + MANO_PARENTS[1:], dtype=np.int64),
                                   np.arange(16, dtype=np.int64)]),
        "hands_components": np.eye(45),
        "hands_mean": np.zeros(45),
    }


def load_mano_pkl(path: str) -> Dict[str, np.ndarray]:
    """Read a real MANO_{RIGHT,LEFT}.pkl (needs `chumpy` importable only if the pickle holds chumpy arrays)."""
    with open(path, "rb") as f:
        dd = pickle.load(f, encoding="latin1")
    out = {}
    for k in ["v_template", "f", "shapedirs", "posedirs", "weights", "kintree_table", "hands_components",
              "hands_mean"]:
        out[k] = np.array(dd[k])
    jr = dd["J_regressor"]
    out["J_regressor"] = np.array(jr.toarray() if hasattr(jr, "toarray") else jr)
    return out


def dump_mano_pkl(model: Dict[str, np.ndarray], path: str) -> None:
    """Write `model` in the on-disk MANO layout (scipy-sparse J_regressor) the reference's loaders expect
    (anakin/postprocess/iknet/manolayer.py:57-66,81-110)."""
    import scipy.sparse as sp

    dd = dict(model)
    dd["J_regressor"] = sp.csc_matrix(model["J_regressor"])
    dd["kintree_table"] = model["kintree_table"]
    with open(path, "wb") as f:
        pickle.dump(dd, f, protocol=2)


# --------------------------------------------------------------------------------------------- objects
def _superquadric(a, e1, e2, rings=90, segs=91, bend=0.0, taper=0.0):
    def spow(c, e):
        return np.sign(c) * np.abs(c) ** e

    verts = [[0.0, 0.0, a[2]]]
    for i in range(1, rings + 1):
        v = np.pi / 2 - np.pi * i / (rings + 1)
        for k in range(segs):
            u = 2 * np.pi * k / segs
            verts.append([a[0] * spow(np.cos(v), e1) * spow(np.cos(u), e2),
                          a[1] * spow(np.cos(v), e1) * spow(np.sin(u), e2),
                          a[2] * spow(np.sin(v), e1)])
    verts.append([0.0, 0.0, -a[2]])
    verts = np.asarray(verts)
    zt = verts[:, 2] / a[2]
    verts[:, :2] *= (1.0 + taper * zt)[:, None]
    verts[:, 0] += bend * a[2] * zt ** 2
    faces = []
    top, bot = 0, len(verts) - 1
    ring = lambda i, k: 1 + (i - 1) * segs + (k % segs)  # noqa: E731
    for k in range(segs):
        faces.append((top, ring(1, k), ring(1, k + 1)))
    for i in range(1, rings):
        for k in range(segs):
            a0, b0, c0, d0 = ring(i, k), ring(i, k + 1), ring(i + 1, k), ring(i + 1, k + 1)
            faces += [(a0, c0, b0), (b0, c0, d0)]
    for k in range(segs):
        faces.append((bot, ring(rings, k + 1), ring(rings, k)))
    return verts, np.asarray(faces, dtype=np.int64)


def signed_volume(verts, faces):
    p0, p1, p2 = verts[faces[:, 0]], verts[faces[:, 1]], verts[faces[:, 2]]
    return float((np.cross(p0, p1) * p2).sum() / 6.0)


def make_synthetic_objects(names: Optional[List[str]] = None, seed: int = 0):
    """-> dict name -> {"vertices" f64[8192,3] bbox-centred, "faces" i64[16380,3] CCW-outward,
    "colors" u8[8192,3], "corners_can" f64[8,3]}."""
    names = list(names) if names is not None else list(YCB_NAMES)
    out = {}
    for name in names:
        idx = YCB_NAMES.index(name) if name in YCB_NAMES else (abs(hash(name)) % 1000)
        rng = np.random.RandomState(seed * 1000 + idx)
        ext = rng.uniform(0.035, 0.09, size=3)  # half extents; rotated bbox lands near 0.1 .. 0.25 m
        ext[2] = max(ext[2], ext[:2].max())
        e1, e2 = rng.choice([0.15, 0.3, 1.0], p=[0.4, 0.3, 0.3]), rng.choice([0.15, 1.0], p=[0.45, 0.55])
        v, f = _superquadric(ext, e1, e2, bend=rng.uniform(-0.15, 0.15), taper=rng.uniform(-0.25, 0.1))
        if signed_volume(v, f) < 0:
            f = f[:, [0, 2, 1]]
        rot = _rand_rotation(rng)
        v = v @ rot.T
        centre = (v.min(0) + v.max(0)) / 2  # transform.center_vert_bbox (transform.py:621-631)
        v = v - centre
        lo, hi = v.min(0), v.max(0)
        corners = np.array([[x, y, z] for x in (lo[0], hi[0]) for y in (lo[1], hi[1]) for z in (lo[2], hi[2])])
        base = rng.uniform(60, 230, size=3)
        stripe = (np.sin(v[:, 2] * rng.uniform(60, 160)) > 0.2)[:, None]
        label = ((np.abs(v[:, 0]) < 0.5 * hi[0]) & (v[:, 1] > 0))[:, None]
        col = base[None] * np.where(stripe, 1.0, 0.65) * np.where(label, [1.1, 0.6, 0.5], 1.0)
        col = np.clip(col + rng.normal(0, 4, size=col.shape), 0, 255).astype(np.uint8)
        assert v.shape == (8192, 3) and f.shape == (16380, 3)
        out[name] = {"vertices": v, "faces": f, "colors": col, "corners_can": corners}
    return out


def _rand_rotation(rng) -> np.ndarray:
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def _rotmat_to_aa_np(R: np.ndarray) -> np.ndarray:
    ang = np.arccos(np.clip((np.trace(R) - 1) / 2, -1, 1))
    if ang < 1e-8:
        return np.zeros(3)
    ax = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]]) / (2 * np.sin(ang))
    return ax * ang


def make_synthetic_grasps(objects: Dict[str, dict], n_grasp: int = 50, seed: int = 0):
    """-> dict name -> list of (hand_pose f64[48], hand_shape f64[10] | None, hand_tsl f64[3]) in the object's
    canonical frame (the layout grasp_engine.py:47-53 unpacks)."""
    out = {}
    for oi, (name, obj) in enumerate(objects.items()):
        rng = np.random.RandomState(seed * 7919 + oi)
        half = np.abs(obj["vertices"]).max(0)
        grasps = []
        for g in range(n_grasp):
            pose = rng.normal(0, 0.12, size=48)
            pose[3:] += np.tile([0.0, 0.0, -0.35], 15) * rng.uniform(0.3, 1.0)  # curl
            R = _rand_rotation(rng)
            pose[:3] = _rotmat_to_aa_np(R)
            d = rng.normal(size=3)
            d /= np.linalg.norm(d)
            surf = d * half
            # put the palm centre (~0.09 m along the hand's +x) near the object surface
            tsl = surf * 1.15 - R @ np.array([0.09, 0.0, 0.0])
            shape = rng.normal(0, 1, size=10) if (g % 3) else None
            grasps.append((pose, shape, tsl))
        out[name] = grasps
    return out


def make_hand_textures(n_tex: int = 51, seed: int = 0, template: Optional[np.ndarray] = None) -> np.ndarray:
    """-> u8[n_tex,778,3] per-vertex RGB; stands in for the 51 HTML textures (hand_texture.py:7-10)."""
    rng = np.random.RandomState(seed + 17)
    tones = np.array([[224, 172, 150], [198, 134, 110], [141, 85, 62], [255, 219, 190], [105, 64, 48]], dtype=np.float64)
    out = np.zeros((n_tex, 778, 3), dtype=np.uint8)
    for t in range(n_tex):
        base = tones[t % len(tones)] * rng.uniform(0.85, 1.1)
        mod = 1.0 + 0.08 * rng.normal(size=(778, 1))
        if template is not None:
            mod = mod + 0.1 * np.sin(template[:, :1] * 90.0 + t)
        out[t] = np.clip(base[None] * mod, 0, 255).astype(np.uint8)
    return out


def make_backgrounds(n_bg: int = 8, height: int = 384, width: int = 384, seed: int = 0) -> np.ndarray:
    """-> u8[n_bg,H,W,3] RGB.  1.5x the render size like renderer.py:98-99."""
    rng = np.random.RandomState(seed + 29)
    yy, xx = np.mgrid[0:height, 0:width].astype(np.float64)
    out = np.zeros((n_bg, height, width, 3), dtype=np.uint8)
    for b in range(n_bg):
        img = np.zeros((height, width, 3))
        for c in range(3):
            f = rng.uniform(0.005, 0.05, size=2)
            ph = rng.uniform(0, 6.28, size=2)
            img[..., c] = 128 + 70 * np.sin(xx * f[0] + ph[0]) * np.cos(yy * f[1] + ph[1]) + rng.uniform(-30, 30)
        img += rng.normal(0, 6, size=img.shape)
        out[b] = np.clip(img, 0, 255).astype(np.uint8)
    return out
