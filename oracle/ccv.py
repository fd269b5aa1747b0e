"""ORACLE (test infrastructure, never imported by the product path).

CPU restatement of the CCV-space sampler, the view engine and the pose generator:

  sample_ovg          anakin/artiboost/ovg_set.py:104-132,161-170 (Categorical over the flat weight map,
                      row_col_calc un-flattening, occurrence count map :172-178)
  view_from_id        anakin/artiboost/view_engine.py:17-86
  random_scrambler    anakin/artiboost/scrambler.py:65-81
  pose_generator      anakin/artiboost/preprocessor.py:20-99 with NullRefine (refiner.py:131-147)
  update_method_1..4  anakin/artiboost/artiboost_loader.py:503-598

Random draws are explicit inputs (the reference mixes np.random / torch.rand / torch.distributions across
processes, so stream parity is impossible; SURVEY.md section 7 "RNG parity").  Pinned by tests/golden/*.npz made by
running the reference's own view_engine.py / ovg_set.py / scrambler.py / preprocessor.py (see
tests/golden/make_golden.py; the MANO layer and the pytorch3d rotations underneath preprocessor.py are
third-party and are shimmed with oracle code there, so for preprocessor.py the pin covers the composition only).
"""
import numpy as np

from . import rotations as rot
from .mano_lbs import ManoLayer


def sample_ovg(weight_map, uniforms):
    """Inverse-CDF categorical draw: idx = first i with cdf[i] > u * total.  -> (obj, persp, grasp) int64."""
    w = np.asarray(weight_map, np.float32)
    n_obj, n_persp, n_grasp = w.shape
    cdf = np.cumsum(w.reshape(-1).astype(np.float64))
    flat = np.searchsorted(cdf, np.asarray(uniforms, np.float64) * cdf[-1], side="right")
    flat = np.minimum(flat, w.size - 1).astype(np.int64)
    return row_col_calc(flat, n_persp, n_grasp)


def row_col_calc(tidx, n_row, n_col):
    tidx = np.asarray(tidx, np.int64)
    return tidx // (n_row * n_col), (tidx // n_col) % n_row, tidx % n_col


def occurrence_count_map(bidx, ridx, cidx, n_b, n_r, n_c):
    res = np.zeros((n_b, n_r, n_c), np.int64)
    np.add.at(res, (bidx, ridx, cidx), 1)
    return res


def align_mat(vec):
    """Rotation taking +z onto `vec`: I + [k]x + [k]x^2/(1 + z.v), k = z cross v (view_engine.py:60-86).
    `vec` is normalised in its own dtype (fp32 on the hot path), the matrix algebra is fp64 (numpy promotes the
    int64 z-axis arrays of the reference with fp32 to fp64)."""
    vec = np.asarray(vec)
    vec = (vec / np.linalg.norm(vec)).astype(np.float64)
    k = np.array([-vec[1], vec[0], 0.0])
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    d = vec[2]
    if d == -1:
        return -np.eye(3)
    if d == 1:
        return np.eye(3)
    return np.eye(3) + K + (K @ K) / (1 + d)


def view_from_id(persp_id, u_bins, theta_bins, z_range, r_u, r_theta, r_roll, r_z):
    """r_* are U[0,1) draws.  -> persp_rotmat f32[3,3], camera_free_transf f32[4,4], z_offset f32[3].

    Precision follows the reference when `persp_id` is a 0-dim torch tensor (as OVGSet.__getitem__ passes it,
    ovg_set.py:138,141): `u_id * u_unit` is then an fp32 tensor, so the bin centre, the jittered u / theta, the
    direction vector and its normalisation are all fp32 (view_engine.py:36-58); only caculate_align_mat's matrix
    algebra is fp64."""
    f32 = np.float32
    u_id, theta_id = persp_id // theta_bins, persp_id % theta_bins
    u_unit, theta_unit = 2 / u_bins, (2 * np.pi) / theta_bins
    u_center = f32(-1 + u_unit / 2) + f32(u_id) * f32(u_unit)
    theta_center = f32(theta_unit / 2) + f32(theta_id) * f32(theta_unit)
    # torch.rand(1) - 0.5 is an fp32 subtraction before float(); the product with the unit is a python float
    u_offset = float(f32(r_u) - f32(0.5)) * u_unit
    theta_offset = float(f32(r_theta) - f32(0.5)) * theta_unit
    u = np.clip(u_center + f32(u_offset), f32(-1), f32(1))
    theta = np.clip(theta_center + f32(theta_offset), f32(0), f32(2 * np.pi))
    s = np.sqrt(f32(1) - u * u)
    rotmat = align_mat(np.array([s * np.cos(theta), s * np.sin(theta), u], dtype=f32))
    roll = r_roll * (2 * np.pi)
    free = np.eye(4)
    free[:3, :3] = [[np.cos(roll), -np.sin(roll), 0], [np.sin(roll), np.cos(roll), 0], [0, 0, 1]]
    # torch Uniform.sample: low + rand * (high - low) on fp32 tensors (view_engine.py:15,29)
    z = f32(z_range[0]) + f32(r_z) * (f32(z_range[1]) - f32(z_range[0]))
    return rotmat.astype(f32), free.astype(f32), np.array([0, 0, z], f32)


def random_scrambler(hand_pose, hand_tsl, n_tsl, n_angle):
    """n_tsl [B,3] ~ N(0,sigma_tsl), n_angle [B,16] ~ N(0,sigma_pose), already scaled."""
    p = np.asarray(hand_pose, np.float32).reshape(-1, 16, 3)
    nrm = np.linalg.norm(p, axis=-1, keepdims=True)
    axis = p / np.maximum(nrm, np.float32(1e-7))
    ang = nrm[..., 0] + np.asarray(n_angle, np.float32)
    return (axis * ang[..., None]).reshape(-1, 48).astype(np.float32), (hand_tsl + n_tsl).astype(np.float32)


def pose_generator(mano_model, hand_pose, hand_shape, hand_tsl, persp_rotmat, camera_free_transf, z_offset,
                   n_tsl=None, n_angle=None, scrambler=None, refiner=None):
    """-> dict(final_obj_pose [B,4,4], final_hand_verts [B,778,3], final_joints [B,21,3], hand_pose, hand_tsl).

    scrambler: optional callable(feed dict of preprocessor.py:66-73) -> (hand_pose, hand_tsl), used instead of the
    `random` scrambler (anatomical scramblers, oracle/refine.py); refiner: optional callable(hand_pose, hand_tsl,
    obj_rot) -> dict(hand_verts, joints) used instead of NullRefine (preprocessor.py:76-82)."""
    f32 = np.float32
    hand_pose, hand_shape, hand_tsl = (np.asarray(x, f32) for x in (hand_pose, hand_shape, hand_tsl))
    Rv = np.asarray(persp_rotmat, f32)
    free = np.asarray(camera_free_transf, f32)
    z_offset = np.asarray(z_offset, f32)
    B = hand_pose.shape[0]
    layer = ManoLayer(mano_model, center_idx=None, dtype=f32)
    out = layer(hand_pose, hand_shape)
    root_R = out.transforms_abs[:, 0, :3, :3]
    joints = out.joints + hand_tsl[:, None]
    Rv_inv = Rv.transpose(0, 2, 1)
    op_offset = np.einsum("bij,bj->bi", Rv_inv, joints[:, 9]) / f32(2.0)
    cam_sys_offset = z_offset - op_offset
    obj_pose = np.zeros((B, 4, 4), f32)
    obj_pose[:, :3, :3] = Rv_inv
    obj_pose[:, :3, 3] = cam_sys_offset
    obj_pose[:, 3, 3] = 1
    obj_pose = free @ obj_pose
    new_root_aa = rot.rotmat_to_aa(Rv_inv @ root_R)
    new_pose = np.concatenate([new_root_aa, hand_pose[:, 3:]], axis=1).astype(f32)
    c = layer.get_rotation_center(hand_shape)
    R0, R1 = rot.aa_to_rotmat(hand_pose[:, :3]), rot.aa_to_rotmat(new_pose[:, :3])
    off0 = c - np.einsum("bij,bj->bi", R0, c)
    off1 = c - np.einsum("bij,bj->bi", R1, c)
    new_tsl = np.einsum("bij,bj->bi", Rv_inv, off0 + hand_tsl) - off1
    if scrambler is not None:
        out2 = layer(new_pose, hand_shape)  # MANO forward #2, for hand_transf (preprocessor.py:62-63)
        feed = {"hand_pose": new_pose, "hand_tsl": new_tsl, "hand_transf": out2.transforms_abs,
                "joints": np.einsum("bij,bvj->bvi", Rv_inv, joints).astype(f32),
                "hand_verts": np.einsum("bij,bvj->bvi", Rv_inv, out.verts + hand_tsl[:, None]).astype(f32)}
        new_pose, new_tsl = scrambler(feed)
    elif n_tsl is not None:
        new_pose, new_tsl = random_scrambler(new_pose, new_tsl, n_tsl, n_angle)
    if refiner is not None:
        r = refiner(new_pose, new_tsl, obj_pose[:, :3, :3])
        verts = r["hand_verts"] + cam_sys_offset[:, None]
        jts = r["joints"] + cam_sys_offset[:, None]
        new_pose, new_tsl = r["hand_pose"], r["hand_tsl"]
    else:
        ref = layer(new_pose)  # NullRefine: betas=None (refiner.py:138)
        verts = ref.verts + new_tsl[:, None] + cam_sys_offset[:, None]
        jts = ref.joints + new_tsl[:, None] + cam_sys_offset[:, None]
    Rf = free[:, :3, :3]
    return {
        "final_obj_pose": obj_pose.astype(f32),
        "final_hand_verts": np.einsum("bij,bvj->bvi", Rf, verts).astype(f32),
        "final_joints": np.einsum("bij,bvj->bvi", Rf, jts).astype(f32),
        "hand_pose": new_pose.astype(f32),
        "hand_tsl": new_tsl.astype(f32),
    }


def update_method_1(weight_map, cells, values, lower=0.1, upper=10.0):
    """cells int[n,3], values f[n] (per-cell mean error).  artiboost_loader.py:503-523."""
    w = np.array(weight_map, np.float32, copy=True)
    v = np.asarray(values, np.float64)  # python floats in the reference
    vmin, vmax = v.min(), v.max()
    conf = (vmax - v) / (vmax - vmin + 1e-8)
    mult = (1.0 / (conf + 0.5)).astype(np.float32)  # torch casts the python scalar to the tensor dtype
    for (o, p, g), m in zip(np.asarray(cells), mult):
        w[o, p, g] *= m
    return np.clip(w, np.float32(lower), np.float32(upper))


def _confidence(values):
    v = np.asarray(values, np.float64)
    return (v.max() - v) / (v.max() - v.min() + 1e-8)


def update_method_2(weight_map, cells, values, lower=0.1, upper=10.0):
    """Incremental mining, artiboost_loader.py:526-545: -0.1 where the confidence exceeds 0.5, +0.1 elsewhere, clamp."""
    w = np.array(weight_map, np.float32, copy=True)
    for (o, p, g), dec in zip(np.asarray(cells), _confidence(values) > 0.5):
        w[o, p, g] += np.float32(-0.1 if dec else 0.1)
    return np.clip(w, np.float32(lower), np.float32(upper))


def update_method_3(weight_map, cells, values, dist_lower=8.0, dist_upper=16.0):
    """Lower-bound deactivation, artiboost_loader.py:548-569: 0 below dist_lower, 1 above dist_upper, halved between;
    no clamp.  -> (weights, dist_lower_ratio)."""
    w = np.array(weight_map, np.float32, copy=True)
    v = np.asarray(values, np.float64)
    low, high = v < dist_lower, v > dist_upper
    for (o, p, g), lo_, hi_ in zip(np.asarray(cells), low, high):
        w[o, p, g] = np.float32(0.0) if lo_ else (np.float32(1.0) if hi_ else w[o, p, g] * np.float32(0.5))
    return w, low.sum() / len(low)


def update_method_4(weight_map, cells, values, epoch_idx, n_epochs, lower=0.1, upper=10.0, dist_lower=8.0, dist_upper=16.0):
    """artiboost_loader.py:572-598: update_method_1 for the first 75 % of the epochs (ratio -1), update_method_3 after."""
    if float(epoch_idx) / n_epochs < 0.75:
        return update_method_1(weight_map, cells, values, lower, upper), -1.0
    return update_method_3(weight_map, cells, values, dist_lower, dist_upper)


def blacklist_map(root_aa, u_bins, theta_bins, rand2=None, threshold=-0.8, return_th=False):
    """ArtiBoostLoader._construct_blacklist_map (artiboost_loader.py:415-500): a cell (object, view, grasp) is blacklisted
    when the view shows the back of the hand, th_sgn = ((Rv^T Rw back) . z) < -0.8 with Rw the wrist rotation of the grasp,
    Rv the view alignment of the cell's (jittered) perspective and back = (1, 0.2, 0) / |.| (:482-492).
    root_aa f[n_obj, n_grasp, 3]; rand2 f[n_obj, n_persp, n_grasp, 2] = the (u, theta) draws of get_view, None = bin centres.
    -> bool [n_obj, n_persp, n_grasp] (and th_sgn)."""
    root_aa = np.asarray(root_aa, np.float64)
    n_obj, n_grasp = root_aa.shape[:2]
    n_persp = u_bins * theta_bins
    back = np.array([1.0, 0.2, 0.0])
    back = back / np.linalg.norm(back)
    out = np.zeros((n_obj, n_persp, n_grasp), bool)
    th = np.zeros((n_obj, n_persp, n_grasp), np.float64)
    for o in range(n_obj):
        Rw = rot.aa_to_rotmat(root_aa[o])                # [n_grasp, 3, 3]
        for v in range(n_persp):
            for g in range(n_grasp):
                ru, rth = (0.5, 0.5) if rand2 is None else rand2[o, v, g]
                Rv, _, _ = view_from_id(v, u_bins, theta_bins, (0.45, 0.55), np.float32(ru), np.float32(rth), 0.0, np.float32(0.0))
                arrow = Rv.astype(np.float64).T @ Rw[g] @ back
                th[o, v, g] = arrow[2]
                out[o, v, g] = arrow[2] < threshold
    return (out, th) if return_th else out


SYNTH_UNIFORMS = 32


def synth_draw(weight_map, uniforms, u_bins, theta_bins, z_range, grasp_table, tsl_sigma, pose_sigma, n_hand_tex,
               light_range, n_bg, bg_hw, frame_wh):
    """What ab_synth_draw (artiboost_b200/csrc/synth.cu) derives from the uniforms of a batch -- the composition of
    sample_ovg, view_from_id, the grasp lookup (grasp_engine.py:47-53), N(0, sigma) scrambler noise (scrambler.py:65-81,
    Box-Muller on uniform pairs) and the renderer's per-view draws (utils/renderer.py:102-104,125-136).
    uniforms f32 [n, 32]; layout: 0 cell | 1..4 view | 5..24 ten Box-Muller pairs | 25 texture | 26 light | 27..30 bg."""
    u = np.asarray(uniforms, np.float32)
    n = u.shape[0]
    o, p, g = sample_ovg(weight_map, u[:, 0])
    views = [view_from_id(int(p[i]), u_bins, theta_bins, z_range, u[i, 1], u[i, 2], u[i, 3], u[i, 4]) for i in range(n)]
    rows = np.asarray(grasp_table, np.float32)[o, g]
    u1, u2 = u[:, 5:25:2].astype(np.float64), u[:, 6:25:2].astype(np.float64)
    r = np.sqrt(-2.0 * np.log(1.0 - u1))
    nrm = np.stack([r * np.cos(2 * np.pi * u2), r * np.sin(2 * np.pi * u2)], -1).reshape(n, 20)
    W, H = frame_wh
    out = {"obj_id": o, "persp_id": p, "grasp_id": g, "hand_pose": rows[:, :48], "hand_shape": rows[:, 48:58],
           "hand_tsl": rows[:, 58:61], "persp_rotmat": np.stack([v[0] for v in views]),
           "camera_free_transf": np.stack([v[1] for v in views]), "z_offset": np.stack([v[2] for v in views]),
           "noise_tsl": (nrm[:, :3] * tsl_sigma).astype(np.float32), "noise_angle": (nrm[:, 3:19] * pose_sigma).astype(np.float32),
           "hand_tex": np.minimum((u[:, 25] * np.float32(n_hand_tex)).astype(np.int64), n_hand_tex - 1),
           "light": (np.float32(light_range[0]) + u[:, 26] * np.float32(light_range[1] - light_range[0])).astype(np.float32)}
    if n_bg > 0:
        bh, bw = bg_hw
        bid = np.minimum((u[:, 27] * np.float32(n_bg)).astype(np.int64), n_bg - 1)
        ch = H + np.minimum((u[:, 28] * np.float32(bh - H + 1)).astype(np.int64), bh - H)
        cw = np.minimum((ch * W) // H, bw)
        y0 = np.minimum((u[:, 29] * (bh - ch + 1).astype(np.float32)).astype(np.int64), bh - ch)
        x0 = np.minimum((u[:, 30] * (bw - cw + 1).astype(np.float32)).astype(np.int64), bw - cw)
        out["bg_sel"] = np.stack([bid, x0, y0, cw, ch], 1)
    return out
