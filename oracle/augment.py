"""TEST INFRASTRUCTURE (oracle) -- CPU restatement of the crop / augment step of the synthetic branch, SURVEY.md 8 a11:
anakin/artiboost/rendered_dataset.py:127-133,155-274 (RenderedDataset.__getitem__), anakin/utils/transform.py:425-470
(get_affine_transform, transform_coords), anakin/utils/img_augment.py:6-80 (colour jitter, AFFINE warp) and
anakin/datasets/hodata.py:161-186 (bbox centre / scale).  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs
may import this file; the product path is artiboost_b200/csrc/augment.cu.

The image arithmetic of the reference lives in Pillow (requirements.txt: Pillow==8.0.1; un-vendored).  Its byte-level
rules are restated here from libImaging (Convert.c rgb2l / rgb2hsv / hsv2rgb, Blend.c, BoxBlur.c, Geometry.c
ImagingScaleAffine / affine_fixed) and ImageEnhance.py / ImageStat.py, and PINNED bit-for-bit against the Pillow that
is installed next to the tests (tests/test_augment_oracle.py runs Pillow itself on random images and on all 2^24
colours).  The composition is pinned against the reference's own RenderedDataset.__getitem__ by
tests/golden/augment.npz (tests/golden/make_golden_augment.py).

Conventions that make the step reproducible on a GPU (the reference leaves them to numpy / LAPACK / libm):
  * every random draw is an explicit input (`draws`), including cos / sin of the in-plane rotation as fp32;
  * projection, bbox, centre / scale jitter and the forward affine matrix are evaluated in fp64 in the order written
    below, the matrix is then rounded to fp32 (transform.py:459: `.astype(np.float32)`);
  * the inverse handed to the warp is the closed-form fp64 inverse of that fp32 matrix, rounded to fp32.  The reference
    calls np.linalg.inv on the fp32 matrix (LAPACK sgesv, rounding order not defined): its coefficients differ from
    ours by a few fp32 ulps, which moves < 0.5 % of the pixels by one source pixel (measured in the golden test).
"""
import numpy as np

OPS = ("brightness", "saturation", "hue", "contrast")  # list-building order of img_augment.apply_jitter (:31-39)


# ----------------------------------------------------------------------------------------- Pillow byte arithmetic
def rgb_to_l(rgb):
    """Convert.c rgb2l: ITU-R 601-2 luma with 16-bit fixed-point weights."""
    r, g, b = (rgb[..., i].astype(np.uint32) for i in range(3))
    return ((r * 19595 + g * 38470 + b * 7471 + 0x8000) >> 16).astype(np.uint8)


def blend(deg, img, alpha):
    """Blend.c ImagingBlend(in1 = deg, in2 = img, alpha as float): in1 + alpha * (in2 - in1), fp32, truncated."""
    a = np.float32(alpha)
    if a == 0.0:
        return deg.copy()
    if a == 1.0:
        return img.copy()
    d = deg.astype(np.float32)
    t = d + a * (img.astype(np.float32) - d)
    if 0.0 <= a <= 1.0:
        return t.astype(np.uint8)
    out = t.astype(np.int32)  # truncation toward zero of in-range values
    out[t <= 0.0] = 0
    out[t >= 255.0] = 255
    return out.astype(np.uint8)


def adjust_brightness(img, factor):
    return blend(np.zeros_like(img), img, factor)  # ImageEnhance.Brightness: degenerate = black


def contrast_mean(img):
    """ImageEnhance.Contrast: int(ImageStat.Stat(image.convert('L')).mean[0] + 0.5)."""
    l = rgb_to_l(img)
    return int(float(l.astype(np.int64).sum()) / float(l.size) + 0.5)


def adjust_contrast(img, factor, mean=None):
    mean = contrast_mean(img) if mean is None else mean
    return blend(np.full_like(img, mean), img, factor)


def adjust_saturation(img, factor):
    l = rgb_to_l(img)
    return blend(np.stack([l, l, l], -1), img, factor)  # ImageEnhance.Color: degenerate = L replicated


def rgb_to_hsv(rgb):
    """Convert.c rgb2hsv_row (follows colorsys.py; float h, s; h / 6.0 + 1.0 and * 255.0 in double)."""
    r, g, b = (rgb[..., i].astype(np.int32) for i in range(3))
    maxc = np.maximum(r, np.maximum(g, b))
    minc = np.minimum(r, np.minimum(g, b))
    cr = (maxc - minc).astype(np.float32)
    safe = np.where(cr == 0, np.float32(1), cr)
    mx = np.where(maxc == 0, 1, maxc).astype(np.float32)
    s = cr / mx
    rc = (maxc - r).astype(np.float32) / safe
    gc = (maxc - g).astype(np.float32) / safe
    bc = (maxc - b).astype(np.float32) / safe
    h = np.where(r == maxc, bc - gc, np.where(g == maxc, (2.0 + rc.astype(np.float64) - bc).astype(np.float32),
                                              (4.0 + gc.astype(np.float64) - rc).astype(np.float32))).astype(np.float32)
    h = np.fmod(h.astype(np.float64) / 6.0 + 1.0, 1.0).astype(np.float32)
    uh = np.clip((h.astype(np.float64) * 255.0).astype(np.int32), 0, 255)
    us = np.clip((s.astype(np.float64) * 255.0).astype(np.int32), 0, 255)
    gray = minc == maxc
    uh = np.where(gray, 0, uh)
    us = np.where(gray, 0, us)
    return np.stack([uh, us, maxc], -1).astype(np.uint8)


def hsv_to_rgb(hsv):
    """Convert.c hsv2rgb (follows colorsys.py)."""
    h, s, v = (hsv[..., i].astype(np.int32) for i in range(3))
    hf = h.astype(np.float32).astype(np.float64) * 6.0 / 255.0
    i = np.floor(hf).astype(np.int32)
    f = (hf - i.astype(np.float32).astype(np.float64)).astype(np.float32)
    fs = (s.astype(np.float32).astype(np.float64) / 255.0).astype(np.float32)
    vf = v.astype(np.float32).astype(np.float64)
    f64, fs64 = f.astype(np.float64), fs.astype(np.float64)

    def rnd(x):  # C round(): half away from zero (arguments are >= 0 here)
        return np.clip(np.floor(x + 0.5).astype(np.int32), 0, 255)

    p = rnd(vf * (1.0 - fs64))
    q = rnd(vf * (1.0 - fs64 * f64))
    t = rnd(vf * (1.0 - fs64 * (1.0 - f64)))
    k = i % 6
    r = np.choose(k, [v, q, p, p, t, v])
    g = np.choose(k, [t, v, v, q, p, p])
    b = np.choose(k, [p, p, t, v, v, q])
    gray = s == 0
    return np.stack([np.where(gray, v, r), np.where(gray, v, g), np.where(gray, v, b)], -1).astype(np.uint8)


def hue_shift_byte(hue_factor):
    """img_augment.py:214: np.uint8(hue_factor * 255) -- truncation toward zero, then wrap to a byte."""
    return int(np.trunc(float(hue_factor) * 255.0)) & 0xFF


def adjust_hue(img, factor):
    hsv = rgb_to_hsv(img)
    hsv[..., 0] = (hsv[..., 0].astype(np.int32) + hue_shift_byte(factor)) & 0xFF  # uint8 addition wraps
    return hsv_to_rgb(hsv)


def gaussian_box_radius(radius, passes=3):
    """BoxBlur.c _gaussian_blur_radius: box radius whose `passes`-fold application approximates the Gaussian."""
    sigma2 = np.float32(radius) * np.float32(radius) / np.float32(passes)
    big_l = np.float32(np.sqrt(np.float64(12.0) * np.float64(sigma2) + 1.0))
    l = np.float32(np.floor((np.float64(big_l) - 1.0) / 2.0))
    a = np.float32((2 * l + 1) * (l * (l + 1) - 3 * sigma2))
    a = np.float32(a / np.float32(6 * (sigma2 - (l + 1) * (l + 1))))
    return np.float32(l + a)


def box_blur_line_weights(float_radius):
    """BoxBlur.c ImagingHorizontalBoxBlur: centre-window weight ww and far-pixel weight fw, 8.24 fixed point."""
    radius = int(float_radius)
    ww = int(np.uint32(np.float32(1 << 24) / np.float32(np.float32(float_radius) * 2 + 1)))
    fw = (((1 << 24) - (radius * 2 + 1) * ww) // 2) & 0xFFFFFFFF
    return radius, ww, fw


def _box_blur_axis(img, float_radius, axis):
    """One pass of ImagingLineBoxBlur along `axis`; edges replicate.  Only box radii < 1 occur (GaussianBlur <= 0.1)."""
    radius, ww, fw = box_blur_line_weights(float_radius)
    if radius != 0:
        raise NotImplementedError("box radius >= 1 is outside the reference's blur range (rendered_dataset.py:66,256)")
    x = img.astype(np.uint64)
    n = img.shape[axis]
    idx = np.arange(n)
    left = np.take(x, np.maximum(idx - 1, 0), axis=axis)
    right = np.take(x, np.minimum(idx + 1, n - 1), axis=axis)
    bulk = (x * ww + (left + right) * fw) & 0xFFFFFFFF  # UINT32 arithmetic
    return (((bulk + (1 << 23)) & 0xFFFFFFFF) >> 24).astype(np.uint8)


def gaussian_blur(img, radius, passes=3):
    """ImageFilter.GaussianBlur(radius): `passes` horizontal box passes, then `passes` vertical ones, each rounded to bytes."""
    if float(radius) == 0.0:
        return img.copy()
    r = gaussian_box_radius(radius, passes)
    out = img
    for axis in (1, 0):
        for _ in range(passes):
            out = _box_blur_axis(out, r, axis)
    return out


def affine_nearest(img, coeffs, out_size):
    """Image.transform(out_size, AFFINE, coeffs) with the default NEAREST filter and fill = 0 (Geometry.c).
    coeffs: 6 doubles (a0..a5): x_in = a0*x + a1*y + a2, y_in = a3*x + a4*y + a5 at output pixel centres."""
    a = [float(c) for c in coeffs]
    wo, ho = out_size
    hi, wi = img.shape[:2]
    out = np.zeros((ho, wo) + img.shape[2:], img.dtype)
    if a[1] == 0.0 and a[3] == 0.0:  # ImagingScaleAffine: running double sums, COORD(v) = v < 0 ? -1 : (int)v
        xs, ys = np.full(wo, -1), np.full(ho, -1)
        xo = a[2] + a[0] * 0.5
        for x in range(wo):
            xs[x] = -1 if xo < 0.0 else int(xo)
            xo += a[0]
        yo = a[5] + a[4] * 0.5
        for y in range(ho):
            ys[y] = -1 if yo < 0.0 else int(yo)
            yo += a[4]
        vx, vy = (xs >= 0) & (xs < wi), (ys >= 0) & (ys < hi)
        out[np.ix_(vy, vx)] = img[np.ix_(ys[vy], xs[vx])]
        return out

    def fix(v):  # 16.16 fixed point: FLOOR(v * 65536.0 + 0.5)
        return int(np.floor(v * 65536.0 + 0.5))

    a0, a1, a3, a4 = fix(a[0]), fix(a[1]), fix(a[3]), fix(a[4])
    a2 = fix(a[2] + a[0] * 0.5 + a[1] * 0.5)
    a5 = fix(a[5] + a[3] * 0.5 + a[4] * 0.5)
    x, y = np.meshgrid(np.arange(wo, dtype=np.int64), np.arange(ho, dtype=np.int64))
    xx, yy = a2 + x * a0 + y * a1, a5 + x * a3 + y * a4
    assert max(abs(int(xx.min())), abs(int(xx.max())), abs(int(yy.min())), abs(int(yy.max()))) < 2 ** 31  # check_fixed
    xin, yin = xx >> 16, yy >> 16
    ok = (xin >= 0) & (xin < wi) & (yin >= 0) & (yin < hi)
    out[ok] = img[yin[ok], xin[ok]]
    return out


def color_jitter(img, draws):
    """img_augment.apply_jitter with the factors and the (shuffled) execution order given explicitly."""
    for op in draws["order"]:
        name = OPS[int(op)]
        if name == "brightness":
            img = adjust_brightness(img, draws["brightness"])
        elif name == "saturation":
            img = adjust_saturation(img, draws["saturation"])
        elif name == "hue":
            img = adjust_hue(img, draws["hue"])
        else:
            img = adjust_contrast(img, draws["contrast"])
    return img


# ------------------------------------------------------------------------------------------------ geometry (fp64)
# Written as individually rounded scalar operations in a fixed order (no BLAS: its FMA use is not defined), so that
# csrc/augment.cu (compiled with -fmad=false) reproduces every integer decision (bbox centre, jitter truncation).
def _dot3(a0, a1, a2, b0, b1, b2):
    return (a0 * b0 + a1 * b1) + a2 * b2


def affine_no_rot(center, scale, res):
    """transform.py:462-470 get_affine_trans_no_rot -> (m00, m11, m02, m12); the other entries are 0 / 1."""
    ratio = float(res[0]) / float(res[1])
    m00 = float(res[0]) / scale
    m11 = float(res[1]) / scale * ratio
    m02 = res[0] * (-float(center[0]) / scale + 0.5)
    m12 = res[1] * (-float(center[1]) / scale * ratio + 0.5)
    return m00, m11, m02, m12


def get_affine_transform(center, scale, optical_center, out_res, cs, sn):
    """transform.py:434-460 with cos / sin supplied.  -> (total fp32 [3,3], post-rotation fp32 [3,3]).
    total = no_rot(R c) . R and post = no_rot(T^-1 R T c): products with the exact zeros / ones of R, T are dropped."""
    cs, sn = float(cs), float(sn)
    cx, cy = float(center[0]), float(center[1])
    ox, oy = float(optical_center[0]), float(optical_center[1])
    orc = (cs * cx + (-sn) * cy, sn * cx + cs * cy)
    dx, dy = cx - ox, cy - oy
    tc = ((cs * dx + (-sn) * dy) + ox, (sn * dx + cs * dy) + oy)
    m00, m11, m02, m12 = affine_no_rot(orc, scale, out_res)
    total = np.array([[m00 * cs, m00 * (-sn), m02], [m11 * sn, m11 * cs, m12], [0, 0, 1]])
    p00, p11, p02, p12 = affine_no_rot(tc, scale, out_res)
    post = np.array([[p00, 0, p02], [0, p11, p12], [0, 0, 1]])
    return total.astype(np.float32), post.astype(np.float32)


def invert_affine(m32):
    """Closed-form fp64 inverse of the fp32 2x3 affine part, rounded to fp32 (see the header)."""
    a, b, c = (float(v) for v in m32[0])
    d, e, f = (float(v) for v in m32[1])
    det = a * e - b * d
    ia, ib, id_, ie = e / det, -b / det, -d / det, a / det
    ic = -(ia * c + ib * f)
    if_ = -(id_ * c + ie * f)
    return np.array([ia, ib, ic, id_, ie, if_], np.float64).astype(np.float32)


def project(K, pts):
    """rendered_dataset.py:127-133: uv = (K . X)[:2] / (Z + 1e-8), fp64."""
    K, pts = np.asarray(K, np.float64), np.asarray(pts, np.float64)
    x, y, z = pts[:, 0], pts[:, 1], pts[:, 2]
    hz = _dot3(K[2, 0], K[2, 1], K[2, 2], x, y, z) + 1e-8
    return np.stack([_dot3(K[0, 0], K[0, 1], K[0, 2], x, y, z) / hz, _dot3(K[1, 0], K[1, 1], K[1, 2], x, y, z) / hz], 1)


def crop_params(joints_2d, corners_2d, draws, cfg):
    """bbox centre / scale (crop_model root_obj | hand_obj | hand; hodata.py:161-186) + jitter (rendered_dataset.py:175-190)."""
    model = cfg.get("crop_model", "root_obj")
    if cfg.get("full_image", False):
        center, scale = np.array([cfg["raw_size"][0] / 2, cfg["raw_size"][1] / 2]), float(cfg["raw_size"][0])
    else:
        pts = {"root_obj": np.concatenate([joints_2d[[0]], corners_2d]), "hand_obj": np.concatenate([joints_2d, corners_2d]),
               "hand": joints_2d}[model]
        mn, mx = pts.min(0), pts.max(0)
        center = np.array([int((mx[0] + mn[0]) / 2), int((mx[1] + mn[1]) / 2)])
        scale = float(max(mx[0] - mn[0], mx[1] - mn[1]))
    f32 = lambda v: float(np.float32(v))  # noqa: E731  configuration factors travel as fp32 (ab_augment_cfg)
    scale *= f32(cfg.get("bbox_expand_ratio", 1.2)) if not cfg.get("full_image", False) else 1.0
    if cfg.get("aug", True):
        off = f32(cfg["center_jit"]) * scale * np.asarray(draws["center_jit"], np.float64)
        center = center + off.astype(int)  # truncation toward zero (rendered_dataset.py:180)
        jit = float(np.clip(float(draws["scale_jit"]) + 1.0, 1 - f32(cfg["scale_jit"]), 1 + f32(cfg["scale_jit"])))
        scale = scale * jit
    return center, scale


def rendered_sample(img, joints, obj_pose, corners_can, cam_intr, draws, cfg):
    """RenderedDataset.__getitem__ (rendered_dataset.py:155-274) for one rendered RGB image + its annotations.
    img uint8 [H,W,3]; joints [21,3]; obj_pose [4,4]; corners_can [8,3]; cam_intr [3,3]; draws: see header; cfg:
    image_size (W,H), raw_size, center_idx, bbox_expand_ratio, crop_model, aug, center_jit, scale_jit."""
    aug = cfg.get("aug", True)
    wo, ho = cfg["image_size"]
    K = np.asarray(cam_intr, np.float64)
    joints = np.asarray(joints, np.float32)
    pose = np.asarray(obj_pose, np.float32)
    corners_can = np.asarray(corners_can, np.float32)
    R, t, cc = pose[:3, :3].astype(np.float64), pose[:3, 3].astype(np.float64), corners_can.astype(np.float64)
    corners_3d = np.stack([_dot3(R[i, 0], R[i, 1], R[i, 2], cc[:, 0], cc[:, 1], cc[:, 2]) + t[i] for i in range(3)], 1)
    j2d, c2d = project(K, joints), project(K, corners_3d)
    center, scale = crop_params(j2d, c2d, draws, cfg)
    cs, sn = (np.float32(draws["rot_cs"][0]), np.float32(draws["rot_cs"][1])) if aug else (np.float32(1), np.float32(0))
    total, post = get_affine_transform(center, scale, (K[0, 2], K[1, 2]), (wo, ho), cs, sn)
    rot3 = np.array([[cs, -sn, 0], [sn, cs, 0], [0, 0, 1]], np.float32)
    P = post.astype(np.float64)
    out = {"cam_intr": np.array([[_dot3(P[i, 0], P[i, 1], P[i, 2], K[0, j], K[1, j], K[2, j]) for j in range(3)]
                                 for i in range(3)]).astype(np.float32)}
    c64, s64 = float(cs), float(sn)

    def rotz(p):
        p = p.astype(np.float64)
        return np.stack([c64 * p[:, 0] + (-s64) * p[:, 1], s64 * p[:, 0] + c64 * p[:, 1], p[:, 2]], 1)

    j3 = rotz(joints)
    root = j3[cfg.get("center_idx", 0)]
    out["root_joint"] = root.astype(np.float32)
    out["joints_3d"] = (j3 - root).astype(np.float32)

    T = total.astype(np.float64)

    def warp_pts(p):
        return np.stack([(T[0, 0] * p[:, 0] + T[0, 1] * p[:, 1]) + T[0, 2], (T[1, 0] * p[:, 0] + T[1, 1] * p[:, 1]) + T[1, 2]], 1).astype(np.float32)

    def vis(raw2d, aug2d, n):
        rw, rh = cfg["raw_size"]
        v_raw = (raw2d[:, 0] >= 0) & (raw2d[:, 0] < rw) & (raw2d[:, 1] >= 0) & (raw2d[:, 1] < rh)
        v_aug = ((aug2d[:, 0] >= 0) & (aug2d[:, 0] < wo) & (aug2d[:, 1] >= 0) & (aug2d[:, 1] < ho)).astype(np.float32)
        if v_raw.sum() < n * 0.4 or v_aug.sum() < n * 0.4:
            return np.zeros(n, np.float32)
        return v_aug

    out["joints_2d"] = warp_pts(j2d)
    out["joints_vis"] = vis(j2d, out["joints_2d"], 21)
    c3 = rotz(corners_3d)
    out["corners_3d"] = (c3 - root).astype(np.float32)
    out["corners_2d"] = warp_pts(c2d)
    out["corners_vis"] = vis(c2d, out["corners_2d"], 8)
    out["corners_can"] = corners_can
    transf = np.eye(4, dtype=np.float32)
    transf[:3, :4] = np.stack([c64 * pose[0, :4] + (-s64) * pose[1, :4], s64 * pose[0, :4] + c64 * pose[1, :4], pose[2, :4]]).astype(np.float32)
    out["obj_transf"] = transf
    out["affine"] = total
    out["inv_affine"] = invert_affine(total)
    if aug:
        img = gaussian_blur(img, draws["blur_radius"])
        img = color_jitter(img, draws)
    warped = affine_nearest(img, out["inv_affine"], (wo, ho))
    out["image_u8"] = warped
    out["image"] = (warped.astype(np.float32) / np.float32(255.0) - np.float32(0.5)).transpose(2, 0, 1)  # to_tensor, -0.5
    return out
