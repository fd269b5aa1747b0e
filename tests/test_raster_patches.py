"""Mesh patches of the tile rasteriser (ab_build_patches_host, host code of the C-ABI library; no device needed) and the
conservativeness of the binning pass.

The tile kernel only ever evaluates real faces with the exact rules, so its image equals the oracle's if and only if, for
every covered pixel, the patch of the WINNING face was (a) not dropped by the patch-level back-face test and (b) placed in
the list of that pixel's tile.  This file restates raster_bin_kernel's arithmetic (artiboost_b200/csrc/raster.cu) in
numpy fp32 and checks (a) and (b) against the winners oracle/raster.c reports, over random CCV views.  It also checks (c):
a culled patch contains no face the rules would draw at all."""
import numpy as np
import pytest

from oracle import ccv, raster
from artiboost_b200 import assets

CFG = dict(width=256, height=256, fx=217.5, fy=217.5, cx=128.0, cy=128.0, znear=0.05, cull_backface=1, ambient=0.8, diffuse=0.25)
TILE = 64
EPS = np.float32(0.0032)
f32 = np.float32


@pytest.fixture(scope="module")
def patches_of(lib_built, objects, mano_model):
    from artiboost_b200.artiboost.patches import build_patches
    out = {name: build_patches(o["vertices"], o["faces"]) for name, o in objects.items()}
    out["hand"] = build_patches(mano_model["v_template"], mano_model["f"])
    return out


def test_patches_partition_the_mesh(patches_of, objects, mano_model):
    meshes = {name: (o["vertices"], np.asarray(o["faces"])) for name, o in objects.items()}
    meshes["hand"] = (mano_model["v_template"], np.asarray(mano_model["f"]))
    for name, (verts, faces) in meshes.items():
        p = patches_of[name]
        n = p["vid"].shape[0]
        prim = p["prim"]
        used = prim[prim >= 0]
        assert sorted(used.tolist()) == list(range(len(faces))), name      # every face exactly once, original ids kept
        assert ((p["face"] == 0xFFFFFFFF) == (prim < 0)).all()
        v32 = np.asarray(verts, np.float32)
        for i in range(n):
            nv = int((p["vid"][i] >= 0).sum())
            assert (p["vid"][i, :nv] >= 0).all() and (p["vid"][i, nv:] == -1).all()
            assert len(set(p["vid"][i, :nv].tolist())) == nv
            np.testing.assert_array_equal(p["pos"][i, :nv, :3], v32[p["vid"][i, :nv]])   # same floats the shader reads
            assert (p["pos"][i, :nv, 3] == 1).all() and (p["pos"][i, nv:, 3] == 0).all()
            for l in range(32):
                if prim[i, l] < 0:
                    continue
                w = int(p["face"][i, l])
                loc = [w & 255, (w >> 8) & 255, (w >> 16) & 255]
                assert max(loc) < nv and (w >> 24) == 0
                assert p["vid"][i, loc].tolist() == faces[prim[i, l]][:3].tolist()
            # bounding sphere holds every vertex
            d = np.linalg.norm(v32[p["vid"][i, :nv]].astype(np.float64) - p["bound"][i, :3].astype(np.float64), axis=1)
            assert d.max() <= p["bound"][i, 3]
            assert p["bound"][i, 10] == nv and p["bound"][i, 11] == (prim[i] >= 0).sum()
        nf = (prim >= 0).sum(1)
        assert nf.mean() > 24, (name, nf.mean())          # lanes are mostly busy
        print(f"{name}: {n} patches, {nf.mean():.1f} faces / {(p['vid'] >= 0).sum(1).mean():.1f} verts per patch, "
              f"radius {p['bound'][:, 3].mean() * 1e3:.1f} mm")


def bin_object_patches(bound, pose, cfg, cull=True):
    """numpy fp32 restatement of the object branch of raster_bin_kernel -> (kept bool[n], px_lo, px_hi, py_lo, py_hi)."""
    M = np.asarray(pose, np.float32)
    c = bound[:, :3]
    s = (c @ M[:3, :3].T + M[:3, 3]).astype(np.float32)
    r = bound[:, 3]
    zmin, zmax = s[:, 2] - r, s[:, 2] + r
    fx, fy, cx, cy = (f32(cfg[k]) for k in ("fx", "fy", "cx", "cy"))
    W, H = cfg["width"], cfg["height"]
    culled = np.zeros(len(bound), bool)
    if cull:
        a = (bound[:, 4:7] @ M[:3, :3].T).astype(np.float32)
        ls = np.sqrt((s * s).sum(1)).astype(np.float32)
        la = np.sqrt((a * a).sum(1)).astype(np.float32)
        cos_phi = ((a * s).sum(1) / (ls * la)).astype(np.float32)
        cc = bound[:, 7]
        sin_phi = np.sqrt(np.maximum(f32(0), f32(1) - cos_phi * cos_phi))
        sin_th = np.sqrt(np.maximum(f32(0), f32(1) - cc * cc))
        bnd = ls * (cos_phi * cc - sin_phi * sin_th) - r
        T = (np.sqrt(s[:, 0] ** 2 + s[:, 1] ** 2) + r) / zmin
        g = max(fx, fy) * (f32(1) + T) / zmin
        with np.errstate(over="ignore", invalid="ignore"):
            need = f32(1.5) * zmax ** 3 / (fx * fy) * (EPS * bound[:, 8] * g + f32(4) * EPS * EPS * bound[:, 9]) + f32(2e-5)
            culled = (cc > 0) & (zmin > 1e-4) & (cos_phi > 0) & (bnd > need)
    ilo, ihi = f32(1) / zmin, f32(1) / zmax
    xl, xh, yl, yh = s[:, 0] - r, s[:, 0] + r, s[:, 1] - r, s[:, 1] + r
    umin = fx * np.minimum(xl * ilo, xl * ihi) + cx
    umax = fx * np.maximum(xh * ilo, xh * ihi) + cx
    vmin = fy * np.minimum(yl * ilo, yl * ihi) + cy
    vmax = fy * np.maximum(yh * ilo, yh * ihi) + cy
    px_lo = np.maximum(np.clip(np.ceil(umin - f32(0.55)), -1e6, 1e6).astype(np.int64), 0)
    px_hi = np.minimum(np.clip(np.floor(umax - f32(0.45)), -1e6, 1e6).astype(np.int64), W - 1)
    py_lo = np.maximum(np.clip(np.ceil(vmin - f32(0.55)), -1e6, 1e6).astype(np.int64), 0)
    py_hi = np.minimum(np.clip(np.floor(vmax - f32(0.45)), -1e6, 1e6).astype(np.int64), H - 1)
    far = zmin > 1e-4
    px_lo, px_hi = np.where(far, px_lo, 0), np.where(far, px_hi, W - 1)
    py_lo, py_hi = np.where(far, py_lo, 0), np.where(far, py_hi, H - 1)
    kept = ~culled & (px_lo <= px_hi) & (py_lo <= py_hi)
    return kept, culled, px_lo, px_hi, py_lo, py_hi


def drawable_faces(verts, faces, pose, cfg):
    """Faces that pass the rules' set-up (oracle/raster.c setup()): all corners valid, snapped area2 < 0 (front) -- numpy
    restatement of project() + area2 with the same fp32 operation order."""
    M = np.asarray(pose, np.float32)
    v = np.asarray(verts, np.float32)
    fma = lambda a, b, c: (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)  # noqa: E731
    cam = []
    for i in range(3):
        acc = fma(np.full(len(v), M[i, 0], np.float32), v[:, 0], np.full(len(v), M[i, 3], np.float32))
        acc = fma(np.full(len(v), M[i, 1], np.float32), v[:, 1], acc)
        cam.append(fma(np.full(len(v), M[i, 2], np.float32), v[:, 2], acc))
    X, Y, Z = cam
    iz = (f32(1) / Z).astype(np.float32)
    u = fma(np.full(len(v), cfg["fx"], np.float32), (X * iz).astype(np.float32), np.full(len(v), cfg["cx"], np.float32))
    w = fma(np.full(len(v), cfg["fy"], np.float32), (Y * iz).astype(np.float32), np.full(len(v), cfg["cy"], np.float32))
    snap = lambda t: np.rint(np.clip((t * f32(256)).astype(np.float32), -4194304.0, 4194304.0)).astype(np.int64)  # noqa: E731
    x, y, ok = snap(u), snap(w), Z >= f32(cfg["znear"])
    f = np.asarray(faces)[:, :3]
    a, b, d = f[:, 0], f[:, 1], f[:, 2]
    area2 = (x[b] - x[a]) * (y[d] - y[a]) - (x[d] - x[a]) * (y[b] - y[a])
    return ok[a] & ok[b] & ok[d] & (area2 < 0)


def test_binning_is_conservative_against_the_oracle_winners(patches_of, objects, mano_model):
    rng = np.random.RandomState(0)
    hand_cols = np.full((778, 4), 200, np.uint8)
    hf = np.asarray(mano_model["f"], np.int32)
    hp = patches_of["hand"]
    hand_patch_of = np.empty(len(hf), np.int64)
    for i in range(hp["prim"].shape[0]):
        hand_patch_of[hp["prim"][i][hp["prim"][i] >= 0]] = i
    culled_frac, dup = [], []
    for name, o in objects.items():
        verts, faces = o["vertices"], np.asarray(o["faces"], np.int32)[:, :3]
        p = patches_of[name]
        patch_of = np.empty(len(faces), np.int64)
        for i in range(p["prim"].shape[0]):
            patch_of[p["prim"][i][p["prim"][i] >= 0]] = i
        cols = np.concatenate([o["colors"][:, :3], np.full((len(verts), 1), 255, np.uint8)], 1)
        for it in range(10):
            rot, free, zoff = ccv.view_from_id(int(rng.randint(288)), 12, 24, (0.45, 0.55), *rng.rand(4))
            pose = np.eye(4, dtype=np.float32)
            pose[:3, :3] = free[:3, :3] @ rot.T
            pose[:3, 3] = zoff + rng.normal(0, 0.03, 3)
            if it >= 7:   # partly off-screen and close to the camera
                pose[:3, 3] = [rng.uniform(-0.25, 0.25), rng.uniform(-0.25, 0.25), rng.uniform(0.2, 0.4)]
            hv = (mano_model["v_template"] + np.array([0.05, 0.0, 0.5]) + rng.normal(0, 0.02, 3)).astype(np.float32)
            _, _, seg, key = raster.render_view(CFG, hv, hf, hand_cols, verts, faces, cols, pose)
            kept, culled, px_lo, px_hi, py_lo, py_hi = bin_object_patches(p["bound"], pose, CFG)
            # (c) a culled patch holds no drawable face
            draw = drawable_faces(verts, faces, pose, CFG)
            assert not draw[culled[patch_of]].any(), (name, it)
            # (a) + (b) on the winners
            ys, xs = np.nonzero(seg == 2)
            win = (key[ys, xs] & np.uint64(0xffffffff)).astype(np.int64)
            pw = patch_of[win]
            assert kept[pw].all(), (name, it)
            t_lo_x, t_hi_x, t_lo_y, t_hi_y = px_lo[pw] // TILE, px_hi[pw] // TILE, py_lo[pw] // TILE, py_hi[pw] // TILE
            assert ((xs // TILE >= t_lo_x) & (xs // TILE <= t_hi_x) & (ys // TILE >= t_lo_y) & (ys // TILE <= t_hi_y)).all(), (name, it)
            assert ((xs >= px_lo[pw]) & (xs <= px_hi[pw]) & (ys >= py_lo[pw]) & (ys <= py_hi[pw])).all(), (name, it)
            if it < 7:
                nfp = (p["prim"] >= 0).sum(1)
                culled_frac.append(nfp[culled].sum() / len(faces))
                ntiles = ((px_hi // TILE - px_lo // TILE + 1) * (py_hi // TILE - py_lo // TILE + 1))[kept]
                dup.append(ntiles.mean())
            # hand patches: exact box of the projected vertices (any face's box lies inside it)
            ys, xs = np.nonzero(seg == 1)
            win = (key[ys, xs] & np.uint64(0xffffffff)).astype(np.int64) - len(faces)
            assert (win >= 0).all()
    print("faces dropped by the patch-level back-face test: mean %.3f (min %.3f, max %.3f); tiles per kept patch %.2f"
          % (np.mean(culled_frac), min(culled_frac), max(culled_frac), np.mean(dup)))
    assert np.mean(culled_frac) > 0.25
