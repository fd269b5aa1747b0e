"""ORACLE (test infrastructure): ctypes wrapper over oracle/raster.c (see that file for the rule set)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class OracleCfg(ctypes.Structure):
    _fields_ = [("width", ctypes.c_int), ("height", ctypes.c_int), ("fx", ctypes.c_float), ("fy", ctypes.c_float),
                ("cx", ctypes.c_float), ("cy", ctypes.c_float), ("znear", ctypes.c_float),
                ("cull_backface", ctypes.c_int), ("ambient", ctypes.c_float), ("diffuse", ctypes.c_float),
                ("bg_r", ctypes.c_int), ("bg_g", ctypes.c_int), ("bg_b", ctypes.c_int)]


def build(force=False):
    so = os.path.join(_HERE, "liboracle_raster.so")
    src = os.path.join(_HERE, "raster.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "liboracle_raster.so"])
    return so


def _lib():
    global _LIB
    if _LIB is None:
        # libgomp reads these when it is loaded; without binding, the sandboxed hosts schedule the team on one core
        os.environ.setdefault("OMP_PROC_BIND", "true")
        os.environ.setdefault("OMP_WAIT_POLICY", "passive")
        _LIB = ctypes.CDLL(build())
        _LIB.ab_oracle_render.restype = None
    return _LIB


def _p(a, t):
    return None if a is None else a.ctypes.data_as(ctypes.POINTER(t))


def render_view(cfg, hand_verts, hand_faces, hand_cols, obj_verts=None, obj_faces=None, obj_cols=None, obj_pose=None,
                light=3.0, bg=None, bg_sel=None):
    """cfg: dict(width,height,fx,fy,cx,cy,znear,cull_backface,ambient,diffuse,bg_rgb).  Colours are u8[...,4].
    -> rgba u8[H,W,4], depth f32[H,W], seg u8[H,W], key u64[H,W]."""
    W, H = int(cfg["width"]), int(cfg["height"])
    bgc = cfg.get("bg_rgb", (128, 128, 128))
    c = OracleCfg(W, H, cfg["fx"], cfg["fy"], cfg["cx"], cfg["cy"], cfg.get("znear", 0.05),
                  int(cfg.get("cull_backface", 1)), cfg.get("ambient", 0.8), cfg.get("diffuse", 0.25), *bgc)
    hv = np.ascontiguousarray(hand_verts, np.float32)
    hf = np.ascontiguousarray(hand_faces, np.int32)
    hc = np.ascontiguousarray(hand_cols, np.uint8)
    assert hc.shape == (hv.shape[0], 4)
    if obj_verts is None:
        ov, of, oc = np.zeros((0, 3), np.float32), np.zeros((0, 3), np.int32), np.zeros((0, 4), np.uint8)
        op = np.eye(4, dtype=np.float32)
    else:
        ov = np.ascontiguousarray(obj_verts, np.float32)
        of = np.ascontiguousarray(obj_faces, np.int32)
        oc = np.ascontiguousarray(obj_cols, np.uint8)
        op = np.ascontiguousarray(obj_pose, np.float32).reshape(4, 4)
        assert oc.shape == (ov.shape[0], 4)
    rgba = np.empty((H, W, 4), np.uint8)
    depth = np.empty((H, W), np.float32)
    seg = np.empty((H, W), np.uint8)
    key = np.empty((H, W), np.uint64)
    bgp = bsel = None
    bh = bw = 0
    if bg is not None:
        bgp = np.ascontiguousarray(bg, np.uint8)
        bh, bw = bgp.shape[:2]
        bsel = np.ascontiguousarray(bg_sel, np.int32)
    _lib().ab_oracle_render(
        ctypes.byref(c), hv.shape[0], _p(hv, ctypes.c_float), hf.shape[0], _p(hf, ctypes.c_int32),
        _p(hc, ctypes.c_uint8), ov.shape[0], _p(ov, ctypes.c_float), of.shape[0], _p(of, ctypes.c_int32),
        _p(oc, ctypes.c_uint8), _p(op, ctypes.c_float), ctypes.c_float(light), _p(bgp, ctypes.c_uint8), bh, bw,
        _p(bsel, ctypes.c_int32), _p(rgba, ctypes.c_uint8), _p(depth, ctypes.c_float), _p(seg, ctypes.c_uint8),
        _p(key, ctypes.c_uint64))
    return rgba, depth, seg, key


def render_views_threaded(cfg, views, n_threads=None):
    """Render a list of kwargs-dicts for render_view on `n_threads` host threads (ctypes releases the GIL during the
    C call, so the threads run on separate cores).  -> list of (rgba, depth, seg, key)."""
    from concurrent.futures import ThreadPoolExecutor
    _lib()
    n_threads = n_threads or os.cpu_count() or 1
    with ThreadPoolExecutor(max_workers=n_threads) as ex:
        return list(ex.map(lambda kw: render_view(cfg, **kw), views))


def render_batch(cfg, scene, hand_verts, hand_tex, obj_id, obj_pose, light, bg_sel=None, n_threads=None, out=None):
    """OpenMP batch driver (the CPU baseline bench.py times).  scene: dict(hand_faces i32[F,3], hand_cols u8[T,V,4],
    obj_verts f32[sumV,3], obj_vert_off i32[n+1], obj_faces i32[sumF,3], obj_face_off i32[n+1], obj_cols u8[sumV,4],
    bgs u8[n,h,w,3] | None).  -> dict(rgba, depth, seg)."""
    lib = _lib()
    W, H = int(cfg["width"]), int(cfg["height"])
    bgc = cfg.get("bg_rgb", (128, 128, 128))
    c = OracleCfg(W, H, cfg["fx"], cfg["fy"], cfg["cx"], cfg["cy"], cfg.get("znear", 0.05),
                  int(cfg.get("cull_backface", 1)), cfg.get("ambient", 0.8), cfg.get("diffuse", 0.25), *bgc)
    n_threads = int(n_threads or os.cpu_count() or 1)
    B = int(hand_verts.shape[0])
    a = lambda x, dt: np.ascontiguousarray(x, dt)  # noqa: E731
    hv, ht, oid = a(hand_verts, np.float32), a(hand_tex, np.int32), a(obj_id, np.int32)
    op, li = a(obj_pose, np.float32), a(light, np.float32)
    hf, hc = a(scene["hand_faces"], np.int32), a(scene["hand_cols"], np.uint8)
    ov, ovo = a(scene["obj_verts"], np.float32), a(scene["obj_vert_off"], np.int32)
    of, ofo, oc = a(scene["obj_faces"], np.int32), a(scene["obj_face_off"], np.int32), a(scene["obj_cols"], np.uint8)
    bgs = None if scene.get("bgs") is None else a(scene["bgs"], np.uint8)
    sel = None if bg_sel is None or bgs is None else a(bg_sel, np.int32)
    if out is None:
        out = {"rgba": np.empty((B, H, W, 4), np.uint8), "depth": np.empty((B, H, W), np.float32),
               "seg": np.empty((B, H, W), np.uint8)}
    key = out.setdefault("_key", np.empty((n_threads, H * W), np.uint64))
    lib.ab_oracle_render_batch.restype = None
    lib.ab_oracle_render_batch(
        ctypes.byref(c), B, hv.shape[1], _p(hv, ctypes.c_float), hf.shape[0], _p(hf, ctypes.c_int32),
        _p(hc, ctypes.c_uint8), _p(ht, ctypes.c_int32), _p(ov, ctypes.c_float), _p(ovo, ctypes.c_int32),
        _p(of, ctypes.c_int32), _p(ofo, ctypes.c_int32), _p(oc, ctypes.c_uint8), _p(oid, ctypes.c_int32),
        _p(op, ctypes.c_float), _p(li, ctypes.c_float), _p(bgs, ctypes.c_uint8), 0 if bgs is None else bgs.shape[1],
        0 if bgs is None else bgs.shape[2], _p(sel, ctypes.c_int32), _p(out["rgba"], ctypes.c_uint8),
        _p(out["depth"], ctypes.c_float), _p(out["seg"], ctypes.c_uint8), _p(key, ctypes.c_uint64), n_threads)
    return out
