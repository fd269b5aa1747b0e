"""Cluster-level back-face culling data for the triangle pass: runs of consecutive faces
with a normal cone and a bounding sphere, built once per mesh on the host.  A whole cluster can be skipped for a view
when every triangle in it is certainly back-facing, before any of its vertices is fetched.  Face order (and with it
the primitive ids of the z-test key, renderer.py:90-93 insertion order) is left untouched, so skipping culled clusters
cannot change a single output bit as long as the test is conservative; tests/test_mesh_clusters.py checks exactly that
against the oracle rasteriser.  Host-side preparation only: the kernel does not consume it yet.
"""
from __future__ import annotations

import numpy as np


def morton_face_order(vertices: np.ndarray, faces: np.ndarray) -> np.ndarray:
    """Permutation that sorts the faces along a 30-bit Morton curve of their centroids (stable): consecutive runs become
    compact surface patches instead of whatever strips the modelling tool emitted.  The rasteriser would walk the faces in
    this order but keep the ORIGINAL index as the primitive id of the z-test key, so ties still break as before."""
    v = np.asarray(vertices, np.float64)
    c = v[np.asarray(faces, np.int64)[:, :3]].mean(1)
    lo, hi = c.min(0), c.max(0)
    q = np.clip(((c - lo) / np.maximum(hi - lo, 1e-12) * 1023.0).astype(np.int64), 0, 1023).astype(np.uint64)

    def spread(x):
        x = (x | (x << np.uint64(16))) & np.uint64(0x030000FF)
        x = (x | (x << np.uint64(8))) & np.uint64(0x0300F00F)
        x = (x | (x << np.uint64(4))) & np.uint64(0x030C30C3)
        x = (x | (x << np.uint64(2))) & np.uint64(0x09249249)
        return x

    code = spread(q[:, 0]) | (spread(q[:, 1]) << np.uint64(1)) | (spread(q[:, 2]) << np.uint64(2))
    return np.argsort(code, kind="stable")


def build_clusters(vertices: np.ndarray, faces: np.ndarray, size: int = 64, order: np.ndarray = None):
    """-> dict(first i32[n], count i32[n], center f32[n,3], radius f32[n], axis f32[n,3], cutoff f32[n][, order]).
    axis / cutoff: unit cone axis and min over the cluster's (non-degenerate) faces of dot(axis, face normal);
    cutoff < 0 means the normals span more than a half-space (such a cluster is never culled).
    order: optional face permutation (morton_face_order); cluster k then holds faces order[first[k] : first[k] + count[k]]."""
    v = np.asarray(vertices, np.float64)
    f = np.asarray(faces, np.int64)[:, :3]
    if order is not None:
        f = f[np.asarray(order)]
    n = np.cross(v[f[:, 1]] - v[f[:, 0]], v[f[:, 2]] - v[f[:, 0]])
    ln = np.linalg.norm(n, axis=1)
    ok = ln > 0
    nn = np.where(ok[:, None], n / np.maximum(ln, 1e-300)[:, None], 0.0)
    first = np.arange(0, len(f), size)
    out = {k: [] for k in ("first", "count", "center", "radius", "axis", "cutoff")}
    for a in first:
        b = min(a + size, len(f))
        pts = v[np.unique(f[a:b])]
        c = (pts.min(0) + pts.max(0)) / 2
        r = np.linalg.norm(pts - c, axis=1).max()
        m = nn[a:b][ok[a:b]]
        if len(m):
            axis = m.sum(0)
            la = np.linalg.norm(axis)
            axis = axis / la if la > 0 else np.array([0.0, 0.0, 1.0])
            cutoff = float((m @ axis).min())
        else:  # only degenerate faces: nothing is ever drawn, always cullable
            axis, cutoff = np.array([0.0, 0.0, 1.0]), 1.0
        for k, val in zip(out, (a, b - a, c, r, axis, cutoff)):
            out[k].append(val)
    res = {"first": np.asarray(out["first"], np.int32), "count": np.asarray(out["count"], np.int32),
           "center": np.asarray(out["center"], np.float32), "radius": np.asarray(out["radius"], np.float32),
           "axis": np.asarray(out["axis"], np.float32), "cutoff": np.asarray(out["cutoff"], np.float32)}
    if order is not None:
        res["order"] = np.asarray(order, np.int64)
    return res


def culled_faces(clusters: dict, culled: np.ndarray) -> np.ndarray:
    """bool per ORIGINAL face index: member of a culled cluster."""
    per_slot = np.repeat(np.asarray(culled, bool), clusters["count"])
    if "order" not in clusters:
        return per_slot
    out = np.zeros(len(per_slot), bool)
    out[clusters["order"]] = per_slot
    return out


def culled_clusters(clusters: dict, pose: np.ndarray, front_sign: float, margin: float = 0.02) -> np.ndarray:
    """bool[n]: clusters whose triangles are all back-facing for the pinhole camera at the origin of the frame `pose`
    (4x4, object -> camera) maps into.  front_sign = +1 / -1: sign of dot(face normal, point) for which the rasteriser's
    rule calls a triangle back-facing (fixed by the mesh winding and the y-down image; see front_sign_of).
    For a normal n within the cone (angle to the axis <= acos(cutoff)) and a point p within the sphere,
        sign * dot(n, p) >= |s| cos(phi + theta) - r,   phi = angle(sign * axis, s), s = sphere centre,
    and the cluster is culled when that bound exceeds margin * |s| (the margin covers the 24.8 vertex snapping, which can
    flip the sign of slivers seen within ~0.2 degrees of edge-on)."""
    R, t = np.asarray(pose, np.float64)[:3, :3], np.asarray(pose, np.float64)[:3, 3]
    s = clusters["center"].astype(np.float64) @ R.T + t
    a = front_sign * (clusters["axis"].astype(np.float64) @ R.T)
    c = clusters["cutoff"].astype(np.float64)
    r = clusters["radius"].astype(np.float64)
    ls = np.linalg.norm(s, axis=1)
    cos_phi = np.einsum("ij,ij->i", a, s) / np.maximum(ls, 1e-30)
    sin_phi = np.sqrt(np.maximum(0.0, 1 - cos_phi ** 2))
    sin_th = np.sqrt(np.maximum(0.0, 1 - c ** 2))
    bound = ls * (cos_phi * c - sin_phi * sin_th) - r
    # cos(phi + theta) is only a lower bound of dot(n, s)/|s| while phi + theta <= pi; cutoff <= 0 clusters are never culled
    valid = (c > 0) & (cos_phi > 0)
    return valid & (bound > margin * ls)


def front_sign_of(vertices: np.ndarray, faces: np.ndarray) -> float:
    """+1 when the rasteriser's rule culls faces whose geometric normal points away from the camera (dot(n, p) > 0), for
    a mesh with outward normals; derived from the rule itself: a triangle is culled when its projected signed area
    (x right, y down) is positive."""
    v = np.asarray(vertices, np.float64)
    f = np.asarray(faces, np.int64)[:, :3]
    # one well-conditioned face, placed 1 m in front of the camera
    n = np.cross(v[f[:, 1]] - v[f[:, 0]], v[f[:, 2]] - v[f[:, 0]])
    i = int(np.argmax(np.linalg.norm(n, axis=1)))
    tri = v[f[i]] - v[f[i]].mean(0) + np.array([0.0, 0.0, 1.0])
    uv = tri[:, :2] / tri[:, 2:3]
    area2 = (uv[1, 0] - uv[0, 0]) * (uv[2, 1] - uv[0, 1]) - (uv[2, 0] - uv[0, 0]) * (uv[1, 1] - uv[0, 1])
    d = float(np.dot(n[i], tri.mean(0)))
    return 1.0 if (area2 > 0) == (d > 0) else -1.0
