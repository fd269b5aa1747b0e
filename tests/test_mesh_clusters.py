"""Cluster-level back-face culling (artiboost_b200/artiboost/mesh_clusters.py) must be conservative: rendering with the
culled clusters' faces removed gives the oracle rasteriser's output bit for bit (colour, depth bits, segmentation and the
z-test keys, i.e. the winning primitive of every pixel)."""
import numpy as np

from oracle import ccv, raster
from artiboost_b200 import assets
from artiboost_b200.artiboost.mesh_clusters import build_clusters, culled_clusters, culled_faces, front_sign_of, morton_face_order

CFG = dict(width=256, height=256, fx=217.5, fy=217.5, cx=128.0, cy=128.0, znear=0.05, cull_backface=1, ambient=0.8, diffuse=0.25)


def test_cluster_culling_is_conservative_and_worth_it(mano_model, objects):
    rng = np.random.RandomState(0)
    hand_cols = np.full((778, 4), 200, np.uint8)
    hv = mano_model["v_template"].astype(np.float32) + np.array([0.05, 0.0, 0.5], np.float32)
    hf = np.asarray(mano_model["f"], np.int32)
    fractions, fractions_patch = [], []
    for name, o in objects.items():
        verts, faces = o["vertices"], np.asarray(o["faces"], np.int32)[:, :3]
        cl = build_clusters(verts, faces, size=32)              # one warp of the triangle pass
        clp = build_clusters(verts, faces, size=32, order=morton_face_order(verts, faces))   # compact patches, ids remapped
        assert cl["first"][0] == 0 and int(cl["count"].sum()) == len(faces)
        sign = front_sign_of(verts, faces)
        cols = np.concatenate([o["colors"][:, :3], np.full((len(verts), 1), 255, np.uint8)], 1) if o["colors"].shape[1] == 3 else o["colors"]
        for _ in range(6):
            rot, free, zoff = ccv.view_from_id(int(rng.randint(288)), 12, 24, (0.45, 0.55), *rng.rand(4))
            pose = np.eye(4, dtype=np.float32)
            pose[:3, :3] = free[:3, :3] @ rot.T
            pose[:3, 3] = zoff + rng.normal(0, 0.03, 3)
            a = raster.render_view(CFG, hv, hf, hand_cols, verts, faces, cols, pose)
            assert (a[2] == 2).sum() > 500                   # the object is in view
            for clusters, acc in ((cl, fractions), (clp, fractions_patch)):
                gone = culled_faces(clusters, culled_clusters(clusters, pose, sign))
                f2 = faces.copy()
                f2[gone] = 0                                 # degenerate (area 0): discarded by the rules, ids preserved
                b = raster.render_view(CFG, hv, hf, hand_cols, verts, f2, cols, pose)
                for x, y, what in zip(a, b, ("rgba", "depth", "seg", "key")):
                    assert np.array_equal(x.view(np.uint8), y.view(np.uint8)), (name, what)
                acc.append(gone.mean())
    # closed meshes seen from outside: about half of the faces are back-facing; whole clusters account for a good part
    assert np.mean(fractions) > 0.25, np.mean(fractions)
    assert np.mean(fractions_patch) > np.mean(fractions)
    print("faces skipped by cluster culling: face order mean %.3f (min %.3f max %.3f), Morton patches mean %.3f (min %.3f max %.3f)"
          % (np.mean(fractions), min(fractions), max(fractions), np.mean(fractions_patch), min(fractions_patch), max(fractions_patch)))


def test_front_sign_matches_the_rasteriser_rule(objects):
    """Calibrates the sign convention against the oracle: with culling on, a single outward-facing triangle seen from outside
    is drawn, its mirror image is not."""
    o = next(iter(objects.values()))
    verts, faces = o["vertices"], np.asarray(o["faces"], np.int32)[:, :3]
    sign = front_sign_of(verts, faces)
    n = np.cross(verts[faces[:, 1]] - verts[faces[:, 0]], verts[faces[:, 2]] - verts[faces[:, 0]])
    cen = verts[faces].mean(1)
    outward = np.einsum("ij,ij->i", n, cen - verts.mean(0)) > 0
    assert outward.mean() > 0.95                               # the synthetic meshes are wound outwards
    pose = np.eye(4, dtype=np.float32)
    pose[2, 3] = 0.5
    p = cen @ pose[:3, :3].T + pose[:3, 3]
    back = sign * np.einsum("ij,ij->i", n @ pose[:3, :3].T, p) > 0
    hv = np.zeros((778, 3), np.float32) + np.array([0, 0, -1.0], np.float32)   # hand behind the camera: not drawn
    hf = np.zeros((1, 3), np.int32)
    cols = np.full((len(verts), 4), 255, np.uint8)
    key = raster.render_view(CFG, hv, hf, np.full((778, 4), 9, np.uint8), verts, faces, cols, pose)[3]
    winners = np.unique(key[key != np.uint64(0xFFFFFFFFFFFFFFFF)] & np.uint64(0xFFFFFFFF)).astype(np.int64)
    assert len(winners) > 100 and not back[winners].any()      # nothing the rule calls back-facing ever wins a pixel
