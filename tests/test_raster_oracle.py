"""Known-answer tests of the raster rule set on the CPU oracle (oracle/raster.c).  The reference renderer
(pyrender/OpenGL) has no golden images and cannot run here: "parity unpinned"; these tests pin the rules themselves."""
import numpy as np

from oracle import raster

CFG = dict(width=16, height=16, fx=16.0, fy=16.0, cx=8.0, cy=8.0, znear=0.05, cull_backface=0, ambient=0.8,
           diffuse=0.25)
WHITE = np.full((8, 4), 255, np.uint8)


def screen_tri(pts, z=1.0):
    """Camera-space triangle whose projection hits the given pixel coordinates exactly (power-of-two intrinsics)."""
    pts = np.asarray(pts, np.float64)
    zz = np.broadcast_to(np.asarray(z, np.float64), (len(pts),))
    return np.stack([(pts[:, 0] - CFG["cx"]) / CFG["fx"] * zz, (pts[:, 1] - CFG["cy"]) / CFG["fy"] * zz, zz], 1).astype(np.float32)


def render(verts, faces, cfg=CFG, cols=None, **kw):
    verts = np.asarray(verts, np.float32)
    cols = np.full((len(verts), 4), 255, np.uint8) if cols is None else cols
    return raster.render_view(cfg, verts, np.asarray(faces, np.int32), cols, **kw)


def test_pixel_centre_sampling_and_top_left_rule():
    # axis-aligned right triangle with corners on pixel CENTRES (2.5,2.5) (6.5,2.5) (2.5,6.5), both windings
    for faces in ([[0, 1, 2]], [[0, 2, 1]]):
        rgba, depth, seg, key = render(screen_tri([(2.5, 2.5), (6.5, 2.5), (2.5, 6.5)]), faces)
        cov = seg > 0
        # top edge (y=2.5, interior below) and left edge (x=2.5) are included; the hypotenuse x+y=9 is a right/bottom
        # edge and is excluded
        exp = np.zeros((16, 16), bool)
        for py in range(16):
            for px in range(16):
                x, y = px + 0.5, py + 0.5
                exp[py, px] = (x >= 2.5) and (y >= 2.5) and (x + y < 9.0)
        np.testing.assert_array_equal(cov, exp)
        assert cov.sum() == 10


def test_shared_edge_is_covered_exactly_once():
    # a quad split along its diagonal: every pixel centre inside the quad belongs to exactly one triangle
    v = screen_tri([(1.5, 1.5), (9.5, 1.5), (9.5, 9.5), (1.5, 9.5)])
    _, _, seg_a, key_a = render(v, [[0, 1, 2]])
    _, _, seg_b, key_b = render(v, [[0, 2, 3]])
    _, _, seg_ab, _ = render(v, [[0, 1, 2], [0, 2, 3]])
    assert not np.any((seg_a > 0) & (seg_b > 0))
    np.testing.assert_array_equal((seg_a > 0) | (seg_b > 0), seg_ab > 0)
    assert (seg_ab > 0).sum() == 64  # centres 1.5..8.5 in both axes: the right/bottom edges at 9.5 are excluded


def test_fronto_parallel_depth_is_exact_and_background_is_zero():
    v = screen_tri([(0.0, 0.0), (16.0, 0.0), (0.0, 16.0)], z=0.5)
    rgba, depth, seg, _ = render(v, [[0, 1, 2]])
    assert np.all(depth[seg > 0] == np.float32(0.5))
    assert np.all(depth[seg == 0] == 0.0) and np.all(rgba[seg == 0, 3] == 0) and np.all(rgba[seg > 0, 3] == 255)
    np.testing.assert_array_equal(rgba[seg == 0, :3], 128)  # bg_color 0.5 (renderer.py:77)


def test_perspective_correct_depth_on_a_slanted_triangle():
    # depth is 1 / (barycentric interpolation of 1/z) in screen space = exact plane depth along the pixel ray
    pts = [(1.0, 1.0), (15.0, 2.0), (3.0, 14.0)]
    zs = np.array([0.4, 0.7, 1.1])
    v = screen_tri(pts, zs)
    _, depth, seg, _ = render(v, [[0, 1, 2]])
    n = np.cross(v[1].astype(np.float64) - v[0], v[2].astype(np.float64) - v[0])
    d = n @ v[0].astype(np.float64)
    ys, xs = np.nonzero(seg)
    assert len(ys) > 30
    ray = np.stack([(xs + 0.5 - CFG["cx"]) / CFG["fx"], (ys + 0.5 - CFG["cy"]) / CFG["fy"], np.ones(len(xs))], 1)
    np.testing.assert_allclose(depth[ys, xs], d / (ray @ n), rtol=2e-6)


def test_z_test_nearest_wins_and_ties_go_to_lower_primitive_id():
    near = screen_tri([(2.0, 2.0), (12.0, 2.0), (2.0, 12.0)], z=0.5)
    far = screen_tri([(2.0, 2.0), (12.0, 2.0), (2.0, 12.0)], z=0.9)
    v = np.concatenate([far, near])
    _, depth, seg, key = render(v, [[0, 1, 2], [3, 4, 5]])
    assert np.all(depth[seg > 0] == np.float32(0.5))
    assert np.all((key[seg > 0] & 0xffffffff) == 1)
    # identical triangles: the lower primitive id wins
    v = np.concatenate([near, near])
    _, _, seg, key = render(v, [[0, 1, 2], [3, 4, 5]])
    assert np.all((key[seg > 0] & 0xffffffff) == 0)
    # hand faces are numbered after object faces: a coincident hand triangle loses to the object (draw order
    # renderer.py:90-93)
    cols = np.full((3, 4), 255, np.uint8)
    _, _, seg, _ = raster.render_view(CFG, near, np.array([[0, 1, 2]], np.int32), cols, near, np.array([[0, 1, 2]], np.int32),
                                      cols, np.eye(4, dtype=np.float32))
    assert set(np.unique(seg)) == {0, 2}


def test_backface_culling_and_near_plane():
    v = screen_tri([(2.0, 2.0), (12.0, 2.0), (2.0, 12.0)])
    cull = dict(CFG, cull_backface=1)
    # x right, y down: (0,1,2) winds clockwise on screen = counter-clockwise seen from the camera side with y up
    n_a = (render(v, [[0, 1, 2]], cfg=cull)[2] > 0).sum()
    n_b = (render(v, [[0, 2, 1]], cfg=cull)[2] > 0).sum()
    assert sorted([n_a, n_b])[0] == 0 and sorted([n_a, n_b])[1] > 0
    # outward-wound closed mesh (assets convention) seen from outside: the visible winding is the kept one
    p0, p1, p2 = v[0], v[2], v[1]
    nrm = np.cross(p1 - p0, p2 - p0)
    assert nrm[2] < 0  # outward normal faces the camera (camera looks down +z)
    assert (render(v, [[0, 2, 1]], cfg=cull)[2] > 0).sum() > 0
    # any vertex closer than znear discards the triangle
    vz = v.copy()
    vz[1, 2] = 0.01
    assert (render(vz, [[0, 1, 2]])[2] > 0).sum() == 0


def test_degenerate_and_offscreen_triangles():
    v = screen_tri([(2.0, 2.0), (6.0, 6.0), (10.0, 10.0)])  # zero area
    assert (render(v, [[0, 1, 2]])[2] > 0).sum() == 0
    v = screen_tri([(-50.0, -50.0), (-10.0, -50.0), (-50.0, -10.0)])
    assert (render(v, [[0, 1, 2]])[2] > 0).sum() == 0
    v = screen_tri([(-100.0, -100.0), (300.0, -100.0), (-100.0, 300.0)])  # covers the whole frame
    assert (render(v, [[0, 1, 2]])[2] > 0).all()
    v = screen_tri([(4.6, 4.6), (4.9, 4.6), (4.6, 4.9)])  # sub-pixel, misses the centre (4.5, 4.5)
    assert (render(v, [[0, 1, 2]])[2] > 0).sum() == 0
    v = screen_tri([(4.4, 4.4), (4.7, 4.4), (4.4, 4.7)])  # sub-pixel, contains the centre
    assert (render(v, [[0, 1, 2]])[2] > 0).sum() == 1


def test_shading_known_answer_and_background_crop():
    # fronto-parallel white triangle at z=1 through the optical axis: cos = 1, d = 1 at the principal point
    cfg = dict(CFG, cx=8.5, cy=8.5)
    v = np.array([[-1, -1, 1], [2, -1, 1], [-1, 2, 1]], np.float32)
    bg = np.arange(24 * 24 * 3, dtype=np.uint32).reshape(24, 24, 3).astype(np.uint8)
    rgba, depth, seg, _ = render(v, [[0, 1, 2]], cfg=cfg, light=2.0, bg=bg, bg_sel=[2, 3, 16, 16])
    assert seg[8, 8] == 1
    assert rgba[8, 8, 0] == 255  # 255 * (0.8 + 0.25 * 2 * 1 / 1) clamps
    rgba2, _, _, _ = render(v, [[0, 1, 2]], cfg=cfg, light=2.0, cols=np.full((3, 4), 100, np.uint8))
    assert rgba2[8, 8, 0] == 130  # 100 * 1.3
    # background: nearest-neighbour crop, identity scale here -> bg[3 + py, 2 + px]
    ys, xs = np.nonzero(seg == 0)
    np.testing.assert_array_equal(rgba[ys, xs, :3], bg[3 + ys, 2 + xs])
