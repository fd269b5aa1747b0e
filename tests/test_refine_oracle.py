"""Pins oracle/refine.py (hand-object refiner, anatomical scramblers, staged pose generator) against fixtures recorded
by running the reference's own refiner.py / scrambler.py / preprocessor.py (tests/golden/make_golden_refine.py)."""
import os
import sys

import numpy as np

from conftest import GOLDEN, golden
from oracle import ccv, refine as orf

sys.path.insert(0, GOLDEN)
import refine_fixture as fx  # noqa: E402


def resampled_objects(seed=5):
    """HORefiner.setup (refiner.py:163-169) on the fixture meshes: subdivide to >= 10 000 vertices, draw 10 000."""
    np.random.seed(seed)
    out = []
    for m in fx.object_meshes().values():
        v, f = np.asarray(m.vertices), np.asarray(m.faces)
        while len(v) < fx.N_SAMPLE:
            v = orf.subdivide(v, f)
            assert len(v) >= fx.N_SAMPLE, "the fixture objects need one subdivision level"
        out.append(v[np.random.choice(len(v), fx.N_SAMPLE, replace=False)].astype(np.float32))
    return np.stack(out)


def test_chamfer_oracle_is_the_nearest_neighbour():
    rng = np.random.RandomState(0)
    x, y = rng.normal(0, 0.1, (97, 3)).astype(np.float32), rng.normal(0, 0.1, (1501, 3)).astype(np.float32)
    y[700] = y[30]  # an exact tie: the first index wins
    x[5] = y[30]
    d, i = orf.chamfer_nn(x, y)
    d64 = np.linalg.norm(x[:, None].astype(np.float64) - y[None].astype(np.float64), axis=2)
    np.testing.assert_allclose(d, d64.min(1), rtol=2e-6, atol=1e-9)
    assert i[5] == 30 and d[5] == 0.0
    assert np.all(np.abs(d64[np.arange(len(x)), i] - d64.min(1)) < 1e-7)


def test_resampled_objects_match_reference_run():
    g = golden("refiner.npz")
    pts = resampled_objects()
    np.testing.assert_array_equal(pts[:, :16], g["pts_head"])
    np.testing.assert_allclose(pts.astype(np.float64).sum(axis=(1, 2)), g["pts_sum"], rtol=0, atol=1e-9)


def test_ho_refiner_oracle_matches_reference(mano_model):
    g = golden("refiner.npz")
    net = orf.RefineNet(fx.refinenet_state(), mano_model, n_iters=3)
    out = orf.ho_refiner(net, resampled_objects(), g["obj_id"], g["pose"], g["tsl"], g["obj_rot"])
    np.testing.assert_allclose(out["h2o"], g["h2o"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(out["hand_tsl"], g["hand_tsl"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(out["hand_pose"], g["hand_pose"], rtol=0, atol=2e-5)
    np.testing.assert_allclose(out["hand_verts"], g["hand_verts"], rtol=0, atol=5e-6)
    np.testing.assert_allclose(out["joints"], g["joints"], rtol=0, atol=5e-6)
    # the refinement does something: the fixture is not a fixed point of the network
    assert np.abs(g["hand_pose"] - g["pose"]).max() > 1e-2


def test_anatomical_scramblers_oracle_matches_reference():
    g = golden("scrambler23.npz")
    nz = fx.scrambler_noise()
    p2 = orf.random_scrambler_2(g["pose"], g["joints"], g["transf"], nz["splay"], nz["bend5"], nz["thumb"])
    p3 = orf.random_scrambler_3(g["pose"], g["joints"], g["transf"], nz["splay"], nz["bend14"], nz["thumb"])
    np.testing.assert_allclose(p2, g["pose2"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(p3, g["pose3"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(g["tsl"] + nz["tsl"], g["tsl2"], rtol=0, atol=1e-7)
    # the root rotation and nothing else is left untouched
    assert np.array_equal(p2[:, :3], g["pose"][:, :3]) and np.abs(p2[:, 3:] - g["pose"][:, 3:]).reshape(6, 15, 3).max(2).min() > 0


def test_axis_layer_is_orthonormal(mano_model):
    g = golden("scrambler23.npz")
    b, u, l = orf.axis_layer(g["joints"], g["transf"])
    for a in (b, u, l):
        np.testing.assert_allclose(np.linalg.norm(a, axis=2), 1.0, atol=1e-6)
    np.testing.assert_allclose(np.einsum("bki,bki->bk", b, l), 0.0, atol=1e-6)
    np.testing.assert_allclose(np.einsum("bki,bki->bk", b, u), 0.0, atol=1e-6)
    np.testing.assert_allclose(np.einsum("bki,bki->bk", u, l), 0.0, atol=1e-6)


def staged_pose_generator_oracle(mano_model, g, nz, pts, iters=2):
    net = orf.RefineNet(fx.refinenet_state(), mano_model, n_iters=iters)

    def scrambler(feed):
        pose = orf.random_scrambler_2(feed["hand_pose"], feed["joints"], feed["hand_transf"], nz["splay"], nz["bend5"], nz["thumb"])
        return pose, (feed["hand_tsl"] + nz["tsl"]).astype(np.float32)

    def refiner(pose, tsl, obj_rot):
        return orf.ho_refiner(net, pts, g["obj_id"], pose, tsl, obj_rot)

    return ccv.pose_generator(mano_model, g["pose"], g["shape"], g["tsl"], g["persp"], g["free"], g["zoff"],
                              scrambler=scrambler, refiner=refiner)


def test_staged_pose_generator_oracle_matches_reference(mano_model):
    g = golden("preprocessor_staged.npz")
    out = staged_pose_generator_oracle(mano_model, g, fx.scrambler_noise(seed=23, B=4), resampled_objects())
    np.testing.assert_allclose(out["final_obj_pose"], g["obj_pose"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(out["final_hand_verts"], g["verts"], rtol=0, atol=1e-5)
    np.testing.assert_allclose(out["final_joints"], g["joints"], rtol=0, atol=1e-5)
