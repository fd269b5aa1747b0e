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


# ---------------------------------------------------------------------- host-side helpers of the product (no kernels)
def test_host_subdivision_and_resampling_match_the_oracle():
    from artiboost_b200.artiboost.refiner import HORefiner, subdivide_mesh
    m = next(iter(fx.object_meshes().values()))
    v, f = subdivide_mesh(m.vertices, m.faces)
    np.testing.assert_array_equal(v, orf.subdivide(m.vertices, m.faces))
    assert f.shape == (4 * len(m.faces), 3) and f.max() == len(v) - 1
    # every new vertex is the midpoint of an edge of the face that produced it
    k = len(m.faces)
    np.testing.assert_allclose(v[f[:k, 1]], (m.vertices[m.faces[:, 0]] + m.vertices[m.faces[:, 1]]) / 2)
    np.random.seed(5)
    pts = np.stack([HORefiner.resample_obj(mm) for mm in fx.object_meshes().values()]).astype(np.float32)
    np.testing.assert_array_equal(pts, resampled_objects(5))


def test_nn_groups_cover_the_cloud():
    from artiboost_b200.artiboost.refiner import build_nn_groups
    pts = resampled_objects()
    sp, pm, bx = build_nn_groups(pts)
    n_obj, P, _ = pts.shape
    assert sp.shape == (n_obj, 10016, 3) and pm.shape == (n_obj, 10016) and bx.shape == (n_obj, 6, 313)
    for o in range(n_obj):
        assert sorted(set(pm[o].tolist())) == list(range(P))              # a permutation (+ padding repeats)
        np.testing.assert_array_equal(sp[o], pts[o][pm[o]])
        assert (pm[o, P:] == pm[o, P - 1]).all()                           # padded with the last sorted point
        g = sp[o].reshape(-1, 32, 3)
        assert (g >= bx[o, :3].T[:, None]).all() and (g <= bx[o, 3:].T[:, None]).all()
        # the groups are spatially tight: a Morton run of 32 surface samples spans a few centimetres at most
        assert np.median((bx[o, 3:] - bx[o, :3]).max(0)) < 0.03


def test_random_scrambler_2_bend_expansion_follows_the_reference_order():
    """expand_bend maps the five per-finger draws onto chain joints (1..12, 14, 15) exactly like scrambler.py:134-170
    (index, middle, RING -> joints 10-12, LITTLE -> joints 7-9, thumb with coefficients 1 / 0.9)."""
    import torch
    from artiboost_b200.artiboost import RandomScrambler2
    s2 = RandomScrambler2({"HAND_TSL_SIGMA": 0.01, "HAND_POSE_SIGMA": 0.1})
    b5 = torch.tensor([[1.0, 2.0, 3.0, 4.0, 5.0]])
    got = s2.expand_bend(b5)[0].numpy()
    link = np.array([1.0, 1.1, 0.9], np.float32)
    want = np.concatenate([1 * link, 2 * link, 4 * link, 3 * link, 5 * link[[0, 2]]]).astype(np.float32)
    np.testing.assert_array_equal(got, want)
    # and the oracle's per-joint angles are the same numbers
    g = golden("scrambler23.npz")
    nz = fx.scrambler_noise()
    _, _, l_axis = orf.axis_layer(g["joints"], g["transf"])
    bend14 = s2.expand_bend(torch.from_numpy(nz["bend5"])).numpy()
    p2 = orf.random_scrambler_3(orf.random_scrambler_2(g["pose"], g["joints"], g["transf"], nz["splay"], np.zeros_like(nz["bend5"]),
                                                       np.zeros_like(nz["thumb"])),
                                g["joints"], g["transf"], np.zeros_like(nz["splay"]), bend14, nz["thumb"])
    np.testing.assert_allclose(p2, g["pose2"], rtol=0, atol=5e-6)


def test_refiner_registries_and_state_dict_names():
    from artiboost_b200.artiboost import HORefiner, NullRefine, Refiner, Scrambler
    assert set(Refiner.build_mapping) == {"null", "hand_obj"}
    assert {"null", "naive", "random", "random_2", "random_3"} <= set(Scrambler.build_mapping)
    with __import__("pytest").raises(KeyError):
        Scrambler.build("nope", {})
    r = HORefiner({"PRETRAINED": None, "ITERS": 3}, mano_model=__import__("artiboost_b200.assets", fromlist=["x"]).make_synthetic_mano(0))
    ours = {k for k in r.refine_net.state_dict() if not k.startswith("mano_layer.")}
    ref = set(fx.refinenet_state()) | {k.replace("running_mean", "num_batches_tracked") for k in fx.refinenet_state()
                                       if k.endswith("running_mean")}
    assert ours == ref
    missing = r.refine_net.load_state_dict({k: __import__("torch").from_numpy(v) for k, v in fx.refinenet_state().items()}, strict=False)
    assert not missing.unexpected_keys
    assert isinstance(Refiner.build("null", None, mano_model=r.refine_net.mano_layer and
                                    __import__("artiboost_b200.assets", fromlist=["x"]).make_synthetic_mano(0)), NullRefine)
