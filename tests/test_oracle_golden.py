"""Pins oracle/ against fixtures produced by running the reference's own code (tests/golden/make_golden.py)."""
import numpy as np

from conftest import golden
from oracle import ccv, rotations
from oracle.mano_lbs import ManoLayer


def test_mano_lbs_matches_in_tree_reference(mano_model):
    """oracle/mano_lbs.py vs anakin/postprocess/iknet/manolayer.py:182-276 run on the same synthetic MANO pickle."""
    g = golden("mano_iknet.npz")
    for cidx in (0, 9):
        layer = ManoLayer(mano_model, center_idx=cidx, dtype=np.float64)
        out = layer(g["pose"], g["betas"])
        # the reference's rodrigues adds 1e-8 before the norm (manolayer.py:163): agreement is ~1e-8, not bitwise
        np.testing.assert_allclose(out.verts, g[f"verts_c{cidx}"], rtol=0, atol=2e-7)
        np.testing.assert_allclose(out.joints, g[f"joints_c{cidx}"], rtol=0, atol=2e-7)


def test_mano_identity_pose_is_shaped_template(mano_model):
    layer = ManoLayer(mano_model, dtype=np.float64)
    betas = np.random.RandomState(0).normal(size=(2, 10))
    out = layer(np.zeros((2, 48)), betas)
    v_shaped = mano_model["v_template"][None] + np.einsum("vdk,bk->bvd", mano_model["shapedirs"], betas)
    np.testing.assert_allclose(out.verts, v_shaped, atol=1e-12)
    np.testing.assert_allclose(out.joints[:, 0], layer.get_rotation_center(betas), atol=1e-12)


def test_view_engine_matches_reference():
    g = golden("view_engine.npz")
    for i in range(len(g["persp_id"])):
        rot, free, zoff = ccv.view_from_id(int(g["persp_id"][i]), int(g["u_bins"]), int(g["theta_bins"]), g["z_range"],
                                           g["r_u"][i], g["r_theta"][i], g["r_roll"][i], g["r_z"][i])
        np.testing.assert_allclose(rot, g["rotmat"][i].astype(np.float32), atol=1e-7)
        np.testing.assert_allclose(free, g["free"][i].astype(np.float32), atol=1e-7)
        np.testing.assert_allclose(zoff, g["z_offset"][i].astype(np.float32), atol=1e-7)
    np.testing.assert_array_equal(ccv.align_mat(np.array([0.0, 0.0, 1.0])), g["pole"][0])
    np.testing.assert_array_equal(ccv.align_mat(np.array([0.0, 0.0, -1.0])), g["pole"][1])


def test_row_col_and_occurrence_match_reference():
    g = golden("ovg_set.npz")
    n_obj, n_persp, n_grasp = (int(x) for x in g["shape"])
    o, p, c = ccv.row_col_calc(g["tidx"], n_persp, n_grasp)
    np.testing.assert_array_equal(o, g["obj"])
    np.testing.assert_array_equal(p, g["persp"])
    np.testing.assert_array_equal(c, g["grasp"])
    occ = ccv.occurrence_count_map(o, p, c, n_obj, n_persp, n_grasp)
    np.testing.assert_array_equal(np.argwhere(occ > 0), g["occ_nonzero"])
    np.testing.assert_array_equal(occ[occ > 0], g["occ_counts"].astype(np.int64))


def test_sample_ovg_inverse_cdf_properties():
    rng = np.random.RandomState(0)
    w = rng.uniform(0.1, 10, size=(3, 7, 5)).astype(np.float32)
    w[1, 2, :] = 0.0
    u = rng.rand(200000)
    o, p, g = ccv.sample_ovg(w, u)
    flat = (o * 7 + p) * 5 + g
    assert not np.any((o == 1) & (p == 2))  # zero-weight cells are never drawn
    freq = np.bincount(flat, minlength=w.size) / len(u)
    np.testing.assert_allclose(freq, w.reshape(-1) / w.sum(), atol=4e-3)
    # edge draws
    o, p, g = ccv.sample_ovg(w, np.array([0.0, np.nextafter(1.0, 0.0)]))
    assert (o[0], p[0], g[0]) == (0, 0, 0) and (o[1], p[1], g[1]) == (2, 6, 4)


def test_scrambler_matches_reference():
    g = golden("scrambler.npz")
    pose, tsl = ccv.random_scrambler(g["pose"], g["tsl"], g["n_tsl"], g["n_ang"])
    np.testing.assert_allclose(pose, g["out_pose"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(tsl, g["out_tsl"], rtol=0, atol=1e-8)


def test_pose_generator_matches_reference(mano_model):
    g = golden("preprocessor.npz")
    for prefix, noise in (("", True), ("clean_", False)):
        out = ccv.pose_generator(mano_model, g["pose"], g["shape"], g["tsl"], g["persp"], g["free"], g["zoff"],
                                 g["n_tsl"] if noise else None, g["n_ang"] if noise else None)
        np.testing.assert_allclose(out["final_obj_pose"], g[prefix + "obj_pose"], atol=1e-6)
        np.testing.assert_allclose(out["final_hand_verts"], g[prefix + "verts"], atol=2e-6)
        np.testing.assert_allclose(out["final_joints"], g[prefix + "joints"], atol=2e-6)


def test_ortho6d_matches_reference():
    g = golden("ortho6d.npz")
    np.testing.assert_allclose(rotations.rotmat_from_ortho6d(g["p6"]), g["R"], atol=1e-6)


def test_rotation_round_trip():
    rng = np.random.RandomState(1)
    aa = rng.normal(0, 1.0, size=(500, 3))
    aa[0] = 0.0
    aa[1] = [1e-9, 0, 0]
    R = rotations.aa_to_rotmat(aa)
    np.testing.assert_allclose(R @ R.transpose(0, 2, 1), np.broadcast_to(np.eye(3), R.shape), atol=1e-12)
    back = rotations.rotmat_to_aa(R)
    np.testing.assert_allclose(rotations.aa_to_rotmat(back), R, atol=1e-9)


def test_update_method_1_matches_reference():
    g = golden("update_method.npz")
    w1 = ccv.update_method_1(g["w0"], g["cells"], g["vals"])
    np.testing.assert_allclose(w1, g["w1"], rtol=1e-6, atol=0)
    assert w1.min() >= np.float32(0.1) and w1.max() <= np.float32(10.0)
    np.testing.assert_array_equal(w1[0, 0, :5], np.float32(0.1))  # blacklisted zeros lifted by the clamp (reference quirk)


def test_update_methods_2_to_4_match_reference():
    """oracle.ccv.update_method_2/3/4 vs the reference's own functions (artiboost_loader.py:526-598)."""
    g = golden("update_method234.npz")
    np.testing.assert_allclose(ccv.update_method_2(g["w0"], g["cells"], g["vals"]), g["w2"], rtol=1e-6, atol=0)
    w3, r3 = ccv.update_method_3(g["w0"], g["cells"], g["vals"])
    np.testing.assert_array_equal(w3, g["w3"])
    assert r3 == float(g["ratio3"]) and 0.0 < r3 < 1.0
    assert (w3 == 0).sum() > 5  # deactivated cells are NOT lifted back by a clamp (reference behaviour)
    w4a, r4a = ccv.update_method_4(g["w0"], g["cells"], g["vals"], 10, 100)
    np.testing.assert_allclose(w4a, g["w4a"], rtol=1e-6, atol=0)
    assert r4a == float(g["ratio4a"]) == -1.0
    w4b, r4b = ccv.update_method_4(g["w0"], g["cells"], g["vals"], 80, 100)
    np.testing.assert_array_equal(w4b, g["w4b"])
    assert r4b == float(g["ratio4b"])


def test_blacklist_oracle_matches_reference_function():
    """oracle.ccv.blacklist_map vs tests/golden/blacklist.npz, recorded by running the reference's own
    ArtiBoostLoader._construct_blacklist_map (artiboost_loader.py:415-500; make_golden_blacklist.py)."""
    from oracle import ccv
    g = golden("blacklist.npz")
    bl, th = ccv.blacklist_map(g["hand_pose"][:, :, :3], int(g["u_bins"]), int(g["theta_bins"]), g["rand2"], return_th=True)
    np.testing.assert_array_equal(bl, g["blacklist"])
    assert 0 < bl.sum() < bl.size and np.abs(th + 0.8).min() > 1e-4   # no cell of the fixture sits on the threshold
