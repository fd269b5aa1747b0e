"""GPU parity: the CUDA path through the C-ABI vs the CPU oracle (and the reference-generated golden fixtures)."""
import numpy as np
import pytest
import torch

from conftest import golden

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def mano_layer(mano_model, lib_built):
    from artiboost_b200.manolayer import ManoLayer
    return ManoLayer(mano_model=mano_model).to(DEV)


def t(a, dtype=torch.float32):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV).to(dtype)


# ------------------------------------------------------------------------------------------------- MANO LBS
@pytest.mark.parametrize("B", [1, 3, 37, 256])
def test_mano_forward_matches_oracle(mano_layer, mano_model, B):
    from oracle.mano_lbs import ManoLayer as OracleMano
    rng = np.random.RandomState(B)
    pose = rng.normal(0, 0.4, size=(B, 48)).astype(np.float32)
    betas = rng.normal(0, 1, size=(B, 10)).astype(np.float32)
    pose[0, 3:6] = 0.0
    ref = OracleMano(mano_model, dtype=np.float64)(pose.astype(np.float64), betas.astype(np.float64))
    out = mano_layer(t(pose), t(betas))
    # tolerance from BASELINE.json north_star: 1e-4 relative on vertex positions (scale: hand ~0.2 m)
    np.testing.assert_allclose(out.verts.cpu().numpy(), ref.verts, rtol=1e-4, atol=2e-6)
    np.testing.assert_allclose(out.joints.cpu().numpy(), ref.joints, rtol=1e-4, atol=2e-6)
    np.testing.assert_allclose(out.transforms_abs.cpu().numpy(), ref.transforms_abs, rtol=1e-4, atol=2e-6)
    # betas=None is the NullRefine call (refiner.py:138)
    ref0 = OracleMano(mano_model, dtype=np.float64)(pose.astype(np.float64))
    out0 = mano_layer(t(pose))
    np.testing.assert_allclose(out0.verts.cpu().numpy(), ref0.verts, rtol=1e-4, atol=2e-6)


def test_mano_forward_matches_reference_golden(mano_model, lib_built):
    """Directly against anakin/postprocess/iknet/manolayer.py outputs (tests/golden/mano_iknet.npz)."""
    from artiboost_b200.manolayer import ManoLayer
    g = golden("mano_iknet.npz")
    for cidx in (0, 9):
        layer = ManoLayer(mano_model=mano_model, center_idx=cidx).to(DEV)
        out = layer(t(g["pose"]), t(g["betas"]))
        np.testing.assert_allclose(out.verts.cpu().numpy(), g[f"verts_c{cidx}"], rtol=1e-4, atol=2e-6)
        np.testing.assert_allclose(out.joints.cpu().numpy(), g[f"joints_c{cidx}"], rtol=1e-4, atol=2e-6)


def test_mano_rejects_bad_input(mano_layer):
    with pytest.raises(ValueError):
        mano_layer(torch.zeros((2, 45), device=DEV))
    from artiboost_b200.lib import AbError
    with pytest.raises(AbError):
        mano_layer(torch.zeros((2, 48)))  # host tensor: there is no CPU path
    out = mano_layer(torch.zeros((0, 48), device=DEV))
    assert out.verts.shape == (0, 778, 3)


# -------------------------------------------------------------------------------------- sampler / view engine
def test_ccv_sample_is_integer_exact(lib_built):
    from artiboost_b200.artiboost import OVGSet
    from oracle import ccv
    from types import SimpleNamespace
    rng = np.random.RandomState(0)
    shape = (4, 288, 50)
    w = rng.uniform(0.1, 10, size=shape).astype(np.float32)
    w[2, 17, :] = 0.0
    n = 34560
    u = rng.rand(n).astype(np.float32)
    u[:3] = [0.0, np.nextafter(np.float32(1), np.float32(0)), 0.5]
    ovg = OVGSet(SimpleNamespace(obj_names=list("abcd")), None, SimpleNamespace(n_persp_center=288), n, 100, 50,
                 torch.zeros(shape, dtype=torch.bool), device=DEV)
    wmap, occ = ovg.update(t(w), torch.zeros(shape, dtype=torch.bool, device=DEV), uniforms=t(u))
    o, p, g = ccv.sample_ovg(w, u)
    np.testing.assert_array_equal(ovg.sampled_obj_idx.cpu().numpy(), o)
    np.testing.assert_array_equal(ovg.sampled_persp_idx.cpu().numpy(), p)
    np.testing.assert_array_equal(ovg.sampled_grasp_idx.cpu().numpy(), g)
    np.testing.assert_array_equal(occ.cpu().numpy(), ccv.occurrence_count_map(o, p, g, *shape) > 0)
    # val mode: uniform without replacement over non-blacklisted cells (ovg_set.py:107-119)
    bl = torch.zeros(shape, dtype=torch.bool)
    bl[0] = True
    ovg = OVGSet(SimpleNamespace(obj_names=list("abcd")), None, SimpleNamespace(n_persp_center=288), n, 5000, 50, bl,
                 device=DEV)
    ovg.val()
    ovg.update(t(w), torch.zeros(shape, dtype=torch.bool, device=DEV))
    assert len(ovg) == 5000 and ovg.sampled_obj_idx.min().item() >= 1
    assert ovg.sampled_idx_tensor.unique().numel() == 5000


def test_view_engine_matches_oracle_and_golden(lib_built):
    from artiboost_b200.artiboost import ViewEngine
    from oracle import ccv
    g = golden("view_engine.npz")
    ve = ViewEngine({"PERSP_U_BINS": 12, "PERSP_THETA_BINS": 24, "CAMERA_Z_RANGE": [0.45, 0.55]})
    rand4 = np.stack([g["r_u"], g["r_theta"], g["r_roll"], g["r_z"]], 1).astype(np.float32)
    rot, free, zoff = ve.get_view_batch(t(g["persp_id"], torch.int32), t(rand4))
    # r_roll is float64 in the reference (np.random.rand) and fp32 here: 2*pi*2^-24 ~ 4e-7 of angle
    np.testing.assert_allclose(rot.cpu().numpy(), g["rotmat"], atol=1e-6)
    np.testing.assert_allclose(free.cpu().numpy(), g["free"], atol=1e-6)
    np.testing.assert_array_equal(zoff.cpu().numpy(), g["z_offset"].astype(np.float32))
    rng = np.random.RandomState(5)
    ids = np.arange(288, dtype=np.int32)
    r4 = rng.rand(288, 4).astype(np.float32)
    rot, free, zoff = ve.get_view_batch(t(ids, torch.int32), t(r4))
    for i in range(288):
        a, b, c = ccv.view_from_id(int(ids[i]), 12, 24, (0.45, 0.55), *r4[i])
        np.testing.assert_allclose(rot[i].cpu().numpy(), a, atol=1e-6)
        np.testing.assert_allclose(free[i].cpu().numpy(), b, atol=1e-6)
        np.testing.assert_array_equal(zoff[i].cpu().numpy(), c)
    R = rot.double()
    # orthogonal up to the reference's own fp32 direction vector: its 6e-8 norm error is amplified by 1/(1 + z.v)
    # next to the south pole (view_engine.py:83-84), so the bound is loose there by construction
    assert torch.allclose(R @ R.transpose(1, 2), torch.eye(3, device=DEV, dtype=torch.float64).expand_as(R), atol=1e-3)


# ------------------------------------------------------------------------------------------- pose generator
def _posegen(mano_model, feed, noise):
    from artiboost_b200.artiboost import NullRefine, PreProcessorPoseGenerator, Scrambler

    class Fixed(Scrambler):
        def sample_noise(self, batch_size, device, generator=None):
            return noise

    refiner = NullRefine(mano_model=mano_model).to(DEV)
    gen = PreProcessorPoseGenerator(refiner, Fixed(), refiner.refine_net.mano_layer, refiner.refine_net.mano_layer)
    return gen(feed)


def test_pose_generator_matches_reference_golden(mano_model, lib_built):
    g = golden("preprocessor.npz")
    feed = {"hand_pose": t(g["pose"]), "hand_shape": t(g["shape"]), "hand_tsl": t(g["tsl"]),
            "persp_rotmat": t(g["persp"]), "camera_free_transf": t(g["free"]), "z_offset": t(g["zoff"])}
    for prefix, noise in (("", (t(g["n_tsl"]), t(g["n_ang"]))), ("clean_", (None, None))):
        out = _posegen(mano_model, feed, noise)
        np.testing.assert_allclose(out["final_obj_pose"].cpu().numpy(), g[prefix + "obj_pose"], rtol=1e-4, atol=2e-6)
        np.testing.assert_allclose(out["final_hand_verts"].cpu().numpy(), g[prefix + "verts"], rtol=1e-4, atol=5e-6)
        np.testing.assert_allclose(out["final_joints"].cpu().numpy(), g[prefix + "joints"], rtol=1e-4, atol=5e-6)


def test_pose_generator_matches_oracle_batch(mano_model, lib_built):
    from oracle import ccv
    rng = np.random.RandomState(11)
    B = 203
    pose = rng.normal(0, 0.35, size=(B, 48)).astype(np.float32)
    shape = rng.normal(0, 1, size=(B, 10)).astype(np.float32)
    tsl = rng.normal(0, 0.08, size=(B, 3)).astype(np.float32)
    views = [ccv.view_from_id(int(rng.randint(288)), 12, 24, (0.45, 0.55), *rng.rand(4)) for _ in range(B)]
    persp, free, zoff = (np.stack([v[i] for v in views]) for i in range(3))
    n_tsl = rng.normal(0, 0.01, size=(B, 3)).astype(np.float32)
    n_ang = rng.normal(0, 0.1, size=(B, 16)).astype(np.float32)
    ref = ccv.pose_generator(mano_model, pose, shape, tsl, persp, free, zoff, n_tsl, n_ang)
    feed = {"hand_pose": t(pose), "hand_shape": t(shape), "hand_tsl": t(tsl), "persp_rotmat": t(persp),
            "camera_free_transf": t(free), "z_offset": t(zoff)}
    out = _posegen(mano_model, feed, (t(n_tsl), t(n_ang)))
    np.testing.assert_allclose(out["final_obj_pose"].cpu().numpy(), ref["final_obj_pose"], rtol=1e-4, atol=2e-6)
    np.testing.assert_allclose(out["final_hand_verts"].cpu().numpy(), ref["final_hand_verts"], rtol=1e-4, atol=5e-6)
    np.testing.assert_allclose(out["final_joints"].cpu().numpy(), ref["final_joints"], rtol=1e-4, atol=5e-6)
