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


# ------------------------------------------------------------------------------------- fused draw / blacklist
def _small_pipe(**kw):
    from artiboost_b200.synth import SynthPipeline
    return SynthPipeline(device=DEV, seed=3, n_hand_tex=7, n_bg=3, **kw)


def test_fused_draw_matches_oracle_composition(lib_built):
    """ab_synth_draw (one launch: CCV cell, view, grasp, scrambler noise, renderer draws) vs oracle.ccv.synth_draw on the
    uniforms the kernel reports: ids / texture / crop integer-exact, views 1e-6, noise and light 1e-5."""
    from oracle import ccv
    pipe = _small_pipe()
    rng = np.random.RandomState(1)
    w = rng.uniform(0.1, 10, size=tuple(pipe.sample_weight_map.shape)).astype(np.float32)
    w[1, 5, :] = 0.0
    pipe.sample_weight_map = t(w)
    n = 700
    for uniforms in (None, t(rng.rand(n, 32).astype(np.float32))):
        occ0 = pipe.occurence_map.clone()
        b = pipe.draw(n, uniforms=uniforms, return_uniforms=True)
        u = b["uniforms"].cpu().numpy()
        if uniforms is not None:
            np.testing.assert_array_equal(u, uniforms.cpu().numpy())
        assert u.min() >= 0.0 and u.max() < 1.0
        r = pipe.renderer
        nb, bh, bw = r.backgrounds.shape[:3]
        ref = ccv.synth_draw(w, u, 12, 24, (0.45, 0.55), pipe.grasp_engine.table.cpu().numpy(), 0.01, 0.1, r.n_hand_tex,
                             (1.0, 5.0), nb, (bh, bw), (r.width, r.height))
        for k in ("obj_id", "persp_id", "grasp_id"):
            np.testing.assert_array_equal(b[k].cpu().numpy(), ref[k], err_msg=k)
        for k in ("hand_pose", "hand_shape", "hand_tsl", "z_offset"):
            np.testing.assert_array_equal(b[k].cpu().numpy(), ref[k], err_msg=k)
        np.testing.assert_allclose(b["persp_rotmat"].cpu().numpy(), ref["persp_rotmat"], atol=3e-6)  # fp32 sincos vs numpy: a few ulp
        np.testing.assert_allclose(b["camera_free_transf"].cpu().numpy(), ref["camera_free_transf"], atol=3e-6)
        n_tsl, n_ang = b["noise"]
        np.testing.assert_allclose(n_tsl.cpu().numpy(), ref["noise_tsl"], atol=1e-6, rtol=1e-4)
        np.testing.assert_allclose(n_ang.cpu().numpy(), ref["noise_angle"], atol=1e-5, rtol=1e-4)
        rr = pipe._render_rand[1]
        np.testing.assert_array_equal(rr["hand_tex"].cpu().numpy(), ref["hand_tex"])
        np.testing.assert_allclose(rr["light"].cpu().numpy(), ref["light"], rtol=1e-6)
        np.testing.assert_array_equal(rr["bg_sel"].cpu().numpy(), ref["bg_sel"])
        # occurrence counts are folded into the map when it is read (ovg_set.py:172-178)
        occ = ccv.occurrence_count_map(ref["obj_id"], ref["persp_id"], ref["grasp_id"], *w.shape) > 0
        np.testing.assert_array_equal(pipe.occurence_map.cpu().numpy(), occ0.cpu().numpy() | occ)
        assert not pipe.occurence_map[1, 5].any()


def test_fused_draw_stream_is_reproducible_and_well_distributed(lib_built):
    a, b = _small_pipe(), _small_pipe()
    ua = a.draw(4096, return_uniforms=True)["uniforms"]
    ub = b.draw(4096, return_uniforms=True)["uniforms"]
    assert torch.equal(ua, ub)                                  # same (seed, offset) -> same batch
    ua2 = a.draw(4096, return_uniforms=True)["uniforms"]
    assert not torch.equal(ua, ua2)                             # the offset advances between calls
    u = torch.cat([ua, ua2]).double()
    assert abs(float(u.mean()) - 0.5) < 5e-3 and abs(float(u.var()) - 1 / 12) < 2e-3
    c = torch.corrcoef(u[:, :8].T)
    assert float((c - torch.eye(8, device=c.device)).abs().max()) < 0.05
    n_tsl, n_ang = a.draw(8192)["noise"]
    assert abs(float(n_ang.mean())) < 5e-3 and abs(float(n_ang.std()) - 0.1) < 3e-3 and abs(float(n_tsl.std()) - 0.01) < 5e-4
    # the synthesis path with the fused draw agrees with the staged path given the same inputs
    pipe = _small_pipe()
    batch = pipe.draw(64)
    fused = pipe.pose_generator(batch)
    staged = pipe.pose_generator._forward_staged(batch, batch["hand_pose"], batch["hand_shape"], batch["hand_tsl"],
                                                 batch["persp_rotmat"], batch["camera_free_transf"], batch["z_offset"], False,
                                                 noise=batch["noise"])
    for k in ("final_obj_pose", "final_hand_verts", "final_joints"):
        torch.testing.assert_close(fused[k], staged[k], rtol=1e-4, atol=2e-6)


def test_prefetching_synthesis_equals_the_plain_sequence(lib_built):
    """synthesise(prefetch=True) issues the next batch's draw + pose generator on a side stream ahead of this batch's
    rasteriser; the sequence of views must be the plain sequence's, bit for bit (same Philox offsets in call order), and a
    change of batch size or drop_prefetch() must fall back to a fresh draw."""
    a, b = _small_pipe(chunk=64), _small_pipe(chunk=64)
    keys = ("rgba", "depth", "seg", "obj_id", "persp_id", "grasp_id", "obj_pose", "hand_verts", "joints")
    for i in range(4):
        va = a.synthesise(48)
        vb = b.synthesise(48, prefetch=True)
        torch.cuda.synchronize()
        for k in keys:
            assert torch.equal(va[k], vb[k]), (i, k)
    assert b._ahead is not None and b._ahead[0] == 48
    # a is one draw behind b now (b drew batch 4 ahead); let it catch up, then both continue with another size
    a.sample_poses(48)
    a._render_rand = None
    va, vb = a.synthesise(16), b.synthesise(16, prefetch=True)   # b discards nothing it needs: size mismatch -> fresh draw
    for k in keys:
        assert torch.equal(va[k], vb[k]), k
    b.drop_prefetch()
    assert b._ahead is None
    a.sample_poses(16)   # the batch b drew ahead and dropped
    torch.cuda.synchronize()
    assert torch.equal(a.occurence_map, b.occurence_map)


def test_blacklist_kernel_matches_oracle_and_reference_fixture(lib_built):
    """ab_ccv_blacklist vs oracle.ccv.blacklist_map on the pipeline's CCV space, and vs the map the reference's own
    _construct_blacklist_map produced (tests/golden/blacklist.npz); blacklisted cells are never drawn."""
    from artiboost_b200 import assets
    from artiboost_b200.synth import DEFAULT_CFG, SynthPipeline
    from oracle import ccv
    g = golden("blacklist.npz")
    names = [str(x) for x in g["names"]]
    cfg = dict(DEFAULT_CFG, VIEW={"PERSP_U_BINS": int(g["u_bins"]), "PERSP_THETA_BINS": int(g["theta_bins"]),
                                  "CAMERA_Z_RANGE": [0.45, 0.55]}, GRASP_NUM=int(g["n_grasp"]))
    pipe = SynthPipeline(obj_names=names, device=DEV, seed=0, cfg=cfg, n_hand_tex=2, n_bg=2)
    np.testing.assert_array_equal(pipe.grasp_engine.table[:, :, :48].cpu().numpy(), g["hand_pose"])
    bl, th = pipe.construct_blacklist_map(rand2=t(g["rand2"]), return_th=True)
    np.testing.assert_array_equal(bl.cpu().numpy(), g["blacklist"])          # the reference function's output
    # the pipeline's own (default) space against the oracle, away from the threshold
    pipe = _small_pipe()
    rng = np.random.RandomState(2)
    rand2 = rng.rand(*pipe.sample_weight_map.shape, 2).astype(np.float32)
    bl, th = pipe.construct_blacklist_map(rand2=t(rand2), return_th=True)
    ref, ref_th = ccv.blacklist_map(pipe.grasp_engine.table[:, :, :3].cpu().numpy(), 12, 24, rand2, return_th=True)
    np.testing.assert_allclose(th.cpu().numpy(), ref_th, atol=2e-6)
    clear = np.abs(ref_th + 0.8) > 1e-5
    np.testing.assert_array_equal(bl.cpu().numpy()[clear], ref[clear])
    assert 0.02 < float(bl.float().mean()) < 0.3
    # applied at construction (artiboost_loader.py:125-130): zero weight, never drawn
    assert bool((pipe.sample_weight_map[pipe.blacklist_map] == 0).all()) and bool(pipe.blacklist_map.any())
    b = pipe.draw(20000)
    assert not bool(pipe.blacklist_map[b["obj_id"].long(), b["persp_id"].long(), b["grasp_id"].long()].any())
