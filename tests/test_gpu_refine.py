"""GPU parity of the hand-object refiner (REFINER.TYPE hand_obj), the anatomical scramblers (random_2 / random_3) and the
staged pose generator: CUDA through the C-ABI vs oracle/refine.py and vs the fixtures recorded from the reference's own
refiner.py / scrambler.py / preprocessor.py (tests/golden/make_golden_refine.py)."""
import sys

import numpy as np
import pytest
import torch

from conftest import GOLDEN, golden
from oracle import refine as orf

sys.path.insert(0, GOLDEN)
import refine_fixture as fx  # noqa: E402
from test_refine_oracle import resampled_objects, staged_pose_generator_oracle  # noqa: E402

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def t(a, dtype=torch.float32):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV).to(dtype)


def state_t():
    return {k: torch.from_numpy(v) for k, v in fx.refinenet_state().items()}


@pytest.fixture(scope="module")
def refiner(mano_model, lib_built):
    from artiboost_b200.artiboost import HORefiner
    r = HORefiner({"PRETRAINED": None, "ITERS": 3}, mano_model=mano_model, state_dict=state_t())
    np.random.seed(5)
    r.setup(fx.object_meshes())
    return r.to(DEV)


# ------------------------------------------------------------------------------------------------ nearest neighbour
@pytest.mark.parametrize("n_x,n_y", [(778, 10000), (5, 13), (833, 2049), (1, 1), (100, 8)])
def test_chamfer_nn_bit_exact_vs_oracle(lib_built, n_x, n_y):
    from artiboost_b200.artiboost.refiner import chamfer_nn
    rng = np.random.RandomState(n_x + n_y)
    B, n_obj = 3, 2
    x = rng.normal(0, 0.08, (B, n_x, 3)).astype(np.float32)
    pts = rng.normal(0, 0.06, (n_obj, n_y, 3)).astype(np.float32)
    if n_y > 40:
        pts[0, 37] = pts[0, 3]      # exact duplicate: the first index must win
        x[0, 0] = pts[0, 3]
    obj_id = np.array([0, 1, 0])
    _, _, R, _ = fx.refiner_inputs(seed=n_x, B=B)
    R[0] = np.eye(3)
    d, i = chamfer_nn(t(x), t(pts), obj_id=t(obj_id, torch.int32), rot=t(R))
    for b in range(B):
        dr, ir = orf.chamfer_nn(x[b], orf.rotate_cloud(R[b], pts[obj_id[b]]))
        np.testing.assert_array_equal(i[b].cpu().numpy(), ir)
        np.testing.assert_array_equal(d[b].cpu().numpy().view(np.uint32), dr.view(np.uint32))
    if n_y > 40:
        assert int(i[0, 0]) == 3 and float(d[0, 0]) == 0.0
    # per-sample clouds (obj_id None), 4x4 poses as the rotation source, folded BatchNorm, strided output
    y = np.stack([orf.rotate_cloud(R[b], pts[obj_id[b]]) for b in range(B)])
    scale, shift = rng.uniform(0.5, 2, n_x).astype(np.float32), rng.normal(0, 1, n_x).astype(np.float32)
    wide = torch.zeros((B, n_x + 7), device=DEV)
    d2, i2 = chamfer_nn(t(x), t(y), scale=t(scale), shift=t(shift), out=wide[:, 3:3 + n_x])
    np.testing.assert_array_equal(i2.cpu().numpy(), i.cpu().numpy())
    np.testing.assert_array_equal(wide[:, 3:3 + n_x].cpu().numpy(), d.cpu().numpy() * scale + shift)
    assert float(wide[:, :3].abs().sum()) == 0 and float(wide[:, 3 + n_x:].abs().sum()) == 0
    P = np.tile(np.eye(4, dtype=np.float32), (B, 1, 1))
    P[:, :3, :3] = R
    P[:, :3, 3] = 9.0  # the translation column must be ignored (refiner.py:190 uses the rotation only)
    d3, i3 = chamfer_nn(t(x), t(pts), obj_id=t(obj_id, torch.int32), rot=t(P))
    assert torch.equal(d3, d) and torch.equal(i3, i)


def test_chamfer_nn_full_batch_properties(lib_built):
    """BASELINE-size batch (512 x 778 x 10 000): distances equal torch's brute force and the returned index attains them."""
    from artiboost_b200.artiboost.refiner import chamfer_nn
    g = torch.Generator(device=DEV).manual_seed(0)
    B = 512
    x = torch.randn((B, 778, 3), device=DEV, generator=g) * 0.08
    pts = torch.randn((4, 10000, 3), device=DEV, generator=g) * 0.06
    obj_id = torch.randint(4, (B,), device=DEV, generator=g, dtype=torch.int32)
    d, i = chamfer_nn(x, pts, obj_id=obj_id)
    for s in range(0, B, 64):
        y = pts[obj_id[s:s + 64].long()]
        ref = torch.cdist(x[s:s + 64].double(), y.double()).min(-1).values
        assert float((d[s:s + 64].double() - ref).abs().max()) < 1e-7
        nn = torch.gather(y, 1, i[s:s + 64].long()[..., None].expand(-1, -1, 3))
        assert float(((x[s:s + 64] - nn).double().norm(dim=-1) - ref).abs().max()) < 1e-7
    assert int(i.min()) >= 0 and int(i.max()) < 10000


@pytest.mark.parametrize("n_x,n_y,B", [(778, 10000, 48), (5, 40, 3), (13, 33, 2), (100, 1, 2), (900, 15360, 4)])
def test_chamfer_nn_grouped_is_bit_identical_to_the_scan(lib_built, n_x, n_y, B):
    """ab_chamfer_nn_grouped (Morton groups + box bounds, cloud in shared memory, one thread per vertex) against ab_chamfer_nn on surface-like and
    volumetric clouds, with rotations, exact duplicates and vertices lying on cloud points."""
    from artiboost_b200.artiboost.refiner import build_nn_groups, chamfer_nn, chamfer_nn_grouped
    rng = np.random.RandomState(n_x * 7 + n_y)
    n_obj = 3
    pts = rng.normal(0, 0.06, (n_obj, n_y, 3)).astype(np.float32)
    d = pts[1] / np.maximum(np.linalg.norm(pts[1], axis=1, keepdims=True), 1e-6)
    pts[1] = (d * np.array([0.08, 0.05, 0.11])).astype(np.float32)          # an ellipsoid surface, like a resampled mesh
    if n_y > 64:
        pts[0, 50] = pts[0, 7]                                               # duplicates far apart in index ...
        pts[2, n_y - 1] = pts[2, 0]
    x = rng.normal(0, 0.09, (B, n_x, 3)).astype(np.float32)
    obj_id = rng.randint(n_obj, size=B)
    _, _, R, _ = fx.refiner_inputs(seed=n_y, B=B)
    R[0] = np.eye(3)
    obj_id[0] = 0
    x[0, 0] = pts[0, min(7, n_y - 1)]                                          # ... hit exactly: the smaller index wins
    x[0, min(1, n_x - 1)] = pts[0, n_y // 2]
    groups = tuple(t(a, dt) for a, dt in zip(build_nn_groups(pts), (torch.float32, torch.int32, torch.float32)))
    oid = t(obj_id, torch.int32)
    d0, i0 = chamfer_nn(t(x), t(pts), obj_id=oid, rot=t(R))
    d1, i1 = chamfer_nn_grouped(t(x), groups, obj_id=oid, rot=t(R))
    assert torch.equal(i0, i1)
    assert torch.equal(d0.view(torch.int32), d1.view(torch.int32))
    if n_y > 64:
        assert int(i1[0, 0]) == 7 and float(d1[0, 0]) == 0.0
    # folded BatchNorm, strided output, no rotation, 4x4 poses as the rotation source
    scale, shift = t(rng.uniform(0.5, 2, n_x).astype(np.float32)), t(rng.normal(0, 1, n_x).astype(np.float32))
    wide = torch.zeros((B, n_x + 5), device=DEV)
    chamfer_nn_grouped(t(x), groups, obj_id=oid, scale=scale, shift=shift, out=wide[:, 2:2 + n_x], return_idx=False)
    ref, _ = chamfer_nn(t(x), t(pts), obj_id=oid, scale=scale, shift=shift)
    assert torch.equal(wide[:, 2:2 + n_x], ref) and float(wide[:, :2].abs().sum()) == 0
    Pm = np.tile(np.eye(4, dtype=np.float32), (B, 1, 1))
    Pm[:, :3, :3] = R
    d2, i2 = chamfer_nn_grouped(t(x), groups, obj_id=oid, rot=t(Pm))
    assert torch.equal(d2, d1) and torch.equal(i2, i1)


def test_refiner_grouped_and_scanned_search_agree(refiner):
    g = golden("refiner.npz")
    names = [fx.OBJ_NAMES[i] for i in g["obj_id"]]
    inp = {"hand_pose": t(g["pose"]), "hand_tsl": t(g["tsl"]), "obj_rot": t(g["obj_rot"])}
    assert hasattr(refiner, "nn_sorted") and refiner.use_groups
    a = refiner(inp, names)
    try:
        refiner.use_groups = False
        b = refiner(inp, names)
    finally:
        refiner.use_groups = True
    for k in ("hand_verts", "joints", "hand_pose", "hand_tsl"):
        assert torch.equal(a[k], b[k]), k


def test_chamfer_nn_argument_errors(lib_built):
    from artiboost_b200.artiboost.refiner import chamfer_nn, point2point_signed
    from artiboost_b200.lib import AbError
    x = torch.zeros((2, 4, 3), device=DEV)
    with pytest.raises(ValueError):
        point2point_signed(x, torch.zeros((3, 5, 3), device=DEV))   # refiner.py:50-51
    with pytest.raises(AbError):
        chamfer_nn(x, torch.zeros((2, 0, 3), device=DEV))
    with pytest.raises(AbError):
        chamfer_nn(torch.zeros((2, 4, 3)), torch.zeros((2, 5, 3)))   # host tensors: there is no CPU path
    d, i = chamfer_nn(torch.zeros((0, 4, 3), device=DEV), torch.zeros((0, 5, 3), device=DEV))
    assert d.shape == (0, 4)


# ------------------------------------------------------------------------------------------------------ fp32 linear
@pytest.mark.parametrize("M,N,K", [(512, 512, 1389), (37, 99, 512), (1, 256, 877), (130, 3, 5), (64, 64, 16)])
def test_linear_f32_matches_fp64(lib_built, M, N, K):
    from artiboost_b200.artiboost.refiner import linear_f32
    rng = np.random.RandomState(M + N + K)
    xw = rng.normal(0, 1, (M, K + 9)).astype(np.float32)
    W, b = rng.normal(0, 1 / np.sqrt(K), (N, K)).astype(np.float32), rng.normal(0, 1, N).astype(np.float32)
    res = rng.normal(0, 1, (M, N)).astype(np.float32)
    xd = t(xw)[:, 4:4 + K]  # a column slice of a wider buffer, unaligned rows
    ref = xw[:, 4:4 + K].astype(np.float64) @ W.astype(np.float64).T + b
    out = torch.full((M, N + 5), -7.0, device=DEV)
    linear_f32(xd, t(W), t(b), out[:, 2:2 + N])
    np.testing.assert_allclose(out[:, 2:2 + N].cpu().numpy(), ref, rtol=1e-5, atol=1e-5)
    assert float((out[:, :2] + 7).abs().sum()) == 0 and float((out[:, 2 + N:] + 7).abs().sum()) == 0
    r = ref + res
    lr = np.where(r > 0, r, 0.2 * r)
    y = t(res)
    linear_f32(xd, t(W), t(b), y, residual=y, leaky=0.2)  # in-place residual, as the output heads use it
    np.testing.assert_allclose(y.cpu().numpy(), lr, rtol=1e-5, atol=1e-5)


# ------------------------------------------------------------------------------------------------ 6D encode / decode
def test_refine_encode_decode(lib_built):
    from artiboost_b200 import lib
    from artiboost_b200.artiboost.refiner import CRot2rotmat, parms_decode
    from oracle import rotations as rot
    rng = np.random.RandomState(3)
    B = 33
    pose = rng.normal(0, 0.5, (B, 48)).astype(np.float32)
    pose[0, 3:6] = 0
    tsl = rng.normal(0, 0.1, (B, 3)).astype(np.float32)
    feat = torch.zeros((B, 120), device=DEV)
    pose_d, tsl_d = t(pose), t(tsl)  # keep the device copies alive across the asynchronous launch
    rc = lib.load().ab_refine_encode(B, lib.ptr(pose_d), lib.ptr(tsl_d), feat[:, 10:].data_ptr(), 120, None)
    lib.check(rc, "ab_refine_encode")
    torch.cuda.synchronize()
    R = rot.aa_to_rotmat(pose.reshape(B, 16, 3).astype(np.float64))
    np.testing.assert_allclose(feat[:, 10:106].cpu().numpy(), R[..., :2].reshape(B, 96), atol=2e-6)
    np.testing.assert_array_equal(feat[:, 106:109].cpu().numpy(), tsl)
    # decode of a perturbed 6D code vs the oracle's CRot2rotmat + rotmat_to_aa; and vs the torch restatement of CRot2rotmat
    code = (feat[:, 10:106].cpu().numpy() + rng.normal(0, 0.05, (B, 96))).astype(np.float32)
    dec = parms_decode(t(code), t(tsl))
    ref_pose, _ = orf.parms_decode(code.astype(np.float64), tsl)
    got = dec["th_pose_coeffs"].cpu().numpy()
    Rg, Rr = rot.aa_to_rotmat(got.reshape(B, 16, 3).astype(np.float64)), rot.aa_to_rotmat(ref_pose.reshape(B, 16, 3))
    np.testing.assert_allclose(Rg, Rr, atol=5e-6)   # compare rotations: axis-angle is ill-conditioned near 0 / pi
    np.testing.assert_allclose(CRot2rotmat(t(code)).cpu().numpy().reshape(B, 16, 3, 3), Rr, atol=5e-6)
    # round trip of an exact code returns the pose
    dec0 = parms_decode(feat[:, 10:106].contiguous(), t(tsl))["th_pose_coeffs"].cpu().numpy()
    np.testing.assert_allclose(rot.aa_to_rotmat(dec0.reshape(B, 16, 3).astype(np.float64)), R, atol=5e-6)


# ----------------------------------------------------------------------------------------------------- HORefiner
def test_ho_refiner_matches_oracle_and_reference(refiner, mano_model):
    g = golden("refiner.npz")
    pts = refiner.resampled_objs_buffer.cpu().numpy()
    np.testing.assert_array_equal(pts[:, :16], g["pts_head"])     # same resampled clouds as the reference run
    names = [fx.OBJ_NAMES[i] for i in g["obj_id"]]
    out = refiner({"hand_pose": t(g["pose"]), "hand_tsl": t(g["tsl"]), "obj_rot": t(g["obj_rot"])}, names)
    net = orf.RefineNet(fx.refinenet_state(), mano_model, n_iters=3)
    ref = orf.ho_refiner(net, pts, g["obj_id"], g["pose"], g["tsl"], g["obj_rot"])
    for src in (ref, g):  # oracle, then the reference's own output
        # north_star tolerance: 1e-4 relative on vertex positions (hand scale ~0.2 m -> 2e-5 m absolute)
        np.testing.assert_allclose(out["hand_verts"].cpu().numpy(), src["hand_verts"], rtol=1e-4, atol=2e-5)
        np.testing.assert_allclose(out["joints"].cpu().numpy(), src["joints"], rtol=1e-4, atol=2e-5)
        np.testing.assert_allclose(out["hand_tsl"].cpu().numpy(), src["hand_tsl"], rtol=0, atol=1e-5)
        np.testing.assert_allclose(out["hand_pose"].cpu().numpy(), src["hand_pose"], rtol=0, atol=1e-4)
    with pytest.raises(AssertionError):
        refiner({"hand_pose": t(g["pose"]), "hand_tsl": t(g["tsl"]), "obj_rot": t(g["obj_rot"])}, names[:2])


def test_refinenet_forward_reference_signature(refiner, mano_model):
    """_RefineNet.forward(h2o_dist, fpose_rhand_rotmat_f, trans_rhand_f, global_orient_rhand_rotmat_f, verts_object)."""
    from oracle import rotations as rot
    g = golden("refiner.npz")
    pts = refiner.resampled_objs_buffer.cpu().numpy()
    R = rot.aa_to_rotmat(g["pose"].reshape(-1, 16, 3)).astype(np.float32)
    vo = np.stack([orf.rotate_cloud(g["obj_rot"][b], pts[g["obj_id"][b]]) for b in range(len(g["pose"]))])
    got = refiner.refine_net(h2o_dist=t(g["h2o"]), fpose_rhand_rotmat_f=t(R[:, 1:]), trans_rhand_f=t(g["tsl"]),
                             global_orient_rhand_rotmat_f=t(R[:, 0]), verts_object=t(vo))
    np.testing.assert_allclose(got["th_tsl"].cpu().numpy(), g["hand_tsl"], rtol=0, atol=1e-5)
    np.testing.assert_allclose(got["th_pose_coeffs"].cpu().numpy(), g["hand_pose"], rtol=0, atol=1e-4)


def test_refinenet_state_dict_names_match_reference(refiner):
    """A GrabNet refinenet.pt must load: same parameter / buffer names as refiner.py:227-319."""
    ours = {k for k in refiner.refine_net.state_dict() if not k.startswith("mano_layer.")}
    ref = set(fx.refinenet_state()) | {k.replace("running_mean", "num_batches_tracked") for k in fx.refinenet_state()
                                       if k.endswith("running_mean")}
    assert ours == ref


# ------------------------------------------------------------------------------------------- anatomical scramblers
def test_anatomical_scramblers_match_oracle_and_reference(lib_built):
    from artiboost_b200.artiboost import AxisLayer, RandomScrambler2, RandomScrambler3
    g = golden("scrambler23.npz")
    nz = fx.scrambler_noise()
    cfg = {"HAND_TSL_SIGMA": 0.01, "HAND_POSE_SIGMA": 0.1}
    feed = {"hand_pose": t(g["pose"]), "hand_tsl": t(g["tsl"]), "joints": t(g["joints"]), "hand_verts": None,
            "hand_transf": t(g["transf"])}
    b, u, l = AxisLayer()(feed["joints"], feed["hand_transf"])
    for got, ref in zip((b, u, l), orf.axis_layer(g["joints"], g["transf"])):
        np.testing.assert_allclose(got.cpu().numpy(), ref, atol=2e-6)
    s2 = RandomScrambler2(cfg)
    r2 = s2(feed, noise=(t(nz["tsl"]), t(nz["splay"]), s2.expand_bend(t(nz["bend5"])), t(nz["thumb"])))
    s3 = RandomScrambler3(cfg)
    r3 = s3(feed, noise=(t(nz["tsl"]), t(nz["splay"]), t(nz["bend14"]), t(nz["thumb"])))
    o2 = orf.random_scrambler_2(g["pose"], g["joints"], g["transf"], nz["splay"], nz["bend5"], nz["thumb"])
    o3 = orf.random_scrambler_3(g["pose"], g["joints"], g["transf"], nz["splay"], nz["bend14"], nz["thumb"])
    for got, orc, ref in ((r2, o2, g["pose2"]), (r3, o3, g["pose3"])):
        np.testing.assert_allclose(got["hand_pose"].cpu().numpy(), orc, rtol=0, atol=1e-5)
        np.testing.assert_allclose(got["hand_pose"].cpu().numpy(), ref, rtol=0, atol=1e-5)
        np.testing.assert_allclose(got["hand_tsl"].cpu().numpy(), g["tsl2"], rtol=0, atol=1e-7)
    # own draws: shapes, determinism under a seeded generator, root rotation untouched
    gen = torch.Generator(device=DEV).manual_seed(3)
    a = s3(feed, generator=gen)["hand_pose"]
    gen.manual_seed(3)
    assert torch.equal(a, s3(feed, generator=gen)["hand_pose"]) and torch.equal(a[:, :3], feed["hand_pose"][:, :3])


# ------------------------------------------------------------------------------------------ staged pose generator
def test_staged_pose_generator_matches_oracle_and_reference(mano_model, lib_built):
    from artiboost_b200.artiboost import HORefiner, PreProcessorPoseGenerator, RandomScrambler2
    g = golden("preprocessor_staged.npz")
    nz = fx.scrambler_noise(seed=23, B=4)
    r = HORefiner({"PRETRAINED": None, "ITERS": 2}, mano_model=mano_model, state_dict=state_t())
    np.random.seed(5)
    r.setup(fx.object_meshes())
    r = r.to(DEV)
    scr = RandomScrambler2({"HAND_TSL_SIGMA": 0.01, "HAND_POSE_SIGMA": 0.1})
    gen = PreProcessorPoseGenerator(r, scr, r.refine_net.mano_layer, r.refine_net.mano_layer).to(DEV)
    feed = {"index": None, "obj_id": t(g["obj_id"], torch.int32), "obj_name": [fx.OBJ_NAMES[i] for i in g["obj_id"]],
            "hand_pose": t(g["pose"]), "hand_shape": t(g["shape"]), "hand_tsl": t(g["tsl"]), "persp_rotmat": t(g["persp"]),
            "camera_free_transf": t(g["free"]), "z_offset": t(g["zoff"])}
    noise = (t(nz["tsl"]), t(nz["splay"]), scr.expand_bend(t(nz["bend5"])), t(nz["thumb"]))
    out = gen._forward_staged(feed, feed["hand_pose"], feed["hand_shape"], feed["hand_tsl"], feed["persp_rotmat"],
                              feed["camera_free_transf"], feed["z_offset"], True, noise=noise)
    ref = staged_pose_generator_oracle(mano_model, g, nz, r.resampled_objs_buffer.cpu().numpy())
    for src in ({"obj_pose": ref["final_obj_pose"], "verts": ref["final_hand_verts"], "joints": ref["final_joints"]}, g):
        np.testing.assert_allclose(out["final_obj_pose"].cpu().numpy(), src["obj_pose"], rtol=1e-4, atol=2e-6)
        np.testing.assert_allclose(out["final_hand_verts"].cpu().numpy(), src["verts"], rtol=1e-4, atol=3e-5)
        np.testing.assert_allclose(out["final_joints"].cpu().numpy(), src["joints"], rtol=1e-4, atol=3e-5)
    # obj_name None: the batched pipeline's obj_id path gives the same result
    feed2 = dict(feed, obj_name=None)
    out2 = gen._forward_staged(feed2, feed["hand_pose"], feed["hand_shape"], feed["hand_tsl"], feed["persp_rotmat"],
                               feed["camera_free_transf"], feed["z_offset"], True, noise=noise)
    assert torch.equal(out2["final_hand_verts"], out["final_hand_verts"])


@pytest.mark.parametrize("scr_type,ref_type", [("random_3", "null"), ("random", "hand_obj"), ("random_2", "hand_obj")])
def test_synth_pipeline_with_refiner_and_scramblers(lib_built, scr_type, ref_type):
    """The whole synthesis pipeline (CCV draw -> pose generator -> rasterise) with the non-default scrambler / refiner."""
    from artiboost_b200.synth import DEFAULT_CFG, SynthPipeline
    cfg = dict(DEFAULT_CFG, SCRAMBLER=dict(DEFAULT_CFG["SCRAMBLER"], TYPE=scr_type),
               REFINER={"TYPE": ref_type, "PRETRAINED": None, "ITERS": 3})
    pipe = SynthPipeline(device=DEV, seed=2, cfg=cfg, n_hand_tex=2, n_bg=2, chunk=8)
    poses = pipe.sample_poses(24)
    v = poses["final_hand_verts"]
    assert v.shape == (24, 778, 3) and bool(torch.isfinite(v).all())
    assert 0.2 < float(v[..., 2].mean()) < 0.9          # in front of the camera, near z_offset
    views = pipe.render(poses)
    seg = views["seg"]
    assert bool((seg == 1).any()) and bool((seg == 2).any())
