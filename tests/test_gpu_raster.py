"""GPU parity of the batched rasteriser (ab_render_batch) vs oracle/raster.c: segmentation, coverage and depth
BIT-exact, colour exact (the rule set fixes every rounding), plus size-independent properties at batch 512."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def pipe(lib_built):
    from artiboost_b200.synth import SynthPipeline
    return SynthPipeline(device=DEV, seed=0, n_hand_tex=5, n_bg=3, chunk=16)


def oracle_views(pipe, poses, rand, idx, cull=1, use_bg=True):
    from oracle import raster
    r = pipe.renderer
    cfg = dict(width=r.width, height=r.height, fx=float(pipe.cam_intr[0, 0]), fy=float(pipe.cam_intr[1, 1]),
               cx=float(pipe.cam_intr[0, 2]), cy=float(pipe.cam_intr[1, 2]), znear=0.05, cull_backface=cull,
               ambient=0.8, diffuse=0.25)
    hf = r.hand_faces.cpu().numpy()[:, :3]
    hcols = r.hand_colors.cpu().numpy()
    bgs = r.backgrounds.cpu().numpy()
    out = []
    for i in idx:
        oid = int(poses["obj_id"][i])
        kw = {}
        if oid >= 0:
            name = pipe.obj_names[oid]
            o = pipe.objects[name]
            kw = dict(obj_verts=o["vertices"], obj_faces=o["faces"],
                      obj_cols=r.obj_colors.cpu().numpy()[r._voff[oid]:r._voff[oid + 1]],
                      obj_pose=poses["final_obj_pose"][i].cpu().numpy())
        sel = rand["bg_sel"][i].cpu().numpy() if use_bg else None
        if sel is not None and sel[0] >= 0:
            kw.update(bg=bgs[sel[0]], bg_sel=sel[1:])
        out.append(raster.render_view(cfg, poses["final_hand_verts"][i].cpu().numpy(), hf,
                                      hcols[int(rand["hand_tex"][i])], light=float(rand["light"][i]), **kw))
    return out


def check(views, ref, idx):
    for j, i in enumerate(idx):
        rgba, depth, seg, _ = ref[j]
        np.testing.assert_array_equal(views["seg"][i].cpu().numpy(), seg, err_msg=f"seg view {i}")
        np.testing.assert_array_equal(views["depth"][i].cpu().numpy().view(np.uint32), depth.view(np.uint32),
                                      err_msg=f"depth bits view {i}")
        np.testing.assert_array_equal(views["rgba"][i].cpu().numpy(), rgba, err_msg=f"rgba view {i}")


def test_render_matches_oracle_on_sampled_views(pipe):
    torch.manual_seed(0)
    B = 40  # not a multiple of the chunk (16): exercises the ragged last chunk
    poses = pipe.sample_poses(B)
    rand = pipe.draw_render_randoms(B)
    views = pipe.render(poses, rand)
    idx = list(range(B))
    ref = oracle_views(pipe, poses, rand, idx)
    check(views, ref, idx)
    seg = views["seg"].cpu().numpy()
    assert (seg == 1).sum() > 500 * B * 0.3 and (seg == 2).sum() > 500 * B * 0.3  # hands and objects are in frame


def test_render_hand_only_no_background_no_cull(pipe):
    B = 9
    poses = pipe.sample_poses(B)
    poses["obj_id"] = poses["obj_id"].clone()
    poses["obj_id"][::2] = -1  # CONST.DUMMY: hand only (renderer.py:107)
    rand = pipe.draw_render_randoms(B)
    rand["bg_sel"][:, 0] = -1  # flat background
    cam = pipe.renderer.camera
    cam.cull_backface = 0
    try:
        views = pipe.render(poses, rand)
    finally:
        cam.cull_backface = 1
    ref = oracle_views(pipe, poses, rand, range(B), cull=0)
    check(views, ref, range(B))
    assert not (views["seg"][0] == 2).any()
    assert (views["rgba"][0][views["seg"][0] == 0][:, :3] == 128).all()


def test_render_is_idempotent_and_order_independent_at_batch_512(pipe):
    B = 512
    poses = pipe.sample_poses(B)
    rand = pipe.draw_render_randoms(B)
    a = {k: v.clone() for k, v in pipe.render(poses, rand).items()}
    b = pipe.render(poses, rand)  # same workspace: the key buffer must have been restored to empty
    for k in ("rgba", "depth", "seg"):
        assert torch.equal(a[k], b[k]), k
    perm = torch.randperm(B, device=DEV)
    poses_p = {k: (v[perm] if torch.is_tensor(v) else v) for k, v in poses.items()}
    rand_p = {k: v[perm] for k, v in rand.items()}
    c = pipe.render(poses_p, rand_p)
    for k in ("rgba", "depth", "seg"):
        assert torch.equal(a[k][perm], c[k]), k
    # composite == per-pixel nearest of (hand only, object only)
    hand_only = dict(poses, obj_id=torch.full_like(poses["obj_id"], -1))
    h = {k: v.clone() for k, v in pipe.render(hand_only, rand).items()}
    dh, da = h["depth"], a["depth"]
    hand_px = a["seg"] == 1
    assert torch.equal(da[hand_px], dh[hand_px])
    obj_px = a["seg"] == 2
    assert bool(((dh[obj_px] == 0) | (dh[obj_px] >= da[obj_px])).all())
    assert bool(((a["seg"] == 0) == (da == 0)).all())
    # a spot check against the oracle at full batch
    idx = [0, 255, 511]
    check(a, oracle_views(pipe, poses, rand, idx), idx)


def test_view_groups_do_not_change_the_result(pipe):
    """A batch larger than the workspace's `chunk` is rendered as consecutive groups of views: any group size, a ragged
    last group and back-to-back calls on the shared workspace must give the same bits."""
    B = 300  # 18 groups of 16 views + one of 12 at the module's chunk size
    poses = pipe.sample_poses(B)
    rand = pipe.draw_render_randoms(B)
    ref = None
    try:
        for n in (16, 300, 7, 512, 64):
            pipe.renderer.set_chunk(n)
            out = {k: v.clone() for k, v in pipe.render(poses, rand).items()}
            out2 = pipe.render(poses, rand)   # immediately again on the same workspace
            if ref is None:
                ref = out
            for k in ("rgba", "depth", "seg"):
                assert torch.equal(ref[k], out[k]) and torch.equal(ref[k], out2[k]), (n, k)
    finally:
        pipe.renderer.set_chunk(16)
    idx = [0, 15, 16, 299]
    check(ref, oracle_views(pipe, poses, rand, idx), idx)


def test_bench_configuration_matches_oracle_on_64_views(lib_built):
    """The exact configuration bench.py times (SynthPipeline(seed=1): 51 hand textures, 8 backgrounds, batch 512 in one
    group): 64 views spread over the batch, bit for bit against oracle/raster.c."""
    from artiboost_b200.synth import SynthPipeline
    bp = SynthPipeline(device=DEV, seed=1)
    B = 512
    poses = bp.sample_poses(B)
    rand = bp.draw_render_randoms(B)
    views = bp.render(poses, rand)
    idx = list(range(0, B, 8))
    assert len(idx) == 64
    check(views, oracle_views(bp, poses, rand, idx), idx)


def _pipe_with_camera(size, f, c, **kw):
    from artiboost_b200.synth import DEFAULT_CFG, SynthPipeline
    cfg = dict(DEFAULT_CFG, RENDER_SIZE=list(size), CAM_PARAM={"FX": f[0], "FY": f[1], "CX": c[0], "CY": c[1]})
    return SynthPipeline(device=DEV, seed=4, cfg=cfg, n_hand_tex=3, n_bg=2, **kw)


@pytest.mark.parametrize("size,f,c", [((250, 190), (217.5, 230.0), (125.0, 95.0)),     # ragged tiles, scalar output path
                                      ((512, 512), (435.0, 435.0), (256.0, 256.0)),    # the reference's render size (yaml:52-61)
                                      ((100, 60), (90.0, 90.0), (50.0, 30.0))])        # smaller than one tile row
def test_other_image_sizes_match_oracle(lib_built, size, f, c):
    p = _pipe_with_camera(size, f, c)
    B = 6
    poses = p.sample_poses(B)
    rand = p.draw_render_randoms(B)
    views = p.render(poses, rand)
    assert views["seg"].shape == (B, size[1], size[0])
    check(views, oracle_views(p, poses, rand, range(B)), range(B))
    assert int((views["seg"] > 0).sum()) > 100


def test_close_up_views_match_oracle(lib_built):
    """Geometry right in front of the camera: triangles tens of pixels across (the warp-cooperative path), extents beyond
    64 px (the int64 edge functions), patches that straddle the near plane and every tile of the frame."""
    p = _pipe_with_camera((256, 256), (217.5, 217.5), (128.0, 128.0))
    B = 8
    poses = p.sample_poses(B)
    rand = p.draw_render_randoms(B)
    z = torch.tensor([0.40, 0.38, 0.36, 0.34, 0.32, 0.30, 0.42, 0.41], device=DEV)
    poses["final_obj_pose"] = poses["final_obj_pose"].clone()
    poses["final_obj_pose"][:, 2, 3] -= z
    poses["final_hand_verts"] = poses["final_hand_verts"] - torch.stack([torch.zeros_like(z), torch.zeros_like(z), z], 1)[:, None]
    for cull in (1, 0):
        p.renderer.camera.cull_backface = cull
        views = p.render(poses, rand)
        check(views, oracle_views(p, poses, rand, range(B), cull=cull), range(B))
    assert float((views["seg"] > 0).float().mean()) > 0.2


def test_render_edge_cases(pipe):
    from artiboost_b200.lib import AbError
    r = pipe.renderer
    empty = r.render_batch(torch.zeros(0, dtype=torch.int32, device=DEV), torch.zeros((0, 4, 4), device=DEV),
                           torch.zeros((0, 778, 3), device=DEV))
    assert empty["rgba"].shape == (0, r.height, r.width, 4)
    with pytest.raises(ValueError):
        r.render_batch(torch.zeros(2, dtype=torch.int32, device=DEV), torch.zeros((2, 4, 4), device=DEV),
                       torch.zeros((2, 700, 3), device=DEV))
    with pytest.raises(AbError):
        r.render_batch(torch.zeros(2, dtype=torch.int32), torch.zeros((2, 4, 4)), torch.zeros((2, 778, 3)))
    # geometry behind the camera / closer than znear produces an all-background frame, not garbage
    poses = pipe.sample_poses(2)
    poses["final_hand_verts"] = poses["final_hand_verts"] - torch.tensor([0, 0, 5.0], device=DEV)
    poses["final_obj_pose"] = poses["final_obj_pose"].clone()
    poses["final_obj_pose"][:, 2, 3] -= 5.0
    v = pipe.render(poses)
    assert int(v["seg"].sum()) == 0 and float(v["depth"].abs().sum()) == 0.0


def test_single_view_call_mirrors_reference_signature(pipe):
    """Renderer.__call__(obj_name, obj_pose, hand_verts) -> uint8[H,W,3] BGR (renderer.py:101-123)."""
    poses = pipe.sample_poses(1)
    name = pipe.obj_names[int(poses["obj_id"][0])]
    img = pipe.renderer(name, poses["final_obj_pose"][0].cpu().numpy(), poses["final_hand_verts"][0].cpu().numpy())
    assert img.shape == (256, 256, 3) and img.dtype == np.uint8
    hand = pipe.renderer("dummy", None, poses["final_hand_verts"][0].cpu().numpy())
    assert hand.shape == (256, 256, 3)
    with pytest.raises(KeyError):
        pipe.renderer("no_such_object", np.eye(4), poses["final_hand_verts"][0].cpu().numpy())


def test_render_provider_queue_protocol(pipe):
    """RendererProvider request/reply protocol (render_infra.py:46-58, rendered_dataset.py:118-123)."""
    from artiboost_b200.artiboost import RendererProvider
    from artiboost_b200.artiboost.renderer import PYRENDER_EXTRINSIC
    prov = RendererProvider(num_workers=2, gpu_render_id=[0], render_size=[256, 256], cam_intr=pipe.cam_intr,
                            cam_extr=PYRENDER_EXTRINSIC, obj_meshes=pipe.obj_engine.obj_trimeshes_mapping,
                            hand_meshes=pipe.hand_meshes[:2], bgs=pipe.backgrounds, lights=None, random_seed=1)
    prov.begin()
    try:
        poses = pipe.sample_poses(6)
        q = prov.get_message_queue()
        for i in range(6):
            q.put({"id": i % 2, "objname": pipe.obj_names[int(poses["obj_id"][i])],
                   "pose": poses["final_obj_pose"][i].cpu().numpy(), "hand_verts": poses["final_hand_verts"][i].cpu().numpy()})
        got = [prov.get_image_queue_list()[i % 2].get(timeout=60) for i in range(6)]
        assert all(g.shape == (256, 256, 3) and g.dtype == np.uint8 for g in got)
    finally:
        prov.end()


def test_many_large_triangles_in_one_tile_match_oracle(lib_built):
    """Triangles of 64 px and more are queued by the patch loop and drawn after it over 8 x 8 blocks.  A disc fan of slivers
    (each ~80 px long, all meeting in one tile) in front of the object: 300 of them fit the tile's queue (512 entries), 700
    overflow it and the tile searches its patch list again.  Both must equal the oracle bit for bit, with and without
    back-face culling."""
    from artiboost_b200 import assets
    names = list(assets.HO3D_TRAIN_OBJS)[:2]
    objs = assets.make_synthetic_objects(names, 4)
    for name, n_fan in zip(names, (300, 700)):
        o = objs[name]
        v, f, c = np.asarray(o["vertices"]), np.asarray(o["faces"]), np.asarray(o["colors"])
        ang = np.linspace(0.0, 2 * np.pi, n_fan, endpoint=False)
        zf = v[:, 2].min() - 0.02
        rim = np.stack([0.16 * np.cos(ang), 0.16 * np.sin(ang), np.full(n_fan, zf)], 1)
        fan_v = np.concatenate([[[0.0, 0.0, zf]], rim])
        base = len(v)
        i = np.arange(n_fan)
        fan_f = np.stack([np.full(n_fan, base), base + 1 + i, base + 1 + (i + 1) % n_fan], 1)
        rng = np.random.RandomState(n_fan)
        objs[name] = dict(o, vertices=np.concatenate([v, fan_v]), faces=np.concatenate([f, fan_f]),
                          colors=np.concatenate([c, rng.randint(40, 255, size=(n_fan + 1, 3)).astype(np.uint8)]))
    p = _pipe_with_camera((256, 256), (217.5, 217.5), (128.0, 128.0), obj_names=names, objects=objs)
    B = 4
    poses = p.sample_poses(B)
    rand = p.draw_render_randoms(B)
    poses["obj_id"] = torch.tensor([0, 1, 0, 1], dtype=torch.int32, device=DEV)
    pose = torch.eye(4, device=DEV).repeat(B, 1, 1)
    pose[:, 0, 3] = torch.tensor([0.0, 0.01, -0.05, 0.06], device=DEV)   # fan centre inside a tile, on a tile corner, off-centre
    pose[:, 1, 3] = torch.tensor([0.0, -0.01, 0.04, 0.0], device=DEV)
    pose[:, 2, 3] = torch.tensor([0.42, 0.40, 0.36, 0.45], device=DEV)
    poses["final_obj_pose"] = pose
    for cull in (0, 1):
        p.renderer.camera.cull_backface = cull
        views = p.render(poses, rand)
        check(views, oracle_views(p, poses, rand, range(B), cull=cull), range(B))
    assert int((views["seg"] == 2).sum()) > 20000
