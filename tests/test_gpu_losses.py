"""Device-resident losses and lookups against the fixtures recorded from the reference (parity gaps):
Criterion / SymCornerLoss on CUDA tensors -- eager and replayed from a captured CUDA graph, the way TrainStep runs them --
vs tests/golden/losses.npz / symloss.npz (anakin/criterions/{criterion,jointloss,ordinal,symcornerloss}.py), and
GraspEngine.gather vs get_obj_grasp (anakin/artiboost/grasp_engine.py:47-53)."""
import json

import numpy as np
import pytest
import torch

from conftest import golden

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _losses_inputs():
    g = golden("losses.npz")
    t = lambda a: torch.from_numpy(a).to(DEV)  # noqa: E731
    preds = {"joints_3d_abs": t(g["pred_joints_3d_abs"]), "corners_3d_abs": t(g["pred_corners_3d_abs"])}
    targs = {k: t(g["targ_" + k]) for k in ("joints_3d", "corners_3d", "root_joint", "joints_vis", "corners_vis")}
    return g, preds, targs


def test_criterion_on_device_matches_reference_losses_eager_and_captured(monkeypatch):
    from artiboost_b200 import criterions as C
    g, preds, targs = _losses_inputs()
    vv = {20: torch.from_numpy(g["vv20"]).to(DEV), 40: torch.from_numpy(g["vv40"]).to(DEV)}
    monkeypatch.setattr(C, "sample_view_vectors", lambda n, device, generator=None: vv[n])
    monkeypatch.setattr(C, "_subsample", lambda n, device, generator: torch.arange(n // 3, device=device))
    crit = C.Criterion(C.DEFAULT_CRITERION_CFG)
    names = ("joints_3d_loss", "corners_3d_loss", "joint_ord_loss", "part_ord_loss", "scene_ord_loss")

    def check(total, parts):
        np.testing.assert_allclose(total.detach().cpu().numpy(), g["total"].reshape(()), rtol=1e-5)
        for k in names:
            np.testing.assert_allclose(parts[k].detach().cpu().numpy(), g["part_" + k], rtol=1e-5, atol=1e-9, err_msg=k)

    total, parts = crit.compute_losses(preds, targs)
    assert total.is_cuda
    check(total, parts)
    # the same arithmetic inside a captured graph (no host syncs, index tables cached on the device), with gradients
    p_static = {k: v.clone().requires_grad_(True) for k, v in preds.items()}
    side = torch.cuda.Stream(DEV)
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            t_, _ = crit.compute_losses(p_static, targs)
            t_.backward()
    torch.cuda.current_stream().wait_stream(side)
    for v in p_static.values():
        v.grad = None
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        total_g, parts_g = crit.compute_losses(p_static, targs)
        total_g.backward()
    eager_grads = {}
    pe = {k: v.clone().requires_grad_(True) for k, v in preds.items()}
    te, _ = crit.compute_losses(pe, targs)
    te.backward()
    eager_grads = {k: v.grad.clone() for k, v in pe.items()}
    for v in p_static.values():
        v.grad.zero_()
    graph.replay()
    torch.cuda.synchronize()
    check(total_g, parts_g)
    for k, v in p_static.items():
        torch.testing.assert_close(v.grad, eager_grads[k], rtol=1e-5, atol=1e-8)


def test_sym_corner_loss_on_device_matches_reference():
    from artiboost_b200 import criterions as C
    g = golden("symloss.npz")
    info = json.loads(str(g["model_info"]))
    t = lambda a: torch.from_numpy(a).to(DEV)  # noqa: E731
    targs = {"obj_idx": t(g["obj_idx"]), "corners_can": t(g["corners_can"]), "obj_transf": t(g["obj_transf"]), "corners_vis": t(g["corners_vis"])}
    for flag in (0, 1):
        loss = C.SymCornerLoss(LAMBDA_SYM_CORNERS_3D=1.0, MODEL_INFO=info, MAX_SYM_DISC_STEP=0.05, USE_HO3D_YCB=bool(flag))
        pred = t(g["pred"]).clone().requires_grad_(True)
        final, parts = loss({"corners_3d_abs": pred}, targs)
        assert final.is_cuda
        np.testing.assert_allclose(parts["sym_corners_3d_loss"].detach().cpu().numpy(), g[f"loss_ho3d{flag}"], rtol=1e-5)
        final.backward()
        assert torch.isfinite(pred.grad).all() and float(pred.grad.abs().sum()) > 0
        # captured
        ps = t(g["pred"]).clone()
        side = torch.cuda.Stream(DEV)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            loss({"corners_3d_abs": ps}, targs)
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            _, parts_g = loss({"corners_3d_abs": ps}, targs)
        graph.replay()
        torch.cuda.synchronize()
        np.testing.assert_allclose(parts_g["sym_corners_3d_loss"].cpu().numpy(), g[f"loss_ho3d{flag}"], rtol=1e-5)


def test_grasp_engine_gather_equals_get_obj_grasp(objects):
    """GraspEngine.gather (one device gather over the packed table) vs the per-sample lookup of the reference interface,
    including the None / 0 placeholders for shape and translation (grasp_engine.py:47-53)."""
    from artiboost_b200 import assets
    from artiboost_b200.artiboost import GraspEngine
    names = list(objects)
    grasps = assets.make_synthetic_grasps(objects, 50, 0)
    # the reference's tables carry placeholders: None shape, scalar 0 translation
    pose, shape, tsl = grasps[names[0]][3]
    grasps[names[0]][3] = (pose, None, 0)
    pose, shape, tsl = grasps[names[1]][7]
    grasps[names[1]][7] = (pose, shape, None)
    eng = GraspEngine(grasps, names, n_grasp=50, device=DEV)
    rng = np.random.RandomState(0)
    oid = np.concatenate([[0, 1], rng.randint(len(names), size=300)])
    gid = np.concatenate([[3, 7], rng.randint(50, size=300)])
    p, s, t = eng.gather(torch.from_numpy(oid).to(DEV), torch.from_numpy(gid).to(DEV))
    assert p.shape == (302, 48) and s.shape == (302, 10) and t.shape == (302, 3) and p.is_contiguous()
    for i in range(len(oid)):
        rp, rs, rt = eng.get_obj_grasp(names[oid[i]], int(gid[i]))
        np.testing.assert_array_equal(p[i].cpu().numpy(), np.asarray(rp, np.float32))
        np.testing.assert_array_equal(s[i].cpu().numpy(), np.asarray(rs, np.float32))
        np.testing.assert_array_equal(t[i].cpu().numpy(), np.asarray(rt, np.float32))
    assert float(s[0].abs().sum()) == 0.0 and float(t[0].abs().sum()) == 0.0 and float(t[1].abs().sum()) == 0.0
    # int32 ids (what the CCV sampler produces) address the same rows
    p32, _, _ = eng.gather(torch.from_numpy(oid.astype(np.int32)).to(DEV), torch.from_numpy(gid.astype(np.int32)).to(DEV))
    assert torch.equal(p32, p)


def test_rendered_dataset_emits_sample_idx(lib_built):
    """Queries.SAMPLE_IDX of the synthetic sample dict (rendered_dataset.py:272, hoquery.py:7)."""
    from artiboost_b200.synth import SynthPipeline
    from artiboost_b200.train import make_augmenter
    pipe = SynthPipeline(device=DEV, seed=0, n_hand_tex=2, n_bg=2)
    views = pipe.synthesise(6)
    aug = make_augmenter(pipe)
    out = aug(dict(views))
    assert out["sample_idx"].dtype == torch.int64 and out["sample_idx"].tolist() == list(range(6))
    out = aug(dict(views, index_base=100))
    assert out["sample_idx"].tolist() == list(range(100, 106))
    out = aug(dict(views, sample_idx=torch.tensor([5, 4, 3, 2, 1, 0])))
    assert out["sample_idx"].tolist() == [5, 4, 3, 2, 1, 0]


# ---------------------------------------------------------------------------------------- fused tail + criterion
def _tail_inputs(B, seed, with_sym=False, n_obj=5):
    g = torch.Generator(device=DEV).manual_seed(seed)
    r = lambda *s: torch.rand(s, device=DEV, generator=g)  # noqa: E731
    n = lambda *s: torch.randn(s, device=DEV, generator=g)  # noqa: E731
    kp = r(B, 22, 3) * 0.8 + 0.1
    r6 = n(B, 6)
    intr = torch.tensor([[610.0, 0.3, 131.0], [0.0, 605.0, 125.0], [0.0, 0.0, 1.0]], device=DEV).repeat(B, 1, 1)
    intr[:, 0, 0] += r(B) * 20
    jv, cv = torch.ones((B, 21), device=DEV), torch.ones((B, 8), device=DEV)
    jv[1, 3:7] = 0
    cv[2, :] = 0
    cv[3, 1] = 0
    inputs = {"image": torch.zeros((B, 3, 256, 256), device=DEV), "root_joint": n(B, 3) * 0.05 + torch.tensor([0.0, 0.0, 0.55], device=DEV),
              "cam_intr": intr, "corners_can": n(B, 8, 3) * 0.06, "joints_3d": n(B, 21, 3) * 0.05, "corners_3d": n(B, 8, 3) * 0.07,
              "joints_vis": jv, "corners_vis": cv}
    if with_sym:
        inputs["obj_idx"] = torch.randint(1, n_obj + 1, (B,), device=DEV, generator=g)
        T = torch.eye(4, device=DEV).repeat(B, 1, 1)
        from oracle import rotations
        T[:, :3, :3] = torch.from_numpy(rotations.aa_to_rotmat(np.random.RandomState(seed).normal(size=(B, 3))).astype(np.float32)).to(DEV)
        T[:, :3, 3] = n(B, 3) * 0.05 + torch.tensor([0.0, 0.0, 0.55], device=DEV)
        inputs["obj_transf"] = T
    return kp, r6, inputs


def _unfused(kp, r6, inputs, crit, center_idx):
    """The torch composition of hybridbaseline.py:41-96 + Criterion.compute_losses (the definition the fused kernel follows)."""
    from artiboost_b200.models.transform import batch_uvd2xyz, compute_rotation_matrix_from_ortho6d
    p = batch_uvd2xyz(kp, inputs["root_joint"], inputs["cam_intr"], [256, 256])
    j, br = p[:, :21], p[:, 21:22]
    R = compute_rotation_matrix_from_ortho6d(r6)
    c = torch.matmul(R, inputs["corners_can"].permute(0, 2, 1)).permute(0, 2, 1) + br
    root = j[:, center_idx]
    c2 = torch.matmul(inputs["cam_intr"], c.permute(0, 2, 1)).permute(0, 2, 1)
    c2 = c2[:, :, :2] / c2[:, :, 2:3]
    c2 = torch.stack((c2[:, :, 0] / 256.0, c2[:, :, 1] / 256.0, torch.zeros_like(c2[:, :, 0])), dim=2)
    preds = {"joints_3d_abs": j, "corners_3d_abs": c, "joints_3d": j - root.unsqueeze(1), "corners_3d": c - root.unsqueeze(1),
             "2d_uvd": torch.cat((kp[:, :21], c2, kp[:, 21:22]), dim=1), "boxroot_3d_abs": br, "box_rot_rotmat": R}
    total, parts = crit.compute_losses(preds, inputs)
    return preds, total, parts


def _sym_info(n_obj):
    return {str(i + 1): ({"symmetries_continuous": [{"axis": [0, 0, 1], "offset": [0, 0, 0]}]} if i % 3 == 0 else
                         {"symmetries_discrete": [[-1, 0, 0, 0, 0, -1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1]]} if i % 3 == 1 else {}) for i in range(n_obj)}


@pytest.mark.parametrize("cfg_name,B,center", [("default", 7, 0), ("default", 128, 9), ("dexycb_sym", 16, 9), ("ho3d_sym", 5, 0)])
def test_fused_tail_criterion_matches_torch_composition(cfg_name, B, center):
    """ab_tail_losses (tail + criterion + gradient in one launch) vs the torch composition it replaces in TrainStep: the seven
    outputs, every loss part, the total and d total / d (kp3d, rot6d).  Both paths draw from generators with the same seed."""
    from artiboost_b200 import criterions as C
    from artiboost_b200.models.fused_tail import FusedTailCriterion
    if cfg_name == "default":
        cfg = C.DEFAULT_CRITERION_CFG
    else:
        cfg = {"LAMBDAS": [1.0, 0.1, 1.0],
               "CRITERION": [{"TYPE": "JointsLoss", "LAMBDA_JOINTS_3D": 1.0, "LAMBDA_CORNERS_3D": 0.0}, {"TYPE": "HandOrdLoss"},
                             {"TYPE": "SymCornerLoss", "LAMBDA_SYM_CORNERS_3D": 1.0, "MODEL_INFO": _sym_info(5), "MAX_SYM_DISC_STEP": 0.05,
                              "USE_HO3D_YCB": cfg_name == "ho3d_sym"}]}
    kp, r6, inputs = _tail_inputs(B, 3 + B, with_sym=cfg_name != "default")
    crit_a = C.Criterion(cfg, generator=torch.Generator(device=DEV).manual_seed(11))
    crit_b = C.Criterion(cfg, generator=torch.Generator(device=DEV).manual_seed(11))
    kp_a, r6_a = kp.clone().requires_grad_(True), r6.clone().requires_grad_(True)
    preds_a, total_a, parts_a = _unfused(kp_a, r6_a, inputs, crit_a, center)
    total_a.backward()
    plan = FusedTailCriterion.plan(crit_b, center, [256, 256])
    assert plan is not None and plan.usable(inputs)
    kp_b, r6_b = kp.clone().requires_grad_(True), r6.clone().requires_grad_(True)
    preds_b, total_b, parts_b = plan(kp_b, r6_b, inputs)
    (total_b * 1.0).backward()
    for k, v in preds_a.items():
        torch.testing.assert_close(preds_b[k], v.detach(), rtol=1e-5, atol=1e-6, msg=lambda m, k=k: f"{k}: {m}")
    assert set(parts_a) == set(parts_b)
    for k, v in parts_a.items():
        if v is None:
            assert parts_b[k] is None
            continue
        torch.testing.assert_close(parts_b[k], v.detach(), rtol=2e-5, atol=1e-8, msg=lambda m, k=k: f"{k}: {m}")
    torch.testing.assert_close(total_b.detach(), total_a.detach(), rtol=2e-5, atol=1e-8)
    # gradients: relative to the largest entry (single terms flip sign on |x| ~ 1e-7 ordinal margins)
    for name, ga, gb in (("kp3d", kp_a.grad, kp_b.grad), ("rot6d", r6_a.grad, r6_b.grad)):
        scale = float(ga.abs().max())
        assert scale > 0
        err = float((ga - gb).abs().max()) / scale
        assert err < 2e-4, (name, err)
    # both generators advanced identically: the fused path consumed the same draws
    assert torch.equal(crit_a.loss_list[1].generator.get_state(), crit_b.loss_list[1].generator.get_state())


def test_fused_tail_is_bit_reproducible_and_refuses_unknown_losses():
    from artiboost_b200 import criterions as C
    from artiboost_b200.models.fused_tail import FusedTailCriterion
    kp, r6, inputs = _tail_inputs(32, 5)
    outs = []
    for _ in range(2):
        crit = C.Criterion(C.DEFAULT_CRITERION_CFG, generator=torch.Generator(device=DEV).manual_seed(3))
        plan = FusedTailCriterion.plan(crit, 0, [256, 256])
        a, b = kp.clone().requires_grad_(True), r6.clone().requires_grad_(True)
        _, total, _ = plan(a, b, inputs)
        total.backward()
        outs.append((total.detach().clone(), a.grad.clone(), b.grad.clone()))
    assert all(torch.equal(x, y) for x, y in zip(*outs))  # fixed-order reductions, no atomics

    class Other:
        def __call__(self, preds, targs, **kw):
            return torch.zeros((), device=DEV), {}
    crit = C.Criterion({"LAMBDAS": [1.0]}, loss_list=[Other()])
    assert FusedTailCriterion.plan(crit, 0, [256, 256]) is None


def test_inference_tail_launch_matches_the_torch_composition(lib_built):
    """HybridBaseline's no-grad forward runs the tail as one ab_tail_losses launch with zero weights; AB_FUSED_TAIL=0 keeps the
    torch composition (hybridbaseline.py:41-96), which defines it: the seven outputs agree to fp32 round-off."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import netcfg
    import artiboost_b200.models as M
    from artiboost_b200.train import real_shaped_batch
    dev = torch.device("cuda", 0)
    arch, preset = netcfg.arch_cfg("ResNet34")
    torch.manual_seed(2)
    model = M.Arch({"ARCH": arch}, M.build_arch_model_list(arch, preset_cfg=preset)).to(dev).eval()
    batch = real_shaped_batch(5, dev, torch.Generator(device=dev).manual_seed(9))
    with torch.no_grad():
        fused = model(batch)
        os.environ["AB_FUSED_TAIL"] = "0"
        try:
            plain = model(batch)
        finally:
            os.environ.pop("AB_FUSED_TAIL")
    fused = fused[next(iter(fused))] if "joints_3d_abs" not in fused else fused
    plain = plain[next(iter(plain))] if "joints_3d_abs" not in plain else plain
    assert set(fused) == set(plain)
    for k in plain:
        torch.testing.assert_close(fused[k].float(), plain[k].float(), rtol=1e-4, atol=2e-5, msg=k)
