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
