"""CPU tests of the host-side logic: losses vs the reference's own criterions (fixture), CCV feedback vs the oracle,
multi-process plumbing on the gloo backend (world_size 2), registries."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import golden


def test_criterion_matches_reference_losses(monkeypatch):
    """artiboost_b200/criterions.py vs anakin/criterions/* (tests/golden/losses.npz, random draws pinned)."""
    from artiboost_b200 import criterions as C
    g = golden("losses.npz")
    t = torch.from_numpy
    preds = {"joints_3d_abs": t(g["pred_joints_3d_abs"]), "corners_3d_abs": t(g["pred_corners_3d_abs"])}
    targs = {k: t(g["targ_" + k]) for k in ("joints_3d", "corners_3d", "root_joint", "joints_vis", "corners_vis")}
    vv = {20: t(g["vv20"]), 40: t(g["vv40"])}
    monkeypatch.setattr(C, "sample_view_vectors", lambda n, device, generator=None: vv[n])
    monkeypatch.setattr(C, "_subsample", lambda n, device, generator: torch.arange(n // 3))
    crit = C.Criterion(C.DEFAULT_CRITERION_CFG)
    total, parts = crit.compute_losses(preds, targs)
    np.testing.assert_allclose(total.numpy(), g["total"].reshape(()), rtol=1e-5)
    for k in ("joints_3d_loss", "corners_3d_loss", "joint_ord_loss", "part_ord_loss", "scene_ord_loss"):
        np.testing.assert_allclose(parts[k].numpy(), g["part_" + k], rtol=1e-5, atol=1e-9, err_msg=k)
    # unpinned draws: finite, differentiable
    crit = C.Criterion(C.DEFAULT_CRITERION_CFG, generator=torch.Generator().manual_seed(0))
    p = {k: v.clone().requires_grad_(True) for k, v in preds.items()}
    total, _ = crit.compute_losses(p, targs)
    total.backward()
    assert torch.isfinite(total) and all(torch.isfinite(v.grad).all() for v in p.values())


@pytest.mark.parametrize("method,epoch,key,ratio_key", [("method_2", 0, "w2", None), ("method_3", 0, "w3", "ratio3"),
                                                         ("method_4", 10, "w4a", "ratio4a"), ("method_4", 80, "w4b", "ratio4b")])
def test_ccv_feedback_update_methods_2_to_4_match_reference(method, epoch, key, ratio_key):
    """CCVFeedback's mining strategies vs the outputs of the reference's own update_method_2/3/4
    (artiboost_loader.py:526-598; tests/golden/make_golden.py gen_update_method) and vs the oracle restatement."""
    from artiboost_b200.train import CCVFeedback
    from oracle import ccv
    g = golden("update_method234.npz")
    fb = CCVFeedback(g["w0"].shape, "cpu", method=method, dist_lower=8.0, dist_upper=16.0, n_epochs=100)
    cells, vals = torch.from_numpy(g["cells"]), torch.from_numpy(g["vals"]).float()
    for delta in (-1.0, 1.0):
        pred = torch.zeros((len(vals), 8, 3))
        pred[:, :, 0] = ((vals + delta) / 1000.0)[:, None]
        fb.feed(pred, torch.zeros_like(pred), cells[:, 0], cells[:, 1], cells[:, 2])
    w = fb.step_eval(torch.from_numpy(g["w0"]), epoch_idx=epoch).numpy()
    np.testing.assert_allclose(w, g[key], rtol=2e-5, atol=1e-7)
    if ratio_key:
        assert abs(fb.dist_lower_ratio - float(g[ratio_key])) < 1e-6
    ref = {"method_2": lambda: (ccv.update_method_2(g["w0"], g["cells"], g["vals"]), None),
           "method_3": lambda: ccv.update_method_3(g["w0"], g["cells"], g["vals"]),
           "method_4": lambda: ccv.update_method_4(g["w0"], g["cells"], g["vals"], epoch, 100)}[method]()
    np.testing.assert_allclose(w, ref[0], rtol=2e-5, atol=1e-7)
    with pytest.raises(KeyError):
        CCVFeedback(g["w0"].shape, "cpu", method="method_5")


def test_ccv_feedback_matches_oracle_update_method_1():
    from artiboost_b200.train import CCVFeedback
    from oracle import ccv
    g = golden("update_method.npz")
    fb = CCVFeedback(g["w0"].shape, "cpu")
    cells, vals = torch.from_numpy(g["cells"]), torch.from_numpy(g["vals"]).float()
    # feed every cell twice with errors whose mean is the recorded value (in metres: feed() converts to millimetres)
    for delta in (-1.0, 1.0):
        err = (vals + delta) / 1000.0
        pred = torch.zeros((len(vals), 8, 3))
        pred[:, :, 0] = err[:, None]
        fb.feed(pred, torch.zeros_like(pred), cells[:, 0], cells[:, 1], cells[:, 2])
    w1 = fb.step_eval(torch.from_numpy(g["w0"]))
    np.testing.assert_allclose(w1.numpy(), g["w1"], rtol=2e-5)  # the reference's own update_method_1 output
    np.testing.assert_allclose(w1.numpy(), ccv.update_method_1(g["w0"], g["cells"], g["vals"]), rtol=2e-5)
    assert float(fb.err_cnt.sum()) == 0.0


def test_shard_range_is_a_partition():
    from artiboost_b200.parallel import shard_range
    for n in (0, 1, 7, 512, 1000):
        for w in (1, 2, 3, 8):
            cuts = [shard_range(n, r, w) for r in range(w)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
            sizes = [hi - lo for lo, hi in cuts]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from artiboost_b200 import parallel
    from artiboost_b200.train import CCVFeedback
    assert parallel.world() == (rank, world)
    # gradient all-reduce on a flat buffer, bucketed
    g = torch.full((1000,), float(rank + 1))
    parallel.allreduce_sum_(g, bucket_elems=256)
    assert torch.equal(g, torch.full((1000,), 3.0))
    # CCV feedback: each rank feeds its own shard of the samples; step_eval must equal the single-process result
    rng = np.random.RandomState(0)
    shape = (4, 12, 5)
    n = 200
    cells = np.stack([rng.randint(s, size=n) for s in shape], 1)
    err = rng.uniform(0.005, 0.06, size=n).astype(np.float32)
    lo, hi = parallel.shard_range(n)
    fb = CCVFeedback(shape, "cpu")
    pred = torch.zeros((hi - lo, 8, 3))
    pred[:, :, 2] = torch.from_numpy(err[lo:hi])[:, None]
    c = torch.from_numpy(cells[lo:hi])
    fb.feed(pred, torch.zeros_like(pred), c[:, 0], c[:, 1], c[:, 2])
    occ = torch.zeros(shape, dtype=torch.bool)
    occ[c[:, 0], c[:, 1], c[:, 2]] = True
    _, _, occ = parallel.allreduce_cell_errors_(torch.zeros(shape), torch.zeros(shape), occ)
    w = fb.step_eval(torch.ones(shape))
    torch.save({"w": w, "occ": occ}, os.path.join(out_dir, f"rank{rank}.pt"))
    dist.destroy_process_group()


def test_two_rank_gloo_allreduce_and_ccv_feedback(tmp_path):
    from artiboost_b200.train import CCVFeedback
    port = 29500 + os.getpid() % 400
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = torch.load(tmp_path / "rank0.pt"), torch.load(tmp_path / "rank1.pt")
    assert torch.equal(r0["w"], r1["w"]) and torch.equal(r0["occ"], r1["occ"])
    rng = np.random.RandomState(0)
    shape, n = (4, 12, 5), 200
    cells = np.stack([rng.randint(s, size=n) for s in shape], 1)
    err = rng.uniform(0.005, 0.06, size=n).astype(np.float32)
    fb = CCVFeedback(shape, "cpu")
    pred = torch.zeros((n, 8, 3))
    pred[:, :, 2] = torch.from_numpy(err)[:, None]
    c = torch.from_numpy(cells)
    fb.feed(pred, torch.zeros_like(pred), c[:, 0], c[:, 1], c[:, 2])
    torch.testing.assert_close(fb.step_eval(torch.ones(shape)), r0["w"], rtol=1e-5, atol=1e-6)
    occ = torch.zeros(shape, dtype=torch.bool)
    occ[c[:, 0], c[:, 1], c[:, 2]] = True
    assert torch.equal(occ, r0["occ"])


def test_model_registry_builds_the_reference_config_on_cpu():
    import artiboost_b200.models as M
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import netcfg
    arch, preset = netcfg.arch_cfg("ResNet34")
    g = golden("network_resnet34.npz")
    torch.manual_seed(netcfg.SEED)
    model = M.Arch({"ARCH": arch}, M.build_arch_model_list(arch, preset_cfg=preset))
    assert sum(p.numel() for p in model.parameters()) == int(g["n_params"]) == 25267734
    from artiboost_b200.lib import AbError
    with torch.no_grad(), pytest.raises(AbError):
        model.eval()(netcfg.make_inputs(1))  # host tensors: there is no CPU path


def test_sym_corner_loss_matches_reference():
    """artiboost_b200.criterions.SymCornerLoss vs anakin/criterions/symcornerloss.py (tests/golden/symloss.npz):
    symmetry tables (identity / discrete / discretised continuous, combined) and the min-over-symmetries loss."""
    import json

    from artiboost_b200 import criterions as C
    g = golden("symloss.npz")
    info = json.loads(str(g["model_info"]))
    t = torch.from_numpy
    targs = {"obj_idx": t(g["obj_idx"]), "corners_can": t(g["corners_can"]), "obj_transf": t(g["obj_transf"]), "corners_vis": t(g["corners_vis"])}
    for flag in (0, 1):
        loss = C.SymCornerLoss(LAMBDA_SYM_CORNERS_3D=1.0, MODEL_INFO=info, MAX_SYM_DISC_STEP=0.05, USE_HO3D_YCB=bool(flag))
        np.testing.assert_allclose(loss.R, g[f"R{flag}"], atol=1e-6)
        np.testing.assert_allclose(loss.t, g[f"t{flag}"], atol=1e-7)
        pred = t(g["pred"]).clone().requires_grad_(True)
        final, parts = loss({"corners_3d_abs": pred}, targs)
        np.testing.assert_allclose(parts["sym_corners_3d_loss"].detach().numpy(), g[f"loss_ho3d{flag}"], rtol=1e-5)
        final.backward()
        assert torch.isfinite(pred.grad).all() and float(pred.grad.abs().sum()) > 0
    off = C.SymCornerLoss(LAMBDA_SYM_CORNERS_3D=0.0, MODEL_INFO=info)
    assert off({"corners_3d_abs": t(g["pred"])}, targs)[1]["sym_corners_3d_loss"] is None
    crit = C.Criterion({"LAMBDAS": [1.0, 1.0], "CRITERION": [{"TYPE": "JointsLoss", "LAMBDA_JOINTS_3D": 1.0, "LAMBDA_CORNERS_3D": 0.0},
                                                            {"TYPE": "SymCornerLoss", "LAMBDA_SYM_CORNERS_3D": 1.0, "MODEL_INFO": info}]})
    assert [type(l).__name__ for l in crit.loss_list] == ["JointsLoss", "SymCornerLoss"]


def test_fused_tail_plan_weights_and_refusals():
    """FusedTailCriterion.plan (host logic of the fused tail + criterion kernel): effective weights = criterion lambda x the
    loss's own lambda, draw order = criterion order, and None for criteria the kernel cannot express."""
    from artiboost_b200 import criterions as C
    from artiboost_b200.models.fused_tail import FusedTailCriterion
    plan = FusedTailCriterion.plan(C.Criterion(C.DEFAULT_CRITERION_CFG), 9, [256, 256])
    assert plan is not None and plan.order == ["JointsLoss", "HandOrdLoss", "SceneOrdLoss"] and plan.center_idx == 9
    w = plan.weights
    assert (w["joints"], w["corners"], w["sym"]) == (0.5 * 1.0, 0.5 * 0.2, 0.0)
    assert w["joint_ord"] == w["part_ord"] == 0.2 and abs(w["scene_ord"] - 0.1) < 1e-12
    assert not plan.usable({"root_joint": 0}) and plan.usable({k: 0 for k in plan.TARGET_KEYS})
    info = {"1": {"symmetries_discrete": [[-1, 0, 0, 0, 0, -1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1]]}, "2": {}}
    cfg = {"LAMBDAS": [1.0, 2.0], "CRITERION": [{"TYPE": "JointsLoss", "LAMBDA_JOINTS_3D": 1.0},
                                                 {"TYPE": "SymCornerLoss", "LAMBDA_SYM_CORNERS_3D": 0.5, "MODEL_INFO": info}]}
    plan = FusedTailCriterion.plan(C.Criterion(cfg), 0, [256, 256])
    assert plan.weights["sym"] == 1.0 and plan.weights["corners"] == 0.0 and plan.sym_t3.shape == (2, 2, 3)
    assert not plan.usable({k: 0 for k in plan.TARGET_KEYS})       # SymCornerLoss also needs obj_idx / obj_transf
    twice = C.Criterion({"LAMBDAS": [1.0, 1.0]}, loss_list=[C.JointsLoss(LAMBDA_JOINTS_3D=1.0), C.JointsLoss(LAMBDA_CORNERS_3D=1.0)])
    assert FusedTailCriterion.plan(twice, 0, [256, 256]) is None   # loss_lambdas is keyed by type: one of each at most


def test_filter_k_padding_and_batch_counter_context():
    """K of the filter / im2col matrices is padded to 16 elements (32-byte rows); BatchNorm batch counters collected inside
    train_ops.batch_counters() are bumped once, together, on exit -- also when the contexts nest."""
    import torch
    from artiboost_b200.models import nhwc, train_ops
    assert [nhwc._pad16(n) for n in (1, 16, 17, 196, 576)] == [16, 16, 32, 208, 576]
    w = torch.arange(2 * 3 * 7 * 7, dtype=torch.float32).reshape(2, 3, 7, 7)
    wp = nhwc.pack_conv_weight(w, cin_pad=4)
    assert wp.shape == (2, 208) and float(wp[:, 196:].abs().sum()) == 0.0
    assert float(wp[1, (2 * 7 + 5) * 4 + 1]) == float(w[1, 1, 2, 5].to(torch.bfloat16))   # K order (ky, kx, ci)
    a, b = torch.zeros((), dtype=torch.long), torch.zeros((), dtype=torch.long)
    assert train_ops._BatchCounters.pending is None
    with train_ops.batch_counters():
        train_ops._BatchCounters.pending.append(a)
        with train_ops.batch_counters():
            train_ops._BatchCounters.pending.append(b)
        assert int(b) == 1 and int(a) == 0
    assert int(a) == 1 and train_ops._BatchCounters.pending is None


def test_stream_override_redirects_library_launches_and_restores():
    """lib.use_stream(s): inside the block every stream handle the library is given is s's (torch's current stream is not
    consulted), nested blocks restore the outer one, and an exception leaves no override behind."""
    from artiboost_b200 import lib

    class FakeStream:
        def __init__(self, handle):
            self.cuda_stream = handle

    a, b = FakeStream(0x1000), FakeStream(0x2000)
    assert lib._stream_override is None
    with lib.use_stream(a):
        assert lib.stream_ptr().value == 0x1000
        with lib.use_stream(b):
            assert lib.stream_ptr("cuda:3").value == 0x2000
        assert lib.stream_ptr().value == 0x1000
    assert lib._stream_override is None
    try:
        with lib.use_stream(b):
            raise RuntimeError("boom")
    except RuntimeError:
        pass
    assert lib._stream_override is None
