"""GPU: fused Adam + gradient clipping vs torch.optim.Adam + clip_grad_norm_; the full synth -> train step runs, learns
and feeds the CCV re-weighting."""
import sys

import pytest
import torch

from conftest import GOLDEN

sys.path.insert(0, GOLDEN)
import netcfg  # noqa: E402

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_fused_adam_matches_torch_adam_with_clipping(lib_built):
    from artiboost_b200.train import FlatParams, FusedAdam
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(37, 19), torch.nn.Linear(19, 5)).to(DEV)
    ref = torch.nn.Sequential(torch.nn.Linear(37, 19), torch.nn.Linear(19, 5)).to(DEV)
    ref.load_state_dict(net.state_dict())
    flat = FlatParams(net)
    opt = FusedAdam(flat, lr=1e-3, max_norm=0.05, weight_decay=0.0)
    topt = torch.optim.Adam(ref.parameters(), lr=1e-3)
    for it in range(5):
        x = torch.randn((16, 37), device=DEV)
        flat.grad.zero_()
        net(x).pow(2).mean().backward()
        opt.step()
        topt.zero_grad()
        ref(x).pow(2).mean().backward()
        total = torch.nn.utils.clip_grad_norm_(ref.parameters(), 0.05)
        topt.step()
        torch.testing.assert_close(opt.grad_norm()[0], total, rtol=1e-5, atol=1e-8)
    for p, q in zip(net.parameters(), ref.parameters()):
        torch.testing.assert_close(p, q, rtol=1e-5, atol=1e-7)
    assert net[0].weight.data_ptr() >= flat.flat.data_ptr()  # parameters live inside the flat buffer


def test_synth_train_step_runs_learns_and_feeds_ccv(lib_built):
    import artiboost_b200.models as M
    from artiboost_b200.synth import SynthPipeline
    from artiboost_b200.train import CCVFeedback, TrainStep, mix_batches, real_shaped_batch, synth_to_batch
    arch, preset = netcfg.arch_cfg("ResNet34")
    torch.manual_seed(1)
    model = M.Arch({"ARCH": arch}, M.build_arch_model_list(arch, preset_cfg=preset)).to(DEV)
    pipe = SynthPipeline(device=DEV, seed=1, n_hand_tex=4, n_bg=2)
    gen = torch.Generator(device=DEV).manual_seed(5)
    step = TrainStep(model, lr=1e-3, grad_clip=1.0, generator=gen)
    fb = CCVFeedback(pipe.sample_weight_map.shape, DEV)
    views = pipe.synthesise(6)
    batch = mix_batches(real_shaped_batch(10, DEV, gen), synth_to_batch(views, pipe))
    assert batch["image"].shape == (16, 3, 256, 256) and float(batch["is_synth"].sum()) == 6
    w0 = model.model_list[0].backbone.conv1.weight.detach().clone()
    losses = []
    for it in range(6):
        loss, preds = step(batch)
        if it == 0:
            first = preds["joints_3d_abs"].detach().clone()
        if it == 1:  # the packed bf16 filters followed the in-place Adam update (no stale copies)
            assert not torch.equal(first, preds["joints_3d_abs"].detach())
        losses.append(float(loss))
        targ = batch["corners_3d"] + batch["root_joint"].unsqueeze(1)
        fb.feed(preds["corners_3d_abs"].detach(), targ, batch["obj_id"], batch["persp_id"], batch["grasp_id"], batch["is_synth"])
    assert all(l == l for l in losses) and losses[-1] < losses[0], losses   # finite, decreasing on a fixed batch
    assert not torch.equal(w0, model.model_list[0].backbone.conv1.weight)
    assert float(fb.err_cnt.sum()) == 36.0
    new_w = fb.step_eval(pipe.sample_weight_map)
    assert new_w.shape == pipe.sample_weight_map.shape and float(new_w.min()) >= 0.1 and float(new_w.max()) <= 10.0
    assert int((new_w != 1.0).sum()) >= 1
    # BatchNorm running statistics moved, and eval-mode inference still works afterwards
    assert float(model.model_list[0].backbone.bn1.running_mean.abs().sum()) > 0
    with torch.no_grad():
        out = model.eval()(batch)["HybridBaseline"]
    assert torch.isfinite(out["joints_3d_abs"]).all()


def test_cuda_graph_step_matches_eager_step(lib_built):
    """The captured-and-replayed step performs the same arithmetic as the eager one.  lr = 0 keeps the two models'
    parameters identical, so every step's loss and flat gradient can be compared -- bit for bit, since no kernel of the
    step uses atomics (column statistics, weight gradients and BatchNorm reductions are two-pass); then a graph-mode run with lr > 0 must learn."""
    import copy

    import artiboost_b200.models as M
    from artiboost_b200.train import TrainStep, real_shaped_batch
    arch, preset = netcfg.arch_cfg("ResNet34")
    torch.manual_seed(1)
    model_a = M.Arch({"ARCH": arch}, M.build_arch_model_list(arch, preset_cfg=preset)).to(DEV)
    model_b = copy.deepcopy(model_a)
    cfg = {"LAMBDAS": [1.0], "CRITERION": [{"TYPE": "JointsLoss", "LAMBDA_JOINTS_3D": 1.0, "LAMBDA_CORNERS_3D": 0.2}]}
    eager = TrainStep(model_a, cfg, lr=0.0, grad_clip=1.0)
    graphed = TrainStep(model_b, cfg, lr=0.0, grad_clip=1.0, use_graph=True, graph_warmup=2)
    gen = torch.Generator(device=DEV).manual_seed(9)
    batches = [real_shaped_batch(8, DEV, gen) for _ in range(5)]
    for i, b in enumerate(batches):
        la, _ = eager(b)
        lb, _ = graphed(b)
        assert (graphed._graph is not None) == (i >= 2)
        assert float(la) == float(lb), (i, float(la), float(lb))
        ga, gb = eager.flat.grad, graphed.flat.grad
        assert torch.equal(ga, gb), "the step has no atomics: gradients are bit-reproducible, eager or replayed"
        torch.testing.assert_close(eager.opt.grad_norm(), graphed.opt.grad_norm(), rtol=1e-5, atol=0)  # one atomic per CTA
    assert graphed.opt.step_count == eager.opt.step_count == 5
    bn_a, bn_b = model_a.model_list[0].backbone.bn1, model_b.model_list[0].backbone.bn1
    torch.testing.assert_close(bn_a.running_var, bn_b.running_var, rtol=1e-6, atol=0)
    assert int(bn_b.num_batches_tracked) == 5
    # graph mode with the full (randomised) criterion learns, and eager eval afterwards sees the trained parameters
    gen2 = torch.Generator(device=DEV).manual_seed(4)
    learner = TrainStep(model_b, lr=1e-3, grad_clip=1.0, generator=gen2, use_graph=True, graph_warmup=1)
    with torch.no_grad():
        before = model_b.eval()(batches[0])["HybridBaseline"]["joints_3d_abs"].clone()
    losses = [float(learner(batches[0])[0]) for _ in range(8)]
    assert learner._graph is not None and losses[-1] < losses[0], losses
    with torch.no_grad():
        after = model_b.eval()(batches[0])["HybridBaseline"]["joints_3d_abs"]
    assert float((after - before).abs().max()) > 1e-4


def test_fused_tail_step_matches_unfused_step(lib_built, monkeypatch):
    """TrainStep with the tail + criterion fused into ab_tail_losses (default) vs the torch composition (AB_FUSED_TAIL=0):
    same weights, same batch, same generator seed -> same loss and the same flat gradient (to fp32 rounding)."""
    import copy

    import artiboost_b200.models as M
    from artiboost_b200.train import TrainStep, real_shaped_batch
    arch, preset = netcfg.arch_cfg("ResNet34")
    torch.manual_seed(2)
    model_a = M.Arch({"ARCH": arch}, M.build_arch_model_list(arch, preset_cfg=preset)).to(DEV)
    model_b = copy.deepcopy(model_a)
    fused = TrainStep(model_a, lr=0.0, grad_clip=1.0, generator=torch.Generator(device=DEV).manual_seed(7))
    monkeypatch.setenv("AB_FUSED_TAIL", "0")
    plain = TrainStep(model_b, lr=0.0, grad_clip=1.0, generator=torch.Generator(device=DEV).manual_seed(7))
    assert fused.fused_tail is not None and plain.fused_tail is None
    batch = real_shaped_batch(16, DEV, torch.Generator(device=DEV).manual_seed(3))
    for _ in range(2):
        la, pa = fused(batch)
        lb, pb = plain(batch)
        torch.testing.assert_close(la, lb, rtol=1e-4, atol=1e-7)
        assert set(pa) == set(pb)
        for k in pa:
            torch.testing.assert_close(pa[k].float(), pb[k].float(), rtol=1e-4, atol=1e-5, msg=lambda m, k=k: f"{k}: {m}")
        ga, gb = fused.flat.grad, plain.flat.grad
        # the two paths hand head_decode_bwd / the MLP the same fp32 gradient to ~1e-4 (test_gpu_losses.py); below them every
        # layer stores its data gradient in bf16, so that difference re-rolls the roundings of all 36 layers: the flat
        # gradients agree to the bf16 noise floor (~0.2 % per layer, uncorrelated), not to fp32 accuracy
        rel = float((ga - gb).norm() / gb.norm())
        cos = float((ga * gb).sum() / (ga.norm() * gb.norm()))
        assert rel < 3e-2 and cos > 0.9995, (rel, cos)


def test_artiboost_loop_synthesises_augments_trains_and_reweights(lib_built):
    """CCV draw -> pose -> rasterise -> crop / augment -> mix -> train step -> per-cell errors -> new sampling weights."""
    import artiboost_b200.models as M
    from artiboost_b200.synth import SynthPipeline
    from artiboost_b200.train import ArtiBoostLoop
    arch, preset = netcfg.arch_cfg("ResNet34")
    torch.manual_seed(1)
    model = M.Arch({"ARCH": arch}, M.build_arch_model_list(arch, preset_cfg=preset)).to(DEV)
    pipe = SynthPipeline(device=DEV, seed=2, n_hand_tex=4, n_bg=2)
    gen = torch.Generator(device=DEV).manual_seed(6)
    loop = ArtiBoostLoop(model, pipe, batch_size=16, generator=gen, lr=1e-3, grad_clip=1.0)
    assert (loop.n_synth, loop.n_real) == (6, 10)
    batch = loop.make_batch()
    assert batch["image"].shape == (16, 3, 256, 256) and float(batch["is_synth"].sum()) == 6
    synth_img = batch["image"][10:]
    assert float(synth_img.min()) >= -0.5 and float(synth_img.max()) <= 0.5 and float(synth_img.std()) > 0.05
    assert set(batch["joints_vis"].unique().tolist()) <= {0.0, 1.0}
    # the crop keeps the grasp in frame: the root joint projects inside the network input with the NEW intrinsics
    K, root = batch["cam_intr"][10:], batch["root_joint"][10:]
    uv = torch.einsum("bij,bj->bi", K, root)
    uv = uv[:, :2] / uv[:, 2:3]
    assert bool(((uv > -32) & (uv < 288)).all())
    losses = [float(loop.step()) for _ in range(3)]
    assert all(l == l for l in losses)
    assert float(loop.feedback.err_cnt.sum()) == 18.0
    w = loop.end_epoch()
    assert w.shape == pipe.sample_weight_map.shape and float(w.min()) >= 0.1 and float(w.max()) <= 10.0


def test_dexycb_style_config_21_objects_sym_corner_loss(lib_built):
    """BASELINE.json configs[4] in miniature: 21-object CCV space, CENTER_IDX 9, JointsLoss (corners off) + HandOrdLoss +
    SymCornerLoss (config_eval/eval_dexycb_clasbased_sym_artiboost.yaml:39,84-91) through the graph-captured loop."""
    import artiboost_b200.models as M
    from artiboost_b200 import assets
    from artiboost_b200.synth import SynthPipeline
    from artiboost_b200.train import DEFAULT_PRESET, ArtiBoostLoop, make_augmenter
    arch, preset = netcfg.arch_cfg("ResNet34")
    arch["DATA_PRESET"] = dict(preset, CENTER_IDX=9)
    torch.manual_seed(1)
    model = M.Arch({"ARCH": arch}, M.build_arch_model_list(arch, preset_cfg=dict(preset, CENTER_IDX=9))).to(DEV)
    pipe = SynthPipeline(obj_names=list(assets.YCB_NAMES), device=DEV, seed=4, n_hand_tex=4, n_bg=2)
    assert tuple(pipe.sample_weight_map.shape) == (21, 288, 50)
    info = {str(i + 1): ({"symmetries_continuous": [{"axis": [0, 0, 1], "offset": [0, 0, 0]}]} if i % 3 == 0 else
                         {"symmetries_discrete": [[-1, 0, 0, 0, 0, -1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1]]} if i % 3 == 1 else {}) for i in range(21)}
    crit = {"LAMBDAS": [1.0, 0.1, 1.0],
            "CRITERION": [{"TYPE": "JointsLoss", "LAMBDA_JOINTS_3D": 1.0, "LAMBDA_CORNERS_3D": 0.0}, {"TYPE": "HandOrdLoss"},
                          {"TYPE": "SymCornerLoss", "LAMBDA_SYM_CORNERS_3D": 1.0, "MODEL_INFO": info, "MAX_SYM_DISC_STEP": 0.05}]}
    gen = torch.Generator(device=DEV).manual_seed(8)
    loop = ArtiBoostLoop(model, pipe, batch_size=8, criterion_cfg=crit, generator=gen, lr=1e-3, grad_clip=1.0, use_graph=True)
    loop.train_step.graph_warmup = 1
    loop.augmenter = make_augmenter(pipe, cfg_preset=dict(DEFAULT_PRESET, CENTER_IDX=9), generator=gen)
    batch = loop.make_batch()
    assert {"obj_idx", "obj_transf"} <= set(batch) and int(batch["obj_idx"].min()) >= 1 and int(batch["obj_idx"].max()) <= 21
    losses = [float(loop.step()) for _ in range(4)]
    assert loop.train_step._graph is not None and all(l == l for l in losses)
    seen = loop.feedback.err_cnt.sum(dim=(1, 2))
    assert float(seen.sum()) == 4 * loop.n_synth and int((seen > 0).sum()) >= 2   # several of the 21 objects were drawn
    w = loop.end_epoch()
    assert tuple(w.shape) == (21, 288, 50)


def test_training_tracks_the_reference_loop(lib_built):
    """The training-equivalence leg of bench.py at a size a test can afford: the same stream and the same initial weights
    through this repo's step (bf16 tensor cores, fused clip + Adam, CUDA graph) and through the reference's loop arithmetic
    (torch.nn fp32, Criterion, clip_grad_norm_, torch.optim.Adam; train/train_artiboost.py:66-96): the per-step losses stay
    together and both networks end at the same Mean3DEPE on a fixed set of rendered samples."""
    import bench
    r = bench.train_equivalence(torch.device(DEV), steps=40, batch=32, eval_samples=256)
    assert r["loss_max_rel_diff"] < 0.08, r
    assert r["mpcpe_rel_diff"] < 0.05 and r["mpjpe_rel_diff"] < 0.10, r
    # 40 steps at lr 5e-5 already move both networks the same way
    assert r["loss_last_steps"]["ours"] < r["loss_first_steps"]["ours"]
    assert r["loss_last_steps"]["reference_loop"] < r["loss_first_steps"]["reference_loop"]
    assert abs(r["mpcpe_before_mm"]["ours"] - r["mpcpe_before_mm"]["reference_loop"]) < 3.0
