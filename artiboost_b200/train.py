"""The training step of train/train_artiboost.py:66-96 and the ArtiBoost feedback loop of
anakin/artiboost/artiboost_loader.py:279-340,503-523, on device:

  synthesise views (CCV draw -> pose generator -> rasterise)  ->  mix with real-shaped samples  ->  HybridBaseline forward
  (batch-statistics BatchNorm)  ->  losses  ->  backward kernels  ->  gradient all-reduce (NCCL)  ->  clip_grad_norm_ + Adam
  as one fused kernel over a flat parameter buffer  ->  per-cell error capture for the CCV re-weighting.
"""
from __future__ import annotations

import os
from typing import Dict, Optional

import torch
import torch.nn as nn

from . import lib, parallel
from .models import nhwc, train_ops
from .criterions import DEFAULT_CRITERION_CFG, Criterion


class FlatParams:
    """Re-homes every parameter of `model` (and its gradient) as a view into one flat fp32 buffer, so the all-reduce,
    the norm and the Adam update are single launches.  state_dict names and shapes are untouched."""

    def __init__(self, model: nn.Module):
        params = [p for p in model.parameters() if p.requires_grad]
        self.params = params
        n = sum(p.numel() for p in params)
        dev = params[0].device
        n_pad = (n + 3) // 4 * 4  # the optimiser kernel works on float4
        self.flat = torch.zeros(n_pad, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(n_pad, dtype=torch.float32, device=dev)
        off = 0
        for p in params:
            k = p.numel()
            self.flat[off:off + k].copy_(p.data.reshape(-1))
            p.data = self.flat[off:off + k].view_as(p.data)
            p.grad = self.grad[off:off + k].view_as(p.data)
            off += k
        self.numel = n_pad

    def zero_grad(self):
        self.grad.zero_()


class FusedAdam:
    """torch.optim.Adam(lr, betas, eps, weight_decay) + clip_grad_norm_(max_norm) in two launches (ab_sumsq, ab_adam_step)."""

    def __init__(self, flat: FlatParams, lr=5e-5, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, max_norm=0.0):
        self.fp, self.lr, self.betas, self.eps, self.wd, self.max_norm = flat, lr, betas, eps, weight_decay, max_norm
        self.m = torch.zeros_like(flat.flat)
        self.v = torch.zeros_like(flat.flat)
        self.sumsq = torch.zeros(1 + lib.SUMSQ_PARTS, device=flat.flat.device)   # [0] = the sum, the rest per-CTA partials
        self.state = torch.zeros(3, device=flat.flat.device)  # {step count, 1 - beta1^t, 1 - beta2^t}, advanced on device

    @property
    def step_count(self) -> int:
        return int(self.state[0].item())

    def step(self, grad_scale: float = 1.0):
        fp, dev = self.fp, self.fp.flat.device
        L = lib.load()
        with torch.cuda.device(dev):
            if self.max_norm > 0:
                lib.check(L.ab_sumsq(fp.grad.data_ptr(), fp.numel, self.sumsq.data_ptr(), lib.stream_ptr(dev)), "ab_sumsq")
            lib.check(L.ab_adam_step(fp.flat.data_ptr(), fp.grad.data_ptr(), self.m.data_ptr(), self.v.data_ptr(), fp.numel,
                                     self.lr, self.betas[0], self.betas[1], self.eps, self.wd, self.state.data_ptr(),
                                     self.sumsq.data_ptr(), float(self.max_norm), float(grad_scale), lib.stream_ptr(dev)),
                      "ab_adam_step")
        nhwc.bump_params()  # the kernel rewrote the parameters in place: packed bf16 filter copies are stale

    def grad_norm(self) -> torch.Tensor:
        return self.sumsq[:1].sqrt()

    # -- torch.optim.Adam checkpoint format, so that `train_param.pth.tar` files move between the reference's loop
    # (anakin/utils/io_utils.py:34,74-84: optimizer.state_dict() / load_state_dict()) and this one
    def state_dict(self) -> dict:
        step = float(self.state[0].item())
        state, off = {}, 0
        for i, p in enumerate(self.fp.params):
            k = p.numel()
            if step > 0:
                state[i] = {"step": torch.tensor(step), "exp_avg": self.m[off:off + k].view_as(p).clone(),
                            "exp_avg_sq": self.v[off:off + k].view_as(p).clone()}
            off += k
        group = {"lr": self.lr, "betas": tuple(self.betas), "eps": self.eps, "weight_decay": self.wd, "amsgrad": False,
                 "maximize": False, "params": list(range(len(self.fp.params)))}
        return {"state": state, "param_groups": [group]}

    def load_state_dict(self, sd: dict) -> None:
        ids = [i for g in sd["param_groups"] for i in g["params"]]
        if len(ids) != len(self.fp.params):
            raise ValueError(f"loaded state dict has {len(ids)} parameters, the optimizer has {len(self.fp.params)}")
        g0 = sd["param_groups"][0]
        self.lr, self.betas, self.eps = float(g0["lr"]), tuple(g0["betas"]), float(g0["eps"])
        self.wd = float(g0.get("weight_decay", 0.0))
        self.m.zero_()
        self.v.zero_()
        step, off = 0.0, 0
        for pid, p in zip(ids, self.fp.params):
            k = p.numel()
            st = sd["state"].get(pid)
            if st is not None:
                if tuple(st["exp_avg"].shape) != tuple(p.shape):
                    raise ValueError(f"optimizer state {pid} has shape {tuple(st['exp_avg'].shape)}, parameter {tuple(p.shape)}")
                self.m[off:off + k].copy_(st["exp_avg"].reshape(-1))
                self.v[off:off + k].copy_(st["exp_avg_sq"].reshape(-1))
                step = max(step, float(st["step"]))
            off += k
        b1, b2 = self.betas
        self.state.copy_(torch.tensor([step, 1.0 - b1 ** step, 1.0 - b2 ** step]))


class CCVFeedback:
    """Per-(object, view, grasp) error capture during training and the mining strategies `update_method_1..4`
    (val_metric.py:76-135 ValMetricMean3DEPE2 + artiboost_loader.py:240-265,292-340,503-598), kept on the device.

    method: the reference's `UPDATE_METHOD` key ("method_1" percentile, "method_2" incremental, "method_3" lower-bound
    deactivation, "method_4" = method_1 for the first 75 % of the epochs, then method_3); lower / upper =
    `WEIGHT_UPDATE.LOWER/UPPER`; dist_lower / dist_upper = `DIST_THRESHOLD.LOWER/UPPER` (millimetres)."""

    METHODS = ("method_1", "method_2", "method_3", "method_4")

    def __init__(self, shape, device, lower=0.1, upper=10.0, method="method_1", dist_lower=8.0, dist_upper=16.0, n_epochs=1):
        if method not in self.METHODS:
            raise KeyError(method)  # the reference's update_method_mapping lookup fails the same way
        self.shape = tuple(shape)
        self.err_sum = torch.zeros(self.shape, device=device)
        self.err_cnt = torch.zeros(self.shape, device=device)
        self.lower, self.upper = lower, upper
        self.method, self.dist_lower, self.dist_upper, self.n_epochs = method, dist_lower, dist_upper, n_epochs
        self.dist_lower_ratio = -1.0  # as reported by method_3 / method_4 (-1: not evaluated)

    @torch.no_grad()
    def feed(self, pred_corners_abs, targ_corners_abs, obj_id, persp_id, grasp_id, is_synth=None):
        """mean corner error in millimetres per sample (meanepe.py:39-70), accumulated in its CCV cell."""
        err = (pred_corners_abs - targ_corners_abs).norm(dim=-1).mean(dim=-1) * 1000.0
        # real samples are weighted out instead of masked out: boolean indexing would force a host sync every step
        w = torch.ones_like(err) if is_synth is None else is_synth.to(err.dtype)
        flat = (obj_id.long() * self.shape[1] + persp_id.long()) * self.shape[2] + grasp_id.long()
        self.err_sum.view(-1).index_add_(0, flat, err.float() * w)
        self.err_cnt.view(-1).index_add_(0, flat, w.float())

    def _confidence(self, val):
        vmax, vmin = val.max(), val.min()
        return (vmax - val) / (vmax - vmin + 1e-8)

    def _method_1(self, new, seen, val):   # artiboost_loader.py:503-523
        new[seen] = new[seen] * (1.0 / (self._confidence(val) + 0.5))
        return torch.clamp(new, self.lower, self.upper)  # also lifts blacklisted zeros to `lower`, like the reference

    def _method_2(self, new, seen, val):   # :526-545
        step = torch.where(self._confidence(val) > 0.5, -0.1, 0.1).to(new.dtype)
        new[seen] = new[seen] + step
        return torch.clamp(new, self.lower, self.upper)

    def _method_3(self, new, seen, val):   # :548-569 (no clamp: deactivated cells stay at 0)
        low, high = val < self.dist_lower, val > self.dist_upper
        cur = new[seen]
        new[seen] = torch.where(low, torch.zeros_like(cur), torch.where(high, torch.ones_like(cur), cur * 0.5))
        self.dist_lower_ratio = float(low.float().mean())
        return new

    @torch.no_grad()
    def step_eval(self, weight_map: torch.Tensor, epoch_idx: int = 0) -> torch.Tensor:
        """All-reduce the cell statistics over ranks, apply the configured update method, reset.  -> new weight map."""
        parallel.allreduce_cell_errors_(self.err_sum, self.err_cnt)
        seen = self.err_cnt > 0
        new = weight_map.clone()
        self.dist_lower_ratio = -1.0
        if bool(seen.any()):
            val = self.err_sum[seen] / self.err_cnt[seen]
            method = self.method
            if method == "method_4":   # :572-598
                method = "method_1" if float(epoch_idx) / self.n_epochs < 0.75 else "method_3"
            new = {"method_1": self._method_1, "method_2": self._method_2, "method_3": self._method_3}[method](new, seen, val)
        elif self.method in ("method_1", "method_2") or (self.method == "method_4" and float(epoch_idx) / self.n_epochs < 0.75):
            new = torch.clamp(new, self.lower, self.upper)
        self.err_sum.zero_()
        self.err_cnt.zero_()
        return new


class TrainStep:
    """One optimisation step on a batch dict with the reference's keys (image, root_joint, cam_intr, corners_can,
    joints_3d, corners_3d, joints_vis, corners_vis).

    use_graph: after `graph_warmup` eager steps the whole step (zero-grad, forward, losses, backward kernels, gradient
    all-reduce, clip + Adam, filter re-packing) is captured ONCE into a CUDA graph and replayed: ~1400 launches become one
    submission.  The batch is then copied into static input buffers, the returned loss / predictions are static
    tensors that the next step overwrites, and the batch size and learning rate are fixed for the graph's lifetime."""

    BATCH_KEYS = ("image", "root_joint", "cam_intr", "corners_can", "joints_3d", "corners_3d", "joints_vis", "corners_vis")

    def __init__(self, arch: nn.Module, criterion_cfg: Optional[dict] = None, lr=5e-5, grad_clip=1e-3, generator=None,
                 use_graph: bool = False, graph_warmup: int = 3):
        self.arch = arch
        self.flat = FlatParams(arch)
        self.opt = FusedAdam(self.flat, lr=lr, max_norm=grad_clip)   # Adam lr 5e-5, clip 1e-3 (yaml:130-142)
        self.criterion = Criterion(criterion_cfg or DEFAULT_CRITERION_CFG, generator=generator)
        self.generator = generator
        self.world = parallel.world()[1]
        self.use_graph, self.graph_warmup = use_graph, graph_warmup
        self.async_wgrad = os.environ.get("AB_ASYNC_WGRAD", "1") != "0"
        self._graph, self._static, self._out, self._eager_steps = None, None, None, 0
        self.fused_tail = self._install_fused_tail() if os.environ.get("AB_FUSED_TAIL", "1") != "0" else None

    def _install_fused_tail(self):
        """A single HybridBaseline with a criterion made of the clasbased losses gets its tail, the criterion and their
        backward as ONE kernel (models/fused_tail.py); anything else keeps the torch composition."""
        from .models.fused_tail import FusedTailCriterion
        from .models.hybridbaseline import HybridBaseline
        heads = [m for m in self.arch.modules() if isinstance(m, HybridBaseline)]
        if len(heads) != 1:
            return None
        plan = FusedTailCriterion.plan(self.criterion, heads[0].center_idx, heads[0].inp_res)
        heads[0].fused_tail = plan
        return plan

    def _eager(self, batch: Dict[str, torch.Tensor]):
        self.arch.train()
        self.flat.grad.zero_()
        with train_ops.batch_counters():   # the BatchNorm layers' num_batches_tracked in one launch
            preds = self.arch(batch)
        preds = preds[next(iter(preds))] if "joints_3d_abs" not in preds else preds
        if "_fused_loss" in preds:
            preds = dict(preds)
            loss, parts = preds.pop("_fused_loss"), preds.pop("_fused_parts")
        else:
            loss, parts = self.criterion.compute_losses(preds, batch)
        if self.async_wgrad:
            with train_ops.async_wgrad():   # weight gradients on a side stream, joined before the all-reduce
                loss.backward()
        else:
            loss.backward()
        parallel.allreduce_sum_(self.flat.grad)
        self.opt.step(grad_scale=1.0 / self.world)
        return loss.detach(), {k: v.detach() for k, v in preds.items()}

    def _check_grads(self):
        lo, hi = self.flat.grad.data_ptr(), self.flat.grad.data_ptr() + 4 * self.flat.numel
        for p in self.flat.params:  # gradients accumulate straight into the flat buffer's views
            if p.grad is None or not lo <= p.grad.data_ptr() < hi:
                raise RuntimeError("a parameter gradient left the flat buffer")

    def _capture(self, batch):
        dev = self.flat.flat.device
        # every tensor of the batch gets a static twin (the losses may read more than BATCH_KEYS, e.g. obj_idx / obj_transf)
        self._static = {k: v.detach().clone() for k, v in batch.items() if torch.is_tensor(v)}
        missing = [k for k in self.BATCH_KEYS if k not in self._static]
        if missing:
            raise KeyError(f"batch lacks {missing}")
        torch.cuda.synchronize(dev)
        graph = torch.cuda.CUDAGraph()
        if self.generator is not None:
            graph.register_generator_state(self.generator)
        with torch.cuda.graph(graph):
            self._out = self._eager(self._static)
        self._graph = graph

    def close(self):
        """Destroys the captured graph (it holds the NCCL all-reduce of the step): call before
        torch.distributed.destroy_process_group(), which otherwise waits on work the graph still references."""
        if self._graph is not None:
            torch.cuda.synchronize(self.flat.flat.device)
            self._graph.reset()
        self._graph, self._static, self._out = None, None, None

    def __call__(self, batch: Dict[str, torch.Tensor]):
        self._check_grads()
        if not self.use_graph:
            return self._eager(batch)
        if self._graph is None:
            if self._eager_steps < self.graph_warmup:
                self._eager_steps += 1
                side = torch.cuda.Stream(self.flat.flat.device)   # warm up off the default stream, as capture will run
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    out = self._eager(batch)
                torch.cuda.current_stream().wait_stream(side)
                return out
            self._capture(batch)
        for k, buf in self._static.items():
            buf.copy_(batch[k], non_blocking=True)
        self._graph.replay()
        nhwc.bump_params()  # the replayed Adam kernel rewrote the parameters: eager users must re-pack their filters
        return self._out


# AUG / AUG_PARAM of the HO3D training config (config/ho3dv2_clasbased_jlol_artiboost2.yaml:85-88) and its DATA_PRESET
# (:98-128) at the 256 x 256 network input of BASELINE.json (the shipped config uses 224 x 224)
DEFAULT_AUG_CFG = {"AUG": True, "AUG_PARAM": {"SCALE_JIT": 0.1, "CENTER_JIT": 0.1, "MAX_ROT": 0.2}}
DEFAULT_PRESET = {"IMAGE_SIZE": [256, 256], "CENTER_IDX": 0, "BBOX_EXPAND_RATIO": 1.2, "FULL_IMAGE": False, "CROP_MODEL": "root_obj"}


def make_augmenter(pipe, cfg_dataset: Optional[dict] = None, cfg_preset: Optional[dict] = None, generator=None):
    """The crop / augment stage that follows the rasteriser (artiboost.RenderedDataset on this pipeline's camera)."""
    from .artiboost import RenderedDataset
    r = pipe.renderer
    return RenderedDataset(pipe.obj_engine.corners_can.cpu().numpy(), pipe.cam_intr, cfg_dataset or DEFAULT_AUG_CFG,
                           cfg_preset or DEFAULT_PRESET, raw_size=(r.width, r.height), device=pipe.device, generator=generator)


def synth_to_batch(views: dict, pipe, center_idx: int = 0, augmenter=None) -> Dict[str, torch.Tensor]:
    """Rendered views + their annotations -> the network's batch dict.  With an `augmenter` (artiboost.RenderedDataset):
    the full RenderedDataset.__getitem__ (crop box, jitter, blur, colour jitter, affine warp, transformed annotations;
    rendered_dataset.py:155-274) on the device.  Without: only the un-augmented part (:127-133,207-272), render size =
    network input size."""
    if augmenter is not None:
        return augmenter({"rgba": views["rgba"], "joints": views["joints"], "obj_pose": views["obj_pose"], "obj_id": views["obj_id"],
                          "persp_id": views["persp_id"], "grasp_id": views["grasp_id"]})
    rgba = views["rgba"]
    B = rgba.shape[0]
    image = rgba[..., :3].permute(0, 3, 1, 2).float() / 255.0 - 0.5           # to_tensor, then -0.5 (:269-270)
    joints = views["joints"]
    pose = views["obj_pose"]
    corners_can = pipe.obj_engine.corners_can[views["obj_id"].long()]
    corners = torch.einsum("bij,bkj->bki", pose[:, :3, :3], corners_can) + pose[:, :3, 3].unsqueeze(1)
    root = joints[:, center_idx]
    K = torch.as_tensor(pipe.cam_intr, device=rgba.device).expand(B, 3, 3).contiguous()
    ones = lambda n: torch.ones((B, n), device=rgba.device)  # noqa: E731
    return {"image": image, "root_joint": root, "cam_intr": K, "corners_can": corners_can,
            "joints_3d": joints - root.unsqueeze(1), "corners_3d": corners - root.unsqueeze(1),
            "joints_vis": ones(21), "corners_vis": ones(8), "is_synth": torch.ones(B, device=rgba.device),
            "obj_id": views["obj_id"], "persp_id": views["persp_id"], "grasp_id": views["grasp_id"]}


def real_shaped_batch(B: int, device, generator=None, size: int = 256) -> Dict[str, torch.Tensor]:
    """Synthetic stand-in for real dataset samples with the schema of anakin/datasets/hodata.py:315-450: random image,
    geometrically consistent random annotations."""
    g = generator
    r = lambda *s: torch.rand(s, device=device, generator=g)  # noqa: E731
    n = lambda *s: torch.randn(s, device=device, generator=g)  # noqa: E731
    root = torch.tensor([0.0, 0.0, 0.5], device=device) + 0.05 * n(B, 3)
    f = 217.5 * size / 256
    K = torch.tensor([[f, 0, size / 2.0], [0, f, size / 2.0], [0, 0, 1]], device=device).expand(B, 3, 3).contiguous()
    return {"image": r(B, 3, size, size) - 0.5, "root_joint": root, "cam_intr": K, "corners_can": 0.2 * r(B, 8, 3) - 0.1,
            "joints_3d": 0.05 * n(B, 21, 3), "corners_3d": 0.08 * n(B, 8, 3), "joints_vis": torch.ones((B, 21), device=device),
            "corners_vis": torch.ones((B, 8), device=device), "is_synth": torch.zeros(B, device=device),
            "obj_id": torch.zeros(B, dtype=torch.int32, device=device), "persp_id": torch.zeros(B, dtype=torch.int32, device=device),
            "grasp_id": torch.zeros(B, dtype=torch.int32, device=device), "obj_idx": torch.ones(B, dtype=torch.int32, device=device),
            "obj_transf": torch.cat([torch.eye(4, device=device)[:, :3].expand(B, 4, 3), torch.cat([root, torch.ones((B, 1), device=device)], 1).unsqueeze(-1)], 2)}


def mix_batches(a: Dict[str, torch.Tensor], b: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """MixedDataset (mixed_dataset.py:32-37) at batch granularity: real-shaped samples followed by synthetic ones."""
    return {k: torch.cat([a[k], b[k]], dim=0) for k in a.keys() & b.keys()}


class ArtiBoostLoop:
    """The per-iteration loop of train/train_artiboost.py:204-227 with ArtiBoostLoader's synthesis inlined
    (artiboost_loader.py:279-340): every step draws `n_synth` CCV cells with the current weights, poses and rasterises
    them on this rank's GPU, mixes them with `n_real` real(-shaped) samples (MixedDataset: N real + SYNTH_FACTOR * N
    synthetic, yaml:2), runs the optimisation step and records the per-cell corner error; `end_epoch()` turns the
    recorded errors into the next epoch's sampling weights (all-reduced over ranks)."""

    def __init__(self, arch: nn.Module, pipe, batch_size: int = 128, synth_factor: float = 0.6, real_source=None,
                 criterion_cfg: Optional[dict] = None, lr=5e-5, grad_clip=1e-3, generator=None, use_graph: bool = False,
                 augment: bool = True, prefetch: bool = True, artiboost_cfg: Optional[dict] = None):
        """artiboost_cfg: the reference's ARTIBOOST node; read here: UPDATE_METHOD, WEIGHT_UPDATE.LOWER/UPPER,
        DIST_THRESHOLD.LOWER/UPPER, EPOCH (artiboost_loader.py:81,240-265)."""
        self.pipe, self.batch_size = pipe, batch_size
        self.prefetch, self._prefetched, self._side = prefetch, None, None
        self.augmenter = make_augmenter(pipe, generator=generator) if augment else None
        self.n_synth = int(round(batch_size * synth_factor / (1.0 + synth_factor)))
        self.n_real = batch_size - self.n_synth
        self.generator = generator
        self.real_source = real_source or (lambda n: real_shaped_batch(n, pipe.device, self.generator, pipe.renderer.width))
        self.train_step = TrainStep(arch, criterion_cfg, lr=lr, grad_clip=grad_clip, generator=generator, use_graph=use_graph)
        ab = artiboost_cfg or {}
        wu, dt = ab.get("WEIGHT_UPDATE", {}), ab.get("DIST_THRESHOLD", {})
        self.feedback = CCVFeedback(pipe.sample_weight_map.shape, pipe.device, lower=wu.get("LOWER", 0.1), upper=wu.get("UPPER", 10.0),
                                    method=ab.get("UPDATE_METHOD", "method_1"), dist_lower=dt.get("LOWER", 8.0),
                                    dist_upper=dt.get("UPPER", 16.0), n_epochs=ab.get("EPOCH", 1))
        self.epoch_idx = 0

    def make_batch(self) -> Dict[str, torch.Tensor]:
        synth = synth_to_batch(self.pipe.synthesise(self.n_synth), self.pipe, augmenter=self.augmenter)
        return mix_batches(self.real_source(self.n_real), synth) if self.n_real else synth

    def step(self, batch: Optional[Dict[str, torch.Tensor]] = None):
        """One iteration.  Without an explicit batch the loop is software-pipelined: the step consumes the batch that was
        synthesised while the previous step's graph was running, then issues the synthesis of the next one on a second
        stream -- its ~60 small kernels (and their host-side launch work) run beside the training graph instead of
        after it.  The only ordering between the two streams: synthesis k+1 starts after everything enqueued before
        step k's graph, and step k+1 waits for synthesis k+1."""
        own = batch is None
        dev = self.pipe.device
        main = torch.cuda.current_stream(dev)
        if own:
            if self._prefetched is not None:
                batch, done = self._prefetched
                main.wait_event(done)
            else:
                batch = self.make_batch()
            self._prefetched = None
        fence = torch.cuda.Event()
        fence.record(main)  # before the step is enqueued: the side stream must not wait for the step itself
        loss, preds = self.train_step(batch)
        targ = batch["corners_3d"] + batch["root_joint"].unsqueeze(1)
        self.feedback.feed(preds["corners_3d_abs"].detach(), targ, batch["obj_id"], batch["persp_id"], batch["grasp_id"],
                           batch["is_synth"])
        if own and self.prefetch:
            if self._side is None:
                # high priority: the synthesis kernels are small and must not queue behind the training graph's
                self._side = torch.cuda.Stream(dev, priority=int(os.environ.get("AB_LOOP_PRIO", "-1")))
            with torch.cuda.stream(self._side):
                self._side.wait_event(fence)
                nxt = self.make_batch()
                for t in nxt.values():
                    t.record_stream(main)  # allocated on the side stream, consumed on the main one
                done = torch.cuda.Event()
                done.record(self._side)
            self._prefetched = (nxt, done)
        return loss

    def close(self):
        """Joins the prefetch stream and destroys the captured training graph (see TrainStep.close)."""
        self._prefetched = None
        if self._side is not None:
            torch.cuda.current_stream(self.pipe.device).wait_stream(self._side)
        self.train_step.close()

    def end_epoch(self):
        self._prefetched = None  # it was drawn with the old weights
        if hasattr(self.pipe, "drop_prefetch"):
            self.pipe.drop_prefetch()
        if self._side is not None:
            torch.cuda.current_stream(self.pipe.device).wait_stream(self._side)
        self.pipe.sample_weight_map = self.feedback.step_eval(self.pipe.sample_weight_map, epoch_idx=self.epoch_idx)
        self.epoch_idx += 1
        return self.pipe.sample_weight_map
