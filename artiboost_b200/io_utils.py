"""Checkpoint / resume compatibility with the reference's on-disk layout (anakin/utils/io_utils.py:19-124,
anakin/utils/recorder.py:68-123,177-226), so that released `.pth.tar` checkpoints load into the B200 modules and
runs started here can be resumed by the reference (state_dict names are identical, tests/golden/make_golden_network.py):

  <dump>/checkpoints/checkpoint/{<ModelClass>.pth.tar, train_param.pth.tar, random_state.pkl}
  <dump>/artiboost/sample_weight/<epoch:03>_train.pkl, <dump>/artiboost/occurence_map/<epoch:03>.pkl, <dump>/artiboost/shutdown
"""
from __future__ import annotations

import os
import pickle
import random
import shutil
from collections import namedtuple
from typing import Optional

import numpy as np
import torch

RandomState = namedtuple("RandomState", ["torch_rng_state", "torch_cuda_rng_state", "torch_cuda_rng_state_all",
                                         "numpy_rng_state", "random_rng_state"])  # anakin/utils/misc.py:11-21
RandomState.__new__.__defaults__ = (None,) * len(RandomState._fields)
# The reference pickles and unpickles this tuple as `anakin.utils.misc.RandomState` (io_utils.py:33-36,54-57): a file
# written under any other module path makes its load_random_state fail (swallowed: the run resumes with a fresh RNG).
# So the class claims the reference's path, dumps go through _dump_random_state (which makes that path resolvable while
# pickle checks it) and loads go through _RefUnpickler (which resolves it without `anakin` being importable).
REF_MODULE = "anakin.utils.misc"
RandomState.__module__ = REF_MODULE


class _RefUnpickler(pickle.Unpickler):

    def find_class(self, module, name):
        if name == "RandomState" and module in (REF_MODULE, __name__):
            return RandomState
        return super().find_class(module, name)


def _dump_random_state(rs: "RandomState", f) -> None:
    import sys
    import types
    added = []
    try:
        parts = REF_MODULE.split(".")
        for i in range(1, len(parts) + 1):   # stub packages only where the real ones are not importable
            name = ".".join(parts[:i])
            if name not in sys.modules:
                try:
                    __import__(name)
                except Exception:  # noqa: BLE001
                    sys.modules[name] = types.ModuleType(name)
                    added.append(name)
        mod = sys.modules[REF_MODULE]
        had = getattr(mod, "RandomState", None)
        if had is None or had is not RandomState:
            rs = RandomState(*rs) if had is None else had(*rs)   # the real class when the reference is importable
            if had is None:
                mod.RandomState = RandomState
        pickle.dump(rs, f)
    finally:
        for name in reversed(added):
            sys.modules.pop(name, None)


def _model_list(model):
    model = model.module if hasattr(model, "module") else model
    return model.model_list


def save_states(state: dict, is_best: bool, checkpoint="checkpoint", foldname="checkpoint", snapshot=None):
    """io_utils.py:19-53.  state: {"epoch", "model_list", "optimizer", "scheduler", "random_state", ["score"]}."""
    foldname = os.path.join(checkpoint, foldname)
    os.makedirs(foldname, exist_ok=True)
    state = dict(state)
    for model in state.pop("model_list"):
        inner = model.module if hasattr(model, "module") else model
        # parameters are views into the flat training buffer (train.FlatParams): clone, so each tensor is saved on its own
        sd = {k: v.detach().clone().cpu() for k, v in inner.state_dict().items()}
        torch.save(sd, os.path.join(foldname, f"{type(model).__name__}.pth.tar"))
    with open(os.path.join(foldname, "random_state.pkl"), "wb") as f:
        _dump_random_state(state.pop("random_state"), f)
    torch.save(state, os.path.join(foldname, "train_param.pth.tar"))
    if snapshot and state["epoch"] % snapshot == 0:
        shutil.copytree(foldname, os.path.join(checkpoint, "checkpoint_{}".format(state["epoch"])))
    if is_best:
        name = f"model_best_{round(state['score'], 3)}" if "score" in state else "model_best"
        shutil.copytree(foldname, os.path.join(checkpoint, name))


def capture_random_state() -> RandomState:
    cuda = torch.cuda.is_available()
    return RandomState(torch_rng_state=torch.get_rng_state(),
                       torch_cuda_rng_state=torch.cuda.get_rng_state() if cuda else None,
                       torch_cuda_rng_state_all=torch.cuda.get_rng_state_all() if cuda else None,
                       numpy_rng_state=np.random.get_state(), random_rng_state=random.getstate())


def load_random_state(resume_path: str) -> bool:
    """io_utils.py:56-72: best effort, like the reference (a failure is reported, not raised)."""
    try:
        with open(resume_path, "rb") as f:
            rs = _RefUnpickler(f).load()
        random.setstate(rs.random_rng_state)
        np.random.set_state(rs.numpy_rng_state)
        torch.set_rng_state(rs.torch_rng_state)
        if torch.cuda.is_available() and rs.torch_cuda_rng_state is not None:
            torch.cuda.set_rng_state(rs.torch_cuda_rng_state)
            torch.cuda.set_rng_state_all(rs.torch_cuda_rng_state_all)
        return True
    except Exception as e:  # noqa: BLE001
        print(f"[io_utils] couldn't resume random state from {resume_path} ({e!r}): the run may not be reproducible")
        return False


def load_train_param(optimizer, scheduler, resume_path: str, map_location=None) -> int:
    """io_utils.py:75-96 -> epoch.  `optimizer` is a torch optimizer or train.FusedAdam (same state_dict format)."""
    try:
        parameters = torch.load(resume_path, map_location=map_location, weights_only=False)
        optimizer.load_state_dict(parameters["optimizer"])
        if scheduler is not None and parameters.get("scheduler") is not None:
            scheduler.load_state_dict(parameters["scheduler"])
        return parameters["epoch"]
    except Exception as e:
        raise ValueError(f"Couldn't resume from {resume_path}: {e!r}") from e


def load_arch(model, resume_path: str, startswith=None, strict=True, as_parallel=False, map_location=None, rank=None):
    """io_utils.py:99-124: one `<ModelClass>.pth.tar` per entry of `model.model_list`; handles the `module.` prefix of
    DataParallel checkpoints and the `startswith` sub-module filter.  map_location defaults like the reference's: "cuda"
    when `rank` is None, else that rank's device (:101-104)."""
    if map_location is None and torch.cuda.is_available():
        map_location = "cuda" if rank is None else f"cuda:{rank}"
    try:
        for m in _model_list(model):
            ckpt = torch.load(os.path.join(resume_path, f"{type(m).__name__}.pth.tar"), map_location=map_location)
            state_dict = ckpt
            first = list(ckpt.keys())[0]
            if as_parallel and "module" not in first:
                state_dict = {"module.{}".format(k): v for k, v in ckpt.items()}
            elif not as_parallel and "module" in first:
                state_dict = {".".join(k.split(".")[1:]): v for k, v in ckpt.items()}
            if startswith is not None:
                state_dict = {".".join(k.split(".")[1:]): v for k, v in state_dict.items() if k.startswith(startswith)}
            with torch.no_grad():
                m.load_state_dict(state_dict, strict=strict)  # copy_ into the existing (possibly flat-buffer) tensors
        from .models import nhwc
        nhwc.bump_params()  # packed bf16 filter copies are stale
    except Exception as e:
        raise ValueError(f"Couldn't resume from {resume_path}: {e!r}") from e


class Recorder:
    """The checkpoint / ArtiBoost-state half of anakin/utils/recorder.py (logging, tensorboard and git bookkeeping are
    out of scope)."""

    def __init__(self, exp_id: str, root_path: str = "exp", rank: Optional[int] = None):
        self.exp_id, self.rank = exp_id, rank
        self.dump_path = os.path.join(root_path, exp_id)
        if not rank:
            os.makedirs(self.dump_path, exist_ok=True)

    def record_checkpoints(self, model, optimizer, scheduler, epoch: int, snapshot: int):
        if self.rank:
            return
        path = os.path.join(self.dump_path, "checkpoints")
        os.makedirs(path, exist_ok=True)
        save_states({"epoch": epoch + 1, "model_list": _model_list(model), "optimizer": optimizer.state_dict(),
                     "scheduler": scheduler.state_dict() if scheduler is not None else None,
                     "random_state": capture_random_state()}, is_best=False, checkpoint=path, snapshot=snapshot)

    def resume_checkpoints(self, model, optimizer, scheduler, resume_path: str, resume_epoch: Optional[int] = None) -> int:
        resume_path = os.path.join(resume_path, "checkpoints", f"checkpoint_{resume_epoch}" if resume_epoch else "checkpoint")
        epoch = load_train_param(optimizer, scheduler, os.path.join(resume_path, "train_param.pth.tar"))
        load_random_state(os.path.join(resume_path, "random_state.pkl"))
        load_arch(model, resume_path, map_location=f"cuda:{self.rank}" if self.rank is not None else None)
        return epoch

    # ---- ArtiBoost sampler state (recorder.py:177-226): numpy pickles of the CCV weight / occurrence maps
    def record_artiboost_loader(self, loader, epoch: int):
        self.record_sample_weight(loader.sample_weight_map, epoch)
        self.record_sample_occurence(loader.occurence_map, epoch)
        if not getattr(loader, "use_synth", True):
            open(os.path.join(self.dump_path, "artiboost", "shutdown"), "w").close()

    def record_sample_weight(self, weight_map: torch.Tensor, epoch: int, is_train: bool = True):
        path = os.path.join(self.dump_path, "artiboost", "sample_weight")
        os.makedirs(path, exist_ok=True)
        with open(os.path.join(path, f"{epoch:0>3}_{'train' if is_train else 'val'}.pkl"), "wb") as f:
            pickle.dump(weight_map.detach().cpu().numpy().copy(), f)

    def record_sample_occurence(self, occurence_map: torch.Tensor, epoch: int, is_train: bool = True):
        path = os.path.join(self.dump_path, "artiboost", "occurence_map")
        os.makedirs(path, exist_ok=True)
        with open(os.path.join(path, f"{epoch:0>3}.pkl"), "wb") as f:
            pickle.dump(occurence_map.detach().cpu().numpy().copy(), f)

    def resume_artiboost_loader(self, loader, resume_epoch: int, resume_path: str):
        epoch = resume_epoch - 1
        with open(os.path.join(resume_path, "artiboost", "sample_weight", f"{epoch:0>3}_train.pkl"), "rb") as f:
            weight_map = torch.from_numpy(pickle.load(f))
        with open(os.path.join(resume_path, "artiboost", "occurence_map", f"{epoch:0>3}.pkl"), "rb") as f:
            occurence_map = torch.from_numpy(pickle.load(f))
        loader.sample_weight_map[:] = weight_map.to(loader.sample_weight_map.device)
        loader.occurence_map[:] = occurence_map.to(loader.occurence_map.device)
        if os.path.exists(os.path.join(resume_path, "artiboost", "shutdown")) and hasattr(loader, "synth_shutdown"):
            loader.synth_shutdown()
