"""Arch: the container train_artiboost.py drives (`Arch(cfg, model_list)`, `.model_list`, `.models_params`,
`forward(batch) -> {model TYPE: outputs}`), with the interface of anakin/models/arch.py:12-72.

cfg["ARCH"] names a small DAG: every entry has a TYPE and the TYPEs whose outputs it consumes (PREVIOUS); exactly one
entry is consumed by nobody and is the one the caller asks for.  The reference walks that graph recursively on every
forward.  Here the graph is resolved ONCE, at construction, into a flat execution plan (topological order restricted to
what the sink needs), so a forward is a plain loop -- which also keeps the captured training graph free of Python
recursion state.  State-dict names are those of the reference (`_model_list.<i>...`).
"""
from typing import Dict, List, Sequence, Tuple

import torch.nn as nn


def _execution_plan(entries: Sequence[dict]) -> Tuple[List[Tuple[int, str, Tuple[str, ...]]], str]:
    """-> ([(index into model_list, TYPE, consumed TYPEs)] in execution order, sink TYPE)."""
    index = {e["TYPE"]: i for i, e in enumerate(entries)}
    if len(index) != len(entries):
        raise ValueError("Arch: duplicate TYPE in cfg['ARCH']")
    feeds = {e["TYPE"]: tuple(e.get("PREVIOUS") or ()) for e in entries}
    for t, prev in feeds.items():
        for p in prev:
            if p not in index:
                raise KeyError(f"Arch: {t} consumes unknown model {p}")
    consumed = {p for prev in feeds.values() for p in prev}
    sinks = [t for t in feeds if t not in consumed]
    if len(sinks) != 1:  # the reference's "multiple roots, a circle or other illegal input" (arch.py:39-41)
        raise ValueError(f"Arch: cfg['ARCH'] must have exactly one model that nobody consumes, found {sinks}")
    plan, state = [], {}   # state: 1 = on the current path, 2 = scheduled

    def schedule(t: str):
        if state.get(t) == 2:
            return
        if state.get(t) == 1:
            raise ValueError(f"Arch: cfg['ARCH'] has a cycle through {t}")
        state[t] = 1
        for p in feeds[t]:
            schedule(p)
        state[t] = 2
        plan.append((index[t], t, feeds[t]))

    schedule(sinks[0])
    return plan, sinks[0]


class Arch(nn.Module):

    def __init__(self, cfg: Dict, model_list: List[nn.Module]):
        super().__init__()
        self._model_list = nn.ModuleList(model_list)
        self._cfg = cfg
        entries = cfg["ARCH"]
        entries = [entries] if isinstance(entries, dict) else list(entries)
        if len(entries) != len(model_list):
            raise ValueError("Arch: cfg['ARCH'] and model_list differ in length")
        self._plan, self.root = _execution_plan(entries)
        self.models = {t: {"id": i, "previous": list(prev)} for i, t, prev in self._plan}

    @property
    def model_list(self) -> nn.ModuleList:
        return self._model_list

    @property
    def models_params(self):
        """Optimizer parameter groups, one per model (arch.py:23-26): trainable parameters only."""
        return [{"params": [p for p in m.parameters() if p.requires_grad]} for m in self._model_list]

    def forward(self, batch: Dict):
        self.outputs = outputs = {}
        for i, mtype, prev in self._plan:
            feed = dict(batch)
            for p in prev:
                feed.update(outputs[p])
            outputs[mtype] = self._model_list[i](feed)
        return outputs
