"""Arch: runs the model DAG described by cfg["ARCH"] (anakin/models/arch.py:12-72)."""
from typing import Dict, List

import torch.nn as nn


class Arch(nn.Module):

    def __init__(self, cfg: Dict, model_list: List[nn.Module]):
        super().__init__()
        self._model_list = nn.ModuleList(model_list)
        self._cfg = cfg
        self.parser()

    @property
    def model_list(self) -> nn.ModuleList:
        return self._model_list

    @property
    def models_params(self):
        return [{"params": filter(lambda p: p.requires_grad, m.parameters())} for m in self._model_list]

    def parser(self):
        items = self._cfg["ARCH"]
        self.models = {}
        if isinstance(items, dict):
            items = [items]
        for i, item in enumerate(items):
            self.models[item["TYPE"]] = {"id": i, "previous": item["PREVIOUS"]}
        outdegree = [0] * len(items)
        for v in self.models.values():
            for p in v["previous"]:
                outdegree[self.models[p]["id"]] += 1
        if outdegree.count(0) != 1:
            raise Exception("Arch has multiple roots, a circle or other illegal input.!")
        self.root = items[outdegree.index(0)]["TYPE"]

    def forward(self, input: Dict):
        self.outputs = {}
        self._forward(self.root, input)
        return self.outputs

    def _forward(self, mtype: str, input: Dict):
        inputs = dict(input)
        for p in self.models[mtype]["previous"]:
            if p not in self.outputs:
                self._forward(p, input)
            inputs.update(self.outputs[p])
        self.outputs[mtype] = self._model_list[self.models[mtype]["id"]](inputs)
