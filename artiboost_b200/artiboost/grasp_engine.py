"""Grasp tables (anakin/artiboost/grasp_engine.py:13-82): per object a list of (hand_pose[48], hand_shape[10] | None,
hand_tsl[3]).  The tables are packed once into device tensors so a batch lookup is one gather."""
import os
import pickle
from typing import Dict, List

import numpy as np
import torch


class GraspEngine:

    def __init__(self, obj_grasps: Dict[str, list], obj_names: List[str], n_grasp: int = None, device="cuda"):
        self._obj_names = list(obj_names)
        self.obj_grasps = {k: obj_grasps[k] for k in self._obj_names}
        n = min(len(v) for v in self.obj_grasps.values())
        self.n_grasp = n if n_grasp is None else min(n, n_grasp)
        tab = np.zeros((len(self._obj_names), self.n_grasp, 61), np.float32)
        for oi, k in enumerate(self._obj_names):
            for gi in range(self.n_grasp):
                pose, shape, tsl = self.get_obj_grasp(k, gi)
                tab[oi, gi, :48], tab[oi, gi, 48:58], tab[oi, gi, 58:61] = pose, shape, tsl
        self.table = torch.from_numpy(tab).to(device)

    @staticmethod
    def from_dir(grasp_dir: str, obj_names: List[str], **kw):
        """Load `<grasp_dir>/<obj>.pkl` like grasp_engine.py:28-32."""
        grasps = {}
        for name in obj_names:
            with open(os.path.join(grasp_dir, name + ".pkl"), "rb") as stream:
                grasps[name] = pickle.load(stream)
        return GraspEngine(grasps, obj_names, **kw)

    @property
    def obj_names(self):
        return self._obj_names

    def has_obj(self, name: str):
        return name in self._obj_names

    def get_obj_grasp(self, obj_name: str, grasp_idx: int):
        hand_pose, hand_shape, hand_tsl = self.obj_grasps[obj_name][grasp_idx]
        if hand_shape is None or (not isinstance(hand_shape, np.ndarray) and not hand_shape):
            hand_shape = np.zeros(10)
        if (isinstance(hand_tsl, (int, float)) and hand_tsl == 0) or hand_tsl is None:
            hand_tsl = np.zeros(3)
        return hand_pose, hand_shape, hand_tsl

    def get_mapping_len(self):
        return {n: len(v) for n, v in self.obj_grasps.items()}

    def gather(self, obj_id: torch.Tensor, grasp_id: torch.Tensor):
        rows = self.table[obj_id.long(), grasp_id.long()]
        return rows[:, :48].contiguous(), rows[:, 48:58].contiguous(), rows[:, 58:61].contiguous()
