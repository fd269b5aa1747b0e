"""Object meshes (anakin/artiboost/object_engine.py:12-91): bbox-centred vertices + 8 canonical corners per object."""
from typing import Dict, List

import numpy as np
import torch

from .renderer import make_mesh


class ObjEngine:

    def __init__(self, objects: Dict[str, dict], obj_names: List[str], device="cuda"):
        """objects: name -> {"vertices", "faces", "colors", "corners_can"} (assets.make_synthetic_objects layout),
        already bbox-centred like object_engine.py:50-54."""
        self._obj_names = list(obj_names)
        self.obj_meshes = {k: objects[k] for k in self._obj_names}
        self.obj_trimeshes_mapping = {k: make_mesh(objects[k]["vertices"], objects[k]["faces"], objects[k]["colors"])
                                      for k in self._obj_names}
        self.corners_can = torch.from_numpy(
            np.stack([np.asarray(objects[k]["corners_can"], np.float32) for k in self._obj_names])).to(device)

    @staticmethod
    def build(dataset_type: str, query_obj: List[str], device="cuda", **paths):
        """object_engine.py:21-27 on the real assets (artiboost_b200/assets_real.py); `paths` override the reference's
        hard-coded locations (obj_root=..., corner_file=...)."""
        from .. import assets_real
        if dataset_type == "HO3D":
            return ObjEngine(assets_real.load_ho3d_objects(query_obj, **paths), query_obj, device=device)
        elif dataset_type == "DexYCB":
            return ObjEngine(assets_real.load_dexycb_objects(query_obj, **paths), query_obj, device=device)
        raise NotImplementedError(dataset_type)

    @property
    def obj_names(self):
        return self._obj_names

    def get_obj_verts_can(self, name):
        return np.asarray(self.obj_meshes[name]["vertices"])

    def get_obj_corners_can(self, name):
        return np.asarray(self.obj_meshes[name]["corners_can"])
