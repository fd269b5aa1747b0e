"""Refiners (anakin/artiboost/refiner.py).  Only `NullRefine` (:118-147) is on the synthetic-benchmark path: the
GrabNet RefineNet weights the `hand_obj` refiner needs (yaml:47-50) are licensed assets that are absent."""
from typing import Callable, Dict, List, Mapping

from torch import nn

from ..manolayer import ManoLayer
from .scrambler import register


class Refiner:
    build_mapping: Mapping[str, Callable] = {}

    @staticmethod
    def build(type, *args, **kwargs):
        return Refiner.build_mapping[type](*args, **kwargs)


class _RefineNet(nn.Module):
    """Holder of the refiner's MANO layer (refiner.py:227-250 keeps it at `refine_net.mano_layer`, where
    ArtiBoostLoader picks it up, artiboost_loader.py:172)."""

    def __init__(self, n_iters=0, mano_model=None, mano_assets_root="assets/mano_v1_2"):
        super().__init__()
        self.n_iters = n_iters
        self.mano_layer = ManoLayer(rot_mode="axisang", side="right", center_idx=None, use_pca=False,
                                    flat_hand_mean=True, mano_assets_root=mano_assets_root, mano_model=mano_model)


@register(reg=Refiner.build_mapping, key="null")
class NullRefine(nn.Module):

    def __init__(self, cfg=None, mano_model=None):
        super().__init__()
        self.cfg = cfg
        self.resampled_objs = []
        self.obj_idx = {}
        self.refine_net = _RefineNet(n_iters=0, mano_model=mano_model)

    def setup(self, obj_meshes: Dict[str, object]):
        pass

    def forward(self, inp, obj_name: List[str] = None):
        hand_pose, hand_tsl = inp["hand_pose"], inp["hand_tsl"]
        mano_out = self.refine_net.mano_layer(hand_pose)  # betas=None (refiner.py:138)
        return {
            "hand_verts": mano_out.verts + hand_tsl.unsqueeze(1),
            "joints": mano_out.joints + hand_tsl.unsqueeze(1),
            "hand_pose": hand_pose,
            "hand_tsl": hand_tsl,
        }
