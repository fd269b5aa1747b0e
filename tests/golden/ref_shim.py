"""Makes the reference's own modules importable in the BUILD container (where /root/reference exists) so that
make_golden*.py can run them and record fixtures.  Never used on the GPU box or by the product.

Third-party packages the reference imports but that are absent here are replaced by stubs:
  * pytorch3d.transforms  -> the oracle's restatement of the pytorch3d functions (oracle/rotations.py), wrapped
                             for torch tensors.  (pytorch3d @ d049cd2e is not installable offline.)
  * manotorch.manolayer   -> a torch wrapper over oracle/mano_lbs.py (manotorch is un-vendored and unpinned).
  * everything else (termcolor, trimesh, pyrender, chamfer_distance, jax ...) -> permissive empty modules; jax.numpy
    is mapped to numpy so that anakin/postprocess/iknet/manolayer.py (the in-tree MANO LBS) runs as written.
"""
import importlib.machinery
import os
import sys
import types

import numpy as np
import torch

REF_ROOT = "/root/reference"
REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

MANO_MODEL = {"model": None}  # set by the caller before building a shim ManoLayer


class _AnyMeta(type):
    def __getattr__(cls, k):
        if k.startswith("__"):
            raise AttributeError(k)
        return cls


class _Anything(metaclass=_AnyMeta):
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        return _Anything()


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, None)
    m.__path__ = []
    m.__getattr__ = lambda k: (_ for _ in ()).throw(AttributeError(k)) if k.startswith("__") else _Anything
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def install():
    if not os.path.isdir(REF_ROOT):
        raise RuntimeError("the reference tree is only available in the build container")
    for p in (REPO, REF_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    if not hasattr(np, "float"):
        np.float = float  # view_engine.py:30 uses the alias removed in numpy 1.24
    from oracle import rotations as rot
    from oracle.mano_lbs import ManoLayer as OracleMano

    def wrap(fn):
        def g(x, *a, **k):
            return torch.from_numpy(np.ascontiguousarray(fn(x.detach().cpu().numpy()))).to(x.dtype)
        return g

    _stub("pytorch3d")
    _stub("pytorch3d.transforms",
          axis_angle_to_matrix=wrap(rot.aa_to_rotmat), axis_angle_to_quaternion=wrap(rot.axis_angle_to_quaternion),
          matrix_to_quaternion=wrap(rot.matrix_to_quaternion), quaternion_to_axis_angle=wrap(rot.quaternion_to_axis_angle),
          quaternion_to_matrix=wrap(rot.quaternion_to_matrix))

    from collections import namedtuple
    MANOOutput = namedtuple("MANOOutput", ["verts", "joints", "center_idx", "center_joint", "full_poses", "betas",
                                           "transforms_abs"])

    class ShimManoLayer(torch.nn.Module):
        def __init__(self, center_idx=None, **kw):
            super().__init__()
            self.layer = OracleMano(MANO_MODEL["model"], center_idx=center_idx, dtype=np.float32)
            self.th_faces = torch.from_numpy(self.layer.faces)

        def get_rotation_center(self, betas=None):
            return torch.from_numpy(self.layer.get_rotation_center(None if betas is None else betas.numpy()))

        def forward(self, pose_coeffs, betas=None, **kw):
            o = self.layer(pose_coeffs.numpy(), None if betas is None else betas.numpy())
            t = torch.from_numpy
            return MANOOutput(t(o.verts), t(o.joints), o.center_idx, t(o.center_joint), t(o.full_poses), t(o.betas),
                              t(o.transforms_abs))

    _stub("manotorch")
    _stub("manotorch.manolayer", ManoLayer=ShimManoLayer, MANOOutput=MANOOutput)
    _stub("manotorch.axislayer")
    _stub("manotorch.utils")
    _stub("manotorch.utils.rodrigues")
    _stub("manotorch.utils.quatutils")
    for name in ["termcolor", "trimesh", "trimesh.base", "pyrender", "pyrender.constants", "pyrender.light",
                 "pyrender.material", "pyrender.platforms", "dex_ycb_toolkit", "dex_ycb_toolkit.dex_ycb",
                 "dex_ycb_toolkit.factory", "deprecated", "deprecated.sphinx", "git", "matplotlib",
                 "matplotlib.pyplot", "matplotlib.patches", "matplotlib.colors", "matplotlib.cm", "mpl_toolkits",
                 "mpl_toolkits.mplot3d", "pytz", "chamfer_distance"]:
        if name not in sys.modules:
            _stub(name)
    sys.modules["termcolor"].colored = lambda s, *a, **k: s
    sys.modules["deprecated.sphinx"].deprecated = lambda *a, **k: (lambda f: f)
    jax = _stub("jax", jit=lambda f, *a, **k: f)
    jax.numpy = np
    sys.modules["jax.numpy"] = np
