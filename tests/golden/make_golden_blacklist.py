"""Records tests/golden/blacklist.npz by running the reference's own ArtiBoostLoader._construct_blacklist_map
(anakin/artiboost/artiboost_loader.py:415-500) -- unbound, on a stand-in `self` that carries the reference's ViewEngine and
a GraspEngine-like lookup over the synthetic grasp tables -- with the torch.rand draws of get_perspective_from_id replayed
into an explicit array.  Build container only (needs /root/reference):  python tests/golden/make_golden_blacklist.py
The third-party rotation underneath aa_to_rotmat (pytorch3d) is the oracle shim of ref_shim.py: the pin covers the
reference's composition (wrist rotation, view alignment, back direction, threshold)."""
import os
import sys
import tempfile
from types import SimpleNamespace

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

ref_shim.install()
from artiboost_b200 import assets  # noqa: E402


def main():
    from anakin.artiboost import artiboost_loader as AL
    from anakin.artiboost.view_engine import ViewEngine
    names = list(assets.HO3D_TRAIN_OBJS)[:2]
    objects = assets.make_synthetic_objects(names, seed=0)
    n_grasp, u_bins, th_bins = 6, 4, 6
    grasps = assets.make_synthetic_grasps(objects, n_grasp, seed=0)
    ve = ViewEngine({"PERSP_U_BINS": u_bins, "PERSP_THETA_BINS": th_bins, "CAMERA_Z_RANGE": [0.45, 0.55]})
    n_persp = u_bins * th_bins

    class GE:  # the two methods _construct_blacklist_map uses (grasp_engine.py:47-53)
        def get_obj_grasp(self, obj_name, gi):
            return grasps[obj_name][gi]

    # replay: the loop calls get_view once per (o, v, g) in itertools.product order; each call draws torch.rand(1) twice
    # (u, theta), np.random.rand() once (roll) and Uniform.sample once
    seed = 7
    torch.manual_seed(seed)
    n_cells = len(names) * n_persp * n_grasp
    rand2 = np.zeros((n_cells, 2), np.float32)
    for i in range(n_cells):
        rand2[i, 0], rand2[i, 1] = float(torch.rand(1)), float(torch.rand(1))
        torch.rand(())  # Uniform(camera_z).sample()
    torch.manual_seed(seed)
    fake = SimpleNamespace(obj_engine=SimpleNamespace(obj_names=names), grasp_engine=GE(), view_engine=ve)

    class Bar(list):
        def set_description(self, *a, **k):
            pass

    AL.etqdm = lambda it, **kw: Bar(it)
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:  # the function caches under ./common/cache
        os.chdir(tmp)
        try:
            # get_perspective_from_id expects a tensor id (torch.div); the loop passes python ints -> wrap
            orig = ve.get_view
            ve.get_view = lambda vi: orig(torch.tensor(vi))
            bl = AL.ArtiBoostLoader._construct_blacklist_map(fake, len(names), n_persp, n_grasp, True)
        finally:
            os.chdir(cwd)
    pose = np.stack([[np.asarray(grasps[n][g][0], np.float32) for g in range(n_grasp)] for n in names])
    np.savez_compressed(os.path.join(HERE, "blacklist.npz"), names=np.array(names), n_grasp=n_grasp, u_bins=u_bins,
                        theta_bins=th_bins, rand2=rand2.reshape(len(names), n_persp, n_grasp, 2), blacklist=bl.numpy(),
                        hand_pose=pose)
    print("wrote blacklist.npz: blacklisted", int(bl.sum()), "/", bl.numel())


if __name__ == "__main__":
    main()
