"""Generates tests/golden/symloss.npz by running the reference's own SymCornerLoss
(anakin/criterions/symcornerloss.py:18-108 over bop_toolkit/bop_misc.py:18-65) on a synthetic models_info table:
object 1 without symmetries, 2 with one discrete symmetry, 3 with a continuous one.
Run in the build container:  python tests/golden/make_golden_symloss.py"""
import json
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

ref_shim.install()
from anakin.criterions.symcornerloss import SymCornerLoss  # noqa: E402

MODEL_INFO = {
    "1": {"diameter": 172.0},
    "2": {"diameter": 250.0, "symmetries_discrete": [[-1, 0, 0, 4.0, 0, -1, 0, -2.0, 0, 0, 1, 0.5, 0, 0, 0, 1]]},
    "3": {"diameter": 120.0, "symmetries_continuous": [{"axis": [0, 0, 1], "offset": [1.5, -2.0, 0.0]}],
          "symmetries_discrete": [[1, 0, 0, 0, 0, -1, 0, 0, 0, 0, -1, 0, 0, 0, 0, 1]]},
}


def main():
    path = os.path.join(tempfile.mkdtemp(), "models_info.json")
    json.dump(MODEL_INFO, open(path, "w"))
    rng = np.random.RandomState(11)
    B = 9
    obj_idx = torch.tensor([1, 2, 3, 3, 2, 1, 3, 2, 1])
    cc = torch.from_numpy(rng.uniform(-0.1, 0.1, size=(B, 8, 3)).astype(np.float32))
    transf = np.tile(np.eye(4, dtype=np.float32), (B, 1, 1))
    for i in range(B):
        q = np.linalg.qr(rng.normal(size=(3, 3)))[0]
        transf[i, :3, :3] = q * np.sign(np.linalg.det(q))
        transf[i, :3, 3] = [rng.uniform(-0.1, 0.1), rng.uniform(-0.1, 0.1), rng.uniform(0.4, 0.6)]
    transf = torch.from_numpy(transf)
    vis = torch.from_numpy((rng.uniform(size=(B, 8)) > 0.2).astype(np.float32))
    pred = torch.einsum("bij,bkj->bki", transf[:, :3, :3], cc) + transf[:, None, :3, 3] + torch.from_numpy(
        rng.normal(0, 0.01, size=(B, 8, 3)).astype(np.float32))
    out = {}
    for flag in (False, True):
        loss = SymCornerLoss(LAMBDA_SYM_CORNERS_3D=1.0, MODEL_INFO_PATH=path, MAX_SYM_DISC_STEP=0.05, USE_HO3D_YCB=flag)
        final, parts = loss({"corners_3d_abs": pred}, {"obj_idx": obj_idx, "corners_can": cc, "obj_transf": transf, "corners_vis": vis})
        out[f"loss_ho3d{int(flag)}"] = parts["sym_corners_3d_loss"].numpy()
        out[f"R{int(flag)}"], out[f"t{int(flag)}"] = loss.R.numpy(), loss.t.numpy()
    np.savez_compressed(os.path.join(HERE, "symloss.npz"), model_info=json.dumps(MODEL_INFO), obj_idx=obj_idx.numpy(), corners_can=cc.numpy(),
                        obj_transf=transf.numpy(), corners_vis=vis.numpy(), pred=pred.numpy(), **out)
    print("wrote symloss.npz", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
