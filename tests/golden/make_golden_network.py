"""Generates tests/golden/network_*.npz by running the REFERENCE's own HybridBaseline (anakin/models/*, imported from
/root/reference through ref_shim) on CPU in fp32.  Weights are not stored (100 MB): both sides build the model under
torch.manual_seed(SEED) -- the construction order and init calls are the same, which this script verifies by comparing
the two state_dicts -- and then apply `randomise_bn` below.  Run in the build container:
    python tests/golden/make_golden_network.py
"""
import os
import sys

import numpy as np
import torch
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

ref_shim.install()

from netcfg import ARCH, DATA_PRESET, SEED, arch_cfg, make_inputs, randomise_bn  # noqa: E402


def check_cfg_matches_reference_yaml():
    with open(os.path.join(ref_shim.REF_ROOT, "config/ho3dv2_clasbased_jlol_artiboost2.yaml")) as f:
        cfg = yaml.safe_load(f)
    assert cfg["ARCH"] == ARCH, "tests/golden/netcfg.py ARCH drifted from the reference yaml"
    assert cfg["DATA_PRESET"] == DATA_PRESET, "tests/golden/netcfg.py DATA_PRESET drifted from the reference yaml"
    assert cfg["TRAIN"]["MANUAL_SEED"] == SEED


def main():
    from anakin.models.arch import Arch as RefArch
    from anakin.utils import builder as ref_builder
    import anakin.models  # noqa: F401  (registers the modules)
    import artiboost_b200.models as ours

    check_cfg_matches_reference_yaml()

    for backbone in ("ResNet34", "ResNet50"):
        arch, preset = arch_cfg(backbone)
        torch.manual_seed(SEED)
        ref = RefArch({"ARCH": arch}, ref_builder.build_arch_model_list(arch, preset_cfg=preset)).eval()
        torch.manual_seed(SEED)
        mine = ours.Arch({"ARCH": arch}, ours.build_arch_model_list(arch, preset_cfg=preset)).eval()
        sd_r, sd_m = ref.state_dict(), mine.state_dict()
        assert list(sd_r.keys()) == list(sd_m.keys()), "state_dict names differ from the reference"
        for k in sd_r:
            assert torch.equal(sd_r[k], sd_m[k]), f"seeded init differs at {k}"
        randomise_bn(ref)
        B = 2
        inp = make_inputs(B)
        with torch.no_grad():
            out = ref(inp)["HybridBaseline"]
            feats = ref.model_list[0].backbone(image=inp["image"])
            head = ref.model_list[0].hybrid_head(feature=feats["res_layer4"])
        arrs = {k: v.numpy() for k, v in out.items()}
        arrs.update(kp3d=head["kp3d"].numpy(), kp3d_confd=head["kp3d_confd"].numpy(),
                    res_layer4_mean=feats["res_layer4_mean"].numpy(),
                    res_layer1_stat=np.array([feats["res_layer1"].mean().item(), feats["res_layer1"].std().item()]),
                    res_layer4_stat=np.array([feats["res_layer4"].mean().item(), feats["res_layer4"].std().item()]),
                    n_params=np.array(sum(p.numel() for p in ref.parameters())))
        path = os.path.join(HERE, f"network_{backbone.lower()}.npz")
        np.savez_compressed(path, **arrs)
        print("wrote", path, {k: v.shape for k, v in arrs.items()})


if __name__ == "__main__":
    main()
