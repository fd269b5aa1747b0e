"""Host logic of SURVEY.md 8(f).4: checkpoint / resume files in the reference's layout (anakin/utils/io_utils.py,
recorder.py) and the loaders for the real assets (object_engine.py, hand_texture.py).  CPU only."""
import os
import pickle

import numpy as np
import pytest
import torch
import torch.nn as nn


class _Net(nn.Module):
    def __init__(self):
        super().__init__()
        self.conv = nn.Conv2d(3, 4, 3)
        self.bn = nn.BatchNorm2d(4)


class HybridBaseline(_Net):  # the file name follows the class name (io_utils.py:29-31)
    pass


class _Arch(nn.Module):
    def __init__(self):
        super().__init__()
        self.model_list = nn.ModuleList([HybridBaseline()])


def test_random_state_pickle_crosses_to_the_reference_and_back(tmp_path, monkeypatch):
    """random_state.pkl must name the class the reference imports (`anakin.utils.misc.RandomState`, io_utils.py:16,54-57):
    a file written here loads in an interpreter that only has the reference's module, and a file the reference wrote loads
    here without `anakin` being importable."""
    import collections
    import pickletools
    import sys
    import types
    from artiboost_b200 import io_utils
    np.random.seed(11)
    path = tmp_path / "random_state.pkl"
    with open(path, "wb") as f:
        io_utils._dump_random_state(io_utils.capture_random_state(), f)
    draw = np.random.rand()
    assert "anakin" not in sys.modules                      # the stub packages used while dumping are gone again
    names = [op[1] for op in pickletools.genops(path.read_bytes()) if isinstance(op[1], str)]
    assert "anakin.utils.misc" in names and "artiboost_b200.io_utils" not in names
    np.random.seed(12)
    assert io_utils.load_random_state(str(path)) and np.random.rand() == draw      # ours -> ours, no `anakin` anywhere
    # the reference's side: only its own module and its own namedtuple exist (anakin/utils/misc.py:11-21)
    ref_cls = collections.namedtuple("RandomState", io_utils.RandomState._fields)
    for name in ("anakin", "anakin.utils", "anakin.utils.misc"):
        monkeypatch.setitem(sys.modules, name, types.ModuleType(name))
    ref_cls.__module__ = "anakin.utils.misc"
    sys.modules["anakin.utils.misc"].RandomState = ref_cls
    rs = pickle.loads(path.read_bytes())                    # what the reference's load_random_state does
    assert type(rs) is ref_cls and rs.numpy_rng_state[0] == "MT19937"
    ref_path = tmp_path / "ref_random_state.pkl"
    ref_path.write_bytes(pickle.dumps(ref_cls(*rs)))       # what the reference's save_states writes
    for name in ("anakin", "anakin.utils", "anakin.utils.misc"):
        monkeypatch.delitem(sys.modules, name)
    np.random.seed(13)
    assert io_utils.load_random_state(str(ref_path)) and np.random.rand() == draw  # reference -> ours


def test_checkpoint_round_trip_in_reference_layout(tmp_path):
    from artiboost_b200 import io_utils
    torch.manual_seed(0)
    arch = _Arch()
    opt = torch.optim.Adam(arch.parameters(), lr=5e-5)
    sched = torch.optim.lr_scheduler.StepLR(opt, 10)
    arch.model_list[0].conv(torch.randn(1, 3, 8, 8)).sum().backward()
    opt.step()
    rec = io_utils.Recorder("exp0", root_path=str(tmp_path))
    np.random.seed(3)
    rec.record_checkpoints(arch, opt, sched, epoch=4, snapshot=5)
    ck = tmp_path / "exp0" / "checkpoints" / "checkpoint"
    assert sorted(os.listdir(ck)) == ["HybridBaseline.pth.tar", "random_state.pkl", "train_param.pth.tar"]
    assert (tmp_path / "exp0" / "checkpoints" / "checkpoint_5").is_dir()       # epoch + 1 = 5 is a snapshot epoch
    sd = torch.load(ck / "HybridBaseline.pth.tar")
    assert set(sd) == set(arch.model_list[0].state_dict())                     # plain state_dict, reference key names
    draw = np.random.rand()
    # resume into a fresh model / optimizer; a DataParallel-style "module." prefix is stripped (io_utils.py:108-111)
    torch.save({"module." + k: v for k, v in sd.items()}, ck / "HybridBaseline.pth.tar")
    arch2 = _Arch()
    opt2 = torch.optim.Adam(arch2.parameters(), lr=1.0)
    sched2 = torch.optim.lr_scheduler.StepLR(opt2, 10)
    epoch = rec.resume_checkpoints(arch2, opt2, sched2, str(tmp_path / "exp0"))
    assert epoch == 5
    for a, b in zip(arch.state_dict().values(), arch2.state_dict().values()):
        assert torch.equal(a, b)
    assert opt2.param_groups[0]["lr"] == 5e-5
    assert np.random.rand() == draw                                            # numpy RNG state restored
    with pytest.raises(ValueError):
        io_utils.load_arch(arch2, str(tmp_path / "nowhere"))


def test_fused_adam_state_dict_is_torch_adam_format():
    """train_param.pth.tar written by the reference's torch.optim.Adam loads into FusedAdam and back (CPU tensors: the
    state_dict plumbing has no kernel call)."""
    from artiboost_b200.train import FlatParams, FusedAdam
    torch.manual_seed(1)
    net = _Net()
    ref_opt = torch.optim.Adam(net.parameters(), lr=3e-4, betas=(0.8, 0.99), eps=1e-7)
    for _ in range(3):
        ref_opt.zero_grad()
        net.bn(net.conv(torch.randn(2, 3, 8, 8))).pow(2).sum().backward()
        ref_opt.step()
    sd = ref_opt.state_dict()
    net2 = _Net()
    fa = FusedAdam(FlatParams(net2), lr=1.0)
    fa.load_state_dict(sd)
    assert (fa.lr, fa.betas, fa.eps) == (3e-4, (0.8, 0.99), 1e-7) and fa.step_count == 3
    np.testing.assert_allclose(fa.state[1].item(), 1 - 0.8 ** 3, rtol=1e-6)
    back = fa.state_dict()
    assert back["param_groups"][0]["params"] == sd["param_groups"][0]["params"]
    for i, st in sd["state"].items():
        assert torch.equal(back["state"][i]["exp_avg"], st["exp_avg"])
        assert torch.equal(back["state"][i]["exp_avg_sq"], st["exp_avg_sq"])
        assert float(back["state"][i]["step"]) == float(st["step"])
    opt3 = torch.optim.Adam(_Net().parameters())
    opt3.load_state_dict(back)                                                 # and the reference's optimizer accepts ours
    with pytest.raises(ValueError):
        fa.load_state_dict({"state": {}, "param_groups": [dict(sd["param_groups"][0], params=[0])]})


def test_artiboost_sampler_state_round_trip(tmp_path):
    from types import SimpleNamespace
    from artiboost_b200 import io_utils
    rec = io_utils.Recorder("exp1", root_path=str(tmp_path))
    g = torch.Generator().manual_seed(0)
    loader = SimpleNamespace(sample_weight_map=torch.rand((4, 288, 50), generator=g), occurence_map=torch.rand((4, 288, 50), generator=g) > 0.5,
                             use_synth=False, shut=False)
    loader.synth_shutdown = lambda: setattr(loader, "shut", True)
    rec.record_artiboost_loader(loader, epoch=6)
    root = tmp_path / "exp1" / "artiboost"
    with open(root / "sample_weight" / "006_train.pkl", "rb") as f:            # recorder.py:183-191: a pickled ndarray
        assert isinstance(pickle.load(f), np.ndarray)
    fresh = SimpleNamespace(sample_weight_map=torch.ones((4, 288, 50)), occurence_map=torch.zeros((4, 288, 50), dtype=torch.bool),
                            shut=False)
    fresh.synth_shutdown = lambda: setattr(fresh, "shut", True)
    rec.resume_artiboost_loader(fresh, resume_epoch=7, resume_path=str(tmp_path / "exp1"))
    assert torch.equal(fresh.sample_weight_map, loader.sample_weight_map)
    assert torch.equal(fresh.occurence_map, loader.occurence_map) and fresh.shut


OBJ = """mtllib m.mtl
v 0 0 0
v 1 0 0
v 1 2 0
v 0 2 0
v 0 0 3
vt 0.1 0.1
vt 0.9 0.1
vt 0.9 0.9
vt 0.1 0.9
f 1/1 2/2 3/3 4/4
f 1/1 2/2 5/3
f -1 -2 -3
"""


def test_obj_loader_and_object_engines(tmp_path):
    from PIL import Image
    from artiboost_b200 import assets_real as ar
    d = tmp_path / "YCB" / "003_cracker_box"
    d.mkdir(parents=True)
    (d / "ds_textured.obj").write_text(OBJ)
    (d / "m.mtl").write_text("newmtl a\nmap_Kd tex.png\n")
    tex = np.zeros((10, 10, 3), np.uint8)
    tex[:5, :5] = [255, 0, 0]      # top-left of the image = (u < .5, v > .5)
    tex[5:, :5] = [0, 255, 0]      # bottom-left = (u < .5, v < .5)
    tex[5:, 5:] = [0, 0, 255]
    Image.fromarray(tex).save(d / "tex.png")
    m = ar.load_obj(str(d / "ds_textured.obj"))
    assert m.vertices.shape == (5, 3) and m.faces.tolist() == [[0, 1, 2], [0, 2, 3], [0, 1, 4], [4, 3, 2]]  # fan + negative ids
    np.testing.assert_allclose(m.uv[0], [0.1, 0.1])                                   # first pairing wins
    np.testing.assert_allclose(m.uv[4], [0.9, 0.9])
    mesh = ar.load_textured_mesh(str(d / "ds_textured.obj"))
    assert mesh.visual.vertex_colors[0].tolist() == [0, 255, 0]                       # (.1, .1): bottom-left texel
    assert mesh.visual.vertex_colors[3].tolist() == [255, 0, 0]                       # (.1, .9): top-left texel
    assert mesh.visual.vertex_colors[1].tolist() == [0, 0, 255]
    corners = {"003_cracker_box": np.array([[x, y, z] for x in (0, 1) for y in (0, 2) for z in (0, 3)], float)}
    with open(tmp_path / "corners.pkl", "wb") as f:
        pickle.dump(corners, f)
    objs = ar.load_ho3d_objects(["003_cracker_box"], obj_root=str(tmp_path / "YCB"), corner_file=str(tmp_path / "corners.pkl"))
    o = objs["003_cracker_box"]
    # object_engine.py:48-60: y and z flipped into the camera convention, then bbox-centred; corners follow
    np.testing.assert_allclose(o["vertices"].min(0), [-0.5, -1.0, -1.5])
    np.testing.assert_allclose(o["vertices"].max(0), [0.5, 1.0, 1.5])
    np.testing.assert_allclose(np.abs(o["corners_can"]), np.tile([0.5, 1.0, 1.5], (8, 1)))
    np.testing.assert_allclose(o["vertices"][4], [-0.5, 1.0, -1.5])                   # (0,0,3) -> (0,0,-3) -> centred
    dd = tmp_path / "Dex" / "003_cracker_box"
    dd.mkdir(parents=True)
    (dd / "textured_simple.obj").write_text(OBJ.replace("mtllib m.mtl\n", ""))
    o2 = ar.load_dexycb_objects(["003_cracker_box"], obj_root=str(tmp_path / "Dex"))["003_cracker_box"]
    np.testing.assert_allclose(o2["vertices"].min(0), [-0.5, -1.0, -1.5])             # no flip for DexYCB (object_engine.py:80-81)
    np.testing.assert_allclose(o2["vertices"][4], [-0.5, -1.0, 1.5])
    assert o2["colors"].shape == (5, 3) and (o2["colors"] == 77).all()                # untextured: pyrender's default grey
    np.testing.assert_allclose(sorted(map(tuple, o2["corners_can"])), sorted(map(tuple, np.abs(o["corners_can"]) * np.array(
        [[sx, sy, sz] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)]))))
