"""Seeded inputs shared by make_golden_refine.py (which feeds them to the REFERENCE's refiner / scrambler / pose
generator) and by the tests (which feed them to the oracle and to the CUDA path).  Nothing here is copied from the
reference: weights are random (the licensed GrabNet refinenet.pt is not redistributable), objects are the synthetic
YCB stand-ins of artiboost_b200/assets.py."""
import numpy as np

N_SAMPLE = 10000  # HORefiner.resample_obj default (refiner.py:172)
OBJ_NAMES = ["010_potted_meat_can", "006_mustard_bottle"]


def refinenet_state(seed=7, in_size=877, h=512, n=256):
    """Random _RefineNet state dict (reference parameter names, refiner.py:227-319) as numpy fp32 arrays: nn.Linear-like
    fan-in scaling, non-trivial BatchNorm running statistics, small output heads so three iterations stay near the
    initial grasp."""
    rng = np.random.RandomState(seed)
    s = {}

    def lin(name, fin, fout, gain=1.0):
        b = gain / np.sqrt(fin)
        s[name + ".weight"] = rng.uniform(-b, b, (fout, fin))
        s[name + ".bias"] = rng.uniform(-b, b, (fout,))

    def bn(name, c, mean=0.0, std=1.0):
        s[name + ".weight"] = rng.uniform(0.5, 1.5, c)
        s[name + ".bias"] = rng.normal(0, 0.1, c)
        s[name + ".running_mean"] = mean + rng.normal(0, 0.1 * std, c)
        s[name + ".running_var"] = (std * rng.uniform(0.7, 1.3, c)) ** 2

    bn("bn1", 778, mean=0.05, std=0.03)   # h2o distances are metres
    for name, fin in (("rb1", in_size), ("rb2", in_size + h), ("rb3", in_size + h)):
        lin(name + ".fc1", fin, n)
        bn(name + ".bn1", n, std=0.5)
        lin(name + ".fc2", n, h)
        bn(name + ".bn2", h, std=0.5)
        lin(name + ".fc3", fin, h)
    lin("out_p", h, 96, gain=0.3)
    lin("out_t", h, 3, gain=0.05)
    return {k: np.asarray(v, np.float32) for k, v in s.items()}


class Mesh:
    """trimesh-like stand-in with the three members HORefiner.resample_obj touches (refiner.py:173-178)."""

    def __init__(self, vertices, faces):
        self.vertices, self.faces = np.asarray(vertices), np.asarray(faces)

    def subdivide(self):
        from artiboost_b200.artiboost.refiner import subdivide_mesh
        return Mesh(*subdivide_mesh(self.vertices, self.faces))


def object_meshes(seed=0):
    from artiboost_b200 import assets
    objs = assets.make_synthetic_objects(OBJ_NAMES, seed)
    return {k: Mesh(objs[k]["vertices"], objs[k]["faces"][:, :3]) for k in OBJ_NAMES}


def refiner_inputs(seed=11, B=4):
    """Hand poses / translations placed near the object surface, object rotations, object ids."""
    rng = np.random.RandomState(seed)
    pose = rng.normal(0, 0.25, (B, 48)).astype(np.float32)
    tsl = (rng.normal(0, 0.02, (B, 3)) + [0.0, 0.0, 0.12]).astype(np.float32)
    q = rng.normal(size=(B, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    w, x, y, z = q.T
    R = np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w), 2 * (x * y + z * w),
                  1 - 2 * (x * x + z * z), 2 * (y * z - x * w), 2 * (x * z - y * w), 2 * (y * z + x * w),
                  1 - 2 * (x * x + y * y)], 1).reshape(B, 3, 3).astype(np.float32)
    obj_id = rng.randint(len(OBJ_NAMES), size=B)
    return pose, tsl, R, obj_id


def scrambler_noise(seed=13, B=6, sigma_tsl=0.01, sigma_pose=0.1):
    rng = np.random.RandomState(seed)
    n = lambda *s: rng.normal(0, 1, s).astype(np.float32)  # noqa: E731
    return {"tsl": n(B, 3) * np.float32(sigma_tsl), "splay": n(B, 4) * np.float32(sigma_pose),
            "bend5": n(B, 5) * np.float32(sigma_pose), "bend14": n(B, 14) * np.float32(sigma_pose),
            "thumb": n(B, 2) * np.float32(sigma_pose)}
