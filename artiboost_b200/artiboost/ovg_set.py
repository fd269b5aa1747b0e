"""Object-Viewpoint-Grasp sampler (anakin/artiboost/ovg_set.py:38-178), on device.

`update` keeps the reference's signature and semantics (train mode: categorical draw with replacement over the
flat weight map, :112-114; val mode: uniform over non-blacklisted cells without replacement, :107-119), and
`get_batch` replaces the `__getitem__` + DataLoader collate of :134-159 with one gather for a whole index range."""
import ctypes as C

import torch

from .. import lib


class OVGSet:

    def __init__(self, obj_engine, grasp_engine, view_engine, config_len_train: int, config_len_val: int,
                 n_grasp: int, blacklist_map: torch.Tensor, device="cuda", generator=None):
        self.obj_engine, self.grasp_engine, self.view_engine = obj_engine, grasp_engine, view_engine
        self.config_len_train, self.config_len_val = config_len_train, config_len_val
        self.train_mode = True
        self.n_obj = len(obj_engine.obj_names)
        self.n_grasp = n_grasp
        self.n_persp_center = view_engine.n_persp_center
        self.n_all_choices = self.n_obj * self.n_persp_center * self.n_grasp
        if self.n_all_choices < self.config_len_val:
            self.config_len_val = self.n_all_choices  # capped like ovg_set.py:73-77
        self.device = torch.device(device)
        self.blacklist_map = blacklist_map.to(self.device)
        self.generator = generator
        self.sampled_idx_tensor = self.sampled_obj_idx = self.sampled_persp_idx = self.sampled_grasp_idx = None

    def __len__(self):
        return self.config_len_train if self.train_mode else self.config_len_val

    def update_len(self, config_len_train=None, config_len_val=None):
        if config_len_train is not None:
            self.config_len_train = config_len_train
        if config_len_val is not None:
            self.config_len_val = config_len_val

    def train(self):
        self.train_mode = True

    def val(self):
        self.train_mode = False

    @torch.no_grad()
    def update(self, global_sample_weight_map: torch.Tensor, global_occurence_map: torch.Tensor, uniforms=None):
        """-> (this_sample_weight_map, global_occurence_map), both on the sampler's device."""
        dev = self.device
        shape = (self.n_obj, self.n_persp_center, self.n_grasp)
        if self.train_mode:
            w = global_sample_weight_map.detach().to(dev).float().clone()
            n = self.config_len_train
        else:
            w = torch.ones(shape, device=dev)
            w[self.blacklist_map] = 0.0
            n = self.config_len_val
        w = w.contiguous()
        assert tuple(w.shape) == shape
        occ = torch.zeros(shape, dtype=torch.int32, device=dev)
        if self.train_mode:
            if uniforms is None:
                uniforms = torch.rand(n, device=dev, generator=self.generator)
            uniforms = uniforms.to(dev).float().contiguous()
            cdf = torch.empty(w.numel(), dtype=torch.float64, device=dev)
            o = torch.empty(n, dtype=torch.int32, device=dev)
            p, g = torch.empty_like(o), torch.empty_like(o)
            with torch.cuda.device(dev):
                rc = lib.load().ab_ccv_sample(lib.ptr(w), *shape, lib.ptr(uniforms), n, lib.ptr(cdf), lib.ptr(o),
                                              lib.ptr(p), lib.ptr(g), lib.ptr(occ), lib.stream_ptr(dev))
            lib.check(rc, "ab_ccv_sample")
            self.sampled_obj_idx, self.sampled_persp_idx, self.sampled_grasp_idx = o, p, g
            self.sampled_idx_tensor = (o.long() * shape[1] + p.long()) * shape[2] + g.long()
        else:
            # without replacement: once per epoch of validation, off the hot path -> torch.multinomial as the reference
            idx = torch.multinomial(w.reshape(-1), num_samples=n, replacement=False, generator=self.generator)
            self.sampled_idx_tensor = idx
            self.sampled_obj_idx = torch.div(idx, shape[1] * shape[2], rounding_mode="floor").int()
            self.sampled_persp_idx = (torch.div(idx, shape[2], rounding_mode="floor") % shape[1]).int()
            self.sampled_grasp_idx = (idx % shape[2]).int()
            occ.view(-1).index_add_(0, idx, torch.ones_like(idx, dtype=torch.int32))
        global_occurence_map = global_occurence_map.to(dev) | (occ > 0)
        return w, global_occurence_map

    @torch.no_grad()
    def get_batch(self, start: int, stop: int, rand4=None):
        """The collated `synth_extend` dict of ovg_set.py:143-157 for samples [start, stop), device tensors."""
        sl = slice(start, stop)
        obj_id, persp_id, grasp_id = self.sampled_obj_idx[sl], self.sampled_persp_idx[sl], self.sampled_grasp_idx[sl]
        hand_pose, hand_shape, hand_tsl = self.grasp_engine.gather(obj_id, grasp_id)
        rot, free, zoff = self.view_engine.get_view_batch(persp_id, rand4=rand4, generator=self.generator)
        return {
            "index": torch.arange(start, start + obj_id.shape[0], device=self.device),
            "obj_id": obj_id, "obj_name": None, "persp_id": persp_id, "grasp_id": grasp_id,
            "hand_pose": hand_pose, "hand_shape": hand_shape, "hand_tsl": hand_tsl,
            "persp_rotmat": rot, "camera_free_transf": free, "z_offset": zoff,
        }

    @staticmethod
    def row_col_calc(tidx, n_row, n_col):
        return (torch.div(tidx, n_row * n_col, rounding_mode="floor"),
                torch.div(tidx, n_col, rounding_mode="floor") % n_row, tidx % n_col)
