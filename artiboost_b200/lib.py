"""ctypes binding of the C-ABI library (include/artiboost_b200.h).  There is no CPU fallback: if the sm_100a
library is missing or a call fails, this raises."""
from __future__ import annotations

import contextlib
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libartiboost_b200.so")

c_float_p = C.POINTER(C.c_float)
c_i32_p = C.POINTER(C.c_int32)
c_u8_p = C.POINTER(C.c_uint8)


class ManoModelStruct(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("v_template", "shapedirs_t", "posedirs_t", "j_template", "j_shapedirs",
                                          "weights")]


class PatchTableStruct(C.Structure):
    _fields_ = [("n_mesh", C.c_int32), ("patch_off_host", c_i32_p), ("pos", C.c_void_p), ("vid", C.c_void_p),
                ("face", C.c_void_p), ("prim", C.c_void_p), ("bound", C.c_void_p)]


class SceneStruct(C.Structure):
    _fields_ = [("n_obj", C.c_int32), ("obj_verts", C.c_void_p), ("obj_faces", C.c_void_p), ("obj_colors", C.c_void_p),
                ("obj_vert_off_host", c_i32_p), ("obj_face_off_host", c_i32_p),
                ("n_hand_verts", C.c_int32), ("n_hand_faces", C.c_int32), ("n_hand_tex", C.c_int32),
                ("hand_faces", C.c_void_p), ("hand_colors", C.c_void_p), ("bgs", C.c_void_p),
                ("n_bg", C.c_int32), ("bg_h", C.c_int32), ("bg_w", C.c_int32), ("bg_channels", C.c_int32),
                ("obj_patches", C.POINTER(PatchTableStruct)), ("hand_patches", C.POINTER(PatchTableStruct))]


class SynthSpaceStruct(C.Structure):
    _fields_ = [("n_obj", C.c_int32), ("n_persp", C.c_int32), ("n_grasp", C.c_int32), ("u_bins", C.c_int32),
                ("theta_bins", C.c_int32), ("z_min", C.c_float), ("z_max", C.c_float), ("grasp_table", C.c_void_p),
                ("tsl_sigma", C.c_float), ("pose_sigma", C.c_float), ("n_hand_tex", C.c_int32), ("light_lo", C.c_float),
                ("light_hi", C.c_float), ("n_bg", C.c_int32), ("bg_h", C.c_int32), ("bg_w", C.c_int32), ("width", C.c_int32),
                ("height", C.c_int32)]


class CameraStruct(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float),
                ("cy", C.c_float), ("znear", C.c_float), ("cull_backface", C.c_int32), ("ambient", C.c_float),
                ("diffuse", C.c_float), ("bg_r", C.c_int32), ("bg_g", C.c_int32), ("bg_b", C.c_int32)]


class AugmentCfgStruct(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("raw_w", "raw_h", "out_w", "out_h", "center_idx", "crop_model", "full_image", "aug")] + \
               [("bbox_expand_ratio", C.c_float), ("center_jit", C.c_float), ("scale_jit", C.c_float), ("K", C.c_float * 9)]


class TailCfgStruct(C.Structure):
    _fields_ = [("batch", C.c_int32), ("center_idx", C.c_int32)] + \
               [(n, C.c_float) for n in ("inp_w", "inp_h", "img_w", "img_h", "depth_range", "w_joints", "w_corners", "w_joint_ord",
                                         "w_part_ord", "w_scene_ord", "w_sym")] + \
               [(n, C.c_int32) for n in ("n_views_hand", "n_pairs_joint", "n_pairs_part", "n_views_scene", "n_pairs_scene", "n_sym",
                                         "sym_ho3d")]


class WgradMapStruct(C.Structure):
    _fields_ = [("row_div", C.c_int32), ("col_div", C.c_int32), ("col_lo_valid", C.c_int32), ("s_row_hi", C.c_int64),
                ("s_row_lo", C.c_int64), ("s_col_hi", C.c_int64), ("s_col_lo", C.c_int64)]


BIG = 1 << 30


def wgrad_map(row_div=BIG, col_div=BIG, col_lo_valid=BIG, s_row_hi=0, s_row_lo=0, s_col_hi=0, s_col_lo=1):
    return WgradMapStruct(row_div, col_div, col_lo_valid, s_row_hi, s_row_lo, s_col_hi, s_col_lo)


EXPORTS = {
    # name: (restype, argtypes)
    "ab_version": (C.c_int, []),
    "ab_last_error": (C.c_char_p, []),
    "ab_launch_count": (C.c_uint64, []),
    "ab_profile_enable": (C.c_int, [C.c_int]),
    "ab_profile_collect": (C.c_int, [C.POINTER(C.c_double), C.POINTER(C.c_int64), C.c_int]),
    "ab_tail_losses_workspace_bytes": (C.c_uint64, [C.c_int]),
    "ab_tail_losses": (C.c_int, [C.POINTER(TailCfgStruct)] + [C.c_void_p] * 30),
    "ab_mano_forward": (C.c_int, [C.POINTER(ManoModelStruct), C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ab_ccv_sample": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ab_view_from_id": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_void_p, C.c_void_p]),
    "ab_ccv_cdf": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "ab_synth_draw": (C.c_int, [C.POINTER(SynthSpaceStruct), C.c_void_p, C.c_int, C.c_uint64, C.c_uint64] + [C.c_void_p] * 18),
    "ab_ccv_blacklist": (C.c_int, [C.POINTER(SynthSpaceStruct), C.c_void_p, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ab_pose_generate_workspace_bytes": (C.c_uint64, [C.c_int]),
    "ab_pose_generate": (C.c_int, [C.POINTER(ManoModelStruct), C.c_int] + [C.c_void_p] * 13),
    "ab_pose_prelude": (C.c_int, [C.POINTER(ManoModelStruct), C.c_int] + [C.c_void_p] * 14),
    "ab_chamfer_nn_workspace_bytes": (C.c_uint64, [C.c_int, C.c_int]),
    "ab_chamfer_nn": (C.c_int, [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ab_chamfer_nn_grouped": (C.c_int, [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                                        C.c_void_p]),
    "ab_linear_f32": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p,
                                C.c_void_p, C.c_int64, C.c_int, C.c_float, C.c_void_p, C.c_int64, C.c_void_p]),
    "ab_refine_encode": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "ab_refine_decode": (C.c_int, [C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_void_p]),
    "ab_scramble_anatomical": (C.c_int, [C.c_int] + [C.c_void_p] * 9),
    "ab_patch_capacity": (C.c_int, [C.c_int]),
    "ab_build_patches_host": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int32)]),
    "ab_render_workspace_bytes": (C.c_uint64, [C.POINTER(SceneStruct), C.POINTER(CameraStruct), C.c_int]),
    "ab_render_batch": (C.c_int, [C.POINTER(SceneStruct), C.POINTER(CameraStruct), C.c_int, C.c_int, C.c_void_p,
                                  C.c_void_p, C.c_void_p, c_i32_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ab_gemm_bf16": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p,
                               C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p,
                               C.c_void_p, C.c_void_p]),
    "ab_conv_stat_rows": (C.c_int, [C.c_int] * 9),
    "ab_conv_bf16_nhwc": (C.c_int, [C.c_void_p] + [C.c_int] * 4 + [C.c_void_p] + [C.c_int] * 5 + [C.c_void_p, C.c_int64,
                                    C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p,
                                    C.c_void_p, C.c_void_p]),
    "ab_wgrad_workspace_bytes": (C.c_uint64, [C.c_int, C.c_int, C.c_int]),
    "ab_wgrad_bf16": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p,
                                C.POINTER(WgradMapStruct), C.c_void_p, C.c_void_p]),
    "ab_conv_wgrad_bf16_nhwc": (C.c_int, [C.c_void_p] + [C.c_int] * 4 + [C.c_void_p] + [C.c_int] * 5 + [C.c_void_p, C.c_int,
                                                                                                   C.c_void_p, C.c_void_p]),
    "ab_pack_conv_filters": (C.c_int, [C.c_void_p] + [C.c_int] * 6 + [C.c_void_p, C.c_void_p, C.c_void_p]),
    "ab_image_to_nhwc": (C.c_int, [C.c_void_p] + [C.c_int] * 5 + [C.c_void_p, C.c_void_p]),
    "ab_im2col_nhwc": (C.c_int, [C.c_void_p] + [C.c_int] * 9 + [C.c_void_p, C.c_void_p]),
    "ab_maxpool3x3s2_nhwc": (C.c_int, [C.c_void_p] + [C.c_int] * 4 + [C.c_void_p, C.c_void_p, C.c_void_p]),
    "ab_maxpool3x3s2_affine_nhwc": (C.c_int, [C.c_void_p] + [C.c_int] * 4 + [C.c_void_p] * 5),
    "ab_avgpool_nhwc": (C.c_int, [C.c_void_p] + [C.c_int] * 3 + [C.c_void_p, C.c_void_p, C.c_void_p]),
    "ab_deconv4x4s2_col2im": (C.c_int, [C.c_void_p] + [C.c_int] * 4 + [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                                                      C.c_void_p, C.c_void_p]),
    "ab_head_decode": (C.c_int, [C.c_void_p] + [C.c_int] * 5 + [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ab_col_stats": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ab_bn_finalize": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_float, C.c_float]
                       + [C.c_void_p] * 7),
    "ab_bn_apply": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "ab_bn_bwd_reduce": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ab_bn_bwd_apply": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ab_affine_relu_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                     C.c_void_p]),
    "ab_maxpool3x3s2_bwd": (C.c_int, [C.c_void_p] * 2 + [C.c_int] * 4 + [C.c_void_p, C.c_void_p]),
    "ab_avgpool_bwd": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "ab_dilate2x": (C.c_int, [C.c_void_p] + [C.c_int] * 6 + [C.c_void_p, C.c_void_p]),
    "ab_deconv4x4s2_gather": (C.c_int, [C.c_void_p] + [C.c_int] * 4 + [C.c_void_p, C.c_void_p]),
    "ab_head_decode_bwd": (C.c_int, [C.c_void_p] * 4 + [C.c_int] * 5 + [C.c_void_p, C.c_void_p]),
    "ab_augment_workspace_bytes": (C.c_uint64, [C.POINTER(AugmentCfgStruct), C.c_int]),
    "ab_crop_augment": (C.c_int, [C.POINTER(AugmentCfgStruct), C.c_int] + [C.c_void_p] * 21),
    "ab_sumsq": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "ab_adam_step": (C.c_int, [C.c_void_p] * 4 + [C.c_int64] + [C.c_float] * 5 + [C.c_void_p, C.c_void_p, C.c_float, C.c_float,
                                                                                  C.c_void_p]),
}

SYNTH_UNIFORMS = 32  # AB_SYNTH_UNIFORMS
SUMSQ_PARTS = 1184  # AB_SUMSQ_PARTS: ab_sumsq writes out[0] and up to this many partials behind it
STAT_PARTS = 1184  # AB_STAT_PARTS: rows of the column-reduction workspace (2 * STAT_PARTS * C floats)
_lib = None


class AbError(RuntimeError):
    pass


def load() -> C.CDLL:
    """Load the library (once).  Raises if it was not built: run `python -m artiboost_b200.build`."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise AbError(f"{LIB_PATH} is missing: build it with `python -m artiboost_b200.build` "
                          "(there is no CPU fallback)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in EXPORTS.items():
            fn = getattr(lib, name)  # AttributeError if the symbol is not exported
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise AbError(f"{what} failed (rc={rc}): {load().ab_last_error().decode()}")


def ptr(t):
    """Device (or host, for *_host arguments) address of a contiguous tensor; None -> NULL."""
    if t is None:
        return None
    assert t.is_contiguous(), "C-ABI arguments must be contiguous"
    return C.c_void_p(t.data_ptr())


_stream_override = None


@contextlib.contextmanager
def use_stream(stream):
    """Inside the block every library launch goes to `stream` while torch's current stream -- and with it the caching
    allocator's bookkeeping -- stays where it is.  For callers that order the two streams themselves with events: buffers
    allocated here belong to the current stream, so no Tensor.record_stream() (slow, and it defers the blocks' reuse) is
    needed as long as the consumer on the current stream waits for the launches and the launches wait for the current
    stream's earlier work."""
    global _stream_override
    prev, _stream_override = _stream_override, stream
    try:
        yield
    finally:
        _stream_override = prev


def stream_ptr(device=None):
    if _stream_override is not None:
        return C.c_void_p(_stream_override.cuda_stream)
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def require_cuda(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise AbError(f"{name} must live on a CUDA device: artiboost_b200 has no CPU path")


def launch_count() -> int:
    return int(load().ab_launch_count())


STAGES = {0: "raster_bin_kernel", 1: "raster_tile_kernel", 2: "synth_draw_kernel", 3: "mano_lbs_kernel",
          4: "posegen_prelude_kernel", 5: "ccv_cdf+draw_kernels", 6: "view_kernel", 7: "gemm_bf16_tn_kernel",
          8: "im2col_kernel", 9: "elementwise_kernels", 10: "head_decode_kernel", 11: "gemm_bf16_tn_kernel<im2col TMA>", 12: "wgrad_bf16_kernel", 13: "train_elementwise_kernels", 14: "optimizer_kernels", 15: "bn_apply_kernel", 16: "bn_bwd_reduce_kernel",
          17: "bn_bwd_apply_kernel", 18: "bn_finalize_kernel", 19: "augment_kernels",
          20: "chamfer_nn_kernel", 21: "linear_f32_kernel", 22: "refine_misc_kernels", 23: "tail_loss_kernel"}


def profile_enable(on: bool) -> None:
    check(load().ab_profile_enable(int(on)), "ab_profile_enable")


def profile_collect() -> dict:
    """-> {stage name: (total ms, launches)} for everything launched since profiling was enabled / last collected."""
    n = 24
    ms, cnt = (C.c_double * n)(), (C.c_int64 * n)()
    check(load().ab_profile_collect(ms, cnt, n), "ab_profile_collect")
    return {STAGES.get(i, f"stage{i}"): (ms[i], int(cnt[i])) for i in range(n) if cnt[i]}
