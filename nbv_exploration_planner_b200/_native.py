"""ctypes loader for the product library ``csrc/libnbvgpu.so`` (C ABI in include/nbvgpu.h).

There is no CPU fallback: if the library is missing it is built with nvcc, and if that fails or no
B200 is visible the constructors raise.  Nothing in this package imports ``oracle``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(CSRC, "libnbvgpu.so")
SYNTH_PATH = os.path.join(CSRC, "libnbvsynth.so")

OK, ERR_INVALID, ERR_NO_DEVICE, ERR_CUDA, ERR_CAPACITY, ERR_STATE, ERR_RANGE = 0, -1, -2, -3, -4, -5, -6
LAYER_TSDF, LAYER_ESDF = 0, 1
PACK_PLANAR, PACK_VOXBLOX = 0, 1


class NbvGpuError(RuntimeError):
    def __init__(self, code: int, msg: str = ""):
        names = {-1: "INVALID", -2: "NO_DEVICE", -3: "CUDA", -4: "CAPACITY", -5: "STATE", -6: "RANGE"}
        super().__init__(f"nbvgpu error {names.get(code, code)}: {msg}")
        self.code = code


class Camera(C.Structure):
    """nbvgpu_camera."""

    _fields_ = [
        ("hfov_deg", C.c_double), ("vfov_deg", C.c_double), ("near_m", C.c_double), ("far_m", C.c_double),
        ("q_CB", C.c_double * 4), ("t_CB", C.c_double * 3),
        ("bbx_min", C.c_double * 3), ("bbx_max", C.c_double * 3),
        ("gain_free", C.c_double), ("gain_occupied", C.c_double), ("gain_unmapped", C.c_double),
        ("sum_all_yaws", C.c_int32), ("reserved", C.c_int32),
    ]


class Stats(C.Structure):
    """nbvgpu_stats."""

    _fields_ = [(k, C.c_uint64) for k in
                ("rays", "voxel_steps", "collision_steps", "blocks_uploaded", "bytes_uploaded", "kernel_launches")]


class ExpandParams(C.Structure):
    """nbvgpu_expand_params."""

    _fields_ = [("root", C.c_double * 3), ("vicinity", C.c_double), ("extension_range", C.c_double), ("overshoot", C.c_double),
                ("robot_radius", C.c_double), ("v_max", C.c_double), ("dyaw_max", C.c_double), ("zero_gain", C.c_double),
                ("sequential", C.c_int32), ("reserved", C.c_int32)]


class Best(C.Structure):
    """nbvgpu_best."""

    _fields_ = [("index", C.c_int64), ("score", C.c_double), ("gain", C.c_double), ("yaw", C.c_double)]


def build(force: bool = False, verbose: bool = False) -> None:
    """Compile libnbvgpu.so / libnbvsynth.so in-tree (nvcc -gencode arch=compute_100a,code=sm_100a)."""
    # make decides what is stale (the Makefile lists every header of the library)
    args = ["make", "-C", CSRC]
    if force:
        args.append("-B")
    out = None if verbose else subprocess.DEVNULL
    # one builder at a time: the ranks of a torchrun launch all come through here
    import fcntl
    with open(os.path.join(CSRC, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            subprocess.check_call(args, stdout=out, stderr=None if verbose else subprocess.STDOUT)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


_lib = None
_synth = None

_u8p, _i32p, _u32p, _i64p, _u64p, _f32p, _f64p = (C.POINTER(t) for t in
                                                   (C.c_uint8, C.c_int32, C.c_uint32, C.c_int64, C.c_uint64, C.c_float, C.c_double))


def lib() -> C.CDLL:
    """Load (building if necessary) the CUDA library.  Raises if it cannot be had."""
    global _lib
    if _lib is not None:
        return _lib
    build()
    if not os.path.exists(LIB_PATH):
        raise NbvGpuError(ERR_NO_DEVICE, f"{LIB_PATH} missing and could not be built; there is no CPU fallback")
    L = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    L.nbvgpu_version.restype = C.c_char_p
    L.nbvgpu_last_error.restype = C.c_char_p
    L.nbvgpu_last_error.argtypes = [vp]
    L.nbvgpu_create.argtypes = [C.c_int, C.POINTER(vp)]
    L.nbvgpu_create_multi.argtypes = [C.c_int, vp, C.POINTER(vp)]
    L.nbvgpu_group_unique_id.argtypes = [vp, C.c_size_t]
    L.nbvgpu_create_rank.argtypes = [C.c_int, C.c_int, C.c_int, vp, C.c_size_t, C.POINTER(vp)]
    L.nbvgpu_group_info.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.nbvgpu_layer_update_pinned.argtypes = [vp, C.c_int, C.c_size_t, vp, vp, C.c_int]
    L.nbvgpu_shared_alloc.argtypes = [vp, C.c_size_t, C.POINTER(vp)]
    L.nbvgpu_shared_free.argtypes = [vp, vp]
    L.nbvgpu_debug_pulled_chunks.argtypes = [vp, C.POINTER(C.c_uint64)]
    L.nbvgpu_evaluate_shard_device.argtypes = [vp, C.c_int, C.c_int, C.c_double, C.c_size_t, C.c_int64, vp, vp, vp, vp, C.c_double,
                                               C.c_double, C.c_double, vp, vp, vp, C.POINTER(Best)]
    L.nbvgpu_destroy.argtypes = [vp]
    L.nbvgpu_destroy.restype = None
    L.nbvgpu_set_stream.argtypes = [vp, vp]
    L.nbvgpu_synchronize.argtypes = [vp]
    L.nbvgpu_layer_create.argtypes = [vp, C.c_int, C.c_float, C.c_int, C.c_size_t, C.POINTER(C.c_int)]
    L.nbvgpu_layer_update.argtypes = [vp, C.c_int, C.c_size_t, vp, vp, C.c_int]
    L.nbvgpu_layer_update_device.argtypes = [vp, C.c_int, C.c_size_t, vp, vp, C.c_int]
    L.nbvgpu_layer_remove.argtypes = [vp, C.c_int, C.c_size_t, vp]
    L.nbvgpu_layer_clear.argtypes = [vp, C.c_int]
    L.nbvgpu_commit.argtypes = [vp]
    L.nbvgpu_layer_num_blocks.argtypes = [vp, C.c_int, C.POINTER(C.c_size_t)]
    L.nbvgpu_layer_read_voxels.argtypes = [vp, C.c_int, C.c_size_t, vp, vp, vp, vp]
    L.nbvgpu_set_camera.argtypes = [vp, C.POINTER(Camera)]
    L.nbvgpu_gain_eval.argtypes = [vp, C.c_int, C.c_size_t, vp, vp, vp, vp, vp, vp]
    L.nbvgpu_gain_eval_device.argtypes = [vp, C.c_int, C.c_size_t, vp, vp, vp, vp, vp, vp]
    L.nbvgpu_optimized_gain.argtypes = [vp, C.c_int, _f64p, _f64p]
    L.nbvgpu_gain_eval_best.argtypes = [vp, C.c_int, C.c_size_t, vp, vp, vp, vp, C.c_double, C.c_double, C.c_double,
                                        C.POINTER(Best), vp, vp]
    L.nbvgpu_gain_eval_best_device.argtypes = [vp, C.c_int, C.c_size_t, vp, vp, vp, vp, C.c_double, C.c_double,
                                               C.c_double, vp, vp, vp]
    L.nbvgpu_evaluate_candidates.argtypes = [vp, C.c_int, C.c_int, C.c_double, C.c_size_t, vp, vp, vp, vp, C.c_double, C.c_double,
                                             C.c_double, vp, C.POINTER(Best), vp, vp]
    L.nbvgpu_evaluate_candidates_device.argtypes = [vp, C.c_int, C.c_int, C.c_double, C.c_size_t, vp, vp, vp, vp, C.c_double,
                                                    C.c_double, C.c_double, vp, vp, vp, vp]
    L.nbvgpu_check_segments.argtypes = [vp, C.c_int, C.c_double, C.c_size_t, vp, vp]
    L.nbvgpu_check_segments_device.argtypes = [vp, C.c_int, C.c_double, C.c_size_t, vp, vp]
    L.nbvgpu_check_motion.argtypes = [vp, C.c_int, C.c_double, _f64p, _f64p, C.POINTER(C.c_int)]
    L.nbvgpu_esdf_distance_gradient.argtypes = [vp, C.c_int, C.c_size_t, vp, vp, vp, vp]
    L.nbvgpu_esdf_distance_gradient_device.argtypes = [vp, C.c_int, C.c_size_t, vp, vp, vp, vp]
    L.nbvgpu_frontier_potential.argtypes = [vp, C.c_int, C.c_size_t, vp, C.c_double, vp, vp, vp, vp]
    L.nbvgpu_debug_round_decimal6.argtypes = [C.c_double, C.POINTER(C.c_uint64)]
    L.nbvgpu_debug_sqrt_threshold.argtypes = [C.c_double, C.POINTER(C.c_double)]
    L.nbvgpu_debug_frontier_hash_redo.argtypes = [vp, C.POINTER(C.c_uint64)]
    L.nbvgpu_debug_dense_clipped_launches.argtypes = [vp, C.POINTER(C.c_uint64)]
    L.nbvgpu_debug_fused_launches.argtypes = [vp, C.POINTER(C.c_uint64)]
    L.nbvgpu_debug_wide_cta_launches.argtypes = [vp, C.POINTER(C.c_uint64)]
    L.nbvgpu_debug_graph_replays.argtypes = [vp, C.POINTER(C.c_uint64)]
    L.nbvgpu_probe_l2_read_bandwidth.argtypes = [vp, C.c_size_t, C.c_int, C.POINTER(C.c_double)]
    L.nbvgpu_get_stats.argtypes = [vp, C.POINTER(Stats)]
    L.nbvgpu_set_kernel_variant.argtypes = [vp, C.c_int]
    L.nbvgpu_debug_classifier_mismatches.argtypes = [vp, C.POINTER(C.c_uint64)]
    L.nbvgpu_layer_load_voxblox_file.argtypes = [vp, C.c_int, C.c_char_p, C.POINTER(C.c_size_t)]
    L.nbvgpu_layer_apply_layer_msg.argtypes = [vp, C.c_int, vp, C.c_size_t, C.POINTER(C.c_size_t)]
    L.nbvgpu_parse_blocks.argtypes = [C.c_int, vp, C.c_size_t, C.c_int, C.c_float, C.c_int, C.POINTER(C.c_size_t), vp, vp,
                                      C.c_size_t, C.POINTER(C.c_int)]
    L.nbvgpu_expand_tree.argtypes = [vp, C.c_int, C.c_int, C.POINTER(ExpandParams), C.c_size_t, vp, vp, C.c_size_t, vp, vp, vp, vp,
                                     vp, vp, vp, vp, C.POINTER(Best)]
    L.nbvgpu_set_profiling.argtypes = [vp, C.c_int]
    L.nbvgpu_profile_collect.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_uint64)]
    _lib = L
    return L


def synth_lib() -> C.CDLL:
    global _synth
    if _synth is None:
        build()
        S = C.CDLL(SYNTH_PATH)
        S.nbvsynth_generate.argtypes = [C.c_float, C.c_int, C.c_size_t, _i32p, C.c_size_t, _f64p, C.c_size_t, _f64p,
                                        C.c_double, _f64p, _f64p, C.c_float, C.c_float, _f32p, _f32p, C.c_int]
        _synth = S
    return _synth


# Every symbol include/nbvgpu.h declares (checked by tests/test_abi_symbols.py against the header).
ABI_SYMBOLS = [
    "nbvgpu_create", "nbvgpu_create_multi", "nbvgpu_group_unique_id", "nbvgpu_create_rank", "nbvgpu_group_info",
    "nbvgpu_layer_update_pinned", "nbvgpu_evaluate_shard_device", "nbvgpu_destroy", "nbvgpu_set_stream", "nbvgpu_synchronize", "nbvgpu_last_error", "nbvgpu_version",
    "nbvgpu_layer_create", "nbvgpu_layer_update", "nbvgpu_layer_update_device", "nbvgpu_layer_remove",
    "nbvgpu_layer_clear", "nbvgpu_commit", "nbvgpu_layer_num_blocks", "nbvgpu_layer_read_voxels",
    "nbvgpu_set_camera", "nbvgpu_gain_eval", "nbvgpu_gain_eval_device", "nbvgpu_optimized_gain",
    "nbvgpu_gain_eval_best", "nbvgpu_gain_eval_best_device", "nbvgpu_check_segments",
    "nbvgpu_check_segments_device", "nbvgpu_check_motion", "nbvgpu_get_stats", "nbvgpu_set_kernel_variant",
    "nbvgpu_debug_classifier_mismatches", "nbvgpu_set_profiling", "nbvgpu_profile_collect",
    "nbvgpu_layer_load_voxblox_file", "nbvgpu_layer_apply_layer_msg", "nbvgpu_parse_blocks", "nbvgpu_expand_tree",
    "nbvgpu_frontier_potential", "nbvgpu_esdf_distance_gradient", "nbvgpu_esdf_distance_gradient_device", "nbvgpu_evaluate_candidates", "nbvgpu_evaluate_candidates_device", "nbvgpu_debug_round_decimal6", "nbvgpu_debug_sqrt_threshold", "nbvgpu_debug_frontier_hash_redo", "nbvgpu_debug_dense_clipped_launches", "nbvgpu_debug_fused_launches", "nbvgpu_debug_wide_cta_launches", "nbvgpu_debug_graph_replays", "nbvgpu_probe_l2_read_bandwidth",
    "nbvgpu_shared_alloc", "nbvgpu_shared_free", "nbvgpu_debug_pulled_chunks",
]
