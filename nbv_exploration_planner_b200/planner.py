"""Host-side mirror of the reference's interface for the viewpoint-gain path, over the C ABI.

The reference reaches the path through three member functions; the classes below
keep their names, argument meaning and failure behaviour so that callers -- and the parity tests --
read like the reference:

  ``RrtGain.optimized_gain(state)`` / ``.gain(state)``   nbveplanner/src/rrt.cpp:351-479 / :481-579
  ``HighResManager.checkMotion(start, end)``              nbveplanner/include/nbveplanner/voxblox_manager.h:89-108
  ``LowResManager.getVoxelStatusByIndex(block, voxel)``   nbveplanner/src/voxblox_manager.cpp:14-33
  ``Layer`` (update / remove / clear / commit)            voxblox/voxblox/include/voxblox/core/layer.h

plus the batched forms the B200 path is built for (``gain_eval``, ``check_segments``,
``gain_eval_best``).  Everything runs on the GPU through ``libnbvgpu.so``; there is no CPU path here
and this module never imports ``oracle``.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np

from . import _native as N

# VoxelStatus (nbveplanner/include/nbveplanner/voxblox_manager.h): values of the 2-bit status tiles.
kUnknown, kFree, kOccupied = 0, 1, 2


@dataclass
class CameraParams:
    """Inputs of CameraModel::setIntrinsicsFromFoV / setExtrinsics / setBoundingBox and Params::gain_*."""

    hfov_deg: float = 90.0
    vfov_deg: float = 60.0
    near_m: float = 0.1
    far_m: float = 5.0
    q_CB: Sequence[float] = (1.0, 0.0, 0.0, 0.0)
    t_CB: Sequence[float] = (0.0, 0.0, 0.0)
    bbx_min: Sequence[float] = (0.0, 0.0, 0.0)
    bbx_max: Sequence[float] = (0.0, 0.0, 0.0)
    gain_free: float = 0.0
    gain_occupied: float = 0.0
    gain_unmapped: float = 1.0
    sum_all_yaws: int = 0

    @staticmethod
    def mount(pitch_deg: float = 15.0, t_BC: Sequence[float] = (0.1, 0.0, 0.0), **kw) -> "CameraParams":
        """Camera mounted at ``t_BC`` in the body frame, pitched ``pitch_deg`` about body +y.

        The reference only ever sees the TF extrinsic T_C_B (nbvep.cpp:48-61; ``camera_pitch_`` is
        read and never used), so the mount is converted to T_C_B = T_B_C^-1 here.
        """
        h = math.radians(pitch_deg) / 2.0
        q_BC = np.array([math.cos(h), 0.0, math.sin(h), 0.0])
        q_CB = q_BC * np.array([1.0, -1.0, -1.0, -1.0])
        t = np.asarray(t_BC, float)
        w, x, y, z = q_CB
        R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                      [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                      [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])
        return CameraParams(q_CB=tuple(q_CB), t_CB=tuple(-(R @ t)), **kw)

    def as_dict(self) -> dict:
        return {k: (list(v) if isinstance(v, (tuple, list, np.ndarray)) else v) for k, v in self.__dict__.items()}

    def as_struct(self) -> N.Camera:
        s = N.Camera()
        for k in ("hfov_deg", "vfov_deg", "near_m", "far_m", "gain_free", "gain_occupied", "gain_unmapped"):
            setattr(s, k, float(getattr(self, k)))
        for k, n in (("q_CB", 4), ("t_CB", 3), ("bbx_min", 3), ("bbx_max", 3)):
            getattr(s, k)[:] = [float(v) for v in getattr(self, k)][:n]
        s.sum_all_yaws = int(self.sum_all_yaws)
        return s


def _vp(a) -> Optional[C.c_void_p]:
    """void* of a numpy array, a torch tensor (device or host) or a raw int address."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return C.c_void_p(a.ctypes.data)
    if hasattr(a, "data_ptr"):
        return C.c_void_p(a.data_ptr())
    return C.c_void_p(int(a))


class NbvGpu:
    """One ``nbvgpu_ctx``: a GPU, its map-mirror layers, a camera and the evaluators."""

    def __init__(self, device: int = 0, _handle=None):
        self._L = N.lib()
        if _handle is None:
            h = C.c_void_p()
            rc = self._L.nbvgpu_create(device, C.byref(h))
            if rc != 0:
                raise N.NbvGpuError(rc, "nbvgpu_create failed (a B200 / sm_100 device is required; no CPU fallback)")
        else:
            h = _handle
        self._h = h
        self.device = device
        self.layers = {}

    # -- multi-GPU contexts: the same object, the same methods ------------------------
    @classmethod
    def multi(cls, device_ids: Sequence[int]) -> "NbvGpu":
        """``nbvgpu_create_multi``: this process drives every GPU in ``device_ids``; batches are sharded over them,
        ``commit`` replicates the map with one NCCL broadcast per chunk."""
        L = N.lib()
        ids = (C.c_int * len(device_ids))(*[int(d) for d in device_ids])
        h = C.c_void_p()
        rc = L.nbvgpu_create_multi(len(device_ids), ids, C.byref(h))
        if rc != 0:
            raise N.NbvGpuError(rc, "nbvgpu_create_multi failed (B200 devices and NCCL are required; no CPU fallback)")
        return cls(int(device_ids[0]), _handle=h)

    @staticmethod
    def unique_id() -> bytes:
        """``nbvgpu_group_unique_id``: 128 bytes rank 0 creates and every rank of a one-process-per-GPU context needs."""
        L = N.lib()
        buf = (C.c_uint8 * 128)()
        rc = L.nbvgpu_group_unique_id(buf, 128)
        if rc != 0:
            raise N.NbvGpuError(rc, "nbvgpu_group_unique_id failed (NCCL is required)")
        return bytes(buf)

    @classmethod
    def rank(cls, device: int, rank: int, world: int, unique_id: bytes) -> "NbvGpu":
        """``nbvgpu_create_rank``: this process is rank ``rank`` of ``world`` and drives GPU ``device``."""
        L = N.lib()
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id[:128])
        h = C.c_void_p()
        rc = L.nbvgpu_create_rank(device, rank, world, buf, 128, C.byref(h))
        if rc != 0:
            raise N.NbvGpuError(rc, "nbvgpu_create_rank failed (a B200 per rank and NCCL are required; no CPU fallback)")
        return cls(device, _handle=h)

    def group_info(self) -> dict:
        r, w, l = C.c_int(0), C.c_int(1), C.c_int(1)
        self._ck(self._L.nbvgpu_group_info(self._h, C.byref(r), C.byref(w), C.byref(l)))
        return {"rank": r.value, "world": w.value, "local_gpus": l.value}

    # -- lifetime ---------------------------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "_h", None):
            self._L.nbvgpu_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc: int) -> None:
        if rc != 0:
            raise N.NbvGpuError(rc, (self._L.nbvgpu_last_error(self._h) or b"").decode())

    def set_stream(self, cuda_stream: Optional[int]) -> None:
        """Run all subsequent work on ``cuda_stream`` (e.g. ``torch.cuda.current_stream().cuda_stream``).
        ``None`` restores the library-owned stream; handle 0 (torch's default stream) is passed as
        ``cudaStreamLegacy`` (0x1), because a NULL pointer means "own stream" at the ABI."""
        if cuda_stream is None:
            self._ck(self._L.nbvgpu_set_stream(self._h, None))
        else:
            self._ck(self._L.nbvgpu_set_stream(self._h, C.c_void_p(cuda_stream if cuda_stream else 1)))

    def synchronize(self) -> None:
        self._ck(self._L.nbvgpu_synchronize(self._h))

    # -- map mirror --------------------------------------------------------------------------
    def layer_create(self, kind: int, voxel_size: float, voxels_per_side: int = 16, max_blocks: int = 4096) -> int:
        out = C.c_int(-1)
        self._ck(self._L.nbvgpu_layer_create(self._h, kind, np.float32(voxel_size), voxels_per_side, max_blocks, C.byref(out)))
        self.layers[out.value] = (kind, np.float32(voxel_size), voxels_per_side)
        return out.value

    def layer_update(self, layer: int, keys: np.ndarray, payload: np.ndarray, packing: int = N.PACK_PLANAR) -> None:
        keys = np.ascontiguousarray(keys, np.int32).reshape(-1, 3)
        kind, _, vps = self.layers[layer]
        nvox = vps ** 3
        if packing == N.PACK_PLANAR:
            payload = np.ascontiguousarray(payload, np.float32)
            per = nvox * (2 if kind == N.LAYER_TSDF else 1)
        else:
            payload = np.ascontiguousarray(payload, np.uint32)
            per = nvox * (3 if kind == N.LAYER_TSDF else 2)
        if payload.size != per * len(keys):
            raise ValueError(f"payload has {payload.size} words, expected {per * len(keys)}")
        self._ck(self._L.nbvgpu_layer_update(self._h, layer, len(keys), _vp(keys), _vp(payload), packing))

    def layer_update_pinned(self, layer: int, keys: np.ndarray, payload, packing: int = N.PACK_PLANAR) -> None:
        """``nbvgpu_layer_update_pinned``: ``payload`` (a pinned torch tensor or anything with ``data_ptr`` / a page-locked
        numpy array) is NOT copied: it must stay untouched until ``commit`` has returned."""
        keys = np.ascontiguousarray(keys, np.int32).reshape(-1, 3)
        self._pinned_keep = getattr(self, "_pinned_keep", []) + [payload]   # keep it alive until the commit
        self._ck(self._L.nbvgpu_layer_update_pinned(self._h, layer, len(keys), _vp(keys), _vp(payload), packing))

    def shared_alloc(self, nbytes: int) -> np.ndarray:
        """``nbvgpu_shared_alloc``: ``nbytes`` of page-locked host memory every GPU of the context can read, as a uint8 array
        (view it as needed).  Collective in a one-process-per-GPU context: every rank calls it with the same size and gets its
        own mapping of the same bytes.  Blocks staged from it (``layer_update_pinned``) are replicated by a parallel pull."""
        p = C.c_void_p()
        self._ck(self._L.nbvgpu_shared_alloc(self._h, int(nbytes), C.byref(p)))
        buf = (C.c_uint8 * int(nbytes)).from_address(p.value)
        a = np.frombuffer(buf, dtype=np.uint8)
        self._shared = getattr(self, "_shared", {})
        self._shared[a.ctypes.data] = p.value
        return a

    def shared_free(self, a: np.ndarray) -> None:
        self._ck(self._L.nbvgpu_shared_free(self._h, C.c_void_p(self._shared[a.ctypes.data])))
        del self._shared[a.ctypes.data]

    def pulled_chunks(self) -> int:
        out = C.c_uint64(0)
        self._ck(self._L.nbvgpu_debug_pulled_chunks(self._h, C.byref(out)))
        return int(out.value)

    def layer_update_device(self, layer: int, n_blocks: int, d_keys, d_payload, packing: int = N.PACK_PLANAR) -> None:
        self._ck(self._L.nbvgpu_layer_update_device(self._h, layer, n_blocks, _vp(d_keys), _vp(d_payload), packing))

    def layer_remove(self, layer: int, keys: np.ndarray) -> None:
        keys = np.ascontiguousarray(keys, np.int32).reshape(-1, 3)
        self._ck(self._L.nbvgpu_layer_remove(self._h, layer, len(keys), _vp(keys)))

    def layer_clear(self, layer: int) -> None:
        self._ck(self._L.nbvgpu_layer_clear(self._h, layer))

    def commit(self) -> None:
        try:
            self._ck(self._L.nbvgpu_commit(self._h))
        finally:
            self._pinned_keep = []

    def layer_num_blocks(self, layer: int) -> int:
        out = C.c_size_t(0)
        self._ck(self._L.nbvgpu_layer_num_blocks(self._h, layer, C.byref(out)))
        return out.value

    def layer_read_voxels(self, layer: int, global_index: np.ndarray):
        g = np.ascontiguousarray(global_index, np.int64).reshape(-1, 3)
        n = len(g)
        vals = np.zeros((n, 2), np.float32)
        found = np.zeros(n, np.uint8)
        status = np.zeros(n, np.uint8)
        self._ck(self._L.nbvgpu_layer_read_voxels(self._h, layer, n, _vp(g), _vp(vals), _vp(found), _vp(status)))
        return vals, found, status

    # -- ingest in the reference's own formats ---------------------------------
    def layer_load_voxblox_file(self, layer: int, path: str) -> int:
        """voxblox::io::LoadBlocksFromFile: load a .voxblox file into `layer`; returns #blocks."""
        n = C.c_size_t(0)
        self._ck(self._L.nbvgpu_layer_load_voxblox_file(self._h, layer, str(path).encode(), C.byref(n)))
        return n.value

    def layer_apply_layer_msg(self, layer: int, msg: bytes) -> int:
        """voxblox::deserializeMsgToLayer on a ROS-1-serialised voxblox_msgs/Layer; returns #blocks."""
        buf = np.frombuffer(msg, np.uint8)
        n = C.c_size_t(0)
        self._ck(self._L.nbvgpu_layer_apply_layer_msg(self._h, layer, _vp(buf), len(buf), C.byref(n)))
        return n.value

    # -- camera ------------------------------------------------------------------------------
    def set_camera(self, cam: CameraParams) -> None:
        s = cam.as_struct()
        self._ck(self._L.nbvgpu_set_camera(self._h, C.byref(s)))
        self.camera = cam

    # -- gain --------------------------------------------------------------------------------
    def gain_eval(self, layer: int, pos: np.ndarray, per_yaw: bool = False, steps: bool = True) -> dict:
        pos = np.ascontiguousarray(pos, np.float64).reshape(-1, 3)
        n = len(pos)
        out = {"gain": np.zeros(n, np.float64), "yaw_deg": np.zeros(n, np.int32), "counts": np.zeros((n, 3), np.uint64),
               "per_yaw": np.zeros((n, 360, 3), np.uint32) if per_yaw else None,
               "steps": np.zeros(n, np.uint64) if steps else None}
        self._ck(self._L.nbvgpu_gain_eval(self._h, layer, n, _vp(pos), _vp(out["gain"]), _vp(out["yaw_deg"]),
                                          _vp(out["counts"]), _vp(out["per_yaw"]), _vp(out["steps"])))
        out["yaw"] = out["yaw_deg"] * math.pi / 180.0
        return out

    def gain_eval_device(self, layer: int, n: int, d_pos, d_gain, d_yaw_deg=None, d_counts=None, d_per_yaw=None,
                         d_steps=None) -> None:
        self._ck(self._L.nbvgpu_gain_eval_device(self._h, layer, n, _vp(d_pos), _vp(d_gain), _vp(d_yaw_deg), _vp(d_counts),
                                                 _vp(d_per_yaw), _vp(d_steps)))

    def gain_eval_best(self, layer: int, pos, parent_yaw, distance, edge_free=None, v_max: float = 1.0,
                       dyaw_max: float = 0.5, zero_gain: float = 0.0) -> dict:
        pos = np.ascontiguousarray(pos, np.float64).reshape(-1, 3)
        n = len(pos)
        parent_yaw = np.ascontiguousarray(parent_yaw, np.float64)
        distance = np.ascontiguousarray(distance, np.float64)
        ef = None if edge_free is None else np.ascontiguousarray(edge_free, np.uint8)
        best = N.Best()
        gain = np.zeros(n, np.float64)
        yaw = np.zeros(n, np.int32)
        self._ck(self._L.nbvgpu_gain_eval_best(self._h, layer, n, _vp(pos), _vp(parent_yaw), _vp(distance), _vp(ef), v_max,
                                               dyaw_max, zero_gain, C.byref(best), _vp(gain), _vp(yaw)))
        return {"index": best.index, "score": best.score, "best_gain": best.gain, "best_yaw": best.yaw, "gain": gain,
                "yaw_deg": yaw}

    def gain_eval_best_device(self, layer: int, n: int, d_pos, d_parent_yaw, d_distance, d_edge_free, v_max, dyaw_max,
                              zero_gain, d_best, d_gain=None, d_yaw_deg=None) -> None:
        self._ck(self._L.nbvgpu_gain_eval_best_device(self._h, layer, n, _vp(d_pos), _vp(d_parent_yaw), _vp(d_distance),
                                                      _vp(d_edge_free), v_max, dyaw_max, zero_gain, _vp(d_best), _vp(d_gain),
                                                      _vp(d_yaw_deg)))

    def evaluate_candidates(self, tsdf_layer: int, esdf_layer: int, robot_radius: float, pos, seg, parent_yaw, distance,
                            v_max: float = 1.0, dyaw_max: float = 0.5, zero_gain: float = 0.0) -> dict:
        """Edge check + gain + node scoring in one call (what ``RrtTree::iterate`` does with a sample after the sampler,
        rrt.cpp:205-246): the same results as ``check_segments`` followed by ``gain_eval_best`` with its verdicts."""
        pos = np.ascontiguousarray(pos, np.float64).reshape(-1, 3)
        seg = np.ascontiguousarray(seg, np.float64).reshape(-1, 6)
        n = len(pos)
        if len(seg) != n:
            raise ValueError("one segment per candidate")
        parent_yaw = np.ascontiguousarray(parent_yaw, np.float64)
        distance = np.ascontiguousarray(distance, np.float64)
        best = N.Best()
        gain, yaw, free = np.zeros(n, np.float64), np.zeros(n, np.int32), np.zeros(n, np.uint8)
        self._ck(self._L.nbvgpu_evaluate_candidates(self._h, tsdf_layer, esdf_layer, float(robot_radius), n, _vp(pos), _vp(seg),
                                                    _vp(parent_yaw), _vp(distance), v_max, dyaw_max, zero_gain, _vp(free),
                                                    C.byref(best), _vp(gain), _vp(yaw)))
        return {"index": best.index, "score": best.score, "best_gain": best.gain, "best_yaw": best.yaw, "gain": gain,
                "yaw_deg": yaw, "free": free}

    def evaluate_candidates_device(self, tsdf_layer: int, esdf_layer: int, robot_radius: float, n: int, d_pos, d_seg,
                                   d_parent_yaw, d_distance, v_max, dyaw_max, zero_gain, d_edge_free, d_best, d_gain=None,
                                   d_yaw_deg=None) -> None:
        self._ck(self._L.nbvgpu_evaluate_candidates_device(self._h, tsdf_layer, esdf_layer, float(robot_radius), n, _vp(d_pos),
                                                           _vp(d_seg), _vp(d_parent_yaw), _vp(d_distance), v_max, dyaw_max,
                                                           zero_gain, _vp(d_edge_free), _vp(d_best), _vp(d_gain), _vp(d_yaw_deg)))

    def evaluate_shard_device(self, tsdf_layer: int, esdf_layer: int, robot_radius: float, n_local: int, index_offset: int, d_pos,
                              d_seg, d_parent_yaw, d_distance, v_max, dyaw_max, zero_gain, d_edge_free, d_gain=None,
                              d_yaw_deg=None) -> dict:
        """``nbvgpu_evaluate_shard_device`` (one process per GPU): this rank's shard from device memory, the best node of
        the WHOLE batch back in host memory -- no collective, no copy, no stream synchronisation."""
        best = N.Best()
        self._ck(self._L.nbvgpu_evaluate_shard_device(self._h, tsdf_layer, esdf_layer, float(robot_radius), n_local, index_offset,
                                                      _vp(d_pos), _vp(d_seg), _vp(d_parent_yaw), _vp(d_distance), v_max, dyaw_max,
                                                      zero_gain, _vp(d_edge_free), _vp(d_gain), _vp(d_yaw_deg), C.byref(best)))
        return {"index": best.index, "score": best.score, "gain": best.gain, "yaw": best.yaw}

    def optimized_gain(self, layer: int, state: np.ndarray) -> float:
        """``double RrtTree::optimized_gain(Pose& state)``: returns the gain, writes ``state[3]``."""
        if not (isinstance(state, np.ndarray) and state.dtype == np.float64 and state.size >= 4 and state.flags.c_contiguous):
            raise TypeError("state must be a contiguous float64 array of 4 elements (Pose)")
        g = C.c_double(0.0)
        self._ck(self._L.nbvgpu_optimized_gain(self._h, layer, state.ctypes.data_as(C.POINTER(C.c_double)), C.byref(g)))
        return g.value

    # -- batched tree expansion --------------------------------------------------
    def expand_tree(self, tsdf_layer: int, esdf_layer: int, params: dict, node_state, node_distance, uniforms,
                    sequential: bool = False) -> dict:
        """One batched pass of the body of ``RrtTree::iterate`` (rrt.cpp:172-246) over uniform triples; ``sequential``:
        the reference's semantics (every sample sees the nodes created by the samples before it)."""
        P = N.ExpandParams()
        P.sequential = 1 if sequential else 0
        P.root[:] = [float(v) for v in params["root"]]
        for k in ("vicinity", "extension_range", "overshoot", "robot_radius", "v_max", "dyaw_max", "zero_gain"):
            setattr(P, k, float(params[k]))
        ns = np.ascontiguousarray(node_state, np.float64).reshape(-1, 4)
        nd = np.ascontiguousarray(node_distance, np.float64)
        u = np.ascontiguousarray(uniforms, np.float64).reshape(-1, 3)
        n = len(u)
        out = {"status": np.zeros(n, np.uint8), "pos": np.zeros((n, 3)), "parent": np.zeros(n, np.int32), "gain": np.zeros(n),
               "yaw": np.zeros(n), "distance": np.zeros(n), "score": np.zeros(n)}
        best = N.Best()
        self._ck(self._L.nbvgpu_expand_tree(self._h, tsdf_layer, esdf_layer, C.byref(P), len(ns), _vp(ns), _vp(nd), n, _vp(u),
                                            _vp(out["status"]), _vp(out["pos"]), _vp(out["parent"]), _vp(out["gain"]),
                                            _vp(out["yaw"]), _vp(out["distance"]), _vp(out["score"]), C.byref(best)))
        out["best_index"] = best.index
        out["best_score"] = best.score
        return out

    # -- collision ---------------------------------------------------------------------------
    def check_segments(self, layer: int, robot_radius: float, seg: np.ndarray) -> np.ndarray:
        seg = np.ascontiguousarray(seg, np.float64).reshape(-1, 6)
        free = np.zeros(len(seg), np.uint8)
        self._ck(self._L.nbvgpu_check_segments(self._h, layer, float(robot_radius), len(seg), _vp(seg), _vp(free)))
        return free

    def check_segments_device(self, layer: int, robot_radius: float, n: int, d_seg, d_free) -> None:
        self._ck(self._L.nbvgpu_check_segments_device(self._h, layer, float(robot_radius), n, _vp(d_seg), _vp(d_free)))

    def check_motion(self, layer: int, robot_radius: float, start, end) -> bool:
        s = (C.c_double * 3)(*[float(v) for v in start][:3])
        e = (C.c_double * 3)(*[float(v) for v in end][:3])
        out = C.c_int(0)
        self._ck(self._L.nbvgpu_check_motion(self._h, layer, float(robot_radius), s, e, C.byref(out)))
        return bool(out.value)

    # -- ESDF interpolation (EsdfMap::getDistanceAndGradientAtPosition, esdf_map.cc:22-52) -----
    def esdf_distance_gradient(self, esdf_layer: int, positions) -> dict:
        pos = np.ascontiguousarray(positions, np.float64).reshape(-1, 3)
        n = len(pos)
        ok = np.zeros(n, np.uint8)
        dist = np.zeros(n)
        grad = np.zeros((n, 3))
        self._ck(self._L.nbvgpu_esdf_distance_gradient(self._h, esdf_layer, n, _vp(pos), _vp(ok), _vp(dist), _vp(grad)))
        return {"ok": ok, "distance": dist, "gradient": grad}

    def esdf_distance_gradient_device(self, esdf_layer: int, n: int, d_positions, d_ok, d_distance, d_gradient) -> None:
        """device-resident buffers (anything _vp accepts: torch tensors, raw pointers); stream-ordered, no synchronisation"""
        self._ck(self._L.nbvgpu_esdf_distance_gradient_device(self._h, esdf_layer, n, _vp(d_positions), _vp(d_ok), _vp(d_distance), _vp(d_gradient)))

    # -- history-graph frontier potential (History::bfs, history.cpp:96-137) -----------------
    def frontier_potential(self, tsdf_layer: int, points, max_distance: float = 6.0) -> dict:
        """potential gain and frontier-voxel count of each vertex position (the reference hard-codes 6.0 m)."""
        pts = np.ascontiguousarray(points, np.float64).reshape(-1, 3)
        gain = np.zeros(len(pts))
        fv = np.zeros(len(pts), np.uint64)
        vis = np.zeros(len(pts), np.uint64)
        chk = np.zeros(len(pts), np.uint64)
        self._ck(self._L.nbvgpu_frontier_potential(self._h, tsdf_layer, len(pts), _vp(pts), float(max_distance), _vp(gain), _vp(fv),
                                                   _vp(vis), _vp(chk)))
        return {"gain": gain, "frontier_voxels": fv, "visited": vis, "cells_checksum": chk}

    def probe_l2_read_bandwidth(self, nbytes: int = 32 << 20, iters: int = 50) -> float:
        """GB/s of 128-bit L1-bypassing reads over an L2-resident buffer (measurement infrastructure)."""
        out = C.c_double(0.0)
        self._ck(self._L.nbvgpu_probe_l2_read_bandwidth(self._h, int(nbytes), int(iters), C.byref(out)))
        return float(out.value)

    def dense_clipped_launches(self) -> int:
        """ray-sweep launches that used the z-clipped dense volume (tests)."""
        out = C.c_uint64(0)
        self._ck(self._L.nbvgpu_debug_dense_clipped_launches(self._h, C.byref(out)))
        return int(out.value)

    def fused_launches(self) -> int:
        """ray-sweep launches that finished with the window scan fused into their last CTA (small batches; tests)."""
        out = C.c_uint64(0)
        self._ck(self._L.nbvgpu_debug_fused_launches(self._h, C.byref(out)))
        return int(out.value)

    def wide_cta_launches(self) -> int:
        """ray-sweep launches that ran as 512-thread CTAs with a half-SM dense volume (tests)."""
        out = C.c_uint64(0)
        self._ck(self._L.nbvgpu_debug_wide_cta_launches(self._h, C.byref(out)))
        return int(out.value)

    def graph_replays(self) -> int:
        """planning steps that ran as a replay of their captured CUDA graph (tests)."""
        out = C.c_uint64(0)
        self._ck(self._L.nbvgpu_debug_graph_replays(self._h, C.byref(out)))
        return int(out.value)

    def frontier_hash_redo(self) -> int:
        """vertices the dense visited set of the flood fill handed back to the hash-keyed one (tests)."""
        out = C.c_uint64(0)
        self._ck(self._L.nbvgpu_debug_frontier_hash_redo(self._h, C.byref(out)))
        return int(out.value)

    # -- introspection -----------------------------------------------------------------------
    def stats(self) -> dict:
        s = N.Stats()
        self._ck(self._L.nbvgpu_get_stats(self._h, C.byref(s)))
        return {k: int(getattr(s, k)) for k, _ in N.Stats._fields_}

    def set_kernel_variant(self, variant: int) -> None:
        self._ck(self._L.nbvgpu_set_kernel_variant(self._h, variant))

    def set_profiling(self, enable: bool) -> None:
        self._ck(self._L.nbvgpu_set_profiling(self._h, int(enable)))

    def profile_collect(self):
        """(summed ray-sweep kernel ms, launches) since the last collect (CUDA events on the context stream)."""
        ms, n = C.c_double(0.0), C.c_uint64(0)
        self._ck(self._L.nbvgpu_profile_collect(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def classifier_mismatches(self) -> int:
        out = C.c_uint64(0)
        self._ck(self._L.nbvgpu_debug_classifier_mismatches(self._h, C.byref(out)))
        return out.value


def parse_blocks(data: bytes, kind: int, voxel_size: float, voxels_per_side: int, fmt: str = "voxblox"):
    """Host-only parse of a .voxblox byte string (fmt="voxblox") or a ROS-1-serialised voxblox_msgs/Layer
    (fmt="msg") into (keys int32[n][3], words uint32[n][words per block], action).  No GPU needed."""
    L = N.lib()
    buf = np.frombuffer(data, np.uint8)
    f = 0 if fmt == "voxblox" else 1
    n, act = C.c_size_t(0), C.c_int(0)
    rc = L.nbvgpu_parse_blocks(f, _vp(buf), len(buf), kind, np.float32(voxel_size), voxels_per_side, C.byref(n), None, None, 0,
                               C.byref(act))
    if rc != 0:
        raise N.NbvGpuError(rc, "could not parse blocks")
    per = voxels_per_side ** 3 * (3 if kind == N.LAYER_TSDF else 2)
    keys = np.zeros((n.value, 3), np.int32)
    words = np.zeros((n.value, per), np.uint32)
    rc = L.nbvgpu_parse_blocks(f, _vp(buf), len(buf), kind, np.float32(voxel_size), voxels_per_side, C.byref(n), _vp(keys),
                               _vp(words), n.value, C.byref(act))
    if rc != 0:
        raise N.NbvGpuError(rc, "could not parse blocks")
    return keys, words, act.value


# ---------------------------------------------------------------------------------------------
# Reference-shaped wrappers.
# ---------------------------------------------------------------------------------------------
class LowResManager:
    """``nbveplanner::LowResManager``: the low-resolution TSDF layer the gain reads (nbvep.cpp:21-27)."""

    def __init__(self, gpu: NbvGpu, voxel_size: float, voxels_per_side: int = 16, max_blocks: int = 4096):
        self.gpu = gpu
        self.layer = gpu.layer_create(N.LAYER_TSDF, voxel_size, voxels_per_side, max_blocks)
        self._vs = np.float32(voxel_size)
        self._vps = voxels_per_side

    def getResolution(self) -> float:  # voxblox_manager.h:35
        return float(self._vs)

    def getVoxelsPerSide(self) -> int:  # voxblox_manager.h:37
        return self._vps

    def isTsdfEmpty(self) -> bool:  # voxblox_manager.h:41-43
        return self.gpu.layer_num_blocks(self.layer) == 0

    def update(self, keys, payload, packing: int = N.PACK_PLANAR, commit: bool = True) -> None:
        self.gpu.layer_update(self.layer, keys, payload, packing)
        if commit:
            self.gpu.commit()

    def clear(self) -> None:  # voxblox_manager.cpp:92-95
        self.gpu.layer_clear(self.layer)

    def getVoxelStatusByIndex(self, block_idx, voxel_idx) -> int:  # voxblox_manager.cpp:14-33
        g = np.asarray(block_idx, np.int64) * self._vps + np.asarray(voxel_idx, np.int64)
        _, _, status = self.gpu.layer_read_voxels(self.layer, g.reshape(1, 3))
        return int(status[0])


class HighResManager:
    """``nbveplanner::HighResManager``: the ESDF layer the collision check reads."""

    def __init__(self, gpu: NbvGpu, voxel_size: float, robot_radius: float, voxels_per_side: int = 16, max_blocks: int = 4096):
        self.gpu = gpu
        self.layer = gpu.layer_create(N.LAYER_ESDF, voxel_size, voxels_per_side, max_blocks)
        self.robot_radius = float(robot_radius)  # params.h:103 system/robot_radius

    def update(self, keys, payload, packing: int = N.PACK_PLANAR, commit: bool = True) -> None:
        self.gpu.layer_update(self.layer, keys, payload, packing)
        if commit:
            self.gpu.commit()

    def clear(self) -> None:  # voxblox_manager.cpp:150-154
        self.gpu.layer_clear(self.layer)

    def checkMotion(self, start, end) -> bool:  # voxblox_manager.h:89-108; True = free
        return self.gpu.check_motion(self.layer, self.robot_radius, start, end)

    def checkMotionBatch(self, segments: np.ndarray) -> np.ndarray:
        return self.gpu.check_segments(self.layer, self.robot_radius, segments)

    def getDistanceAndGradientAtPosition(self, pos):  # voxblox_manager.cpp:143-148 -> (ok, distance, gradient)
        r = self.gpu.esdf_distance_gradient(self.layer, np.asarray(pos, float).reshape(1, 3))
        return bool(r["ok"][0]), float(r["distance"][0]), r["gradient"][0]


class History:
    """The map-reading part of ``nbveplanner::History`` (history.h): ``bfs`` / ``recalculatePotential``."""

    def __init__(self, gpu: NbvGpu, manager_lowres: "LowResManager", zero_frontier_voxels: float = 10.0):
        self.gpu = gpu
        self.manager_lowres = manager_lowres
        # params.h:42,105-106: nbvep/graph/zero_frontier_voxels, default 10.0; a vertex whose potential is not
        # above it is retired (history.cpp:213) -- that decision stays with the caller
        self.zero_frontier_voxels = float(zero_frontier_voxels)

    def bfs(self, point) -> float:  # history.cpp:103-137
        return float(self.gpu.frontier_potential(self.manager_lowres.layer, np.asarray(point, float).reshape(1, 3))["gain"][0])

    def recalculatePotential(self, positions) -> np.ndarray:  # history.cpp:211-222, all vertices of a pass at once
        return self.gpu.frontier_potential(self.manager_lowres.layer, positions)["gain"]


@dataclass
class RrtGain:
    """The gain members of ``nbveplanner::RrtTree`` (rrt.h:62-64) bound to a TSDF layer + camera."""

    gpu: NbvGpu
    manager_lowres: LowResManager
    camera: CameraParams = field(default_factory=CameraParams)

    def __post_init__(self):
        self._mode = None

    def _ensure(self, sum_all: int) -> None:
        if self._mode != sum_all:
            cam = CameraParams(**{**self.camera.__dict__, "sum_all_yaws": sum_all})
            self.gpu.set_camera(cam)
            self._mode = sum_all

    def optimized_gain(self, state: np.ndarray) -> float:  # rrt.cpp:351-479
        self._ensure(0)
        return self.gpu.optimized_gain(self.manager_lowres.layer, state)

    def gain(self, state: np.ndarray) -> float:  # rrt.cpp:481-579 (state[3] = 0)
        self._ensure(1)
        return self.gpu.optimized_gain(self.manager_lowres.layer, state)

    def gain_eval(self, positions: np.ndarray, per_yaw: bool = False) -> dict:
        self._ensure(int(self.camera.sum_all_yaws))
        return self.gpu.gain_eval(self.manager_lowres.layer, positions, per_yaw=per_yaw)
