"""ctypes binding of ``oracle/libnbv_oracle.so`` (the CPU restatement of the
reference's gain / collision path).  TEST INFRASTRUCTURE ONLY -- the product
package ``nbv_exploration_planner_b200`` never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libnbv_oracle.so")


def build(force: bool = False) -> str:
    """Compile the oracle (and oracle/_ref when /root/reference exists)."""
    src = os.path.join(_HERE, "nbv_oracle.cc")
    stale = (not os.path.exists(_LIB_PATH)) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src)
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "libnbv_oracle.so"], stdout=subprocess.DEVNULL)
    if os.path.isdir("/root/reference") and not os.path.exists(os.path.join(_HERE, "_ref", "libkdtree_ref.so")):
        subprocess.check_call(["make", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


class OracleCamera(C.Structure):
    """Same field order as the ABI's nbvgpu_camera (declared independently)."""

    _fields_ = [
        ("hfov_deg", C.c_double), ("vfov_deg", C.c_double), ("near_m", C.c_double), ("far_m", C.c_double),
        ("q_CB", C.c_double * 4), ("t_CB", C.c_double * 3),
        ("bbx_min", C.c_double * 3), ("bbx_max", C.c_double * 3),
        ("gain_free", C.c_double), ("gain_occupied", C.c_double), ("gain_unmapped", C.c_double),
        ("sum_all_yaws", C.c_int32), ("reserved", C.c_int32),
    ]


class OracleExpandParams(C.Structure):
    _fields_ = [("root", C.c_double * 3), ("vicinity", C.c_double), ("extension_range", C.c_double), ("overshoot", C.c_double),
                ("robot_radius", C.c_double), ("v_max", C.c_double), ("dyaw_max", C.c_double), ("zero_gain", C.c_double)]


def _ptr(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.nbvo_create.restype = C.c_void_p
        L.nbvo_destroy.argtypes = [C.c_void_p]
        L.nbvo_layer_create.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_int]
        L.nbvo_layer_update.argtypes = [C.c_void_p, C.c_int, C.c_size_t, C.POINTER(C.c_int32), C.POINTER(C.c_float)]
        L.nbvo_layer_remove.argtypes = [C.c_void_p, C.c_int, C.c_size_t, C.POINTER(C.c_int32)]
        L.nbvo_layer_clear.argtypes = [C.c_void_p, C.c_int]
        L.nbvo_layer_num_blocks.argtypes = [C.c_void_p, C.c_int]
        L.nbvo_layer_num_blocks.restype = C.c_size_t
        L.nbvo_set_camera.argtypes = [C.c_void_p, C.POINTER(OracleCamera)]
        L.nbvo_gain_eval.argtypes = [C.c_void_p, C.c_int, C.c_size_t, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                     C.POINTER(C.c_int32), C.POINTER(C.c_double), C.POINTER(C.c_uint64),
                                     C.POINTER(C.c_uint32), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.c_int]
        L.nbvo_check_segments.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_size_t, C.POINTER(C.c_double),
                                          C.POINTER(C.c_uint8), C.POINTER(C.c_uint64), C.c_int]
        L.nbvo_expand_tree.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(OracleExpandParams), C.c_size_t, C.POINTER(C.c_double),
                                       C.POINTER(C.c_double), C.c_size_t, C.POINTER(C.c_double), C.POINTER(C.c_uint8),
                                       C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_double), C.POINTER(C.c_double),
                                       C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int64), C.c_int]
        L.nbvo_expand_tree_seq.argtypes = L.nbvo_expand_tree.argtypes
        L.nbvo_cast_ray.argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_int64), C.c_size_t]
        L.nbvo_cast_ray.restype = C.c_size_t
        L.nbvo_yaw_frame.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_double),
                                     C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int32)]
        L.nbvo_grid_index_from_point.argtypes = [C.POINTER(C.c_float), C.c_float, C.POINTER(C.c_int64)]
        L.nbvo_block_and_local.argtypes = [C.POINTER(C.c_int64), C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        L.nbvo_linear_index.argtypes = [C.c_int] * 4
        L.nbvo_linear_index.restype = C.c_uint64
        L.nbvo_block_hash.argtypes = [C.c_int] * 3
        L.nbvo_block_hash.restype = C.c_uint64
        L.nbvo_voxel_status.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int64)]
        L.nbvo_esdf_distance.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int64)]
        L.nbvo_esdf_distance.restype = C.c_float
        _lib = L
    return _lib


def camera_struct(cam: dict) -> OracleCamera:
    s = OracleCamera()
    for k in ("hfov_deg", "vfov_deg", "near_m", "far_m", "gain_free", "gain_occupied", "gain_unmapped"):
        setattr(s, k, float(cam[k]))
    for k, n in (("q_CB", 4), ("t_CB", 3), ("bbx_min", 3), ("bbx_max", 3)):
        getattr(s, k)[:] = [float(v) for v in cam[k]][:n]
    s.sum_all_yaws = int(cam.get("sum_all_yaws", 0))
    return s


class Oracle:
    """One oracle context: layers + camera, CPU evaluation of gain / collision."""

    TSDF, ESDF = 0, 1

    def __init__(self):
        self._l = lib()
        self._c = C.c_void_p(self._l.nbvo_create())

    def __del__(self):
        if getattr(self, "_c", None):
            self._l.nbvo_destroy(self._c)
            self._c = None

    def layer_create(self, kind: int, voxel_size: float, vps: int = 16) -> int:
        r = self._l.nbvo_layer_create(self._c, kind, np.float32(voxel_size), vps)
        if r < 0:
            raise ValueError("bad layer parameters")
        return r

    def layer_update(self, layer: int, keys: np.ndarray, data: np.ndarray) -> None:
        keys = np.ascontiguousarray(keys, np.int32)
        data = np.ascontiguousarray(data, np.float32)
        self._l.nbvo_layer_update(self._c, layer, len(keys), _ptr(keys, C.c_int32), _ptr(data, C.c_float))

    def layer_remove(self, layer: int, keys: np.ndarray) -> None:
        keys = np.ascontiguousarray(keys, np.int32)
        self._l.nbvo_layer_remove(self._c, layer, len(keys), _ptr(keys, C.c_int32))

    def layer_clear(self, layer: int) -> None:
        self._l.nbvo_layer_clear(self._c, layer)

    def num_blocks(self, layer: int) -> int:
        return self._l.nbvo_layer_num_blocks(self._c, layer)

    def set_camera(self, cam: dict) -> None:
        s = camera_struct(cam)
        self._l.nbvo_set_camera(self._c, C.byref(s))

    def gain_eval(self, layer: int, pos: np.ndarray, per_yaw: bool = False, threads: int = 1) -> dict:
        pos = np.ascontiguousarray(pos, np.float64).reshape(-1, 3)
        n = len(pos)
        out = {
            "gain": np.zeros(n, np.float64), "yaw_deg": np.zeros(n, np.int32), "yaw": np.zeros(n, np.float64),
            "counts": np.zeros((n, 3), np.uint64), "steps": np.zeros(n, np.uint64), "rays": np.zeros(n, np.uint64),
            "per_yaw": np.zeros((n, 360, 3), np.uint32) if per_yaw else None,
        }
        r = self._l.nbvo_gain_eval(self._c, layer, n, _ptr(pos, C.c_double), _ptr(out["gain"], C.c_double),
                                   _ptr(out["yaw_deg"], C.c_int32), _ptr(out["yaw"], C.c_double),
                                   _ptr(out["counts"], C.c_uint64), _ptr(out["per_yaw"], C.c_uint32),
                                   _ptr(out["steps"], C.c_uint64), _ptr(out["rays"], C.c_uint64), threads)
        if r != 0:
            raise RuntimeError(f"oracle gain_eval failed: {r}")
        return out

    def check_segments(self, layer: int, radius: float, seg: np.ndarray, threads: int = 1) -> dict:
        seg = np.ascontiguousarray(seg, np.float64).reshape(-1, 6)
        n = len(seg)
        free = np.zeros(n, np.uint8)
        steps = np.zeros(n, np.uint64)
        r = self._l.nbvo_check_segments(self._c, layer, float(radius), n, _ptr(seg, C.c_double),
                                        _ptr(free, C.c_uint8), _ptr(steps, C.c_uint64), threads)
        if r != 0:
            raise RuntimeError(f"oracle check_segments failed: {r}")
        return {"free": free, "steps": steps}

    def expand_tree(self, tsdf_layer: int, esdf_layer: int, params: dict, node_state, node_distance, uniforms, threads: int = 1,
                    sequential: bool = False) -> dict:
        """sequential: the reference's own semantics -- every sample sees the nodes created by the samples before it
        (nbvo_expand_tree_seq); otherwise parents are the given nodes only."""
        P = OracleExpandParams()
        P.root[:] = [float(v) for v in params["root"]]
        for k in ("vicinity", "extension_range", "overshoot", "robot_radius", "v_max", "dyaw_max", "zero_gain"):
            setattr(P, k, float(params[k]))
        ns = np.ascontiguousarray(node_state, np.float64).reshape(-1, 4)
        nd = np.ascontiguousarray(node_distance, np.float64)
        u = np.ascontiguousarray(uniforms, np.float64).reshape(-1, 3)
        n = len(u)
        out = {"status": np.zeros(n, np.uint8), "pos": np.zeros((n, 3)), "parent": np.zeros(n, np.int32), "gain": np.zeros(n),
               "yaw": np.zeros(n), "distance": np.zeros(n), "score": np.zeros(n)}
        best = C.c_int64(-1)
        fn = self._l.nbvo_expand_tree_seq if sequential else self._l.nbvo_expand_tree
        r = fn(self._c, tsdf_layer, esdf_layer, C.byref(P), len(ns), _ptr(ns, C.c_double), _ptr(nd, C.c_double),
                                     n, _ptr(u, C.c_double), _ptr(out["status"], C.c_uint8), _ptr(out["pos"], C.c_double),
                                     _ptr(out["parent"], C.c_int32), _ptr(out["gain"], C.c_double), _ptr(out["yaw"], C.c_double),
                                     _ptr(out["distance"], C.c_double), _ptr(out["score"], C.c_double), C.byref(best), threads)
        if r != 0:
            raise RuntimeError(f"oracle expand_tree failed: {r}")
        out["best_index"] = best.value
        return out

    def yaw_frame(self, layer: int, pos, yaw_deg: int) -> dict:
        pos = np.ascontiguousarray(pos, np.float64)
        planes = np.zeros((6, 4)); cam = np.zeros(3); pp = np.zeros((3, 3)); um = C.c_int32(0)
        self._l.nbvo_yaw_frame(self._c, layer, _ptr(pos, C.c_double), yaw_deg, _ptr(planes, C.c_double),
                               _ptr(cam, C.c_double), _ptr(pp, C.c_double), C.byref(um))
        return {"planes": planes, "cam": cam, "pp": pp, "u_max": um.value}

    def frontier_potential(self, layer: int, points, max_distance: float = 6.0, threads: int = 1) -> dict:
        """History::bfs (history.cpp:103-137) per vertex: potential gain, frontier-voxel count, visited cells."""
        pts = np.ascontiguousarray(points, np.float64).reshape(-1, 3)
        n = len(pts)
        gain = np.zeros(n); fv = np.zeros(n, np.uint64); vis = np.zeros(n, np.uint64); chk = np.zeros(n, np.uint64)
        self._l.nbvo_frontier_potential.argtypes = [C.c_void_p, C.c_int, C.c_size_t, C.POINTER(C.c_double), C.c_double,
                                                    C.POINTER(C.c_double), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64),
                                                    C.POINTER(C.c_uint64), C.c_int]
        rc = self._l.nbvo_frontier_potential(self._c, layer, n, _ptr(pts, C.c_double), float(max_distance), _ptr(gain, C.c_double),
                                             _ptr(fv, C.c_uint64), _ptr(vis, C.c_uint64), _ptr(chk, C.c_uint64), threads)
        if rc != 0:
            raise RuntimeError(f"nbvo_frontier_potential -> {rc}")
        return {"gain": gain, "frontier_voxels": fv, "visited": vis, "cells_checksum": chk}

    def layer_update_esdf_observed(self, layer: int, keys, dist, observed) -> None:
        keys = np.ascontiguousarray(keys, np.int32).reshape(-1, 3)
        dist = np.ascontiguousarray(dist, np.float32)
        observed = np.ascontiguousarray(observed, np.uint8)
        self._l.nbvo_layer_update_esdf_observed.argtypes = [C.c_void_p, C.c_int, C.c_size_t, C.POINTER(C.c_int32), C.POINTER(C.c_float),
                                                            C.POINTER(C.c_uint8)]
        rc = self._l.nbvo_layer_update_esdf_observed(self._c, layer, len(keys), _ptr(keys, C.c_int32), _ptr(dist, C.c_float), _ptr(observed, C.c_uint8))
        if rc != 0:
            raise RuntimeError(f"nbvo_layer_update_esdf_observed -> {rc}")

    def distance_and_gradient(self, layer: int, positions, threads: int = 1) -> dict:
        """EsdfMap::getDistanceAndGradientAtPosition (esdf_map.cc:22-52) per query."""
        pos = np.ascontiguousarray(positions, np.float64).reshape(-1, 3)
        n = len(pos)
        ok = np.zeros(n, np.uint8); dist = np.zeros(n); grad = np.zeros((n, 3))
        self._l.nbvo_distance_and_gradient.argtypes = [C.c_void_p, C.c_int, C.c_size_t, C.POINTER(C.c_double), C.POINTER(C.c_uint8),
                                                       C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int]
        rc = self._l.nbvo_distance_and_gradient(self._c, layer, n, _ptr(pos, C.c_double), _ptr(ok, C.c_uint8), _ptr(dist, C.c_double),
                                                _ptr(grad, C.c_double), threads)
        if rc != 0:
            raise RuntimeError(f"nbvo_distance_and_gradient -> {rc}")
        return {"ok": ok, "distance": dist, "gradient": grad}

    def voxel_status_at(self, layer: int, p) -> int:
        p = np.ascontiguousarray(p, np.float64)
        self._l.nbvo_voxel_status_at.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double)]
        return self._l.nbvo_voxel_status_at(self._c, layer, _ptr(p, C.c_double))

    def voxel_status(self, layer: int, g) -> int:
        g = np.ascontiguousarray(g, np.int64)
        return self._l.nbvo_voxel_status(self._c, layer, _ptr(g, C.c_int64))

    def esdf_distance(self, layer: int, g) -> float:
        g = np.ascontiguousarray(g, np.int64)
        return self._l.nbvo_esdf_distance(self._c, layer, _ptr(g, C.c_int64))


def cast_ray(start, end, cap: int = 1 << 16) -> np.ndarray:
    s = np.ascontiguousarray(start, np.float32)
    e = np.ascontiguousarray(end, np.float32)
    out = np.zeros((cap, 3), np.int64)
    n = lib().nbvo_cast_ray(_ptr(s, C.c_float), _ptr(e, C.c_float), _ptr(out, C.c_int64), cap)
    return out[: min(n, cap)].copy()


def grid_index_from_point(p, grid_size_inv: float) -> np.ndarray:
    p = np.ascontiguousarray(p, np.float32)
    out = np.zeros(3, np.int64)
    lib().nbvo_grid_index_from_point(_ptr(p, C.c_float), np.float32(grid_size_inv), _ptr(out, C.c_int64))
    return out


def block_and_local(g, vps: int):
    g = np.ascontiguousarray(g, np.int64)
    b = np.zeros(3, np.int32); l = np.zeros(3, np.int32)
    lib().nbvo_block_and_local(_ptr(g, C.c_int64), vps, _ptr(b, C.c_int32), _ptr(l, C.c_int32))
    return b, l


def linear_index(vps: int, x: int, y: int, z: int) -> int:
    return lib().nbvo_linear_index(vps, x, y, z)


def block_hash(x: int, y: int, z: int) -> int:
    return lib().nbvo_block_hash(x, y, z)
