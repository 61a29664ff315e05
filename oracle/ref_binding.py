"""ctypes binding of ``oracle/_ref/libnbv_ref.so`` -- the REFERENCE'S OWN hot-path sources compiled where
they lie under /root/reference (oracle/Makefile target ``ref``, oracle/ref_wrap.cc).  TEST INFRASTRUCTURE:
pins the oracle restatement (tests/test_ref_pinning.py) and serves as bench.py's reference arm.

The reference's gain functions keep ``static const`` locals (voxel size, voxels per side, half hfov:
rrt.cpp:352-356,452), fixed at the first call in a process, so every distinct configuration gets its own
private copy of the shared object (``RefLib()``)."""
from __future__ import annotations

import ctypes as C
import os
import shutil
import tempfile

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_LIB = os.path.join(_HERE, "_ref", "libnbv_ref.so")


def available() -> bool:
    return os.path.exists(REF_LIB)


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


class RefLib:
    """A private copy of libnbv_ref.so (fresh function-local statics)."""

    def __init__(self):
        if not available():
            raise FileNotFoundError(REF_LIB)
        fd, path = tempfile.mkstemp(prefix="nbv_ref_", suffix=".so")
        os.close(fd)
        shutil.copyfile(REF_LIB, path)
        self.path = path
        L = C.CDLL(path)
        os.unlink(path)          # the mapping stays valid
        L.nbvr_create.restype = C.c_void_p
        L.nbvr_create.argtypes = [C.c_float, C.c_int, C.c_float, C.c_int]
        L.nbvr_destroy.argtypes = [C.c_void_p]
        dp, fp, ip, lp, up = (C.POINTER(t) for t in (C.c_double, C.c_float, C.c_int32, C.c_int64, C.c_uint8))
        L.nbvr_tsdf_update.argtypes = [C.c_void_p, C.c_size_t, ip, fp]
        L.nbvr_esdf_update.argtypes = [C.c_void_p, C.c_size_t, ip, fp]
        L.nbvr_tsdf_remove.argtypes = [C.c_void_p, C.c_size_t, ip]
        L.nbvr_set_camera.argtypes = [C.c_void_p] + [C.c_double] * 4 + [dp] * 4 + [C.c_double] * 4
        L.nbvr_gain_eval.argtypes = [C.c_void_p, C.c_size_t, dp, C.c_int, dp, dp]
        L.nbvr_check_segments.argtypes = [C.c_void_p, C.c_size_t, dp, up]
        L.nbvr_cast_ray.argtypes = [fp, fp, lp, C.c_size_t]
        L.nbvr_cast_ray.restype = C.c_size_t
        L.nbvr_yaw_frame.argtypes = [C.c_void_p, dp, C.c_int, dp, dp, dp]
        L.nbvr_in_view.argtypes = [C.c_void_p, dp]
        L.nbvr_block_and_local.argtypes = [lp, C.c_int, ip, ip]
        L.nbvr_grid_index_from_point.argtypes = [fp, C.c_float, lp]
        L.nbvr_block_hash.argtypes = [C.c_int] * 3
        L.nbvr_block_hash.restype = C.c_uint64
        L.nbvr_voxel_status.argtypes = [C.c_void_p, ip, ip]
        self.L = L

    def cast_ray(self, start, end, cap: int = 1 << 16) -> np.ndarray:
        s = np.ascontiguousarray(start, np.float32)
        e = np.ascontiguousarray(end, np.float32)
        out = np.zeros((cap, 3), np.int64)
        n = self.L.nbvr_cast_ray(_p(s, C.c_float), _p(e, C.c_float), _p(out, C.c_int64), cap)
        return out[:min(n, cap)].copy()

    def block_and_local(self, g, vps: int):
        g = np.ascontiguousarray(g, np.int64)
        b, l = np.zeros(3, np.int32), np.zeros(3, np.int32)
        self.L.nbvr_block_and_local(_p(g, C.c_int64), vps, _p(b, C.c_int32), _p(l, C.c_int32))
        return b, l

    def grid_index_from_point(self, p, inv: float) -> np.ndarray:
        p = np.ascontiguousarray(p, np.float32)
        out = np.zeros(3, np.int64)
        self.L.nbvr_grid_index_from_point(_p(p, C.c_float), np.float32(inv), _p(out, C.c_int64))
        return out

    def block_hash(self, x: int, y: int, z: int) -> int:
        return self.L.nbvr_block_hash(x, y, z)


class Reference:
    """One planner instance of the reference: low-res TSDF manager, high-res ESDF manager, camera, tree."""

    def __init__(self, tsdf_voxel_size, tsdf_vps, esdf_voxel_size=None, esdf_vps=None, lib: RefLib | None = None):
        self.lib = lib or RefLib()
        self.L = self.lib.L
        self._c = C.c_void_p(self.L.nbvr_create(np.float32(tsdf_voxel_size), tsdf_vps,
                                                np.float32(esdf_voxel_size or tsdf_voxel_size), esdf_vps or tsdf_vps))
        self.sum_all = 0

    def __del__(self):
        if getattr(self, "_c", None):
            self.L.nbvr_destroy(self._c)
            self._c = None

    def tsdf_update(self, keys, data):
        keys = np.ascontiguousarray(keys, np.int32); data = np.ascontiguousarray(data, np.float32)
        self.L.nbvr_tsdf_update(self._c, len(keys), _p(keys, C.c_int32), _p(data, C.c_float))

    def esdf_update(self, keys, data):
        keys = np.ascontiguousarray(keys, np.int32); data = np.ascontiguousarray(data, np.float32)
        self.L.nbvr_esdf_update(self._c, len(keys), _p(keys, C.c_int32), _p(data, C.c_float))

    def tsdf_remove(self, keys):
        keys = np.ascontiguousarray(keys, np.int32)
        self.L.nbvr_tsdf_remove(self._c, len(keys), _p(keys, C.c_int32))

    def set_camera(self, cam: dict, robot_radius: float = 0.5):
        arr = {k: np.ascontiguousarray(cam[k], np.float64) for k in ("q_CB", "t_CB", "bbx_min", "bbx_max")}
        self.L.nbvr_set_camera(self._c, float(cam["hfov_deg"]), float(cam["vfov_deg"]), float(cam["near_m"]), float(cam["far_m"]),
                               _p(arr["q_CB"], C.c_double), _p(arr["t_CB"], C.c_double), _p(arr["bbx_min"], C.c_double),
                               _p(arr["bbx_max"], C.c_double), float(cam["gain_free"]), float(cam["gain_occupied"]),
                               float(cam["gain_unmapped"]), float(robot_radius))
        self.sum_all = int(cam.get("sum_all_yaws", 0))

    def gain_eval(self, pos):
        pos = np.ascontiguousarray(pos, np.float64).reshape(-1, 3)
        gain, yaw = np.zeros(len(pos)), np.zeros(len(pos))
        self.L.nbvr_gain_eval(self._c, len(pos), _p(pos, C.c_double), self.sum_all, _p(gain, C.c_double), _p(yaw, C.c_double))
        return {"gain": gain, "yaw": yaw}

    def check_segments(self, seg):
        seg = np.ascontiguousarray(seg, np.float64).reshape(-1, 6)
        free = np.zeros(len(seg), np.uint8)
        self.L.nbvr_check_segments(self._c, len(seg), _p(seg, C.c_double), _p(free, C.c_uint8))
        return free

    def yaw_frame(self, pos, yaw_deg: int):
        pos = np.ascontiguousarray(pos, np.float64)
        planes, cam, pp = np.zeros((6, 4)), np.zeros(3), np.zeros((3, 3))
        self.L.nbvr_yaw_frame(self._c, _p(pos, C.c_double), yaw_deg, _p(planes, C.c_double), _p(cam, C.c_double), _p(pp, C.c_double))
        return {"planes": planes, "cam": cam, "pp": pp}

    def frontier_potential(self, points):
        """The reference's own History::bfs (history.cpp:103-137; 6.0 m hard-coded there)."""
        pts = np.ascontiguousarray(points, np.float64).reshape(-1, 3)
        gain = np.zeros(len(pts))
        self.L.nbvr_frontier_potential.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        self.L.nbvr_frontier_potential.restype = None
        self.L.nbvr_frontier_potential(self._c, len(pts), _p(pts, C.c_double), _p(gain, C.c_double))
        return gain

    def iterate_stream(self, params: dict, node_state, node_distance, uniforms) -> dict:
        """The reference's own RrtTree::iterate (rrt.cpp:163-264) with its own kd-tree, called until the stream of uniform
        triples is used up (nbvr_iterate_stream); same outputs as the oracle's expand_tree(sequential=True)."""
        ns = np.ascontiguousarray(node_state, np.float64).reshape(-1, 4)
        nd = np.ascontiguousarray(node_distance, np.float64)
        u = np.ascontiguousarray(uniforms, np.float64).reshape(-1, 3)
        n = len(u)
        out = {"status": np.zeros(n, np.uint8), "pos": np.zeros((n, 3)), "parent": np.zeros(n, np.int32), "gain": np.zeros(n),
               "yaw": np.zeros(n), "distance": np.zeros(n), "score": np.zeros(n)}
        best = C.c_int64(-1)
        root = np.ascontiguousarray(params["root"], np.float64)
        dp = C.POINTER(C.c_double)
        f = self.L.nbvr_iterate_stream
        f.argtypes = [C.c_void_p, dp] + [C.c_double] * 6 + [C.c_size_t, dp, dp, C.c_size_t, dp, C.POINTER(C.c_uint8), dp,
                                                            C.POINTER(C.c_int32), dp, dp, dp, dp, C.POINTER(C.c_int64)]
        rc = f(self._c, _p(root, C.c_double), float(params["vicinity"]), float(params["extension_range"]), float(params["overshoot"]),
               float(params["v_max"]), float(params["dyaw_max"]), float(params["zero_gain"]), len(ns), _p(ns, C.c_double), _p(nd, C.c_double),
               n, _p(u, C.c_double), _p(out["status"], C.c_uint8), _p(out["pos"], C.c_double), _p(out["parent"], C.c_int32),
               _p(out["gain"], C.c_double), _p(out["yaw"], C.c_double), _p(out["distance"], C.c_double), _p(out["score"], C.c_double),
               C.byref(best))
        if rc != 0:
            raise RuntimeError(f"nbvr_iterate_stream failed: {rc}")
        out["best_index"] = best.value
        return out

    def voxel_status_at(self, p) -> int:
        p = np.ascontiguousarray(p, np.float64)
        self.L.nbvr_voxel_status_at.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        return self.L.nbvr_voxel_status_at(self._c, _p(p, C.c_double))

    def esdf_update_observed(self, keys, esdf, observed) -> None:
        keys = np.ascontiguousarray(keys, np.int32).reshape(-1, 3)
        esdf = np.ascontiguousarray(esdf, np.float32)
        observed = np.ascontiguousarray(observed, np.uint8)
        self.L.nbvr_esdf_update_observed.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_int32), C.POINTER(C.c_float), C.POINTER(C.c_uint8)]
        self.L.nbvr_esdf_update_observed.restype = None
        self.L.nbvr_esdf_update_observed(self._c, len(keys), _p(keys, C.c_int32), _p(esdf, C.c_float), _p(observed, C.c_uint8))

    def distance_and_gradient(self, positions) -> dict:
        """The reference's own EsdfMap::getDistanceAndGradientAtPosition (esdf_map.cc:22-52) per query."""
        pos = np.ascontiguousarray(positions, np.float64).reshape(-1, 3)
        n = len(pos)
        ok = np.zeros(n, np.uint8); dist = np.zeros(n); grad = np.zeros((n, 3))
        self.L.nbvr_distance_and_gradient.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_double), C.POINTER(C.c_uint8),
                                                      C.POINTER(C.c_double), C.POINTER(C.c_double)]
        self.L.nbvr_distance_and_gradient.restype = None
        self.L.nbvr_distance_and_gradient(self._c, n, _p(pos, C.c_double), _p(ok, C.c_uint8), _p(dist, C.c_double), _p(grad, C.c_double))
        return {"ok": ok, "distance": dist, "gradient": grad}

    def voxel_status(self, block, voxel) -> int:
        b = np.ascontiguousarray(block, np.int32); v = np.ascontiguousarray(voxel, np.int32)
        return self.L.nbvr_voxel_status(self._c, _p(b, C.c_int32), _p(v, C.c_int32))
