"""Synthetic inputs for the BASELINE.json configurations (maps, candidate sets, replay streams).

Input data only -- nothing here evaluates gain or collision.  Maps follow the reference's own
fixture recipe (voxblox SimulationWorld, simulation_world_inl.h:13-86: analytic SDF objects sampled
at voxel centres, TSDF truncated at 4 voxels, weight 1 where observed); the heavy lifting is
``csrc/synth_gen.cc``.  Candidates follow the sampling rule of ``RrtTree::iterate``
(nbveplanner/src/rrt.cpp:172-212) with fixed seeds instead of the reference's clock seed.
"""
from __future__ import annotations

import ctypes as C
import hashlib
import math
import os
from dataclasses import dataclass, field
from typing import Callable, List, Optional

import numpy as np

from . import _native as N
from .planner import CameraParams

SPHERE, CUBE, PLANE, CYLINDER = 0, 1, 2, 3


@dataclass
class SynthConfig:
    name: str
    bounds_min: tuple
    bounds_max: tuple
    voxel_size: float
    far_m: float
    n_candidates: int
    world: str                 # "pillars" | "office" | "outdoor"
    seed_map: int
    seed_cand: int
    root: tuple
    obs_radius: float = 6.0
    traj_points: int = 12
    traj_step: float = 4.0
    vicinity: float = 5.0
    extension_range: float = 1.0
    overshoot: float = 0.5
    robot_radius: float = 0.5 * math.sqrt(0.5 ** 2 + 0.5 ** 2 + 0.3 ** 2)  # half diagonal of 0.5 x 0.5 x 0.3 m
    hfov_deg: float = 90.0
    vfov_deg: float = 60.0
    near_m: float = 0.1
    pitch_deg: float = 15.0
    vps: int = 16
    with_esdf: bool = True


# the five BASELINE.json configurations.
CONFIGS = {
    "cfg1": SynthConfig("cfg1", (0, 0, 0), (20, 20, 5), 0.15, 5.0, 100, "pillars", 1001, 2001, (10.0, 10.0, 1.5), with_esdf=True),
    "cfg2": SynthConfig("cfg2", (0, 0, 0), (20, 20, 5), 0.15, 5.0, 1000, "pillars", 1001, 2002, (10.0, 10.0, 1.5)),
}
CONFIGS["cfg3"] = SynthConfig("cfg3", (0, 0, 0), (50, 50, 10), 0.10, 5.0, 10000, "office", 1003, 2003, (27.5, 27.5, 1.5),
                              traj_points=60, traj_step=5.0, vicinity=12.0)
CONFIGS["cfg4"] = SynthConfig("cfg4", (0, 0, 0), (100, 100, 20), 0.05, 8.0, 10000, "outdoor", 1004, 2004, (50.0, 50.0, 2.0),
                              obs_radius=8.0, traj_points=160, traj_step=6.0, vicinity=20.0)
CONFIGS["cfg5"] = SynthConfig("cfg5", (0, 0, 0), (20, 20, 5), 0.15, 5.0, 5000, "pillars", 1001, 2005, (10.0, 10.0, 1.5),
                              obs_radius=5.0, traj_points=200, traj_step=0.6)
# Small cases for tests.
CONFIGS["tiny"] = SynthConfig("tiny", (0, 0, 0), (8, 8, 3), 0.2, 3.0, 16, "pillars", 1100, 2100, (4.0, 4.0, 1.2),
                              obs_radius=3.5, traj_points=4, traj_step=2.0, vicinity=2.5)


def camera_for(c: SynthConfig, **kw) -> CameraParams:
    """Camera of a configuration: README default mount (0.1 m forward, pitched about +y), exploration
    box = map bounds, gains (free, occupied, unmapped) = (0, 0, 1) unless overridden."""
    return CameraParams.mount(pitch_deg=c.pitch_deg, t_BC=(0.1, 0.0, 0.0), hfov_deg=c.hfov_deg, vfov_deg=c.vfov_deg,
                              near_m=c.near_m, far_m=c.far_m, bbx_min=c.bounds_min, bbx_max=c.bounds_max, **kw)


@dataclass
class SynthMap:
    cfg: SynthConfig
    voxel_size: np.float32
    vps: int
    keys: np.ndarray                 # [nb][3] int32
    tsdf: np.ndarray                 # [nb][vps^3][2] float32
    esdf: Optional[np.ndarray]       # [nb][vps^3] float32
    objects: np.ndarray              # [no][7] float64
    trajectory: np.ndarray           # [nt][3] float64

    def camera(self, **kw) -> CameraParams:
        return camera_for(self.cfg, **kw)

    def sha256(self) -> str:
        h = hashlib.sha256()
        for a in (self.keys, self.tsdf, self.esdf):
            if a is not None:
                h.update(np.ascontiguousarray(a).tobytes())
        return h.hexdigest()


def _world_objects(cfg: SynthConfig, rng: np.random.Generator) -> np.ndarray:
    lo, hi = np.array(cfg.bounds_min, float), np.array(cfg.bounds_max, float)
    root = np.array(cfg.root, float)
    objs: List[List[float]] = []
    objs.append([PLANE, 0, 0, lo[2], 0, 0, 1])                      # floor / ground
    if cfg.world in ("pillars", "office"):
        objs.append([PLANE, 0, 0, hi[2], 0, 0, -1])                 # ceiling
    if cfg.world == "pillars":
        n = 0
        while n < 14:
            c = np.array([rng.uniform(lo[0] + 1, hi[0] - 1), rng.uniform(lo[1] + 1, hi[1] - 1), 0.0])
            if np.hypot(*(c[:2] - root[:2])) < 2.5:
                continue
            if n % 2 == 0:
                h = rng.uniform(2.0, hi[2])
                objs.append([CYLINDER, c[0], c[1], lo[2] + h / 2, rng.uniform(0.3, 0.8), h, 0])
            else:
                s = rng.uniform(0.6, 2.5, 3)
                objs.append([CUBE, c[0], c[1], lo[2] + s[2] / 2, s[0], s[1], s[2]])
            n += 1
        # two long walls with gaps
        objs.append([CUBE, 6.0, 14.0, 2.5, 8.0, 0.3, 5.0])
        objs.append([CUBE, 15.0, 5.0, 2.5, 0.3, 7.0, 5.0])
    elif cfg.world == "office":
        # walls = cube slabs on a 5 m grid with 1.2 m door gaps (each 5 m cell side: two slabs)
        t = 0.2
        H = hi[2] - lo[2]
        for gx in np.arange(lo[0] + 5, hi[0], 5.0):
            for cy in np.arange(lo[1], hi[1], 5.0):
                door = cy + rng.uniform(1.0, 3.0)
                for (a, b) in ((cy, door), (door + 1.2, cy + 5.0)):
                    if b - a > 0.05:
                        objs.append([CUBE, gx, (a + b) / 2, lo[2] + H / 2, t, b - a, H])
        for gy in np.arange(lo[1] + 5, hi[1], 5.0):
            for cx in np.arange(lo[0], hi[0], 5.0):
                door = cx + rng.uniform(1.0, 3.0)
                for (a, b) in ((cx, door), (door + 1.2, cx + 5.0)):
                    if b - a > 0.05:
                        objs.append([CUBE, (a + b) / 2, gy, lo[2] + H / 2, b - a, t, H])
    elif cfg.world == "outdoor":
        n = 0
        while n < 260:
            c = np.array([rng.uniform(lo[0] + 2, hi[0] - 2), rng.uniform(lo[1] + 2, hi[1] - 2)])
            if np.hypot(*(c - root[:2])) < 4.0:
                continue
            if n % 3 != 0:
                h = rng.uniform(3.0, 15.0)
                objs.append([CYLINDER, c[0], c[1], lo[2] + h / 2, rng.uniform(0.3, 1.5), h, 0])
            else:
                s = np.array([rng.uniform(2, 8), rng.uniform(2, 8), rng.uniform(2, 10)])
                objs.append([CUBE, c[0], c[1], lo[2] + s[2] / 2, s[0], s[1], s[2]])
            n += 1
    return np.array(objs, np.float64)


def _trajectory(cfg: SynthConfig, rng: np.random.Generator) -> np.ndarray:
    lo, hi = np.array(cfg.bounds_min, float), np.array(cfg.bounds_max, float)
    p = np.array(cfg.root, float)
    pts = [p.copy()]
    heading = rng.uniform(0, 2 * math.pi)
    for _ in range(cfg.traj_points - 1):
        heading += rng.normal(0.0, 0.6)
        step = np.array([math.cos(heading), math.sin(heading), 0.0]) * cfg.traj_step
        q = p + step
        for k in range(2):  # reflect at the walls
            if q[k] < lo[k] + 1.0 or q[k] > hi[k] - 1.0:
                heading = math.pi - heading if k == 0 else -heading
                q = p + np.array([math.cos(heading), math.sin(heading), 0.0]) * cfg.traj_step
        q[:2] = np.clip(q[:2], lo[:2] + 1.0, hi[:2] - 1.0)
        q[2] = np.clip(p[2] + rng.normal(0, 0.15), lo[2] + 0.8, min(hi[2] - 0.8, lo[2] + 3.0))
        p = q
        pts.append(p.copy())
    return np.array(pts)


def blocks_touching(voxel_size: float, vps: int, centres: np.ndarray, radius: float, bounds_min, bounds_max) -> np.ndarray:
    """Block indices whose cube intersects a sphere of `radius` around any centre, clipped to the bounds."""
    bs = float(np.float32(voxel_size)) * vps
    lo = np.floor(np.array(bounds_min, float) / bs).astype(np.int64)
    hi = np.floor((np.array(bounds_max, float) - 1e-9) / bs).astype(np.int64)
    found = set()
    for c in np.atleast_2d(centres):
        a = np.maximum(np.floor((c - radius) / bs).astype(np.int64), lo)
        b = np.minimum(np.floor((c + radius) / bs).astype(np.int64), hi)
        if np.any(b < a):
            continue
        gx, gy, gz = np.meshgrid(np.arange(a[0], b[0] + 1), np.arange(a[1], b[1] + 1), np.arange(a[2], b[2] + 1), indexing="ij")
        idx = np.stack([gx.ravel(), gy.ravel(), gz.ravel()], 1)
        near = np.clip(c, idx * bs, (idx + 1) * bs)
        keep = np.linalg.norm(near - c, axis=1) <= radius
        found.update(map(tuple, idx[keep]))
    out = np.array(sorted(found), np.int32).reshape(-1, 3)
    return out


def generate_blocks(cfg: SynthConfig, keys: np.ndarray, objects: np.ndarray, trajectory: np.ndarray,
                    want_tsdf: bool = True, want_esdf: bool = True, threads: Optional[int] = None):
    """TSDF / ESDF payloads of `keys` given the observed region spanned by `trajectory`."""
    S = N.synth_lib()
    vs = np.float32(cfg.voxel_size)
    nvox = cfg.vps ** 3
    keys = np.ascontiguousarray(keys, np.int32)
    nb = len(keys)
    tsdf = np.empty((nb, nvox, 2), np.float32) if want_tsdf else None
    esdf = np.empty((nb, nvox), np.float32) if want_esdf else None
    objects = np.ascontiguousarray(objects, np.float64)
    traj = np.ascontiguousarray(trajectory, np.float64)
    bmin = np.array(cfg.bounds_min, np.float64)
    bmax = np.array(cfg.bounds_max, np.float64)
    f32p, f64p, i32p = C.POINTER(C.c_float), C.POINTER(C.c_double), C.POINTER(C.c_int32)
    S.nbvsynth_generate(vs, cfg.vps, nb, keys.ctypes.data_as(i32p), len(objects), objects.ctypes.data_as(f64p), len(traj),
                        traj.ctypes.data_as(f64p), float(cfg.obs_radius), bmin.ctypes.data_as(f64p), bmax.ctypes.data_as(f64p),
                        np.float32(4.0 * float(vs)), np.float32(2.0),  # truncation 4 voxels (ros_params.h:66-67); ESDF max 2 m
                        tsdf.ctypes.data_as(f32p) if want_tsdf else None, esdf.ctypes.data_as(f32p) if want_esdf else None,
                        threads or os.cpu_count() or 1)
    return tsdf, esdf


def make_map(cfg_name: str, traj_prefix: Optional[int] = None) -> SynthMap:
    cfg = CONFIGS[cfg_name] if isinstance(cfg_name, str) else cfg_name
    rng = np.random.Generator(np.random.MT19937(cfg.seed_map))
    objects = _world_objects(cfg, rng)
    traj = _trajectory(cfg, rng)
    if traj_prefix is not None:
        traj = traj[:traj_prefix]
    keys = blocks_touching(cfg.voxel_size, cfg.vps, traj, cfg.obs_radius, cfg.bounds_min, cfg.bounds_max)
    tsdf, esdf = generate_blocks(cfg, keys, objects, traj, True, cfg.with_esdf)
    return SynthMap(cfg, np.float32(cfg.voxel_size), cfg.vps, keys, tsdf, esdf, objects, traj)


# ---------------------------------------------------------------------------------------------
# Candidates: RrtTree::iterate sampling rule (rrt.cpp:172-212) and node bookkeeping (:220-234).
# ---------------------------------------------------------------------------------------------
@dataclass
class Candidates:
    positions: np.ndarray     # [n][3]  newState[0..2]
    segments: np.ndarray      # [n][6]  checkMotion(origin, direction + origin + normalized(direction) * overshoot)
    parent: np.ndarray        # [n]     index of the parent node (-1 = root)
    parent_yaw: np.ndarray    # [n]     filled in by the caller when the parent's yaw is known (0 for the root)
    distance: np.ndarray      # [n]     Node::distance_
    attempts: int = 0


def make_candidates(cfg: SynthConfig, n: int, check_fn: Optional[Callable[[np.ndarray], np.ndarray]] = None,
                    seed: Optional[int] = None, batch: int = 256) -> Candidates:
    """Grow a tree of `n` nodes by the reference's rule.  `check_fn(segments[m][6]) -> free[m]` is the
    collision oracle used for acceptance (the GPU path in bench.py, the CPU oracle in tests); None
    accepts every sample.  Samples are drawn in batches against the tree as of the batch start."""
    rng = np.random.Generator(np.random.MT19937(cfg.seed_cand if seed is None else seed))
    root = np.array(cfg.root, float)
    lo, hi = np.array(cfg.bounds_min, float), np.array(cfg.bounds_max, float)
    nodes = [root]
    node_dist = [0.0]
    pos, seg, par, dist = [], [], [], []
    attempts = 0
    while len(pos) < n:
        P = np.array(nodes)
        D = np.array(node_dist)
        s = 2.0 * cfg.vicinity * (rng.random((batch, 3)) - 0.5)             # :176-178
        s = s[(s ** 2).sum(1) <= cfg.vicinity ** 2] + root                   # :179-182
        s = s[np.all((s >= lo) & (s <= hi), axis=1)]                         # :183-185
        if len(s) == 0:
            continue
        attempts += len(s)
        d2 = ((s[:, None, :] - P[None, :, :]) ** 2).sum(2)
        pi = d2.argmin(1)                                                    # kd_nearest3 (:190-197)
        origin = P[pi]
        direction = s - origin
        nrm = np.sqrt((direction ** 2).sum(1))
        ok = nrm > 1e-9
        s, pi, origin, direction, nrm = s[ok], pi[ok], origin[ok], direction[ok], nrm[ok]
        far = nrm > cfg.extension_range
        unit = direction / nrm[:, None]
        direction = np.where(far[:, None], cfg.extension_range * unit, direction)  # :204-206
        new = origin + direction                                             # :207-209
        dn = np.sqrt((direction ** 2).sum(1))
        end = (direction + origin) + (direction / dn[:, None]) * cfg.overshoot    # :210-212
        segs = np.concatenate([origin, end], 1)
        free = np.ones(len(segs), bool) if check_fn is None else np.asarray(check_fn(segs)).astype(bool)
        for i in np.nonzero(free)[0]:
            if len(pos) >= n:
                break
            pos.append(new[i]); seg.append(segs[i]); par.append(int(pi[i]) - 1)
            dist.append(D[pi[i]] + dn[i])                                    # :223
            nodes.append(new[i]); node_dist.append(D[pi[i]] + dn[i])
        if attempts > 200 * n + 10000:
            raise RuntimeError("candidate generation does not converge (map too cluttered around the root?)")
    return Candidates(np.array(pos), np.array(seg), np.array(par, np.int64), np.zeros(len(pos)), np.array(dist), attempts)
