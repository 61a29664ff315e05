#!/usr/bin/env python
"""bench.py -- viewpoint-gain hot path benchmark (BASELINE.json metric: candidate viewpoints/s).

A "step" is one planning-step pass of the hot path over one batch of synthetic candidates: segment collision check of
every tree edge (HighResManager::checkMotion), gain evaluation behind every free edge (RrtTree::optimized_gain: 360-yaw ray
sweep + best-yaw window), node scoring and best-candidate selection (rrt.cpp:220-246), with the best node of the WHOLE batch
known on the host of every rank at the end of the step.

Workload (no flags): BASELINE.json configs[2] ("cfg3": 50x50x10 m office-like TSDF/ESDF, 0.10 m voxels, 90x60 deg FOV, 5 m
range, 10,000 candidates with their parent edges) -- the 10k-candidate workload north_star's scaling target names; it fits
one GPU.  STRONG scaling: with N GPUs the same 10,000 candidates are sharded inside libnbvgpu.so (one process per GPU,
nbvgpu_create_rank), the map is replicated by nbvgpu_commit (page-locked H2D on rank 0 + ncclBroadcast).  `extra` carries
configs[1] (cfg2), configs[4] (cfg5, incremental replay) and configs[3] (cfg4: 100x100x20 m at 0.05 m, 8 m range, the 7 GB map
replicated inside every one of its three steps) measured in the same run; `--config cfgN` measures one configuration alone.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config cfgX] [--extra cfgA,cfgB]

Prints ONE JSON line (rank 0).  `value` = candidates of all ranks / device time (CUDA events on the launching stream around
each step, max over ranks) with inputs resident in HBM; `e2e` = same through the host-buffer C ABI (H2D of the candidates and
D2H of the results inside the timed region).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

ALGO_BYTES_PER_STEP = 8.0  # one {distance, weight} pair per voxel-step


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg3", help="cfg1..cfg5 (BASELINE.json configurations; cfg3, the 10,000-candidate workload, is the "
                    "headline), or one of the side measurements: frontier (History::bfs), interp (ESDF interpolation), shipped "
                    "(the reference's shipped geometry)")
    ap.add_argument("--extra", default=None, help="comma-separated configurations measured as sub-records (default: cfg2,cfg5,cfg4 with "
                    "the default headline, none otherwise)")
    ap.add_argument("--no-l2-probe", action="store_true")
    ap.add_argument("--no-shared-map", action="store_true", help="stage the map from rank 0's private page-locked memory (broadcast) "
                    "instead of shared memory every GPU pulls its slice from")
    ap.add_argument("--map", default="cfg1", help="map of the frontier / interp side measurements")
    ap.add_argument("--candidates", type=int, default=0, help="candidates per GPU (default: the config's)")
    ap.add_argument("--variant", type=int, default=0, help="gain kernel variant (A/B measurement)")
    ap.add_argument("--cpu-sample", type=int, default=0, help="candidates in the cpu_baseline sample (default 1000)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--scaling", default="", choices=["", "weak", "strong"], help="default: weak for cfg1/2, strong otherwise")
    ap.add_argument("--selfcheck", action="store_true", help="re-run one step with the plain FP64 kernel and compare")
    ap.add_argument("--traj-prefix", type=int, default=0, help="use only the first N trajectory points for the map")
    a = ap.parse_args()
    if a.extra is None:
        a.extra = "cfg2,cfg5,cfg4" if (a.config == "cfg3" and not a.candidates and a.impl == "ours") else ""
    return a


def cpu_model() -> str:
    """Model name of the host CPU the baseline legs run on (SURVEY.md 8d asks for it next to the core count)."""
    try:
        for line in open("/proc/cpuinfo"):
            if line.lower().startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def l2_info(g, achieved_gbs):
    """the same algorithmic bytes normalised by an L2 read peak measured on this box in this run
    (nbvgpu_probe_l2_read_bandwidth: 128-bit ld.global.cg reads from every SM over a 64 MiB buffer, untimed warm-up
    pass, 40 timed passes, CUDA events; best of 3).  None if the probe fails."""
    try:
        peak = max(g.probe_l2_read_bandwidth(64 << 20, 40) for _ in range(3))
    except Exception as e:  # measurement extra, never fatal
        print(f"bench: L2 probe failed: {e}", file=sys.stderr)
        return None
    return {"peak": peak, "unit": "GB/s", "frac": achieved_gbs / peak,
            "peak_source": "measured in this run: 64 MiB buffer, 128-bit L1-bypassing reads, best of 3"}


def ncu_traffic_per_launch(config):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the ray-sweep kernel for THIS configuration, from the committed
    `ncu --set full` capture of the same launch shape (profiles/ncu_traffic.json: {"<cfg>": {"k_gain_dram_bytes_per_launch": ...}},
    written by profiles/summarize_ncu.py), or None when that configuration has no capture."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        d = json.load(open(p))
    except Exception:
        return None
    rec = d.get(config)
    return rec.get("k_gain_dram_bytes_per_launch") if isinstance(rec, dict) else None


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples, self.reasons, self.max_mhz = index, False, [], set(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _sample(self):
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        try:
            self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
            r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            for bit, name in names.items():
                if r & bit:
                    self.reasons.add(name)
        except Exception:
            pass

    def run(self):
        if self.nv is None:
            return
        while True:
            # sparse on purpose (one rank per GPU: a step ends when the slowest rank's does, so whatever disturbs one rank's
            # host thread disturbs every step): the first sample 5 ms into the timed region, then four per second
            time.sleep(0.005 if not self.samples else 0.25)
            if self.stop_flag:
                break
            self._sample()

    def summary(self):
        if self.nv is not None and not self.samples:
            self._sample()   # a timed region shorter than 5 ms: one sample right behind it
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def load_oracle_ctx(m, cam):
    from oracle.binding import Oracle  # cpu_baseline / reference arm only
    o = Oracle()
    lt = o.layer_create(0, m.voxel_size, m.vps)
    le = o.layer_create(1, m.voxel_size, m.vps)
    o.layer_update(lt, m.keys, m.tsdf)
    o.layer_update(le, m.keys, m.esdf)
    o.set_camera(cam.as_dict())
    return o, lt, le


def time_oracle(o, lt, le, cfg, pos, seg, threads):
    t0 = time.perf_counter()
    o.check_segments(le, cfg.robot_radius, seg, threads=threads)
    r = o.gain_eval(lt, pos, threads=threads)
    dt = time.perf_counter() - t0
    return dt, int(r["steps"].sum())


class CpuReference:
    """The reference's own CPU implementation of the path on the host cores.

    kind "reference": oracle/_ref/libnbv_ref.so -- the reference's integrator_utils.cc, camera_model.cpp and
    the gain / collision bodies of rrt.cpp and voxblox_manager.{h,cpp}, compiled where they lie (oracle/Makefile
    `ref`).  The reference is single-threaded (one planner thread, and optimized_gain mutates its camera model),
    so "all host threads" means one independent planner instance per thread, candidates split between them.
    kind "port": the oracle restatement (about 2x faster than the real code: no heap-materialised rays, no
    shared_ptr copies), used when oracle/_ref is absent."""

    def __init__(self, m, cam, cfg, threads):
        from oracle import ref_binding as rb
        self.cfg, self.threads = cfg, max(1, threads)
        if rb.available():
            self.kind = "reference"
            lib = rb.RefLib()
            self.ctxs = []
            for _ in range(self.threads):
                r = rb.Reference(m.voxel_size, m.vps, lib=lib)
                r.tsdf_update(m.keys, m.tsdf)
                r.esdf_update(m.keys, m.esdf)
                r.set_camera(cam.as_dict(), cfg.robot_radius)
                self.ctxs.append(r)
        else:
            self.kind = "port"
            self.o, self.lt, self.le = load_oracle_ctx(m, cam)

    def run(self, pos, seg, threads=None):
        """One pass (collision check + gain) over the sample; returns wall seconds."""
        t = min(threads or self.threads, self.threads)
        t0 = time.perf_counter()
        if self.kind == "port":
            self.o.check_segments(self.le, self.cfg.robot_radius, seg, threads=t)
            self.o.gain_eval(self.lt, pos, threads=t)
        else:
            def work(i):
                self.ctxs[i].check_segments(seg[i::t])
                self.ctxs[i].gain_eval(pos[i::t])
            ths = [threading.Thread(target=work, args=(i,)) for i in range(t)]
            for x in ths:
                x.start()
            for x in ths:
                x.join()
        return time.perf_counter() - t0


# ---------------------------------------------------------------------------------------------
def run_reference(args):
    """Reference arm: the reference's own CPU implementation of the path (oracle/_ref when the reference compiled, else
    the oracle port) on the host cores, all threads, on the SAME configuration and candidate set as the default arm, the
    steps and warm-up steps asked for; each step a bounded sample of the candidates (a fixed prefix, stated as a fraction)
    sized so that the whole run ends within a few minutes.  Throughput is per candidate, so the sample size does not bias it."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from nbv_exploration_planner_b200 import synth
    name = args.config if args.config in synth.CONFIGS else "cfg3"
    cfg = synth.CONFIGS[name]
    m = synth.make_map(name)
    cam = m.camera()
    o, lt, le = load_oracle_ctx(m, cam)
    cores = os.cpu_count() or 1
    strong = (args.scaling or ("weak" if name in ("cfg1", "cfg2") else "strong")) == "strong"
    n_total = (args.candidates or cfg.n_candidates) * (1 if strong else max(1, args.gpus))
    c = synth.make_candidates(cfg, n_total, lambda s: o.check_segments(le, cfg.robot_radius, s, threads=cores)["free"])
    ref = CpuReference(m, cam, cfg, cores)
    # size the per-step sample from a first untimed pass: about 100 s for the whole run, at least one candidate per thread
    probe = min(n_total, cores)
    t_probe = ref.run(c.positions[:probe], c.segments[:probe])
    n_steps_all = max(1, args.steps) + max(0, args.warmup)
    sample = int(max(cores, min(n_total, 256, 100.0 / n_steps_all / max(t_probe / probe, 1e-6))))
    pos, seg = c.positions[:sample], c.segments[:sample]
    steps_per_vp = float(o.gain_eval(lt, pos, threads=cores)["steps"].mean())
    for _ in range(max(0, args.warmup)):
        ref.run(pos, seg)
    steps = max(1, args.steps)
    total = sum(ref.run(pos, seg) for _ in range(steps))
    value = sample * steps / total
    out = {
        "impl": "reference", "metric": "candidate viewpoints/s", "value": value, "unit": "viewpoints/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": max(0, args.warmup), "ms_per_step": 1e3 * total / steps, "higher_is_better": True,
        "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f64 frustum + f32 DDA + i64 voxel index", "data": "synthetic",
        "voxel_steps_per_s": value * steps_per_vp,
        "config": {"workload": workload_text(cfg, name, n_total, n_total), "candidates_total": int(n_total), "candidates_per_step": sample,
                   "fraction_of_candidates_per_step": sample / n_total, "map_blocks": int(len(m.keys))},
        "cpu_baseline": {"value": value, "unit": "viewpoints/s", "cores": cores, "cpu_model": cpu_model(), "kind": ref.kind,
                         "sample": f"first {sample} of {n_total} candidates per step ({100.0 * sample / n_total:.2f} %), {steps} steps, {cores} host "
                                   f"threads (one reference planner instance per thread)"},
        "e2e": {"value": value, "unit": "viewpoints/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(out)


def emit(obj):
    out = globals().get("_json_out") or sys.stdout
    out.write(json.dumps(obj) + "\n")
    out.flush()


# ---------------------------------------------------------------------------------------------
class Rig:
    """The process's place in the launch (one process per GPU, RANK / LOCAL_RANK / WORLD_SIZE from torchrun) and the
    plumbing torch.distributed provides around the library: the NCCL unique id and the candidate set travel through it
    once per configuration; nothing on the per-step path does."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a B200: the product path has no CPU fallback")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=self.dev)
        self.stream = torch.cuda.Stream(device=self.dev)   # one explicit stream: torch events and the library's kernels
        torch.cuda.set_stream(self.stream)
        self.flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=self.dev)  # > 126 MB L2

    def make_ctx(self):
        """One GPU: a one-device multi context (the shared-block result path without NCCL); N ranks: nbvgpu_create_rank."""
        from nbv_exploration_planner_b200 import dist as nd
        return nd.create_context(self.local, self.stream.cuda_stream)

    def bcast_int(self, v):
        if self.world == 1:
            return int(v)
        t = self.torch.tensor([int(v)], dtype=self.torch.int64, device=self.dev)
        self.dist.broadcast(t, 0)
        return int(t.item())

    def bcast_array(self, a, shape):
        """float64 array of `shape` from rank 0 to every rank (numpy in, numpy out)."""
        torch = self.torch
        t = torch.from_numpy(np.ascontiguousarray(a, np.float64)).to(self.dev) if self.rank == 0 else torch.empty(shape, dtype=torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.broadcast(t, 0)
        return t.cpu().numpy()

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def reduce(self, values, op="max"):
        torch = self.torch
        t = torch.tensor([float(v) for v in values], dtype=torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX if op == "max" else self.dist.ReduceOp.SUM)
        return [float(x) for x in t.cpu()]

    def gather(self, values):
        """[world][len(values)] floats: every rank's values on every rank (diagnostics: per-rank step / kernel times)."""
        torch = self.torch
        t = torch.tensor([float(v) for v in values], dtype=torch.float64, device=self.dev)
        if self.world == 1:
            return [[float(x) for x in t.cpu()]]
        out = torch.empty((self.world, len(values)), dtype=torch.float64, device=self.dev)
        self.dist.all_gather_into_tensor(out, t)
        return [[float(x) for x in row] for row in out.cpu()]

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


class PageLocked:
    """A numpy array page-locked in place (cudaHostRegister) so that nbvgpu_layer_update_pinned can copy straight from it;
    release() undoes the registration before the array may be freed (a pinned copy is the fallback)."""

    def __init__(self, a: np.ndarray):
        import torch
        self.array = np.ascontiguousarray(a)
        self.registered = False
        self.buffer = self.array
        try:
            self.registered = int(torch.cuda.cudart().cudaHostRegister(self.array.ctypes.data, self.array.nbytes, 0)) == 0
        except Exception:
            self.registered = False
        if not self.registered:
            self.buffer = torch.from_numpy(self.array).pin_memory()

    def release(self):
        import torch
        if self.registered:
            torch.cuda.cudart().cudaHostUnregister(self.array.ctypes.data)
            self.registered = False


def shard(n, world, rank):
    per = (n + world - 1) // world
    lo = min(n, rank * per)
    return lo, min(n, lo + per)


def workload_text(cfg, name, n_all, n_local):
    return (f"{name}: {cfg.bounds_max[0]}x{cfg.bounds_max[1]}x{cfg.bounds_max[2]} m synthetic TSDF+ESDF, {cfg.voxel_size} m voxels, "
            f"{cfg.vps}^3 blocks, {cfg.hfov_deg}x{cfg.vfov_deg} deg FOV, {cfg.far_m} m range, {int(n_all)} candidates "
            f"({n_local} on this GPU) with ESDF edge check (radius {cfg.robot_radius:.4f} m, overshoot {cfg.overshoot} m)")


def bench_gain_config(rig, args, name, steps, warmup, strong, with_cpu_baseline, include_replication=False):
    """One BASELINE.json gain configuration (cfg1..cfg4): map on rank 0 -> nbvgpu_commit (pinned H2D + NCCL broadcast inside
    the library) -> K timed planning steps.  A step = edge check of every candidate's parent edge, ray sweep behind the
    free ones, node scoring, best node of the WHOLE batch known on the host of every rank.
      value: shards resident in HBM (nbvgpu_evaluate_shard_device), CUDA events on the launching stream around each step --
             the closing event is recorded after the call has returned, i.e. after the slowest rank's record arrived --,
             L2 flushed between steps, max over ranks.
      e2e:   the whole batch in host memory on every rank (nbvgpu_evaluate_candidates): shard H2D, results D2H, wall clock.
    include_replication (config 4, "map broadcast cost included"): every step first clears the layers and replicates the
    whole map again, in both legs."""
    import hashlib
    torch = rig.torch
    from nbv_exploration_planner_b200 import LAYER_ESDF, LAYER_TSDF, synth
    cfg = synth.CONFIGS[name]
    world, rank, dev = rig.world, rig.rank, rig.dev
    n_total = (args.candidates or cfg.n_candidates) * (1 if strong else world)
    g = rig.make_ctx()
    g.set_kernel_variant(args.variant)
    t_setup = time.perf_counter()
    m = synth.make_map(name, traj_prefix=args.traj_prefix or None) if rank == 0 else None
    nb = rig.bcast_int(len(m.keys) if rank == 0 else 0)
    lt = g.layer_create(LAYER_TSDF, cfg.voxel_size, cfg.vps, nb + 64)
    le = g.layer_create(LAYER_ESDF, cfg.voxel_size, cfg.vps, nb + 64)
    locked = []
    nvox = cfg.vps ** 3
    map_bytes = nb * nvox * 12
    # The map lives in memory every rank's GPU can read (nbvgpu_shared_alloc: collective; plain page-locked memory on one GPU):
    # rank 0 -- the map's owner -- fills it, and a commit of that much data has every GPU pull its slice of each chunk over its
    # own PCIe link before an all-gather over NVLink completes it.  --no-shared-map keeps the map in rank 0's private memory
    # (one link + ncclBroadcast), for the A/B.
    shared = None
    use_shared = not args.no_shared_map
    if use_shared and world > 1:   # across processes the memory is a POSIX shared-memory segment: it has to fit /dev/shm
        try:
            vfs = os.statvfs("/dev/shm")
            use_shared = vfs.f_bavail * vfs.f_frsize > 1.25 * map_bytes
        except OSError:
            use_shared = False
        use_shared = bool(rig.bcast_int(1 if use_shared else 0))
    shared_note = None
    if use_shared:
        try:
            shared = g.shared_alloc(map_bytes + 256)
        except Exception as e:   # fails on every rank alike (the library agrees on it across ranks): private memory + broadcast then
            shared_note = str(e)
            print(f"bench: rank {rank}: {e}; the map stays in rank 0's private page-locked memory", file=sys.stderr)
    if shared is not None:
        if rank == 0:
            p_tsdf = shared[:nb * nvox * 8].view(np.float32).reshape(m.tsdf.shape)
            p_esdf = shared[nb * nvox * 8:nb * nvox * 12].view(np.float32).reshape(m.esdf.shape)
            p_tsdf[...] = m.tsdf
            p_esdf[...] = m.esdf
    if shared is None and rank == 0:
        locked = [PageLocked(m.tsdf), PageLocked(m.esdf)]
        p_tsdf, p_esdf = locked[0].buffer, locked[1].buffer

    def replicate():
        """rank 0: stage the whole map from page-locked memory; all ranks: commit.  Returns wall ms (commit is synchronous)."""
        t0 = time.perf_counter()
        if rank == 0:
            g.layer_update_pinned(lt, m.keys, p_tsdf)
            g.layer_update_pinned(le, m.keys, p_esdf)
        g.commit()
        return (time.perf_counter() - t0) * 1e3

    rig.barrier()
    cold_repl_ms = replicate()          # first commit of the context: staging buffers, NCCL connections come into being
    g.layer_clear(lt)
    g.layer_clear(le)
    rig.barrier()
    repl_ms = [replicate()]             # the same again, warm: what a planner pays when it replaces the whole map
    setup_repl_ms = repl_ms[0]
    g.set_camera(synth.camera_for(cfg))
    # candidates: grown on rank 0 with the GPU collision check as the acceptance test, then shared
    if rank == 0:
        c = synth.make_candidates(cfg, n_total, lambda s: g.check_segments(le, cfg.robot_radius, s))
        pack = np.concatenate([c.positions, c.segments, c.distance[:, None]], 1)
        mb = sum(int(np.asarray(a).nbytes) for a in (m.keys, m.tsdf, m.esdf))
        inputs_sha = {"candidates": hashlib.sha256(np.ascontiguousarray(pack).tobytes()).hexdigest(),
                      "map": m.sha256() if mb <= (1 << 30) else f"not hashed ({mb >> 20} MiB)",
                      "seeds": {"map": cfg.seed_map, "candidates": cfg.seed_cand}}
    else:
        pack, inputs_sha = None, None
    pack = rig.bcast_array(pack, (n_total, 10))
    lo, hi = shard(n_total, world, rank)
    n = hi - lo
    h_pos, h_seg, h_dist = (np.ascontiguousarray(pack[:, 0:3]), np.ascontiguousarray(pack[:, 3:9]), np.ascontiguousarray(pack[:, 9]))
    h_pyaw = np.zeros(n_total)                                  # flat candidate set: parents face yaw 0
    d_pos = torch.from_numpy(h_pos[lo:hi].copy()).to(dev)
    d_seg = torch.from_numpy(h_seg[lo:hi].copy()).to(dev)
    d_dist = torch.from_numpy(h_dist[lo:hi].copy()).to(dev)
    d_pyaw = torch.zeros(n, dtype=torch.float64, device=dev)
    d_free = torch.empty(max(n, 1), dtype=torch.uint8, device=dev)
    d_gain = torch.empty(max(n, 1), dtype=torch.float64, device=dev)
    d_yaw = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
    v_max, dyaw_max, zero_gain = 1.0, 0.5, 0.0
    setup_s = time.perf_counter() - t_setup

    def redo_map():
        g.layer_clear(lt)
        g.layer_clear(le)
        return replicate()

    def step_device():
        if include_replication:
            repl_ms.append(redo_map())
        return g.evaluate_shard_device(lt, le, cfg.robot_radius, n, lo, d_pos, d_seg, d_pyaw, d_dist, v_max, dyaw_max, zero_gain,
                                       d_free, d_gain, d_yaw)

    for _ in range(max(warmup, 3)):
        step_device()
    rig.barrier()
    # ---- timed region: K steps, CUDA events on the launching stream, L2 flushed between steps ------
    sampler = ClockSampler(rig.local)
    st0 = g.stats()
    g.set_profiling(True)
    repl_ms.clear()
    rig.barrier()
    sampler.start()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    winner = None
    for a, b in evs:
        rig.flush.zero_()
        a.record()
        winner = step_device()
        b.record()
    rig.barrier()
    sampler.stop_flag = True
    dev_ms = sum(a.elapsed_time(b) for a, b in evs)
    kern_ms, kern_launches = g.profile_collect()
    g.set_profiling(False)
    st1 = g.stats()
    launches = st1["kernel_launches"] - st0["kernel_launches"]
    vsteps = st1["voxel_steps"] - st0["voxel_steps"]
    rays = st1["rays"] - st0["rays"]
    csteps = st1["collision_steps"] - st0["collision_steps"]
    repl_timed = list(repl_ms)

    # ---- e2e: the same step through the host-buffer C ABI (H2D of the shard + D2H of its results inside) ------------
    e2e_steps = max(3, min(steps, 30))

    def step_host():
        if include_replication:
            redo_map()
        return g.evaluate_candidates(lt, le, cfg.robot_radius, h_pos, h_seg, h_pyaw, h_dist, v_max, dyaw_max, zero_gain)

    for _ in range(2):
        r_host = step_host()
    rig.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        rig.flush.zero_()
        r_host = step_host()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    rig.barrier()
    same = (r_host["index"], r_host["score"]) == (winner["index"], winner["score"]) and \
        bool(np.array_equal(r_host["gain"][lo:hi], d_gain[:n].cpu().numpy()))
    h2d = n * (48 + 24 + 8 + 8) + (map_bytes + nb * 64 if include_replication and rank == 0 else 0)
    d2h = n * (1 + 8 + 4)          # verdicts, gain, yaw of the shard; the best records arrive by GPU stores into host memory

    # ---- drop-in latency: one candidate through the n = 1 shims (what the unmodified iterate() would call)
    lat_us = None
    if rank == 0 and n > 0:
        st = np.array([h_pos[0, 0], h_pos[0, 1], h_pos[0, 2], 0.0])
        for _ in range(5):
            g.optimized_gain(lt, st)
            g.check_motion(le, cfg.robot_radius, h_seg[0, :3], h_seg[0, 3:])
        t1 = time.perf_counter()
        for _ in range(50):
            g.optimized_gain(lt, st)
            g.check_motion(le, cfg.robot_radius, h_seg[0, :3], h_seg[0, 3:])
        lat_us = (time.perf_counter() - t1) / 50 * 1e6

    selfcheck = None
    if args.selfcheck:   # the plain FP64 transcription kernel must give the same gains on the whole shard
        ref_gain, ref_yaw = d_gain.clone(), d_yaw.clone()
        g.set_kernel_variant(1)
        g.evaluate_shard_device(lt, le, cfg.robot_radius, n, lo, d_pos, d_seg, d_pyaw, d_dist, v_max, dyaw_max, zero_gain, d_free, d_gain, d_yaw)
        torch.cuda.synchronize()
        selfcheck = {"variant_1_equals_default": bool(torch.equal(ref_gain, d_gain) and torch.equal(ref_yaw, d_yaw)), "candidates": n}
        g.set_kernel_variant(args.variant)

    repl_med = float(np.median(repl_timed)) if repl_timed else setup_repl_ms
    dev_ms_max, e2e_s_max, kern_ms_max, repl_max = rig.reduce([dev_ms, e2e_s, kern_ms, repl_med], "max")
    vsteps_all, rays_all, launches_all, n_all, csteps_all, same_all = rig.reduce([vsteps, rays, launches, n, csteps, 0.0 if same else 1.0], "sum")
    per_rank = rig.gather([dev_ms / steps, kern_ms / max(1, kern_launches), vsteps / max(1, kern_launches), n])
    pulled = g.pulled_chunks()
    clocks = sampler.summary()
    out = None
    if rank == 0:
        value = n_all * steps / (dev_ms_max * 1e-3)
        peak, peak_src = measured_peaks()
        steps_per_launch = vsteps / max(1, kern_launches)
        ach = steps_per_launch * ALGO_BYTES_PER_STEP / (kern_ms / max(1, kern_launches) * 1e-3) / 1e9
        first_repl = repl_max
        out = {
            "metric": "candidate viewpoints/s", "value": value, "unit": "viewpoints/s", "n_gpus": world, "steps": steps,
            "warmup": max(warmup, 3), "ms_per_step": dev_ms_max / steps, "higher_is_better": True,
            "scaling": "strong" if strong else "weak", "vs_baseline": None,
            "dtype": "f64 frustum planes + f32 DDA + i32 voxel index (bit-exact vs reference arithmetic)", "data": "synthetic",
            "config": {"workload": workload_text(cfg, name, n_all, n), "candidates_total": int(n_all), "map_blocks": nb,
                       "l2": "flushed between steps (256 MiB write)", "kernel_variant": args.variant,
                       "parallelism": f"candidates sharded over {world} GPU(s) inside libnbvgpu.so, map replicated by nbvgpu_commit "
                                      + (f"(the map lies in shared page-locked memory, nbvgpu_shared_alloc: per chunk every GPU copies 1/{world} over its own "
                                         f"PCIe link, ncclAllGather over NVLink completes it; {pulled} chunks so far)" if pulled else
                                         "(page-locked H2D on rank 0 + ncclBroadcast of {descriptor table | payload} per chunk)")
                                      + "; best records through GPU stores into shared host memory (no collective per step)",
                       "map_replication": {"ms": first_repl, "bytes": map_bytes, "gb_per_s": map_bytes / (first_repl * 1e-3) / 1e9 if first_repl else None,
                                           "first_commit_ms": cold_repl_ms, "inside_timed_steps": bool(include_replication),
                                           "pulled_chunks": pulled,
                                           "what": "host wall clock of layer_clear x2 excluded; rank 0 stages TSDF + ESDF from page-locked "
                                                   "memory, every rank commits: H2D (its slice when pulled) + NCCL + scatter kernel, max over ranks"},
                       "setup_s": setup_s, "inputs_sha256": inputs_sha},
            "rays_per_s": rays_all / (dev_ms_max * 1e-3), "voxel_steps_per_s": vsteps_all / (dev_ms_max * 1e-3),
            "voxel_steps_per_viewpoint": vsteps_all / max(1.0, n_all * steps),
            "edges_per_s": n_all * steps / (dev_ms_max * 1e-3), "collision_steps_per_s": csteps_all / (dev_ms_max * 1e-3),
            "e2e": {"value": n_all * e2e_steps / e2e_s_max, "unit": "viewpoints/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "steps": e2e_steps, "equals_device_leg": same_all == 0.0,
                    "l2": "flushed between steps"},
            "gpu_launches": int(launches_all),
            "roofline": {"bound": "hbm", "kernel": "k_gain (ray sweep)", "achieved": ach, "peak": peak, "unit": "GB/s",
                         "frac": ach / peak, "traffic": ncu_traffic_per_launch(name), "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": steps_per_launch * ALGO_BYTES_PER_STEP,
                         "avg_launch_ms": kern_ms / max(1, kern_launches),
                         "kernel_share_of_step": kern_ms / dev_ms if dev_ms > 0 else None,
                         "step_minus_kernel_us": (dev_ms - kern_ms) / steps * 1e3,   # edge check, window scan, scoring, record to host, launch gaps
                         "graph_replays": g.graph_replays(),
                         "l2": l2_info(g, ach) if not args.no_l2_probe else None,
                         "note": "8 B per executed voxel-step (SURVEY.md 8d); the status tiles the sweep reads are L2/L1 resident, so real DRAM traffic is far lower"},
            "clocks": clocks, "best": winner,
            # per rank: device time per step (what the max is taken over), the sweep's own time per launch, voxel-steps and candidates
            "per_rank": {"ms_per_step": [r[0] for r in per_rank], "sweep_ms": [r[1] for r in per_rank],
                         "voxel_steps_per_launch": [r[2] for r in per_rank], "candidates": [int(r[3]) for r in per_rank]},
            "single_candidate_latency_us": lat_us,  # checkMotion + optimized_gain, n = 1, host buffers
        }
        if selfcheck is not None:
            out["selfcheck"] = selfcheck
        if with_cpu_baseline and not args.no_cpu_baseline and world == 1:
            out["cpu_baseline"] = cpu_baseline(args, name, cfg, h_pos, h_seg)
    g.close()
    for pl in locked:
        pl.release()
    return out


def bench_replay(rig, args, name, steps):
    """BASELINE config 5: incremental exploration replay.  Per planning step: the blocks within the observation sphere of
    the next trajectory point change (rank 0 stages them from page-locked memory: TSDF + ESDF), nbvgpu_commit replicates them
    (one H2D, one broadcast, one kernel), then a fresh UNFILTERED set of candidates around the new position is checked,
    swept behind its free edges and scored, the whole batch in host memory on every rank (nbvgpu_evaluate_candidates).
    Everything per step is inside the timed region (host wall clock, max over ranks); the map WRITER (block generation) and
    the candidate sampler run before it."""
    torch = rig.torch
    from nbv_exploration_planner_b200 import LAYER_ESDF, LAYER_TSDF, synth
    world, rank = rig.world, rig.rank
    cfg = synth.CONFIGS[name]
    n_total = args.candidates or cfg.n_candidates
    n_steps = min(steps, cfg.traj_points - 1)
    g = rig.make_ctx()
    nvox = cfg.vps ** 3
    rng = np.random.Generator(np.random.MT19937(cfg.seed_map))
    objects = synth._world_objects(cfg, rng)
    traj = synth._trajectory(cfg, rng)
    all_keys = synth.blocks_touching(cfg.voxel_size, cfg.vps, traj, cfg.obs_radius, cfg.bounds_min, cfg.bounds_max)
    lt = g.layer_create(LAYER_TSDF, cfg.voxel_size, cfg.vps, len(all_keys) + 64)
    le = g.layer_create(LAYER_ESDF, cfg.voxel_size, cfg.vps, len(all_keys) + 64)
    g.set_camera(synth.camera_for(cfg))
    updates = []
    if rank == 0:   # per-step changed blocks in page-locked memory (map writing is not the path being measured)
        for k in range(n_steps + 1):
            keys = synth.blocks_touching(cfg.voxel_size, cfg.vps, traj[k:k + 1], cfg.obs_radius, cfg.bounds_min, cfg.bounds_max)
            tsdf, esdf = synth.generate_blocks(cfg, keys, objects, traj[:k + 1])
            updates.append((keys, torch.from_numpy(tsdf).pin_memory(), torch.from_numpy(esdf).pin_memory()))
    # candidate offsets around the current position (star-shaped expansion from the current pose), per step, up front
    crng = np.random.Generator(np.random.MT19937(cfg.seed_cand))
    off = 2.0 * cfg.vicinity * (crng.random((4 * n_total, 3)) - 0.5)
    off = off[(off ** 2).sum(1) <= cfg.vicinity ** 2][:n_total]
    lo_b, hi_b = np.array(cfg.bounds_min, float) + 0.3, np.array(cfg.bounds_max, float) - 0.3
    cands = []
    for k in range(n_steps + 1):
        p = traj[k]
        pos = np.clip(p + off, lo_b, hi_b)
        d = pos - p
        nrm = np.maximum(np.linalg.norm(d, axis=1, keepdims=True), 1e-9)
        seg = np.concatenate([np.broadcast_to(p, pos.shape), pos + d / nrm * cfg.overshoot], 1)
        cands.append((np.ascontiguousarray(pos), np.ascontiguousarray(seg), np.ascontiguousarray(nrm[:, 0])))
    pyaw = np.zeros(n_total)
    lo, hi = shard(n_total, world, rank)

    def apply_update(k):
        if rank == 0:
            keys, pt, pe = updates[k]
            g.layer_update_pinned(lt, keys, pt)
            g.layer_update_pinned(le, keys, pe)
        g.commit()
        return len(updates[k][0]) if rank == 0 else 0

    def plan(k):
        pos, seg, dist = cands[k]
        return g.evaluate_candidates(lt, le, cfg.robot_radius, pos, seg, pyaw, dist, 1.0, 0.5, 0.0)

    apply_update(0)
    plan(0)                                   # warm-up on the first pose
    st0 = g.stats()
    sampler = ClockSampler(rig.local)
    rig.barrier()
    sampler.start()
    t0 = time.perf_counter()
    up_ms, blocks, nfree = 0.0, 0, 0
    best = None
    for k in range(1, n_steps + 1):
        tu = time.perf_counter()
        blocks += apply_update(k)             # synchronous: returns when the snapshot is in place on this rank's GPU
        up_ms += (time.perf_counter() - tu) * 1e3
        best = plan(k)
        nfree += int(best["free"][lo:hi].sum())
    torch.cuda.synchronize()
    total_s = time.perf_counter() - t0
    rig.barrier()
    sampler.stop_flag = True
    st1 = g.stats()
    total_s, up_ms = rig.reduce([total_s, up_ms], "max")
    vsteps, launches, nfree, blocks, up_bytes = rig.reduce([st1["voxel_steps"] - st0["voxel_steps"], st1["kernel_launches"] - st0["kernel_launches"],
                                                            nfree, blocks, st1["bytes_uploaded"] - st0["bytes_uploaded"]], "sum")
    out = None
    if rank == 0:
        value = n_total * n_steps / total_s
        peak, _ = measured_peaks()
        bpb = nvox * 12 + nvox // 4 + nvox // 8      # tile bytes written per (TSDF + ESDF) block: float2 + float tiles, status + observed bits
        out = {"metric": "candidate viewpoints/s", "value": value, "unit": "viewpoints/s", "n_gpus": world, "steps": n_steps,
               "warmup": 1, "ms_per_step": 1e3 * total_s / n_steps, "higher_is_better": True, "scaling": "strong",
               "vs_baseline": None, "dtype": "f64 frustum planes + f32 DDA + i32 voxel index", "data": "synthetic",
               "config": {"workload": f"{name}: incremental exploration replay, {n_steps} planning steps, per-step changed-block "
                                      f"commit + {n_total} unfiltered candidates, {cfg.voxel_size} m voxels, {cfg.far_m} m range",
                          "blocks_uploaded_per_step": blocks / n_steps, "map_update_ms_per_step": up_ms / n_steps,
                          "map_update_gb_per_s": (blocks / n_steps) * nvox * 12 / (up_ms / n_steps * 1e-3) / 1e9,
                          "free_edges_per_step": nfree / n_steps,
                          "note": "a candidate whose edge collides is not swept (rrt.cpp:210-216)",
                          "timing": "host wall clock over the whole replay, host buffers, max over ranks"},
               "voxel_steps_per_s": vsteps / total_s, "gpu_launches": int(launches), "clocks": sampler.summary(),
               "update_kernel_bytes_written_per_step": (blocks / n_steps) * bpb,
               "e2e": {"value": value, "unit": "viewpoints/s",
                       "h2d_bytes_per_step": (hi - lo) * 88 + int(blocks / n_steps) * (nvox * 12 + 64),
                       "d2h_bytes_per_step": (hi - lo) * 13},
               "best": {k: best[k] for k in ("index", "score", "best_gain", "best_yaw")}}
    g.close()
    return out


def run_ours(args):
    """Headline: BASELINE.json configs[2] ("cfg3": 50x50x10 m office map, 0.10 m voxels, 10,000 candidates), STRONG scaling --
    the workload north_star's scaling target names.  `extra` carries configs[1] (cfg2: 1,000 candidates, the largest
    single-GPU configuration of round 1) and configs[4] (cfg5: the incremental replay) measured in the same process.
    --config cfgN runs one configuration alone (cfg4: every step includes the map replication)."""
    rig = Rig()
    head = args.config
    strong = (args.scaling or ("weak" if head in ("cfg1", "cfg2") else "strong")) == "strong"
    if head == "cfg5":
        out = bench_replay(rig, args, head, args.steps)
    else:
        out = bench_gain_config(rig, args, head, args.steps, args.warmup, strong, True, include_replication=(head == "cfg4"))
    extras = [e for e in (args.extra.split(",") if args.extra else []) if e and e != head and e != "none"]
    extra = {}
    for e in extras:
        sub = argparse.Namespace(**vars(args))
        sub.candidates = 0
        if e == "cfg5":
            r = bench_replay(rig, sub, e, 200)   # BASELINE configs[4]: the whole 200-pose trajectory (199 updates + evaluations)
        else:
            # config 4: 16.7 M voxel-steps per viewpoint and 7 GB of map replicated inside every step: three steps are enough
            r = bench_gain_config(rig, sub, e, 3 if e == "cfg4" else min(args.steps, 20), 1 if e == "cfg4" else args.warmup, True, False,
                                  include_replication=(e == "cfg4"))
        if r is not None:
            for k in ("dtype", "data", "higher_is_better", "vs_baseline", "metric", "unit"):
                r.pop(k, None)
            extra[e] = r
    if rig.rank == 0:
        if extra:
            out["extra"] = extra
        emit(out)
    rig.close()


# ---- the widened rows and the reference's shipped geometry: one JSON line each --------------------
# Host wall clock around the C-ABI call with host buffers (plus a device-resident figure for the interpolation); their
# cpu_baseline legs time the reference's own code from oracle/_ref (one thread: the reference runs these on its planner
# thread) and the oracle port on a bounded sample.  Not the headline metric: `python bench.py` without --config is that.
def run_frontier(args):
    """--config frontier: history-graph frontier potential, History::bfs (history.cpp:96-137); --candidates = vertices per call."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    n_items = args.candidates or 148
    cpu_sample = 0 if args.no_cpu_baseline else (args.cpu_sample or 3)
    import torch
    from nbv_exploration_planner_b200 import LAYER_TSDF, NbvGpu, synth
    m = synth.make_map(args.map)
    cam = m.camera()
    g = NbvGpu(0)
    lt = g.layer_create(LAYER_TSDF, m.voxel_size, m.vps, len(m.keys) + 8)
    g.layer_update(lt, m.keys, m.tsdf)
    g.commit()
    g.set_camera(cam)
    pts = synth.make_candidates(m.cfg, n_items, None).positions
    for _ in range(args.warmup):
        r = g.frontier_potential(lt, pts, 6.0)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r = g.frontier_potential(lt, pts, 6.0)
    dt = (time.perf_counter() - t0) / args.steps
    one = []
    for i in range(min(8, len(pts))):
        t1 = time.perf_counter()
        g.frontier_potential(lt, pts[i:i + 1], 6.0)
        one.append(time.perf_counter() - t1)
    out = {"metric": "history-graph vertices/s (frontier potential, History::bfs)", "value": len(pts) / dt, "unit": "vertices/s",
           "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
           "timing": "host wall clock around the C-ABI call (host buffers in, host buffers out)",
           "config": {"workload": f"{args.map} map, {len(pts)} vertices per call, max_distance {6.0} m, "
                                  f"voxel {m.voxel_size} m", "map_blocks": int(len(m.keys))},
           "visited_cells_per_vertex": float(r["visited"].mean()), "cells_per_s": float(r["visited"].sum()) / dt,
           "single_vertex_latency_ms": float(np.median(one) * 1e3)}
    # CPU side on a bounded sample: the reference's own code (1 thread) and the oracle port
    k = min(cpu_sample, len(pts))
    if k > 0:
        from oracle import binding as ob, ref_binding as rb
        o = ob.Oracle()
        lo = o.layer_create(0, m.voxel_size, m.vps)
        o.layer_update(lo, m.keys, m.tsdf)
        o.set_camera(cam.as_dict())
        t1 = time.perf_counter()
        ro = o.frontier_potential(lo, pts[:k], 6.0, threads=1)
        t_or = (time.perf_counter() - t1) / k
        ok = bool(np.array_equal(ro["gain"], r["gain"][:k]) and np.array_equal(ro["cells_checksum"], r["cells_checksum"][:k]))
        cpu = {"oracle_port_ms_per_vertex": t_or * 1e3, "sample": f"first {k} vertices, 1 thread", "gpu_equals_oracle_on_sample": ok}
        if rb.available() and 6.0 == 6.0:
            ref = rb.Reference(m.voxel_size, m.vps)
            ref.tsdf_update(m.keys, m.tsdf)
            ref.set_camera(cam.as_dict(), 0.4)
            t1 = time.perf_counter()
            rr = ref.frontier_potential(pts[:k])
            t_ref = (time.perf_counter() - t1) / k
            cpu.update({"kind": "reference", "reference_ms_per_vertex": t_ref * 1e3, "value": 1.0 / t_ref, "unit": "vertices/s", "cores": 1,
                        "gpu_equals_reference_on_sample": bool(np.array_equal(rr, r["gain"][:k]))})
        else:
            cpu.update({"kind": "port", "value": 1.0 / t_or, "unit": "vertices/s", "cores": 1})
        out["cpu_baseline"] = cpu
    emit(out)
    g.close()



def run_interp(args):
    """--config interp: ESDF interpolated distance + gradient, EsdfMap::getDistanceAndGradientAtPosition (esdf_map.cc:22-52);
    --candidates = queries per call."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    n_items = args.candidates or 1000000
    cpu_sample = 0 if args.no_cpu_baseline else (args.cpu_sample or 200000)
    import torch
    from nbv_exploration_planner_b200 import LAYER_ESDF, PACK_VOXBLOX, NbvGpu, synth
    from voxblox_fixture import esdf_words
    m = synth.make_map(args.map)
    obs = (m.tsdf[..., 1] > 0).astype(np.uint8)
    g = NbvGpu(0)
    le = g.layer_create(LAYER_ESDF, m.voxel_size, m.vps, len(m.keys) + 8)
    g.layer_update(le, m.keys, esdf_words(m.esdf, flags=obs.astype(np.uint32)), PACK_VOXBLOX)
    g.commit()
    rng = np.random.default_rng(5)
    lo, hi = np.array(m.cfg.bounds_min, float), np.array(m.cfg.bounds_max, float)
    q = rng.uniform(lo, hi, (n_items, 3))
    for _ in range(args.warmup):
        r = g.esdf_distance_gradient(le, q)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r = g.esdf_distance_gradient(le, q)
    dt = (time.perf_counter() - t0) / args.steps
    t1 = time.perf_counter()
    for i in range(200):
        g.esdf_distance_gradient(le, q[i:i + 1])
    one = (time.perf_counter() - t1) / 200
    # inputs and outputs resident in HBM: nbvgpu_esdf_distance_gradient_device on an explicit stream, CUDA events
    stream = torch.cuda.Stream()
    g.set_stream(stream.cuda_stream)
    with torch.cuda.stream(stream):
        d_q = torch.from_numpy(q).cuda()
        d_ok = torch.empty(len(q), dtype=torch.uint8, device="cuda")
        d_dist = torch.empty(len(q), dtype=torch.float64, device="cuda")
        d_grad = torch.empty((len(q), 3), dtype=torch.float64, device="cuda")
        for _ in range(args.warmup):
            g.esdf_distance_gradient_device(le, len(q), d_q, d_ok, d_dist, d_grad)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            g.esdf_distance_gradient_device(le, len(q), d_q, d_ok, d_dist, d_grad)
        e1.record()
        stream.synchronize()
        dev_ms = e0.elapsed_time(e1) / args.steps
        same_dev = bool(np.array_equal(d_ok.cpu().numpy(), r["ok"]) and np.array_equal(d_dist.cpu().numpy(), r["distance"]))
    g.set_stream(None)
    out = {"metric": "ESDF interpolated distance+gradient queries/s (EsdfMap::getDistanceAndGradientAtPosition)", "value": len(q) / dt,
           "unit": "queries/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
           "timing": "host wall clock around the C-ABI call (host buffers in, host buffers out)",
           "config": {"workload": f"{args.map} map, {len(q)} uniform random positions inside the map bounds per call, voxel {m.voxel_size} m",
                      "map_blocks": int(len(m.keys))},
           "success_fraction": float(r["ok"].mean()), "single_query_latency_us": one * 1e6,
           "device_resident": {"value": len(q) / (dev_ms * 1e-3), "unit": "queries/s", "ms_per_step": dev_ms,
                               "timing": "CUDA events on the launching stream, buffers in HBM (nbvgpu_esdf_distance_gradient_device)",
                               "equals_host_path": same_dev},
           "h2d_bytes_per_step": len(q) * 24, "d2h_bytes_per_step": len(q) * 33}
    k = min(cpu_sample, len(q))
    if k > 0:
        from oracle import binding as ob, ref_binding as rb
        o = ob.Oracle()
        lo_e = o.layer_create(1, m.voxel_size, m.vps)
        o.layer_update_esdf_observed(lo_e, m.keys, m.esdf, obs)
        t1 = time.perf_counter()
        ro = o.distance_and_gradient(lo_e, q[:k], threads=1)
        t_or = time.perf_counter() - t1
        same = bool(np.array_equal(ro["ok"], r["ok"][:k]) and np.array_equal(ro["distance"], r["distance"][:k]) and
                    np.array_equal(ro["gradient"], r["gradient"][:k]))
        cpu = {"oracle_port_queries_per_s": k / t_or, "sample": f"first {k} queries, 1 thread", "gpu_equals_oracle_on_sample": same}
        if rb.available():
            ref = rb.Reference(m.voxel_size, m.vps)
            ref.esdf_update_observed(m.keys, m.esdf, obs)
            t1 = time.perf_counter()
            rr = ref.distance_and_gradient(q[:k])
            t_ref = time.perf_counter() - t1
            cpu.update({"kind": "reference", "value": k / t_ref, "unit": "queries/s", "cores": 1,
                        "note": "the reference's interpolator lines compiled against the Eigen stand-in (scalar products)",
                        "gpu_equals_reference_on_sample": bool(np.array_equal(rr["ok"], r["ok"][:k]) and np.array_equal(rr["distance"], r["distance"][:k])
                                                               and np.array_equal(rr["gradient"], r["gradient"][:k]))})
        else:
            cpu.update({"kind": "port", "value": k / t_or, "unit": "queries/s", "cores": 1})
        out["cpu_baseline"] = cpu
    emit(out)
    g.close()



def run_shipped(args):
    """--config shipped: gain evaluation on the geometry of the reference's shipped indoor configuration
    (resource/exploration_indoor.yaml, launch/voxblox_exploration_indoor.launch): 69 x 42 deg, 3.0 m range, 0.3 m voxels."""
    n = args.candidates or 1000
    import torch
    from nbv_exploration_planner_b200 import LAYER_TSDF, NbvGpu, synth
    cfg = synth.SynthConfig("shipped-indoor", (-4, -1, 0), (10, 14, 3), 0.3, 3.0, n, "pillars", 4001, 4002, (3.0, 6.0, 1.0), obs_radius=5.0,
                            traj_points=8, traj_step=3.0, vicinity=6.0, hfov_deg=69.0, vfov_deg=42.0, pitch_deg=0.0, near_m=0.1)
    m = synth.make_map(cfg)
    cam = m.camera()
    g = NbvGpu(0)
    lt = g.layer_create(LAYER_TSDF, m.voxel_size, m.vps, len(m.keys) + 8)
    g.layer_update(lt, m.keys, m.tsdf)
    g.commit()
    g.set_camera(cam)
    pos = synth.make_candidates(cfg, n, None).positions
    for _ in range(3):
        r = g.gain_eval(lt, pos, steps=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        r = g.gain_eval(lt, pos, steps=True)
    dt = (time.perf_counter() - t0) / 20
    out = {"metric": "candidate viewpoints/s, shipped indoor geometry", "value": n / dt, "unit": "viewpoints/s", "ms_per_step": dt * 1e3,
           "timing": "host wall clock around nbvgpu_gain_eval (host buffers)", "n_candidates": n,
           "voxel_steps_per_viewpoint": float(r["steps"].mean()), "voxel_steps_per_s": float(r["steps"].sum()) / dt,
           "config": {"workload": "69x42 deg, 3.0 m range, 0.3 m voxels, 16^3 blocks, box (-4,-1,0)..(10,14,3), pillars world", "map_blocks": int(len(m.keys))}}
    st = np.zeros(4)
    for i in range(8):
        st[:3] = pos[i]
        g.optimized_gain(lt, st)
    t1 = time.perf_counter()
    for i in range(100):
        st[:3] = pos[i]
        g.optimized_gain(lt, st)
    out["single_candidate_latency_us"] = (time.perf_counter() - t1) / 100 * 1e6          # nbvgpu_optimized_gain: the drop-in call
    t1 = time.perf_counter()
    for i in range(100):
        g.gain_eval(lt, pos[i:i + 1])
    out["single_candidate_with_counts_and_steps_latency_us"] = (time.perf_counter() - t1) / 100 * 1e6
    from oracle import ref_binding as rb
    if rb.available() and not args.no_cpu_baseline:
        ref = rb.Reference(m.voxel_size, m.vps)
        ref.tsdf_update(m.keys, m.tsdf)
        ref.set_camera(cam.as_dict(), 0.4)
        k = min(n, 200)
        t1 = time.perf_counter()
        rr = ref.gain_eval(pos[:k])
        t_ref = time.perf_counter() - t1
        out["cpu_baseline"] = {"kind": "reference", "value": k / t_ref, "unit": "viewpoints/s", "cores": 1, "sample": f"first {k} candidates, 1 thread",
                               "gpu_equals_reference_on_sample": bool(np.array_equal(rr["gain"], r["gain"][:k]))}
    emit(out)
    g.close()



def cpu_baseline(args, name, cfg, h_pos, h_seg):
    """The reference's CPU implementation timed on the host cores of this box on a bounded sample of the
    same workload: all threads and one thread (the reference itself plans on a single thread); the faster
    oracle port is reported alongside."""
    from nbv_exploration_planner_b200 import synth
    m = synth.make_map(name)
    cam = m.camera()
    cores = os.cpu_count() or 1
    ref = CpuReference(m, cam, cfg, cores)
    sample = min(len(h_pos), args.cpu_sample or max(cores * 8, 128))
    one = min(len(h_pos), 12)
    dt1 = ref.run(h_pos[:one], h_seg[:one], threads=1)
    dta = min(ref.run(h_pos[:sample], h_seg[:sample]) for _ in range(2))
    o, lt, le = load_oracle_ctx(m, cam)
    dtp, stp = time_oracle(o, lt, le, cfg, h_pos[:sample], h_seg[:sample], cores)
    dtp1, _ = time_oracle(o, lt, le, cfg, h_pos[:one], h_seg[:one], 1)
    return {"value": sample / dta, "unit": "viewpoints/s", "cores": cores, "cpu_model": cpu_model(), "kind": ref.kind,
            "sample": f"first {sample} candidates of the same workload, {cores} host threads (one reference planner instance per "
                      f"thread), best of 2",
            "voxel_steps_per_s": (sample / dta) * stp / sample,
            "single_thread": {"value": one / dt1, "sample": f"first {one} candidates, 1 thread"},
            "oracle_port": {"value": sample / dtp, "single_thread_value": one / dtp1, "cores": cores,
                            "note": "restatement of the same arithmetic without the reference's per-step heap / shared_ptr overheads"}}


if __name__ == "__main__":
    # stdout carries exactly one JSON line: anything else written to fd 1 (NCCL banners, library
    # chatter) is sent to stderr; the JSON goes to the saved copy of the original stdout.
    _json_fd = os.dup(1)
    os.dup2(2, 1)
    _json_out = os.fdopen(_json_fd, "w")
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    elif a.config == "frontier":
        run_frontier(a)
    elif a.config == "interp":
        run_interp(a)
    elif a.config == "shipped":
        run_shipped(a)
    else:
        run_ours(a)
