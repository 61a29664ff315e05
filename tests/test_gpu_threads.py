"""The reference's threading model on the boundary: the planner thread
evaluates gains, the planner and history threads both call checkMotion, and the map (ROS spin)
thread uploads changed blocks -- concurrently.  Every ABI call is serialised on the context and
ordered on one stream, so each evaluation must see the map exactly as of one of the commits."""
import threading

import numpy as np
import pytest

from conftest import load_oracle

pytestmark = pytest.mark.gpu


def test_concurrent_evaluators_and_map_updates(oracle_mod, tiny_map):
    from nbv_exploration_planner_b200 import LAYER_ESDF, LAYER_TSDF, NbvGpu, synth
    cfg = tiny_map.cfg
    cam = tiny_map.camera()
    g = NbvGpu(0)
    lt = g.layer_create(LAYER_TSDF, cfg.voxel_size, 16, 64)
    le = g.layer_create(LAYER_ESDF, cfg.voxel_size, 16, 64)
    g.set_camera(cam)
    # two map states: A = the tiny map, B = the same blocks with everything unknown / ESDF 0 (collides)
    tsdf_b = np.zeros_like(tiny_map.tsdf)
    esdf_b = np.zeros_like(tiny_map.esdf)
    g.layer_update(lt, tiny_map.keys, tiny_map.tsdf); g.layer_update(le, tiny_map.keys, tiny_map.esdf)
    g.commit()
    c = synth.make_candidates(cfg, 16, None)
    # expected results for both states from the oracle
    o, lt_o, le_o = load_oracle(oracle_mod, tiny_map, cam)
    gain_a = o.gain_eval(lt_o, c.positions)["gain"]
    free_a = o.check_segments(le_o, cfg.robot_radius, c.segments)["free"]
    o.layer_update(lt_o, tiny_map.keys, tsdf_b); o.layer_update(le_o, tiny_map.keys, esdf_b)
    gain_b = o.gain_eval(lt_o, c.positions)["gain"]
    free_b = o.check_segments(le_o, cfg.robot_radius, c.segments)["free"]
    assert not np.array_equal(gain_a, gain_b) and not np.array_equal(free_a, free_b)

    errors, stop = [], threading.Event()

    def map_thread():
        try:
            for k in range(40):
                if k % 2 == 0:
                    g.layer_update(lt, tiny_map.keys, tsdf_b); g.layer_update(le, tiny_map.keys, esdf_b)
                else:
                    g.layer_update(lt, tiny_map.keys, tiny_map.tsdf); g.layer_update(le, tiny_map.keys, tiny_map.esdf)
                g.commit()
        except Exception as e:  # pragma: no cover
            errors.append(e)
        finally:
            stop.set()

    def planner_thread():
        try:
            while not stop.is_set():
                r = g.gain_eval(lt, c.positions)["gain"]
                if not (np.array_equal(r, gain_a) or np.array_equal(r, gain_b)):
                    errors.append(AssertionError("gain saw a half-applied TSDF update"))
        except Exception as e:  # pragma: no cover
            errors.append(e)

    def history_thread():
        try:
            while not stop.is_set():
                f = g.check_segments(le, cfg.robot_radius, c.segments)
                if not (np.array_equal(f, free_a) or np.array_equal(f, free_b)):
                    errors.append(AssertionError("checkMotion saw a half-applied ESDF update"))
                g.check_motion(le, cfg.robot_radius, c.segments[0, :3], c.segments[0, 3:])
        except Exception as e:  # pragma: no cover
            errors.append(e)

    ts = [threading.Thread(target=f) for f in (map_thread, planner_thread, history_thread, history_thread)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(timeout=120)
    assert not errors, errors[:3]
    g.close()


def test_torch_stream_ordering(tiny_map):
    """nbvgpu_set_stream: work is ordered on the caller's torch stream, so torch events bracket it."""
    import torch
    from nbv_exploration_planner_b200 import LAYER_TSDF, NbvGpu
    g = NbvGpu(0)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        g.set_stream(s.cuda_stream)
        lt = g.layer_create(LAYER_TSDF, tiny_map.voxel_size, 16, 64)
        g.layer_update(lt, tiny_map.keys, tiny_map.tsdf)
        g.commit()
        g.set_camera(tiny_map.camera())
        pos = torch.tensor([[4.0, 4.0, 1.2]] * 64, dtype=torch.float64, device="cuda")
        gain = torch.empty(64, dtype=torch.float64, device="cuda")
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.gain_eval_device(lt, 64, pos, gain)
        e1.record()
        s.synchronize()
        assert e0.elapsed_time(e1) > 0.02          # the kernels ran between the two events
        assert float(gain.min()) == float(gain.max()) and float(gain[0]) >= 0.0
        g.set_stream(None)                          # back to the library-owned stream
        g.set_stream(0)                             # torch's default stream handle (cudaStreamLegacy)
        g.gain_eval_device(lt, 64, pos, gain)
        torch.cuda.synchronize()
    g.close()


def test_check_motion_does_not_queue_behind_a_gain_batch(oracle_mod, cfg1_map):
    """SURVEY.md 8b / history.cpp:249,461,555: the history thread calls checkMotion while the planner thread has a large gain
    batch in flight.  The batch holds the context's own lane; the small call takes a lane context (own high-priority stream,
    own staging) and must come back in tens of microseconds, not after the batch -- with the right answers on both sides."""
    import time
    from nbv_exploration_planner_b200 import synth
    m, cfg = cfg1_map, synth.CONFIGS["cfg2"]
    cam = m.camera()
    from conftest import load_gpu
    g, lt, le = load_gpu(m, cam)
    o, lt_o, le_o = load_oracle(oracle_mod, m, cam)
    big = synth.make_candidates(cfg, 10000, None).positions          # ~14 ms of ray sweep per call
    c = synth.make_candidates(cfg, 64, None)
    free_o = o.check_segments(le_o, cfg.robot_radius, c.segments)["free"]
    gain_ref = g.gain_eval(lt, big[:2000])["gain"]
    # idle latency of the single call
    for i in range(20):
        g.check_motion(le, cfg.robot_radius, c.segments[i, :3], c.segments[i, 3:])
    idle = []
    for i in range(200):
        t0 = time.perf_counter()
        g.check_motion(le, cfg.robot_radius, c.segments[i % 64, :3], c.segments[i % 64, 3:])
        idle.append(time.perf_counter() - t0)
    stop, errors, batch_ms = threading.Event(), [], []

    def planner():
        try:
            while not stop.is_set():
                t0 = time.perf_counter()
                r = g.gain_eval(lt, big, steps=False)
                batch_ms.append((time.perf_counter() - t0) * 1e3)
                if not np.array_equal(r["gain"][:2000], gain_ref):
                    errors.append(AssertionError("gain batch disturbed"))
        except Exception as e:  # pragma: no cover
            errors.append(e)

    th = threading.Thread(target=planner)
    th.start()
    time.sleep(0.05)                                                  # the batch loop is running
    busy, wrong = [], 0
    for i in range(400):
        t0 = time.perf_counter()
        f = g.check_motion(le, cfg.robot_radius, c.segments[i % 64, :3], c.segments[i % 64, 3:])
        busy.append(time.perf_counter() - t0)
        wrong += int(f != bool(free_o[i % 64]))
        time.sleep(0.0005)
    stop.set()
    th.join()
    assert not errors, errors
    assert wrong == 0
    med_idle, med_busy = np.median(idle) * 1e6, np.median(busy) * 1e6
    print(f"check_motion latency: idle {med_idle:.1f} us, with a {np.median(batch_ms):.1f} ms gain batch in flight {med_busy:.1f} us "
          f"(p90 {np.percentile(busy, 90) * 1e6:.1f} us), {len(batch_ms)} batches")
    assert len(batch_ms) >= 5 and np.median(batch_ms) > 5.0           # there really was a long batch in flight
    assert med_busy < 40.0, (med_idle, med_busy)                      # < 40 us (a queued call would take milliseconds)
    g.close()
