"""Small batches (what the reference's own, sequential RrtTree::iterate makes of the drop-in: one candidate per call).
Up to 64 candidates the host entry points run without copy operations (kernels read and write mapped pinned memory in
place) and the dense ray sweep finishes with the best-yaw window scan in the last CTA of each candidate (one launch).
Results must equal the oracle's and the large-batch path's bit for bit, call after call (the arrival counters reset)."""
import numpy as np
import pytest

from conftest import load_gpu, load_oracle

pytestmark = pytest.mark.gpu


def _setup(oracle_mod, name="tiny", variant=0, **cam_kw):
    from nbv_exploration_planner_b200 import synth
    m = synth.make_map(name)
    cam = m.camera(**cam_kw)
    o, lt_o, le_o = load_oracle(oracle_mod, m, cam)
    g, lt_g, le_g = load_gpu(m, cam, variant)
    c = synth.make_candidates(m.cfg, 200, None)
    return m, o, lt_o, le_o, g, lt_g, le_g, c


@pytest.mark.parametrize("variant", [0, 2])
def test_small_batches_equal_oracle_and_large_batch(oracle_mod, variant):
    m, o, lt_o, le_o, g, lt_g, le_g, c = _setup(oracle_mod, variant=variant)
    pos = c.positions
    ro = o.gain_eval(lt_o, pos, threads=8)
    big = g.gain_eval(lt_g, pos, steps=False)            # 200 candidates: two launches, staged copies
    assert g.fused_launches() == 0
    assert np.array_equal(big["gain"], ro["gain"]) and np.array_equal(big["yaw_deg"], ro["yaw_deg"])
    at = 0
    for n in (1, 1, 2, 3, 7, 1, 33, 64, 1):
        r = g.gain_eval(lt_g, pos[at:at + n], steps=False)
        for k in ("gain", "yaw_deg", "counts"):
            assert np.array_equal(r[k], big[k][at:at + n]), (k, n, at)
        at += n
    assert g.fused_launches() == 9, "the fused sweep was not the one launched"
    # the same candidate again and again: the arrival counters must be back at zero after every launch
    for _ in range(5):
        st = np.array([*pos[5], 0.0])
        gain = g.optimized_gain(lt_g, st)
        assert gain == ro["gain"][5] and st[3] == ro["yaw_deg"][5] * np.pi / 180.0
    if variant == 2:
        assert g.classifier_mismatches() == 0
    g.close()


def test_small_batch_with_steps_or_per_yaw_keeps_their_outputs(oracle_mod):
    m, o, lt_o, le_o, g, lt_g, le_g, c = _setup(oracle_mod)
    pos = c.positions[:5]
    ro = o.gain_eval(lt_o, pos, per_yaw=True, threads=4)
    r = g.gain_eval(lt_g, pos, per_yaw=True, steps=True)
    for k in ("gain", "yaw_deg", "steps", "per_yaw"):
        assert np.array_equal(r[k], ro[k]), k
    g.close()


def test_sum_all_yaws_small_batch(oracle_mod):
    """RrtTree::gain (rrt.cpp:481-579) through the fused finish: total over all yaws, yaw 0."""
    m, o, lt_o, le_o, g, lt_g, le_g, c = _setup(oracle_mod, sum_all_yaws=1)
    pos = c.positions[:9]
    ro = o.gain_eval(lt_o, pos, threads=4)
    for i in range(9):
        r = g.gain_eval(lt_g, pos[i:i + 1], steps=False)
        assert r["gain"][0] == ro["gain"][i] and r["yaw_deg"][0] == 0
        assert np.array_equal(r["counts"][0], ro["counts"][i])
    assert g.fused_launches() == 9
    g.close()


def test_out_of_range_candidate_in_device_batch_finishes(oracle_mod):
    """A device-resident small batch with a candidate the sweep skips (non-finite): its CTAs still arrive, the window scan
    runs and reports gain 0 for it; the others are untouched."""
    import torch
    m, o, lt_o, le_o, g, lt_g, le_g, c = _setup(oracle_mod)
    pos = c.positions[:6].copy()
    ro = o.gain_eval(lt_o, pos, threads=4)
    pos[2] = [np.nan, 0.0, 0.0]
    d_pos = torch.from_numpy(pos).cuda()
    d_gain = torch.full((6,), -1.0, dtype=torch.float64, device="cuda")
    d_yaw = torch.full((6,), -7, dtype=torch.int32, device="cuda")
    for _ in range(2):
        g.gain_eval_device(lt_g, 6, d_pos.data_ptr(), d_gain.data_ptr(), d_yaw.data_ptr())
        g.synchronize()
        torch.cuda.synchronize()
        gain, yaw = d_gain.cpu().numpy(), d_yaw.cpu().numpy()
        keep = [0, 1, 3, 4, 5]
        assert np.array_equal(gain[keep], ro["gain"][keep]) and np.array_equal(yaw[keep], ro["yaw_deg"][keep])
        assert gain[2] == 0.0 and yaw[2] == 0
    assert g.fused_launches() == 2
    g.close()


def test_check_segments_small_equals_large(oracle_mod):
    m, o, lt_o, le_o, g, lt_g, le_g, c = _setup(oracle_mod)
    seg = c.segments
    fo = o.check_segments(le_o, m.cfg.robot_radius, seg)["free"]
    big = g.check_segments(le_g, m.cfg.robot_radius, seg)
    assert np.array_equal(big, fo)
    assert 0 < int(fo.sum()) < len(fo)
    at = 0
    for n in (1, 1, 5, 64, 1, 17):
        assert np.array_equal(g.check_segments(le_g, m.cfg.robot_radius, seg[at:at + n]), fo[at:at + n])
        at += n
    for i in range(20):
        assert g.check_motion(le_g, m.cfg.robot_radius, seg[i, :3], seg[i, 3:]) == bool(fo[i])
    g.close()


def test_interleaved_single_calls_stress(oracle_mod):
    """400 single-item and small-batch calls of all three kinds back to back (arrival counters, completion tickets and the
    shared staging area are reused call after call): every result equals the large-batch one."""
    m, o, lt_o, le_o, g, lt_g, le_g, c = _setup(oracle_mod)
    pos, seg = c.positions, c.segments
    big = g.gain_eval(lt_g, pos, steps=False)
    free = g.check_segments(le_g, m.cfg.robot_radius, seg)
    q = np.random.default_rng(5).uniform(m.cfg.bounds_min, m.cfg.bounds_max, (200, 3))
    interp = g.esdf_distance_gradient(le_g, q)
    rng = np.random.default_rng(11)
    for it in range(400):
        kind = int(rng.integers(0, 4))
        n = 1 if rng.random() < 0.6 else int(rng.integers(2, 65))
        at = int(rng.integers(0, 200 - n + 1))
        if kind == 0:
            r = g.gain_eval(lt_g, pos[at:at + n], steps=False)
            assert np.array_equal(r["gain"], big["gain"][at:at + n]) and np.array_equal(r["yaw_deg"], big["yaw_deg"][at:at + n]), (it, n, at)
            assert np.array_equal(r["counts"], big["counts"][at:at + n])
        elif kind == 1:
            assert np.array_equal(g.check_segments(le_g, m.cfg.robot_radius, seg[at:at + n]), free[at:at + n]), (it, n, at)
        elif kind == 2:
            r = g.esdf_distance_gradient(le_g, q[at:at + n])
            for k in ("ok", "distance", "gradient"):
                assert np.array_equal(r[k], interp[k][at:at + n]), (it, k, n, at)
        else:
            st = np.array([*pos[at], 0.0])
            assert g.optimized_gain(lt_g, st) == big["gain"][at] and st[3] == big["yaw_deg"][at] * np.pi / 180.0
    g.close()


def test_long_segments_one_warp_each(oracle_mod):
    """Segments far longer than 32 voxels, mostly through observed free space on a low radius (several rounds of 32
    look-ups before the first failing voxel), plus degenerate ones: flag and executed ESDF steps equal the oracle's."""
    from nbv_exploration_planner_b200 import synth
    m = synth.make_map("cfg1")
    cam = m.camera()
    o, lt_o, le_o = load_oracle(oracle_mod, m, cam)
    g, lt_g, le_g = load_gpu(m, cam, 0)
    rng = np.random.default_rng(77)
    lo, hi = np.array(m.cfg.bounds_min, float), np.array(m.cfg.bounds_max, float)
    traj = np.asarray(m.trajectory, float)
    a = traj[rng.integers(0, len(traj), 300)] + rng.uniform(-0.3, 0.3, (300, 3))
    b = traj[rng.integers(0, len(traj), 300)] + rng.uniform(-0.3, 0.3, (300, 3))
    seg = np.concatenate([a, b], axis=1)
    seg[:10, 3:] = seg[:10, :3]                      # zero length
    seg[10:20, 3] = seg[10:20, 0]                    # axis-aligned in x
    seg[20:30, 3:] = seg[20:30, :3] + [0.0, 0.0, 1.0]
    seg = np.clip(seg, np.tile(lo, 2) - 1.0, np.tile(hi, 2) + 1.0)
    radius = 0.05
    ro = o.check_segments(le_o, radius, seg)
    assert int(ro["steps"].max()) > 40 and 0 < int(ro["free"].sum()) < len(seg)
    before = g.stats()["collision_steps"]
    for at in range(0, 300, 50):
        assert np.array_equal(g.check_segments(le_g, radius, seg[at:at + 50]), ro["free"][at:at + 50]), at
    assert g.stats()["collision_steps"] - before == int(ro["steps"].sum())
    big = g.check_segments(le_g, radius, seg)         # one thread per segment
    assert np.array_equal(big, ro["free"])
    assert g.stats()["collision_steps"] - before == 2 * int(ro["steps"].sum())
    for i in (0, 15, 25, 100, 101, 102):
        assert g.check_motion(le_g, radius, seg[i, :3], seg[i, 3:]) == bool(ro["free"][i])
    g.close()


@pytest.mark.parametrize("env", ["NBVGPU_NO_FUSE", "NBVGPU_NO_SMALL_PATH", "NBVGPU_NO_POLL"])
def test_paths_switched_off_give_the_same_results(oracle_mod, env, monkeypatch):
    m, o, lt_o, le_o, g, lt_g, le_g, c = _setup(oracle_mod)
    pos = c.positions[:12]
    ref = [g.gain_eval(lt_g, pos[i:i + 3], steps=False) for i in range(0, 12, 3)]
    assert g.fused_launches() == 4
    g.close()
    monkeypatch.setenv(env, "1")
    g2, lt2, le2 = load_gpu(m, m.camera(), 0)
    for i, r in zip(range(0, 12, 3), ref):
        r2 = g2.gain_eval(lt2, pos[i:i + 3], steps=False)
        for k in ("gain", "yaw_deg", "counts"):
            assert np.array_equal(r2[k], r[k]), k
    if env == "NBVGPU_NO_FUSE":
        assert g2.fused_launches() == 0
    g2.close()


@pytest.mark.parametrize("n", [1, 10, 64, 65, 400])
def test_evaluate_candidates_equals_the_two_calls(oracle_mod, n):
    """nbvgpu_evaluate_candidates (one call) == nbvgpu_check_segments followed by nbvgpu_gain_eval_best, host and device
    variants, small and large batches; verdicts equal the oracle's, and so do the gains behind the free edges (a colliding
    edge gets no gain: rrt.cpp:210-216)."""
    import torch
    from nbv_exploration_planner_b200 import synth
    m, o, lt_o, le_o, g, lt_g, le_g, c0 = _setup(oracle_mod)
    c = synth.make_candidates(m.cfg, n, None)
    pos, seg, dist = c.positions, c.segments, c.distance
    pyaw = np.linspace(-3.0, 3.0, n)
    ro = o.gain_eval(lt_o, pos, threads=8)
    fo = o.check_segments(le_o, m.cfg.robot_radius, seg)["free"]
    want_gain, want_yaw = np.where(fo == 1, ro["gain"], 0.0), np.where(fo == 1, ro["yaw_deg"], 0)
    free = g.check_segments(le_g, m.cfg.robot_radius, seg)
    two = g.gain_eval_best(lt_g, pos, pyaw, dist, free, 1.0, 0.5, 0.0)
    for _ in range(2):
        one = g.evaluate_candidates(lt_g, le_g, m.cfg.robot_radius, pos, seg, pyaw, dist, 1.0, 0.5, 0.0)
        assert np.array_equal(one["free"], fo) and np.array_equal(one["gain"], want_gain) and np.array_equal(one["yaw_deg"], want_yaw)
        for k in ("index", "score", "best_gain", "best_yaw"):
            assert one[k] == two[k], k
    d = lambda a, t: torch.from_numpy(np.ascontiguousarray(a, t)).cuda()
    d_pos, d_seg, d_py, d_di = d(pos, np.float64), d(seg, np.float64), d(pyaw, np.float64), d(dist, np.float64)
    d_free = torch.full((n,), 9, dtype=torch.uint8, device="cuda")
    d_gain = torch.zeros(n, dtype=torch.float64, device="cuda")
    d_yaw = torch.zeros(n, dtype=torch.int32, device="cuda")
    d_best = torch.zeros(32, dtype=torch.uint8, device="cuda")
    stream = torch.cuda.Stream()
    g.set_stream(stream.cuda_stream)
    with torch.cuda.stream(stream):
        for _ in range(3):
            g.evaluate_candidates_device(lt_g, le_g, m.cfg.robot_radius, n, d_pos, d_seg, d_py, d_di, 1.0, 0.5, 0.0, d_free, d_best,
                                         d_gain, d_yaw)
        stream.synchronize()
    g.set_stream(None)
    assert np.array_equal(d_free.cpu().numpy(), fo) and np.array_equal(d_gain.cpu().numpy(), want_gain)
    rec = np.frombuffer(d_best.cpu().numpy().tobytes(), dtype=[("index", "<i8"), ("score", "<f8"), ("gain", "<f8"), ("yaw", "<f8")])[0]
    assert rec["index"] == two["index"] and rec["score"] == two["score"] and rec["gain"] == two["best_gain"] and rec["yaw"] == two["best_yaw"]
    g.close()


@pytest.mark.parametrize("n", [40, 700])
def test_planning_step_graph_replays_equal_eager(oracle_mod, n, monkeypatch):
    """The planning step as a CUDA graph: first call eager, second captured, then replays -- host-buffer and device-resident
    forms, with profiling events inside the graph -- all bit-equal to the eager path (NBVGPU_NO_GRAPH=1) and to new inputs
    written into the same buffers between replays."""
    import torch
    from nbv_exploration_planner_b200 import synth
    m, o, lt_o, le_o, g, lt_g, le_g, c0 = _setup(oracle_mod)
    c = synth.make_candidates(m.cfg, 2 * n, None)
    pyaw = np.linspace(-3.0, 3.0, n)
    sets = [(c.positions[:n], c.segments[:n], c.distance[:n]), (c.positions[n:], c.segments[n:], c.distance[n:])]
    monkeypatch.setenv("NBVGPU_NO_GRAPH", "1")
    ge, lt_e, le_e = _setup(oracle_mod)[4:7]
    want = [ge.evaluate_candidates(lt_e, le_e, m.cfg.robot_radius, p, s, pyaw, d, 1.0, 0.5, 0.0) for p, s, d in sets]
    assert ge.graph_replays() == 0
    ge.close()
    monkeypatch.delenv("NBVGPU_NO_GRAPH")
    for rep in range(6):                      # host-buffer form: alternate the inputs, the shape stays
        p, s, d = sets[rep % 2]
        r = g.evaluate_candidates(lt_g, le_g, m.cfg.robot_radius, p, s, pyaw, d, 1.0, 0.5, 0.0)
        w = want[rep % 2]
        for k in ("index", "score", "best_gain", "best_yaw"):
            assert r[k] == w[k], (rep, k)
        for k in ("gain", "yaw_deg", "free"):
            assert np.array_equal(r[k], w[k]), (rep, k)
    assert g.graph_replays() >= 4
    # device-resident form on a caller's stream, with the sweep's profiling events inside the graph
    dv = lambda a, t: torch.from_numpy(np.ascontiguousarray(a, t)).cuda()
    d_pos, d_seg, d_py, d_di = dv(sets[0][0], np.float64), dv(sets[0][1], np.float64), dv(pyaw, np.float64), dv(sets[0][2], np.float64)
    d_free = torch.zeros(n, dtype=torch.uint8, device="cuda")
    d_gain = torch.zeros(n, dtype=torch.float64, device="cuda")
    d_yaw = torch.zeros(n, dtype=torch.int32, device="cuda")
    d_best = torch.zeros(32, dtype=torch.uint8, device="cuda")
    stream = torch.cuda.Stream()
    g.set_stream(stream.cuda_stream)
    before = g.graph_replays()
    g.set_profiling(True)
    with torch.cuda.stream(stream):
        for rep in range(6):
            src = sets[rep % 2]
            d_pos.copy_(dv(src[0], np.float64)); d_seg.copy_(dv(src[1], np.float64)); d_di.copy_(dv(src[2], np.float64))
            g.evaluate_candidates_device(lt_g, le_g, m.cfg.robot_radius, n, d_pos, d_seg, d_py, d_di, 1.0, 0.5, 0.0, d_free, d_best, d_gain, d_yaw)
            stream.synchronize()
            assert np.array_equal(d_gain.cpu().numpy(), want[rep % 2]["gain"]) and np.array_equal(d_free.cpu().numpy(), want[rep % 2]["free"])
    ms, launches = g.profile_collect()
    g.set_profiling(False)
    g.set_stream(None)
    assert g.graph_replays() - before >= 4
    assert launches == 6 and 0.0 < ms < 50.0, (ms, launches)      # every step's sweep was timed, replayed or not
    g.close()
