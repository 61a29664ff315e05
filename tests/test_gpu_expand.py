"""batched tree expansion on device (the body of RrtTree::iterate, rrt.cpp:172-246)
against the oracle's restatement: same rejected samples, parents, clipped positions, collision verdicts,
gains, yaws, distances, scores and best node -- bit for bit."""
import numpy as np
import pytest

from conftest import load_gpu, load_oracle

pytestmark = pytest.mark.gpu


def _tree(cfg, o, le_o, lt_o, n_nodes, rng):
    """An existing tree: root + nodes grown by the generator, each with the yaw its gain evaluation chose."""
    from nbv_exploration_planner_b200 import synth
    c = synth.make_candidates(cfg, n_nodes - 1, lambda s: o.check_segments(le_o, cfg.robot_radius, s)["free"])
    yaws = o.gain_eval(lt_o, c.positions)["yaw"]
    state = np.zeros((n_nodes, 4))
    state[0, :3] = cfg.root
    state[0, 3] = rng.uniform(-3, 3)
    state[1:, :3] = c.positions
    state[1:, 3] = yaws
    dist = np.concatenate([[0.0], c.distance])
    return state, dist


def _check(a, b):
    assert np.array_equal(a["status"], b["status"])
    assert np.array_equal(a["parent"], b["parent"])
    for k in ("pos", "gain", "yaw", "distance", "score"):
        assert np.array_equal(a[k], b[k], equal_nan=True), k   # a zero-length edge scores 0/0 = NaN on both sides
    assert a["best_index"] == b["best_index"]


@pytest.mark.parametrize("which", ["tiny", "cfg1"])
def test_expand_tree_matches_oracle(oracle_mod, tiny_map, cfg1_map, which):
    m = tiny_map if which == "tiny" else cfg1_map
    cfg = m.cfg
    cam = m.camera()
    o, lt_o, le_o = load_oracle(oracle_mod, m, cam)
    g, lt_g, le_g = load_gpu(m, cam)
    rng = np.random.default_rng(7)
    state, dist = _tree(cfg, o, le_o, lt_o, 40, rng)
    params = {"root": cfg.root, "vicinity": cfg.vicinity, "extension_range": cfg.extension_range, "overshoot": cfg.overshoot,
              "robot_radius": cfg.robot_radius, "v_max": 1.0, "dyaw_max": 0.75, "zero_gain": 0.0}
    u = rng.random((300 if which == "cfg1" else 200, 3))
    ro = o.expand_tree(lt_o, le_o, params, state, dist, u, threads=16)
    rg = g.expand_tree(lt_g, le_g, params, state, dist, u)
    _check(rg, ro)
    assert set(np.unique(ro["status"])) == {0, 1, 2}       # rejected samples, colliding edges and new nodes all occur
    assert ro["best_index"] >= 0 and ro["status"][ro["best_index"]] == 2
    # every accepted node obeys the tree rule
    ok = ro["status"] == 2
    d = np.linalg.norm(rg["pos"][ok] - state[rg["parent"][ok], :3], axis=1)
    assert np.all(d <= cfg.extension_range * (1 + 1e-12))
    # n = 1 calls == one iterate() each; a high zero_gain leaves no best node
    for i in np.nonzero(ok)[0][:5]:
        one = g.expand_tree(lt_g, le_g, params, state, dist, u[i:i + 1])
        assert one["status"][0] == 2 and one["gain"][0] == ro["gain"][i] and one["score"][0] == ro["score"][i]
        assert one["best_index"] == (0 if ro["score"][i] > 0 else -1)
    hi = dict(params, zero_gain=1e30)
    assert g.expand_tree(lt_g, le_g, hi, state, dist, u)["best_index"] == -1
    g.close()


def test_expand_tree_sum_all_mode_and_edge_cases(oracle_mod, tiny_map):
    cfg = tiny_map.cfg
    cam = tiny_map.camera(sum_all_yaws=1)               # RrtTree::gain: yaw 0 for every node
    o, lt_o, le_o = load_oracle(oracle_mod, tiny_map, cam)
    g, lt_g, le_g = load_gpu(tiny_map, cam)
    rng = np.random.default_rng(11)
    state = np.array([[*cfg.root, 0.3]])
    dist = np.array([0.0])
    params = {"root": cfg.root, "vicinity": cfg.vicinity, "extension_range": 0.6, "overshoot": 0.25, "robot_radius": 0.3,
              "v_max": 1.0, "dyaw_max": 0.5, "zero_gain": 1.0}
    u = rng.random((128, 3))
    u[:4] = [[0.5, 0.5, 0.5], [0.0, 0.0, 0.0], [1.0 - 1e-16, 0.5, 0.5], [0.5, 0.5 + 1e-9, 0.5]]   # centre, corner, rim, tiny step
    _check(g.expand_tree(lt_g, le_g, params, state, dist, u), o.expand_tree(lt_o, le_o, params, state, dist, u, threads=8))
    assert len(g.expand_tree(lt_g, le_g, params, state, dist, np.zeros((0, 3)))["status"]) == 0
    g.close()


@pytest.mark.parametrize("which,n", [("tiny", 64), ("tiny", 400), ("cfg1", 300)])
def test_sequential_expand_tree_matches_oracle(oracle_mod, tiny_map, cfg1_map, which, n):
    """nbvgpu_expand_params.sequential = 1: n consecutive iterate() calls on one stream of triples -- every sample sees the
    nodes the samples before it created (as parent, and through the parent's chosen yaw in its score).  Bit-equal to the
    oracle's sequential restatement, which tests/test_ref_pinning.py pins against the reference's own rrt.cpp:163-264 with its
    own kd-tree."""
    m = tiny_map if which == "tiny" else cfg1_map
    cfg = m.cfg
    cam = m.camera()
    o, lt_o, le_o = load_oracle(oracle_mod, m, cam)
    g, lt_g, le_g = load_gpu(m, cam)
    rng = np.random.default_rng(19 + n)
    state, dist = _tree(cfg, o, le_o, lt_o, 6, rng)
    params = {"root": cfg.root, "vicinity": cfg.vicinity, "extension_range": cfg.extension_range, "overshoot": cfg.overshoot,
              "robot_radius": cfg.robot_radius, "v_max": 1.0, "dyaw_max": 0.75, "zero_gain": 0.0}
    u = rng.random((n, 3))
    ro = o.expand_tree(lt_o, le_o, params, state, dist, u, threads=16, sequential=True)
    rg = g.expand_tree(lt_g, le_g, params, state, dist, u, sequential=True)
    _check(rg, ro)
    made = ro["status"] == 2
    assert (ro["parent"][made] >= len(state)).sum() >= made.sum() // 3      # many parents were created by the same call
    # scores of children of same-batch parents really use that parent's yaw
    kids = np.nonzero(made & (ro["parent"] >= len(state)))[0]
    created = np.nonzero(made)[0]
    for i in kids[:10]:
        p_triple = created[ro["parent"][i] - len(state)]
        dy = ro["yaw"][i] - ro["yaw"][p_triple]
        dy = dy - 2 * np.pi if dy > np.pi else (dy + 2 * np.pi if dy < -np.pi else dy)
        assert rg["score"][i] == ro["gain"][i] / max(dy / 0.75, ro["distance"][i] / 1.0)
    # the independent mode gives a different tree from the same stream; n = 1 is the same in both modes
    ind = g.expand_tree(lt_g, le_g, params, state, dist, u)
    assert not np.array_equal(ind["parent"], rg["parent"])
    one_s, one_i = g.expand_tree(lt_g, le_g, params, state, dist, u[:1], sequential=True), g.expand_tree(lt_g, le_g, params, state, dist, u[:1])
    _check(one_s, one_i)
    g.close()
