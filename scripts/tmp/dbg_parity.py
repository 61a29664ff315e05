import sys, os, numpy as np
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
from nbv_exploration_planner_b200 import synth
from conftest import load_gpu, load_oracle
from oracle import binding
binding.build()
m = synth.make_map("cfg1"); cam = m.camera(); cfg = m.cfg
o, lt_o, le_o = load_oracle(binding, m, cam)
c = synth.make_candidates(cfg, 100, None)
ro = o.gain_eval(lt_o, c.positions, per_yaw=True, threads=16)
for variant in (0, 2, 4, 1):
    g, lt, le = load_gpu(m, cam, variant)
    rg = g.gain_eval(lt, c.positions, per_yaw=True)
    d = (rg["per_yaw"].astype(np.int64) - ro["per_yaw"].astype(np.int64))
    bad = np.argwhere(d.any(axis=2))
    print("variant", variant, "mismatching (cand,yaw):", len(bad), "steps equal", np.array_equal(rg["steps"], ro["steps"]), "mism counter", g.classifier_mismatches())
    for v, y in bad[:12]:
        print("   cand", v, "yaw", y - 180, "gpu", rg["per_yaw"][v, y], "oracle", ro["per_yaw"][v, y], "steps", rg["steps"][v], ro["steps"][v])
    g.close()
