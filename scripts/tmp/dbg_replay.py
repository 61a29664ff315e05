import sys, os, numpy as np, torch
sys.path.insert(0, os.getcwd())
from nbv_exploration_planner_b200 import LAYER_ESDF, LAYER_TSDF, NbvGpu, synth
cfg = synth.CONFIGS["cfg5"]
rng = np.random.Generator(np.random.MT19937(cfg.seed_map))
objects = synth._world_objects(cfg, rng); traj = synth._trajectory(cfg, rng)
all_keys = synth.blocks_touching(cfg.voxel_size, cfg.vps, traj, cfg.obs_radius, cfg.bounds_min, cfg.bounds_max)
for mode in ("plain", "multi"):
    g = NbvGpu(0) if mode == "plain" else NbvGpu.multi([0])
    lt = g.layer_create(LAYER_TSDF, cfg.voxel_size, cfg.vps, len(all_keys) + 64)
    le = g.layer_create(LAYER_ESDF, cfg.voxel_size, cfg.vps, len(all_keys) + 64)
    for k in range(60):
        keys = synth.blocks_touching(cfg.voxel_size, cfg.vps, traj[k:k + 1], cfg.obs_radius, cfg.bounds_min, cfg.bounds_max)
        tsdf, esdf = synth.generate_blocks(cfg, keys, objects, traj[:k + 1])
        pt, pe = torch.from_numpy(tsdf).pin_memory(), torch.from_numpy(esdf).pin_memory()
        try:
            g.layer_update_pinned(lt, keys, pt); g.layer_update_pinned(le, keys, pe); g.commit()
        except Exception as e:
            print(mode, "FAIL at", k, len(keys), hex(pt.data_ptr()), hex(pe.data_ptr()), pt.is_pinned(), pt.shape, e); break
    else:
        print(mode, "ok", g.layer_num_blocks(lt))
    g.close()
