#!/bin/bash
# Runs the multi-GPU configurations of BASELINE.json on one 8-GPU box (use under gpurun --gpus 8).
# Usage: scripts/run_multi_gpu.sh <N> <tag>
N=${1:-8}; TAG=${2:-r1}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
mkdir -p gpurun_out
$TR --master-port 29601 bench.py --gpus $N --config cfg2 --steps 30 --warmup 3 > gpurun_out/${TAG}_cfg2_n$N.json 2> gpurun_out/${TAG}_cfg2_n$N.err
$TR --master-port 29602 bench.py --gpus $N --config cfg3 --steps 10 --warmup 3 --selfcheck > gpurun_out/${TAG}_cfg3_n$N.json 2> gpurun_out/${TAG}_cfg3_n$N.err
$TR --master-port 29603 bench.py --gpus $N --config cfg5 --steps 200 > gpurun_out/${TAG}_cfg5_n$N.json 2> gpurun_out/${TAG}_cfg5_n$N.err
$TR --master-port 29604 bench.py --gpus $N --config cfg4 --steps 3 --warmup 3 --selfcheck > gpurun_out/${TAG}_cfg4_n$N.json 2> gpurun_out/${TAG}_cfg4_n$N.err
for c in cfg2 cfg3 cfg5 cfg4; do
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${TAG}_${c}_n$N.json"))
    print("$c N=$N value", round(d["value"]), "ms/step", round(d["ms_per_step"], 3), "Gsteps/s", round(d.get("voxel_steps_per_s", 0) / 1e9, 1),
          "e2e", round(d["e2e"]["value"]), "selfcheck", d.get("selfcheck"), "bcast_ms", d["config"].get("map_broadcast_ms_setup"),
          "blocks", d["config"].get("map_blocks"), "incl_bcast", round(d.get("value_including_map_broadcast", 0)))
except Exception as e:
    print("$c failed:", e)
    print(open("gpurun_out/${TAG}_${c}_n$N.err").read()[-1500:])
PY
done
