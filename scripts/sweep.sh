#!/bin/bash
# usage: scripts/sweep.sh   -- a few bench runs with experiment knobs, prints value per run
run() { env $1 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$2', round(d['value']), round(d['ms_per_step'],3), round(d['voxel_steps_per_s']/1e9,1))"; }
run "A=1" base
for yc in 20 30 60 72; do run "NBVGPU_YC=$yc NBVGPU_DENSE_KB=100" yc$yc; done
for pad in 16 40 60; do run "NBVGPU_SMEM_PAD_KB=$pad NBVGPU_DENSE_KB=200" pad$pad; done
