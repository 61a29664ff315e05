#!/bin/bash
# usage: scripts/ab.sh <tag> "<ENV=.. ENV=..>|<bench args>" ...   -- A/B runs of bench.py on one box; prints value / frac per run
tag=$1; shift
mkdir -p gpurun_out
i=0
for spec in "$@"; do
  envs=${spec%%|*}; args=${spec#*|}
  i=$((i+1))
  env $envs timeout 900 python bench.py $args --no-cpu-baseline --no-l2-probe --extra= > gpurun_out/${tag}_$i.json 2> gpurun_out/${tag}_$i.err
  python - "$spec" gpurun_out/${tag}_$i.json <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[2]).read().strip().splitlines()[-1]); r=d.get('roofline',{})
    print(sys.argv[1], '->', round(d['value']), 'vp/s', round(d['ms_per_step'],3), 'ms  frac', round(r.get('frac',0),4), 'kernel ms', round(r.get('avg_launch_ms',0),3), 'gap us', r.get('step_minus_kernel_us'))
except Exception as e:
    print(sys.argv[1], 'FAILED', e); print(open(sys.argv[2].replace('.json','.err')).read()[-800:])
P
done
