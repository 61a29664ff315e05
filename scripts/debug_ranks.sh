#!/bin/bash
# usage (gpurun --gpus 2): scripts/debug_ranks.sh  -- the two-rank worker of tests/test_gpu_multi.py with a stack dump if it hangs
mkdir -p gpurun_out /tmp/dbg
rm -f /tmp/dbg/*
export NBVGPU_PULL_MIN_MB=0
for r in 0 1; do
  timeout 170 python -X faulthandler tests/multi_rank_worker.py $r 2 /tmp/dbg/uid.bin /tmp/dbg/res 1001 > gpurun_out/dbg_rank$r.log 2>&1 &
done
wait
for r in 0 1; do echo "== rank $r"; tail -25 gpurun_out/dbg_rank$r.log; done
ls -la /tmp/dbg
