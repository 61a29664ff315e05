#!/bin/bash
# usage: profiles/collect_round.sh <gpurun_out tag> <profiles prefix>     e.g. profiles/collect_round.sh r2m r2
# Turns the scratch files of scripts/gpu_evidence.sh / scripts/scale_run.sh (gpurun_out/<tag>_*) into the tracked records
# profiles/<prefix>_*: bench lines as they were printed, the ncu launch list with its shares, and per capture the raw-page
# summary, the per-source-line shares and the per-region shares of the ray sweep; ncu_traffic.json gets the dram bytes per launch.
set -u
tag=$1; pre=$2
cd "$(dirname "$0")/.."
g=gpurun_out
cp $g/${tag}_bench_default.json profiles/${pre}_bench_default.json
cp $g/${tag}_bench_reference.json profiles/${pre}_bench_reference.json
for n in 2 4 8; do [ -s $g/${tag}_scale_n$n.json ] && cp $g/${tag}_scale_n$n.json profiles/${pre}_scale_n$n.json; done
[ -s $g/${tag}_multi_tests_n2.log ] && cp $g/${tag}_multi_tests_n2.log profiles/${pre}_multi_gpu_tests_n2.log
cat $g/${tag}_pytest.log $g/${tag}_smoke.log > profiles/${pre}_gpu_tests.log
grep -v "^==PROF==\|^==WARNING==" $g/${tag}_launches.csv > profiles/${pre}_launches.csv
python profiles/launch_shares.py profiles/${pre}_launches.csv "python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-l2-probe --extra=cfg2" > profiles/${pre}_launch_shares.txt
for c in cfg3 cfg2 cfg4; do
  rep=$g/${tag}_k_gain_$c.ncu-rep
  [ -s $rep ] || continue
  python profiles/ncu_summary.py $rep > profiles/${pre}_k_gain_${c}_summary.txt
  ncu -i $rep --page source --csv --print-source sass,cuda > /tmp/_src_$c.csv 2>/dev/null
  python profiles/ncu_lines.py /tmp/_src_$c.csv 40 > profiles/${pre}_k_gain_${c}_lines.txt
  python profiles/ncu_regions.py /tmp/_src_$c.csv nbv_exploration_planner_b200/csrc/gain_kernels.cuh >> profiles/${pre}_k_gain_${c}_lines.txt
  python profiles/summarize_ncu.py $rep $c "capture of the launch bench.py times for $c ($(basename $rep))" > /dev/null
done
rep=$g/${tag}_k_apply_cfg3.ncu-rep
if [ -s $rep ]; then
  python profiles/ncu_summary.py $rep > profiles/${pre}_k_apply_blocks_summary.txt
fi
ls -la profiles | grep " ${pre}_"
