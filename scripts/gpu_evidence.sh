#!/bin/bash
# usage (on the GPU box, through gpurun): scripts/gpu_evidence.sh <tag> [tests] [bench] [ref] [launches] [ncu3] [ncu2] [ncu4] [ncuup]
# Collects the round's measured evidence into gpurun_out/<tag>_*: the GPU test suite, the default bench line, the reference arm,
# the ncu launch list of a short default-shaped run, and one `ncu --set full` capture per launch shape of the ray sweep
# (config 3: 512-thread dense, config 2: 512-thread 90-yaw dense, config 4: slot grid / cluster) and of the commit's scatter kernel.
# Every ncu pass runs only after the same command has exited 0 without ncu; numbers printed under ncu are never bench values.
set -u
tag=$1; shift
what=" $* "
[ "$what" = "  " ] && what=" tests bench ref launches ncu3 ncu2 ncu4 ncuup "
out=gpurun_out
mkdir -p $out
has() { case "$what" in *" $1 "*) return 0;; *) return 1;; esac; }
short="--no-cpu-baseline --no-l2-probe --extra="

if has tests; then
  timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $out/${tag}_pytest.log
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; echo "smoke rc=$?" | tee -a $out/${tag}_smoke.log
fi
if has bench; then
  ( time timeout 1500 python bench.py > $out/${tag}_bench_default.json 2> $out/${tag}_bench_default.err ) 2> $out/${tag}_bench_default.time; echo "bench rc=$?"; tail -3 $out/${tag}_bench_default.time
fi
if has ref; then
  timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > $out/${tag}_bench_reference.json 2> $out/${tag}_bench_reference.err; echo "ref rc=$?"
fi
if has launches; then
  cmd="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-l2-probe --extra=cfg2"
  timeout 600 $cmd > $out/${tag}_launches_plain.json 2>&1 && \
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file $out/${tag}_launches.csv $cmd > $out/${tag}_launches_ncu.log 2>&1
  echo "launches rc=$?"
fi
cap() {  # cap <name> <kernel regex> <skip> <bench args...>
  name=$1; kre=$2; skip=$3; shift 3
  timeout 600 python bench.py "$@" $short > $out/${tag}_${name}_plain.json 2>&1 || { echo "$name plain run failed"; return; }
  timeout 1200 ncu --set full --clock-control none --import-source on -k "regex:$kre" -s $skip -c 1 -f -o $out/${tag}_${name} \
     python bench.py "$@" $short > $out/${tag}_${name}_ncu.log 2>&1
  echo "$name ncu rc=$?"
}
has ncu3 && cap k_gain_cfg3 '^k_gain' 3 --config cfg3 --steps 1 --warmup 3
has ncu2 && cap k_gain_cfg2 '^k_gain' 3 --config cfg2 --steps 1 --warmup 3
has ncu4 && cap k_gain_cfg4 '^k_gain' 1 --config cfg4 --steps 1 --warmup 1
has ncuup && cap k_apply_cfg3 '^k_apply_blocks' 1 --config cfg3 --candidates 500 --steps 1 --warmup 1
ls -la $out | grep " ${tag}_" | awk '{print $5, $9}'
