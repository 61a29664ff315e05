#!/bin/bash
# usage (under gpurun --gpus N): scripts/scale_run.sh <N> <tag> [tests]
# The default bench (config 3 strong scaling + config 2 / 5 / 4 sub-records) launched the way the driver launches it, one rank per
# GPU; with `tests` also the multi-GPU hardware tests (N-GPU == 1-GPU bit for bit, one process on all GPUs and one process per GPU).
N=$1; tag=$2; shift 2
mkdir -p gpurun_out
if [ "${1:-}" = tests ]; then
  timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_cpp_dropin.py -m gpu -q > gpurun_out/${tag}_multi_tests_n$N.log 2>&1
  echo "multi-GPU tests rc=$?"; tail -3 gpurun_out/${tag}_multi_tests_n$N.log
fi
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 \
  bench.py --gpus $N > gpurun_out/${tag}_scale_n$N.json 2> gpurun_out/${tag}_scale_n$N.err
echo "bench rc=$?"
python scripts/show_bench.py gpurun_out/${tag}_scale_n$N.json || tail -20 gpurun_out/${tag}_scale_n$N.err
