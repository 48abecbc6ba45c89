#!/bin/bash
# quick single-GPU check of a code change: GPU parity tests of the hot path, then stage clocks of the four workloads
# usage: gpurun -- 'bash profiles/tools/quick_bench.sh <tag> [pytest -k expression]'
cd "$(dirname "$0")/../.."
T=${1:-quick}
K=${2:-"golden or oracle or real_image or hints"}
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "$K" > $O/${T}_tests.log 2>&1; echo tests rc=$?; tail -2 $O/${T}_tests.log
for w in volume_2048 volume_1024 image_16384 batch_512; do
  timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --stages --no-cpu-baseline --no-e2e --no-others > $O/${T}_bench_$w.json 2> $O/${T}_bench_$w.err
  echo "$w rc=$? $(python -c "import json,sys; d=json.loads(open('$O/${T}_bench_$w.json').read().strip().splitlines()[-1]); print(round(d['ms_per_step'],4), 'ms  job frac', round(d['roofline']['job']['frac_of_n_gpus'],4))")"
  grep -E "^  [a-z_]+ +[0-9.]+ ms" $O/${T}_bench_$w.err | tr -s ' ' | tr '\n' ';'; echo
done
