#!/bin/bash
# round-1 final captures on one GPU: tests, bench lines, ncu launch list, ncu --set full of one pass
cd "$(dirname "$0")/../.."
O=gpurun_out
timeout 250 python -m pytest tests -m gpu -x -q > $O/r01c_tests.log 2>&1; tail -2 $O/r01c_tests.log
timeout 200 python bench.py --stages > $O/r01c_bench_volume_2048_n1.json 2> $O/r01c_bench_volume_2048_n1.err; echo bench rc=$?
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > $O/r01c_bench_reference.json 2> $O/r01c_bench_reference.err; echo ref rc=$?
for w in volume_1024 image_16384 batch_512; do timeout 150 python bench.py --workload $w --stages --no-cpu-baseline > $O/r01c_bench_${w}_n1.json 2> $O/r01c_bench_${w}_n1.err; echo $w rc=$?; done
B="python bench.py --workload volume_2048 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e"
timeout 100 $B > $O/r01c_plain.log 2>&1; echo plain rc=$?
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/r01c_launches_volume_2048.csv $B > $O/r01c_ncu_list.log 2>&1; echo list rc=$?
SKIP=$(python - <<'PY'
import csv
lines = [l for l in open('gpurun_out/r01c_launches_volume_2048.csv') if not l.startswith('==')]
rows = list(csv.DictReader(lines))
idx = [i for i, r in enumerate(rows) if 'skb_pack' in r['Kernel Name']]
print(idx[-2], idx[-1] - idx[-2])
PY
)
set -- $SKIP
echo "skip $1 count $2"
timeout 400 ncu --set full --clock-control none --import-source on --launch-skip $1 -c $2 -f -o $O/r01c_prof $B > $O/r01c_ncu_full.log 2>&1; echo full rc=$?
