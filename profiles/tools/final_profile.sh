#!/bin/bash
# final captures of a round on one GPU: bench line, ncu launch list, ncu --set full of one pass, reference arm,
# then the other workloads (in that order: what comes last is what a short GPU budget loses)
cd "$(dirname "$0")/../.."
O=gpurun_out
T=${1:-r01g}
timeout 200 python bench.py --stages > $O/${T}_bench_volume_2048_n1.json 2> $O/${T}_bench_volume_2048_n1.err; echo bench rc=$?
B="python bench.py --workload volume_2048 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-others"
timeout 100 $B > $O/${T}_plain.log 2>&1; echo plain rc=$?
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/${T}_launches_volume_2048.csv $B > $O/${T}_ncu_list.log 2>&1; echo list rc=$?
SKIP=$(T=$T python - <<'PY'
import csv, os
lines = [l for l in open(f'gpurun_out/{os.environ["T"]}_launches_volume_2048.csv') if not l.startswith('==')]
rows = list(csv.DictReader(lines))
idx = [i for i, r in enumerate(rows) if 'skb_pack' in r['Kernel Name']]
print(idx[-2], idx[-1] - idx[-2])
PY
)
set -- $SKIP
echo "skip $1 count $2"
# (ncu's kernel replay fails on launches with the programmatic-serialisation attribute: the capture runs with plain launches)
SKAN_B200_PDL=0 timeout 300 ncu --set full --clock-control none --import-source on --launch-skip $1 -c $2 -f -o $O/${T}_prof $B > $O/${T}_ncu_full.log 2>&1; echo full rc=$?
timeout 120 python bench.py --impl reference --steps 2 --warmup 1 > $O/${T}_bench_reference.json 2> $O/${T}_bench_reference.err; echo ref rc=$?
for w in volume_1024 image_16384 batch_512; do timeout 100 python bench.py --workload $w --stages --no-cpu-baseline > $O/${T}_bench_${w}_n1.json 2> $O/${T}_bench_${w}_n1.err; echo $w rc=$?; done
