#!/bin/bash
# ncu launch list and `--set full` capture of one steady-state pass of the default workload (plain launches: ncu's kernel
# replay fails on launches with the programmatic-serialisation attribute): bash profiles/tools/ncu_full.sh <tag>
cd "$(dirname "$0")/../.."
O=gpurun_out
T=${1:-r02}
export SKAN_B200_PDL=0
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
timeout 400 ncu --set full --clock-control none --import-source on --launch-skip $1 -c $2 -f -o $O/${T}_prof $B > $O/${T}_ncu_full.log 2>&1; echo full rc=$?
tail -3 $O/${T}_ncu_full.log
