#!/bin/bash
# the round's multi-GPU lines on N GPUs of one box: volume_2048 in N Z-slabs (host timeline, e2e) and batch_512 in N shares
# bash profiles/tools/nN_final.sh <N> <tag>
cd "$(dirname "$0")/../.."
N=$1; T=$2
bash profiles/tools/nN_bench.sh $N $T
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 20 --warmup 3 --workload batch_512 --no-e2e > gpurun_out/${T}_bench_batch_n$N.json 2> gpurun_out/${T}_bench_batch_n$N.err; echo batch rc=$?
python - <<PY
import json
d=json.loads(open("gpurun_out/${T}_bench_batch_n$N.json").read().strip().splitlines()[-1])
print("batch_512 ms", d["ms_per_step"], "median", d["ms_median"])
PY
