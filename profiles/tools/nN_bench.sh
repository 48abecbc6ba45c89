#!/bin/bash
# bench.py on N GPUs of one box with the host timeline of the slab runner: bash profiles/tools/nN_bench.sh <N> <tag> [extra bench args]
cd "$(dirname "$0")/../.."
N=$1; T=$2; shift 2
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 3 --timeline "$@" > gpurun_out/${T}_bench_n$N.json 2> gpurun_out/${T}_bench_n$N.err; echo bench rc=$?
grep -E "timeline|PARITY" gpurun_out/${T}_bench_n$N.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${T}_bench_n$N.json").read().strip().splitlines()[-1])
p=d.get("parity") or {}
print("ms", d["ms_per_step"], "median", d["ms_median"], "digest", p.get("digest"), "parity_ok", p.get("ok"), "e2e ms", (d.get("e2e") or {}).get("ms_per_step"), "sweep ms", d["roofline"]["kernel_ms"])
PY
