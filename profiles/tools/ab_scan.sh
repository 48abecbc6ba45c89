#!/bin/bash
# A/B of the scan's tile size and look-back window: rebuilds the library on the GPU box for every variant, checks the scan
# against torch.cumsum and times it (profiles/tools/time_scan.py); the default build is restored at the end.
cd "$(dirname "$0")/../.."
T=${1:-r02}
F="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC -shared -I skan_b200/csrc -I include"
for v in "16 0" "32 0" "16 1" "32 1"; do
  set -- $v
  nvcc $F -DSKB_SCAN_ITEMS=$1 -DSKB_SCAN_WIDE=$2 -o skan_b200/lib/libskan_b200.so skan_b200/csrc/skb_cuda.cu 2>/dev/null || { echo "build failed $v"; continue; }
  echo "== items $1 wide $2"
  timeout 120 python profiles/tools/time_scan.py 2>&1 | tail -5
  timeout 200 python bench.py --workload volume_2048 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-others 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('volume_2048', round(d['ms_per_step'],4))"
done > gpurun_out/${T}_ab_scan.txt 2>&1
python -c "import __graft_entry__ as g; g.build(force=True)" > /dev/null 2>&1
cat gpurun_out/${T}_ab_scan.txt
