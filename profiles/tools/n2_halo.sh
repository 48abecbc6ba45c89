SKAN_B200_REQUIRE_RANKS=2 python -m pytest tests/test_gpu_slab.py -m gpu -x -q -k "halo_exchange" > gpurun_out/r02k_slab_tests.log 2>&1; echo tests rc=$?; tail -3 gpurun_out/r02k_slab_tests.log
for m in exchange input; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 --halo-mode $m --no-cpu-baseline > gpurun_out/r02k_bench_n2_$m.json 2> gpurun_out/r02k_bench_n2_$m.err; echo bench $m rc=$?
  python - <<PY
import json
d=json.loads(open("gpurun_out/r02k_bench_n2_$m.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["parity"]["digest"], d["e2e"]["ms_per_step"], d["e2e"]["h2d_bytes_per_step"])
PY
done
