#!/bin/bash
# slab-mode check on N GPUs of one box: the GPU slab tests (world sizes up to N are REQUIRED to run), then the bench line
# with the host timeline: bash profiles/tools/nN_check.sh <N> <tag> [extra bench args]
cd "$(dirname "$0")/../.."
N=$1; T=$2; shift 2
SKAN_B200_REQUIRE_RANKS=$N timeout 900 python -m pytest tests/test_gpu_slab.py -m gpu -x -q > gpurun_out/${T}_slab_tests.log 2>&1; echo tests rc=$?; tail -3 gpurun_out/${T}_slab_tests.log
bash profiles/tools/nN_bench.sh $N $T "$@"
