#!/bin/bash
# A/B of programmatic dependent launch on one GPU: the option on / off on the four workloads (step time + stage clock)
cd "$(dirname "$0")/../.."
T=${1:-r02}
for w in volume_1024 volume_2048 image_16384 batch_512; do
  AB_ONLY="default,plain" timeout 300 python profiles/tools/ab_walks.py $w > gpurun_out/${T}_ab_pdl_$w.txt 2> gpurun_out/${T}_ab_pdl_$w.err; echo "$w rc=$?"
  cut -c1-90 gpurun_out/${T}_ab_pdl_$w.txt
done
