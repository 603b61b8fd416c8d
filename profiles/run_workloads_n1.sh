#!/bin/bash
# the other BASELINE configs on one GPU, final code
mkdir -p gpurun_out
for w in C2 C4 C5; do
  timeout 200 python bench.py --workload $w --steps 100 --warmup 6 --no-mapping --no-tracking --no-loop --no-cpu > gpurun_out/final_ours_${w}_n1.json 2> gpurun_out/final_ours_${w}_n1.err
  echo $w; grep -o "\"ms_per_step\": [0-9.]*" gpurun_out/final_ours_${w}_n1.json | head -3 | tr '\n' ' '
done
