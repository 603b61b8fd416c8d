#!/bin/bash
# full ncu capture of the plan / emit kernels of one emulated rank (world 8, contiguous tile runs, two-pass projection)
mkdir -p gpurun_out
PROBE_WORLDS=8 PROBE_LAYOUTS=bands ncu --set full --clock-control none --import-source on \
  -k regex:'k_surfel_candidates|k_surfel_forward|k_emit' -s 9 -c 3 -f -o gpurun_out/shard_plan_full \
  python profiles/shard_plan_probe.py C3 > gpurun_out/shard_plan_full.log 2>&1
tail -3 gpurun_out/shard_plan_full.log
ls -la gpurun_out/*.ncu-rep
