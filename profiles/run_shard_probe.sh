#!/bin/bash
# one-GPU emulation of one rank's plan / render at world 2,4,8 (profiles/shard_plan_probe.py), both projection variants
mkdir -p gpurun_out
for tp in 0 default; do
  if [ $tp = default ]; then unset EGS_SHARD_TWO_PASS; else export EGS_SHARD_TWO_PASS=$tp; fi
  python profiles/shard_plan_probe.py C3 2>&1 | grep -E "world|two_pass|Error|error"
done | tee gpurun_out/shard_plan_probe.txt
