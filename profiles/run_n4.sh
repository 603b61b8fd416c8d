#!/bin/bash
# 8-GPU check + bench
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --steps 102 --warmup 6 --no-loop > gpurun_out/final_ours_n4.json 2> gpurun_out/final_ours_n4.err
grep -E "bench:|Error|error" gpurun_out/final_ours_n4.err | head -5
grep -o "\"ms_per_step\": [0-9.]*" gpurun_out/final_ours_n4.json | head -5 | tr '\n' ' '; grep -o "\"stage_ms\": {[^}]*}" gpurun_out/final_ours_n4.json | head -2
