#!/bin/bash
# 4-GPU check + bench exactly as the driver launches it
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_check.py C2 1 2>&1 | grep -E "MGPU|Error|error|assert" | head
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --steps 100 --warmup 5 > gpurun_out/final_ours_n4.json 2> gpurun_out/final_ours_n4.err
grep -E "bench:|Error|error" gpurun_out/final_ours_n4.err | head -5
grep -o "\"ms_per_step\": [0-9.]*" gpurun_out/final_ours_n4.json | head -5 | tr '\n' ' '; grep -o "\"stage_ms\": {[^}]*}" gpurun_out/final_ours_n4.json | head -1
timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 4 --steps 20 --warmup 3 --no-loop 2>/dev/null | cut -c1-300
