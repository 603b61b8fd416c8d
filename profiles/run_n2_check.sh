#!/bin/bash
# 2-GPU check + bench (default layout)
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_check.py C2 2 2>&1 | grep -E "MGPU|Error|error" | head
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 102 --warmup 6 --no-loop --no-c4 > gpurun_out/n2.log 2> gpurun_out/n2.err
grep -E "bench:|Error|error" gpurun_out/n2.err | head -5
grep -o "\"ms_per_step\": [0-9.]*" gpurun_out/n2.log | head -4 | tr '\n' ' '; grep -o "\"e2e\": {[^}]*}" gpurun_out/n2.log | cut -c1-300; grep -o "\"mapping_iter\": {[^}]*}" gpurun_out/n2.log | cut -c1-200
