#!/bin/bash
# 2-GPU check + bench (default layout) + the single-GPU tests touched by this change
mkdir -p gpurun_out
python -m pytest tests/test_parity_gpu.py tests/test_mapping_gpu.py -q -k "sharded_projection or fused_mapper or autograd_level" 2>&1 | tail -2
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_check.py C2 3 2>&1 | grep -E "MGPU|Error|error" | head
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 102 --warmup 6 --no-loop > gpurun_out/n2.log 2>&1
grep -o "\"ms_per_step\": [0-9.]*" gpurun_out/n2.log | head -4 | tr '\n' ' '; grep -o "\"stage_ms\": {[^}]*}" gpurun_out/n2.log | head -2
python profiles/push_probe.py 8 2>&1 | grep -E "world|Error|error"
