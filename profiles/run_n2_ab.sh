#!/bin/bash
# 2-GPU check + bench, both tile layouts
mkdir -p gpurun_out
python -m pytest tests/test_parity_gpu.py -q -k "sharded_projection" 2>&1 | tail -2
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_check.py C2 3 2>&1 | grep -E "MGPU|Error|error" | head
for lay in rows bands; do
  extra="--no-c4"; [ $lay = bands ] && extra=""
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 102 --warmup 6 --no-loop --partition $lay $extra > gpurun_out/n2_$lay.log 2>&1
  echo $lay; grep -o "\"ms_per_step\": [0-9.]*" gpurun_out/n2_$lay.log | head -4 | tr '\n' ' '; grep -o "\"stage_ms\": {[^}]*}" gpurun_out/n2_$lay.log | head -2
done
