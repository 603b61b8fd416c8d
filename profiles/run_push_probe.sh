#!/bin/bash
mkdir -p gpurun_out
python profiles/push_probe.py 8 2>&1 | grep -E "world|Error|error"
python profiles/push_probe.py 2 2>&1 | grep -E "world|Error|error"
ncu --set full --clock-control none --import-source on -k regex:'k_push_rows|k_fold_inbox' -s 8 -c 2 -f -o gpurun_out/push_full python profiles/push_probe.py 8 > gpurun_out/push_full.log 2>&1
tail -2 gpurun_out/push_full.log
