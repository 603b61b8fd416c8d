#!/bin/bash
# Final single-GPU pass of the round: GPU test suite, launch lists of the final kernels, both bench arms.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/final_tests.txt
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,l1tex__data_pipe_lsu_wavefronts.sum
ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r02_launches_ours_C3.csv python profiles/prof_step.py C3 3 > gpurun_out/r02_prof1.log 2>&1
ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r02_launches_mapping_C3.csv python profiles/prof_mapping.py C3 3 > gpurun_out/r02_prof5.log 2>&1
timeout 400 python bench.py --impl reference --steps 60 --warmup 4 > gpurun_out/final_ref_n1.json 2> gpurun_out/final_ref_n1.err
timeout 600 python bench.py > gpurun_out/final_ours_n1.json 2> gpurun_out/final_ours_n1.err
tail -c 600 gpurun_out/final_ref_n1.json; echo; grep -o "\"ms_per_step\": [0-9.]*" gpurun_out/final_ours_n1.json | head -3
