# Round-2 profiling pass (gpurun, 1 GPU): launch lists with device time + DRAM bytes, full captures of the compositing
# kernels (ours and the reference's), per-workload traffic of the dominant kernels.
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,l1tex__data_pipe_lsu_wavefronts.sum
ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r02_launches_ours_C3.csv python profiles/prof_step.py C3 3 > gpurun_out/r02_prof1.log 2>&1
ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r02_launches_reference_C3.csv python profiles/prof_reference.py C3 2 > gpurun_out/r02_prof2.log 2>&1
for w in C2 C4 C5; do
  ncu --metrics $M --clock-control none -k regex:k_render --csv --log-file gpurun_out/r02_launches_ours_$w.csv python profiles/prof_step.py $w 2 > gpurun_out/r02_prof_$w.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:k_render -s 4 -c 2 -f -o gpurun_out/r02_render_ours python profiles/prof_step.py C3 3 > gpurun_out/r02_prof3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:renderCUDA -s 2 -c 2 -f -o gpurun_out/r02_render_reference python profiles/prof_reference.py C3 2 > gpurun_out/r02_prof4.log 2>&1
ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r02_launches_mapping_C3.csv python profiles/prof_mapping.py C3 3 > gpurun_out/r02_prof5.log 2>&1
