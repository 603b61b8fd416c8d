# A/B harness used during round 2 (gpurun): parity tests, then bench lines of the kernel variants
python -m pytest tests/test_parity_gpu.py -m gpu -q 2>&1 | tail -40 > gpurun_out/r2_t2.txt
python bench.py --steps 100 --warmup 10 --no-mapping --no-tracking --no-cpu > gpurun_out/r2_b_lane.json 2> gpurun_out/r2_b_lane.err
EGS_BWD_KERNEL=warp python bench.py --steps 100 --warmup 10 --no-mapping --no-tracking --no-cpu --no-e2e > gpurun_out/r2_b_warp.json 2> gpurun_out/r2_b_warp.err
