# A/B harness used during round 2 (gpurun): parity tests, then bench lines of the kernel variants
python -m pytest tests/test_parity_gpu.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r2_t2.txt
python bench.py --steps 100 --warmup 10 --no-mapping --no-tracking --no-cpu > gpurun_out/r2_b_lane.json 2> gpurun_out/r2_b_lane.err
