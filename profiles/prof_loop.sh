for arm in ours reference; do
python - <<PY > gpurun_out/r2_loopprof_$arm.txt 2>&1
import cProfile, pstats, sys, io, runpy
sys.argv = ["ref_loop.py", "--arm", "$arm", "--config", "tum", "--frames", "20", "--out", "/tmp/x_$arm.npz"]
pr = cProfile.Profile()
pr.enable()
try:
    runpy.run_path("tests/ref_loop.py", run_name="__main__")
except SystemExit:
    pass
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45)
print(s.getvalue()[-9000:])
PY
done
