#!/bin/bash
# per-kernel durations of one emulated rank's plan at world 8, contiguous tile runs, two-pass projection
mkdir -p gpurun_out
PROBE_WORLDS=8 PROBE_LAYOUTS=bands ncu --metrics gpu__time_duration.sum --clock-control none -c 140 --csv \
  --log-file gpurun_out/plan_probe_launches.csv python profiles/shard_plan_probe.py C3 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/plan_probe_launches.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
for r in rows[60:90]:
    print(r[ki][:50], r[vi])
PY
