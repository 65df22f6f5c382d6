#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
VLR_WAVE_DEBUG=1 python scripts/prof_wave.py 65536 1 2>&1 | tail -10
python scripts/prof_wave.py 65536 3 2>&1 | tail -2
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_wave.csv python scripts/prof_wave.py 65536 2 > gpurun_out/prof_wave.log 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/launches_wave.csv')))
hdr=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
cols=rows[hdr]
ki=cols.index('Kernel Name'); vi=cols.index('Metric Value')
print(" ".join("%s:%s"%(r[ki][19:24], r[vi]) for r in rows[hdr+1+15:] if len(r)>vi))
PY
