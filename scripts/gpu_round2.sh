#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
python scripts/prof_wave.py 65536 3 2>&1 | tail -3
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_wave.csv python scripts/prof_wave.py 65536 2 > gpurun_out/prof_wave.log 2>&1
tail -2 gpurun_out/prof_wave.log
# full capture of round 0 and round 3 of the second call (launch ids: 1 call = memset-free list: prep, 9 rounds, finish, generic = 12)
ncu --set full --clock-control none --import-source on -k regex:vlr_wave_round_kernel -s 9 -c 5 -o gpurun_out/ncu_wave_round -f python scripts/prof_wave.py 65536 2 > gpurun_out/ncu_round.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:vlr_wave_prep_kernel -s 1 -c 1 -o gpurun_out/ncu_wave_prep -f python scripts/prof_wave.py 65536 2 > gpurun_out/ncu_prep.log 2>&1
ls -la gpurun_out
