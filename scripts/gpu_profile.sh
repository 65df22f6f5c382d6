#!/bin/bash
# The GPU session behind profiles/ (run under gpurun, one GPU): parity tests, smoke, bench lines, per-kernel launch list
# with DRAM traffic, and ncu --set full captures of the dominant kernels. Numbers printed under ncu are never bench values.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/bench_cfg2_1m.json 2> gpurun_out/bench_cfg2_1m.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2>/dev/null
python bench.py --config 5 --loci 200000 > gpurun_out/bench_cfg5_200k.json 2>/dev/null
python bench.py --config 3 > gpurun_out/bench_cfg3_1m.json 2>/dev/null
# launch list + DRAM bytes of one device-entry call over 65536 config-2 loci (second call of the script)
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 200 --csv \
    --log-file gpurun_out/traffic_wave.csv python scripts/prof_wave.py 65536 2 > gpurun_out/prof_wave.log 2>&1
# full captures: round 1 of the second call (warp-per-4-lcs round kernel), the coefficient kernel
ncu --set full --clock-control none --import-source on -k regex:vlr_wave_round_warp_kernel -s 12 -c 1 \
    -o gpurun_out/ncu_wave_roundw -f python scripts/prof_wave.py 65536 2 > gpurun_out/ncu_round.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:vlr_wave_coef_kernel -s 1 -c 1 \
    -o gpurun_out/ncu_wave_coef_final -f python scripts/prof_wave.py 65536 2 > gpurun_out/ncu_coef.log 2>&1
# read here with: ncu -i gpurun_out/X.ncu-rep --page raw --csv   /   --page source --csv --print-source sass
# contamination estimator: launch list and one full capture of the likelihood kernel (second launch =
# the 100 000-observation call; the first is the 10-observation warm-up)
python tests/tools/prof_contamination.py 100000 2>&1 | tail -1
ncu --metrics gpu__time_duration.sum --clock-control none -c 30 --csv --log-file gpurun_out/launches_contam.csv \
    python tests/tools/prof_contamination.py 100000 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:vlr_contam_likelihood_kernel -s 1 -c 1 \
    -o gpurun_out/ncu_contam -f python tests/tools/prof_contamination.py 100000 > gpurun_out/ncu_contam.log 2>&1
