#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/bench_cfg2_1m.json 2> gpurun_out/bench_cfg2_1m.err; tail -c 400 gpurun_out/bench_cfg2_1m.json; tail -2 gpurun_out/bench_cfg2_1m.err
python bench.py --config 5 --loci 100000 > gpurun_out/bench_cfg5_100k.json 2> gpurun_out/bench_cfg5.err; python -c "
import json; d=json.load(open('gpurun_out/bench_cfg5_100k.json')); print('cfg5', d['value'], d['e2e']['value'], d.get('cpu_baseline'))"
python bench.py --config 3 > gpurun_out/bench_cfg3_1m.json 2> gpurun_out/bench_cfg3.err; python -c "
import json; d=json.load(open('gpurun_out/bench_cfg3_1m.json')); print('cfg3', d['value'], d['e2e']['value'], d.get('cpu_baseline'))"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 200 --csv --log-file gpurun_out/traffic_wave.csv python scripts/prof_wave.py 65536 2 > gpurun_out/prof_wave.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:vlr_wave_round_kernel -s 12 -c 1 -o gpurun_out/ncu_wave_round_final -f python scripts/prof_wave.py 65536 2 > gpurun_out/ncu_round.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:vlr_wave_coef_kernel -s 1 -c 1 -o gpurun_out/ncu_wave_coef_final -f python scripts/prof_wave.py 65536 2 > gpurun_out/ncu_coef.log 2>&1
ls gpurun_out | tail -20
