#!/bin/bash
# round 2, last session (one GPU): compute-sanitizer over every round-2 kernel, the full -m gpu suite, the default bench
# line, and the captures of the final build behind profiles/*_r2f_*. Numbers printed under ncu are never bench values.
mkdir -p gpurun_out
{ cat /sys/devices/system/node/online; nvidia-smi topo -m; grep -i allowed_list /proc/self/status
  VLR_NUMA_DEBUG=1 python -c "from varlociraptor_b200 import engine; import numpy as np; a = engine.pinned_empty(1 << 20, np.float32); print('pinned ok', a.nbytes)"; } > gpurun_out/r2f_numa.log 2>&1
tail -4 gpurun_out/r2f_numa.log
for tool in memcheck synccheck; do
  timeout 240 compute-sanitizer --tool $tool --error-exitcode 9 python scripts/sanitize_r2.py > gpurun_out/r2f_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|bitwise" gpurun_out/r2f_$tool.log | tail -4
done
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/r2f_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2f_pytest_gpu.log
timeout 600 python bench.py > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2f_bench_ref.json 2>> gpurun_out/r2f_bench.err
# launch lists + DRAM bytes of one device-entry call (second call of the script)
for c in 2 5 3; do
  n=65536; [ $c = 5 ] && n=16384
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
      --log-file gpurun_out/r2f_traffic_cfg$c.csv python scripts/prof_wave.py $n 2 $c > gpurun_out/r2f_prof_cfg$c.log 2>&1
  python scripts/sum_launches.py gpurun_out/r2f_traffic_cfg$c.csv > gpurun_out/r2f_launches_cfg$c.txt
done
ncu --set full --clock-control none --import-source on -k regex:vlr_wave_resident_kernel -s 1 -c 1 \
    -o gpurun_out/ncu_r2f_resident_final -f python scripts/prof_wave.py 65536 2 2 > gpurun_out/r2f_ncu_res.log 2>&1
cat gpurun_out/r2f_launches_cfg2.txt
cut -c1-400 gpurun_out/r2f_bench.json
