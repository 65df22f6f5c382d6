#!/bin/bash
# round 2, closing session (one GPU) after the last kernel change: full -m gpu suite, smoke, default bench line, and the
# capture of the all-Set lc kernel with its MAPQ table. Numbers printed under ncu are never bench values.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/r2i_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2i_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err; echo "bench rc=$?"
cut -c1-300 gpurun_out/r2i_bench.json
python scripts/ab_libs.py 131072 3 2 prev new 2>&1 | tail -2
python scripts/ab_libs.py 100000 3 5 prev new 2>&1 | tail -2
