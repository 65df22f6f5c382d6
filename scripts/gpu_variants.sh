#!/bin/bash
for v in g20 g16 g24; do echo "variant $v"; VLR_ENGINE_LIB=$PWD/build_variants/libvlr_$v.so python scripts/prof_wave.py 65536 3 2 2>&1 | tail -1; done
echo baseline; python scripts/prof_wave.py 65536 3 2 2>&1 | tail -1
