#!/bin/bash
for v in g16_c3 g12_c5 g8_c8; do echo "variant $v"; VLR_ENGINE_LIB=$PWD/build_variants/libvlr_$v.so python scripts/prof_wave.py 65536 3 2 2>&1 | tail -1; done
echo baseline; python scripts/prof_wave.py 65536 3 2 2>&1 | tail -1
