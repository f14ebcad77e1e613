#!/bin/bash
# Dev build (knobs + stamps) -> per-unit stamps of the config-2 distance GEMM -> production build restored.
make -C pero_pretraining_b200/csrc DEV=1 -j 8 > /dev/null 2>&1 || { echo "dev build failed"; exit 1; }
timeout 120 python scratch/timeline_c2.py "$@"
