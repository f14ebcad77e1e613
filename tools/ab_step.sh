#!/bin/bash
# A/B of step variants through environment knobs of bench.py (DeviceStep): prints us/step per variant.
# usage: tools/ab_step.sh "VAR=1 VAR2=0" "VAR=0" ...
B="python bench.py --steps 30 --warmup 5 --skip-cpu --skip-e2e --skip-configs --skip-gpu-baseline"
for v in "$@"; do
  for i in 1 2; do
    env $v timeout 300 $B 2>/dev/null | grep "^{" | python -c "
import json,sys;d=json.loads(sys.stdin.read());r=d['roofline'];print('$v', round(d['ms_per_step']*1e3,1), 'us  frac', round(r['step_frac_of_sustained'],3), ' gemm', round(r['kernel_us'],2), round(r['frac'],3))"
  done
done
