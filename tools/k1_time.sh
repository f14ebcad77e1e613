#!/bin/bash
# distance GEMM + arg-min at configs[1] (roofline leg of bench.py) and the step time, production build
for i in 1 2; do
timeout 300 python bench.py --steps 30 --warmup 5 --skip-cpu --skip-configs --skip-gpu-baseline --skip-e2e 2>/dev/null | grep "^{" | python -c "
import json,sys;d=json.loads(sys.stdin.read());r=d['roofline'];print('step', round(d['ms_per_step']*1e3,1), 'us  frac', round(r['step_frac_of_sustained'],3), ' gemm', round(r['kernel_us'],2), round(r['kernel_us_min'],2), round(r['frac'],3))"
done
