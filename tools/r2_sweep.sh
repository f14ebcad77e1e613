#!/bin/bash
# Round-2 A/B sweep on one B200 (dev build of the library: PERO_* knobs are honoured).  Output: gpurun_out/r2s_*.
out=gpurun_out
mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > $out/r2s_tests.log
tail -3 $out/r2s_tests.log
B="python bench.py --steps 30 --warmup 5 --skip-cpu --skip-e2e --skip-configs --skip-gpu-baseline"
run() { name=$1; shift; env "$@" timeout 300 $B > $out/r2s_$name.json 2> $out/r2s_$name.err; python - "$out/r2s_$name.json" "$name" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); r=d['roofline']
    print(f"{sys.argv[2]:14s} step {d['ms_per_step']*1e3:7.1f} us  gemm {r['kernel_us']:6.2f} us frac {r['frac']:.3f}  step_frac {r['step_frac_of_sustained']:.3f} kernels {d['kernels_per_step']}")
except Exception as e:
    print(sys.argv[2], 'FAILED', e)
PY
}
run default PERO_X=0
run default2 PERO_X=0
run scatter_pdl PERO_SCATTER_PDL=1
run ema_hi PERO_EMA_PRIO=-1
run ema_hi2 PERO_EMA_PRIO=-1
PERO_EMA_PRIO=-1 timeout 300 $B --timeline $out/r2s_timeline.txt > $out/r2s_tl.json 2> $out/r2s_tl.err
tail -70 $out/r2s_timeline.txt
