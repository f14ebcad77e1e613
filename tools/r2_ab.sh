#!/bin/bash
# A/B pass: stream priorities of the EMA / quantize chains.  Output: gpurun_out/ab_*.
out=gpurun_out; mkdir -p $out
B="python bench.py --steps 30 --warmup 5 --skip-cpu --skip-configs --skip-gpu-baseline --skip-e2e"
show() { python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); r=d['roofline']
    print(f"{sys.argv[2]:22s} step {d['ms_per_step']*1e3:7.1f} us frac {r['step_frac_of_sustained']:.3f} gemm {r['kernel_us']:6.2f}")
except Exception as ex:
    print(sys.argv[2], 'FAILED', ex)
PY
}
run() { name=$1; shift; env "$@" timeout 300 $B > $out/ab_$name.json 2> $out/ab_$name.err; show $out/ab_$name.json $name; }
run gside1 PERO_STEP_GATHER_SIDE=1
run gside0 PERO_STEP_GATHER_SIDE=0
run gside1_ema2 PERO_STEP_GATHER_SIDE=1 PERO_EMA_PRIO=-2
run gside0_ema2 PERO_STEP_GATHER_SIDE=0 PERO_EMA_PRIO=-2
run gside0_ema3 PERO_STEP_GATHER_SIDE=0 PERO_EMA_PRIO=-3
run gside1_ema3 PERO_STEP_GATHER_SIDE=1 PERO_EMA_PRIO=-3
run gside1_b PERO_STEP_GATHER_SIDE=1
PERO_STEP_GATHER_SIDE=1 timeout 300 $B --timeline $out/ab_timeline_gside1.txt > /dev/null 2>&1
tail -22 $out/ab_timeline_gside1.txt | cut -c1-110
PERO_STEP_GATHER_SIDE=1 PERO_EMA_PRIO=-3 timeout 300 $B --timeline $out/ab_timeline_gside1_ema3.txt > /dev/null 2>&1
tail -22 $out/ab_timeline_gside1_ema3.txt | cut -c1-110
