#!/bin/bash
# A/B pass: e2e variants, hardware-queue count.  Output: gpurun_out/ab_*.
out=gpurun_out; mkdir -p $out
B="python bench.py --steps 30 --warmup 5 --skip-cpu --skip-configs --skip-gpu-baseline"
show() { python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); r=d['roofline']; e=d.get('e2e') or {}
    print(f"{sys.argv[2]:14s} step {d['ms_per_step']*1e3:7.1f} us frac {r['step_frac_of_sustained']:.3f} gemm {r['kernel_us']:6.2f} | e2e {e.get('ms_per_step')} all {e.get('ms_per_step_all')} blocking {e.get('ms_per_step_blocking_loss_read')}")
except Exception as ex:
    print(sys.argv[2], 'FAILED', ex)
PY
}
timeout 300 $B > $out/ab_e2e_default.json 2> $out/ab_e2e_default.err; show $out/ab_e2e_default.json default
PERO_E2E_ST=0 timeout 300 $B > $out/ab_e2e_mt.json 2> $out/ab_e2e_mt.err; show $out/ab_e2e_mt.json autograd_mt
for n in 32 32; do
CUDA_DEVICE_MAX_CONNECTIONS=$n timeout 300 $B --skip-e2e > $out/ab_conn$n.json 2> $out/ab_conn$n.err; show $out/ab_conn$n.json conn$n
done
CUDA_DEVICE_MAX_CONNECTIONS=32 timeout 300 $B --skip-e2e --timeline $out/ab_timeline_conn32.txt > /dev/null 2>&1
tail -22 $out/ab_timeline_conn32.txt | cut -c1-110
