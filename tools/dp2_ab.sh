#!/bin/bash
# 2-GPU A/B of the data-parallel step through environment variables; prints us/step per variant (twice each)
run() { for i in 1; do env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2955$i bench.py --gpus 2 --steps 30 --warmup 5 --skip-cpu --skip-configs --skip-gpu-baseline --skip-e2e 2>/dev/null | grep "^{" | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('$*', round(d['ms_per_step']*1e3,1))"; done; }
for v in "$@"; do run $v; done
