#!/bin/bash
# usage: tools/dp_sweep.sh <n_gpus>   — data-parallel bench under a few exchange-kernel settings (tuning aid)
n=${1:-2}
port=29600
run() {
    port=$((port + 1))
    echo "=== $*"
    tools/run_with_deadline.sh 120 gpurun_out/sweep_tmp.log env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $n \
        --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n --steps 30 --warmup 5 --skip-cpu --skip-e2e
    grep -o '"ms_per_step": [0-9.]*\|"exchange": "[^"]*"' gpurun_out/sweep_tmp.log | head -3
    grep -i "error\|Traceback\|DEADLINE" gpurun_out/sweep_tmp.log | head -5
}
run PERO_DP_CHUNKS=1
run PERO_DP_CHUNKS=2
run PERO_DP_CHUNKS=4
run PERO_DP_CHUNKS=2 PERO_PEER_BLOCKS=24
run PERO_DP_CHUNKS=2 PERO_PEER_BLOCKS=32 PERO_PEER_THREADS=128
run PERO_DP_CHUNKS=2 PERO_PEER_MULTICAST=0 PERO_PEER_THREADS=512 PERO_PEER_BLOCKS=48
