#!/bin/bash
# Final evidence pass of round 2 on one B200: full -m gpu suite, the default bench line, ncu launch list of one eager step,
# ncu --set full of the distance GEMM (-> profiles/roofline_traffic.json) and of the masked-CE chain.
out=gpurun_out; mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -6 > $out/final_tests.log; tail -2 $out/final_tests.log
timeout 900 python bench.py > $out/final_bench.json 2> $out/final_bench.err; echo bench rc $?
B="python bench.py --steps 2 --warmup 1 --no-graph --skip-cpu --skip-e2e --skip-configs --skip-gpu-baseline"
timeout 200 $B > /dev/null 2>&1 && timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/r2_launches.csv $B > $out/final_ncu.log 2>&1; echo launches rc $?
timeout 120 python scratch/ncu_assign.py > /dev/null 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tn -c 2 -f -o $out/assign_r2 python scratch/ncu_assign.py > $out/ncu_assign.log 2>&1; echo ncu assign rc $?
timeout 120 python scratch/ncu_ce2.py > /dev/null 2>&1 && timeout 900 ncu --set full --clock-control none --cache-control none --import-source on -k regex:'gemm_|ce_' --launch-skip 10 -c 6 -f -o $out/ce_r2 python scratch/ncu_ce2.py > $out/ncu_ce.log 2>&1; echo ncu ce rc $?
timeout 300 python bench.py --steps 30 --warmup 5 --skip-cpu --skip-e2e --skip-configs --skip-gpu-baseline --timeline $out/r2_timeline.txt > /dev/null 2>&1; echo timeline rc $?
