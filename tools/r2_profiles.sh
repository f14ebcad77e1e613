#!/bin/bash
# Round-2 evidence pass on one B200: step timeline, ncu --set full of the distance GEMM and the masked-CE chain,
# compute-sanitizer logs of every kernel at small shapes.  Output: gpurun_out/ (summarised into profiles/ afterwards).
out=gpurun_out
mkdir -p $out
B="python bench.py --steps 30 --warmup 5 --skip-cpu --skip-e2e --skip-configs --skip-gpu-baseline"
timeout 300 $B --timeline $out/r2_timeline.txt > $out/r2_tl.json 2> $out/r2_tl.err; echo timeline rc $?
timeout 120 python scratch/ncu_assign.py > /dev/null 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tn -c 2 -f -o $out/assign_r2 python scratch/ncu_assign.py > $out/ncu_assign.log 2>&1; echo ncu assign rc $?
timeout 120 python scratch/ncu_ce2.py > /dev/null 2>&1 && timeout 900 ncu --set full --clock-control none --cache-control none --import-source on -k regex:'gemm_|ce_' --launch-skip 10 -c 6 -f -o $out/ce_r2 python scratch/ncu_ce2.py > $out/ncu_ce.log 2>&1; echo ncu ce rc $?
# compute-sanitizer: refused by this GPU pool ("compute-sanitizer is closed on this pool and stays closed", exit 86, 3 tools tried
# in round 2); tests/test_gpu_canaries.py (guard bands around every buffer of every ABI entry point) stands in for memcheck.
