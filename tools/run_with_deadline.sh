#!/bin/bash
# usage: tools/run_with_deadline.sh <seconds> <logfile> <cmd...>
# Runs <cmd> in its own session/process group and kills the WHOLE group at the deadline (torchrun workers
# otherwise survive a killed launcher and keep the GPU box busy).
secs=$1; log=$2; shift 2
setsid "$@" > "$log" 2>&1 &
pid=$!
for ((i = 0; i < secs * 2; i++)); do
    if ! kill -0 "$pid" 2>/dev/null; then wait "$pid"; echo "exit=$? (finished)" >> "$log"; exit 0; fi
    sleep 0.5
done
echo "DEADLINE ${secs}s reached: killing process group $pid" >> "$log"
kill -TERM -- "-$pid" 2>/dev/null; sleep 3; kill -KILL -- "-$pid" 2>/dev/null
exit 124
