#!/bin/bash
# ncu --set full captures: $1 = kernel regex, $2 = skip count, $3 = capture count, $4 = output name
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$1" -s "$2" -c "$3" \
   -o "gpurun_out/$4" -f python bench.py --steps 1 --warmup 1 --no-graph --no-cpu > "gpurun_out/ncu_$4.log" 2>&1
echo "=== ncu $4 rc=$?"; tail -3 "gpurun_out/ncu_$4.log"
